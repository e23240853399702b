#!/usr/bin/env python
"""Requests/s of the coordination store: Python server (socketserver threads) vs the native C++ / epoll server,
same Python clients.  `python tools/bench_store.py [--clients 8] [--ops 4000]`"""
import argparse
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edl_b200.store import KVClient, KVServer, NativeKVServer, native_server  # noqa: E402


def run(server_cls, clients, ops):
    srv = server_cls().start()
    try:
        lat = []

        def worker(i):
            c = KVClient(srv.endpoint)
            lease = c.lease(30)
            t0 = time.perf_counter()
            for k in range(ops):
                key = "/bench/%d/%d" % (i, k % 64)
                if k % 4 == 0:
                    c.put(key, b"x" * 64, lease.id)
                elif k % 4 == 1:
                    c.get(key)
                elif k % 4 == 2:
                    c.txn([{"key": key, "target": "version", "op": ">", "value": 0}], [{"op": "put", "key": key, "value": "y"}])
                else:
                    c.get_prefix("/bench/%d/" % i)
            lat.append((time.perf_counter() - t0) / ops)
            c.close()

        watcher = KVClient(srv.endpoint)
        seen = [0]
        watcher.add_watch_prefix_callback("/bench/", lambda evs, rev: seen.__setitem__(0, seen[0] + len(evs)))
        ts = [threading.Thread(target=worker, args=(i,)) for i in range(clients)]
        t0 = time.perf_counter()
        [t.start() for t in ts]
        [t.join() for t in ts]
        dt = time.perf_counter() - t0
        time.sleep(0.3)
        watcher.close()
        return {"requests_per_s": clients * ops / dt, "mean_latency_us": 1e6 * sum(lat) / len(lat), "watch_events": seen[0]}
    finally:
        srv.stop()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--clients", type=int, default=8)
    ap.add_argument("--ops", type=int, default=4000)
    a = ap.parse_args()
    out = {"python": run(KVServer, a.clients, a.ops)}
    if native_server.available():
        out["native"] = run(NativeKVServer, a.clients, a.ops)
    print(json.dumps(out, indent=1))
