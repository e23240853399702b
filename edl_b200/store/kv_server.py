"""A small self-contained key-value server with etcd-v3 semantics.

The reference keeps ALL cluster state in an external etcd (python/edl/utils/constants.py:15-39,
discovery/etcd_client.py) -- ``etcd3`` / ``etcd`` are not installable here, so the framework ships
its own store with the exact primitives EDL relies on:

* revisions (global, per-key create/mod revision, version),
* leases with TTL, keep-alive and revoke; keys attached to a lease vanish when it expires,
* ``put_if_not_exists`` and general compare-and-swap transactions (leader election, leader-guarded
  writes: utils/leader_pod.py:57-66, utils/cluster_generator.py:223-250, utils/state.py:186-200),
* prefix range reads returning the header revision (race-free get-then-watch,
  discovery/etcd_client.py:103-120),
* prefix watches from a start revision, delivered as ordered put/delete events.

Wire format: 4-byte big-endian length + msgpack map.  Requests carry ``id``; responses echo it;
watch events are pushed with ``watch_id``.  Run standalone with
``python -m edl_b200.store.kv_server --port 2379``.
"""
from __future__ import annotations

import argparse
import os
import logging
import queue
import socket
import socketserver
import struct
import threading
import time
from collections import deque
from typing import Dict, List, Optional, Tuple

import msgpack

logger = logging.getLogger("edl.store")

_HDR = struct.Struct("!I")


def send_msg(sock: socket.socket, obj, lock: Optional[threading.Lock] = None) -> None:
    data = msgpack.packb(obj, use_bin_type=True)
    buf = _HDR.pack(len(data)) + data
    if lock is not None:
        with lock:
            sock.sendall(buf)
    else:
        sock.sendall(buf)


def _recv_exact(sock: socket.socket, n: int) -> Optional[bytes]:
    chunks = []
    while n > 0:
        b = sock.recv(n)
        if not b:
            return None
        chunks.append(b)
        n -= len(b)
    return b"".join(chunks)


def recv_msg(sock: socket.socket):
    hdr = _recv_exact(sock, 4)
    if hdr is None:
        return None
    (n,) = _HDR.unpack(hdr)
    data = _recv_exact(sock, n)
    if data is None:
        return None
    return msgpack.unpackb(data, raw=False)


class _KeyValue:
    __slots__ = ("value", "create_rev", "mod_rev", "version", "lease")

    def __init__(self, value, create_rev, mod_rev, version, lease):
        self.value, self.create_rev, self.mod_rev, self.version, self.lease = (
            value, create_rev, mod_rev, version, lease)

    def to_wire(self, key):
        return {"key": key, "value": self.value, "create_revision": self.create_rev,
                "mod_revision": self.mod_rev, "version": self.version, "lease": self.lease}


class _Lease:
    __slots__ = ("id", "ttl", "expiry", "keys")

    def __init__(self, lid, ttl):
        self.id, self.ttl, self.expiry, self.keys = lid, ttl, time.monotonic() + ttl, set()


class KVState:
    """The data model; every public method takes the big lock."""

    def __init__(self, history: int = 100000):
        self.lock = threading.RLock()
        self.kv: Dict[str, _KeyValue] = {}
        self.rev = 1
        self.leases: Dict[int, _Lease] = {}
        self._next_lease = int(time.time() * 1000) % (1 << 30) + 1
        self.events = deque(maxlen=history)      # (rev, type, key, kv-wire)
        self.watchers: Dict[Tuple[int, int], "_Watcher"] = {}

    # -- durability ------------------------------------------------------------------------
    def snapshot(self) -> bytes:
        """Consistent image of keys, revision counter and leases (remaining TTLs); watchers / history are not
        part of it -- clients re-watch from their last seen revision and get a compaction-style full resync."""
        with self.lock:
            now = time.monotonic()
            return msgpack.packb({
                "rev": self.rev, "next_lease": self._next_lease,
                "kv": [[k, v.value, v.create_rev, v.mod_rev, v.version, v.lease] for k, v in self.kv.items()],
                "leases": [[l.id, l.ttl, max(0.0, l.expiry - now)] for l in self.leases.values()],
            }, use_bin_type=True)

    def restore(self, blob: bytes, lease_grace: float = 0.0):
        """Load a snapshot.  Leases get at least ``lease_grace`` seconds so that clients that survived the
        store restart can refresh them before their keys expire."""
        d = msgpack.unpackb(blob, raw=False)
        with self.lock:
            self.rev = int(d["rev"])
            self._next_lease = max(self._next_lease, int(d["next_lease"]))
            self.kv = {k: _KeyValue(val, cr, mr, ver, lease) for k, val, cr, mr, ver, lease in d["kv"]}
            self.leases = {}
            now = time.monotonic()
            for lid, ttl, remaining in d["leases"]:
                le = _Lease(lid, ttl)
                le.expiry = now + max(remaining, lease_grace)
                self.leases[lid] = le
            for k, v in self.kv.items():
                if v.lease:
                    if v.lease in self.leases:
                        self.leases[v.lease].keys.add(k)
                    else:
                        v.lease = 0

    # -- mutations (call with lock held) ---------------------------------------------------
    def _emit(self, typ: str, key: str, wire: dict):
        ev = {"type": typ, "key": key, "kv": wire, "revision": self.rev}
        self.events.append(ev)
        for w in list(self.watchers.values()):
            if key.startswith(w.prefix) and (w.end is None or key < w.end):
                w.push([ev], self.rev)

    def _put(self, key: str, value: bytes, lease: int) -> dict:
        if lease and lease not in self.leases:
            raise KeyError("lease %d not found" % lease)
        self.rev += 1
        old = self.kv.get(key)
        if old is not None and old.lease and old.lease != lease and old.lease in self.leases:
            self.leases[old.lease].keys.discard(key)
        if old is None:
            cur = _KeyValue(value, self.rev, self.rev, 1, lease)
        else:
            cur = _KeyValue(value, old.create_rev, self.rev, old.version + 1, lease)
        self.kv[key] = cur
        if lease:
            self.leases[lease].keys.add(key)
        wire = cur.to_wire(key)
        self._emit("put", key, wire)
        return wire

    def _delete(self, key: str) -> int:
        old = self.kv.pop(key, None)
        if old is None:
            return 0
        self.rev += 1
        if old.lease and old.lease in self.leases:
            self.leases[old.lease].keys.discard(key)
        self._emit("delete", key, {"key": key, "value": b"", "create_revision": 0,
                                   "mod_revision": self.rev, "version": 0, "lease": 0})
        return 1

    def _range(self, prefix: str, end: Optional[str]) -> List[dict]:
        keys = sorted(k for k in self.kv if k.startswith(prefix) and (end is None or k < end))
        return [self.kv[k].to_wire(k) for k in keys]

    def _compare(self, c: dict) -> bool:
        kv = self.kv.get(c["key"])
        target = c.get("target", "value")
        op = c.get("op", "==")
        if target == "version":
            lhs, rhs = (kv.version if kv else 0), int(c.get("value", 0))
        elif target == "create":
            lhs, rhs = (kv.create_rev if kv else 0), int(c.get("value", 0))
        elif target == "mod":
            lhs, rhs = (kv.mod_rev if kv else 0), int(c.get("value", 0))
        else:
            if kv is None:
                return op == "!="
            lhs, rhs = kv.value, c.get("value", b"")
        return {"==": lhs == rhs, "!=": lhs != rhs, ">": lhs > rhs, "<": lhs < rhs}[op]

    def _apply(self, op: dict):
        t = op["op"]
        if t == "put":
            return {"put": self._put(op["key"], op.get("value", b""), int(op.get("lease", 0)))}
        if t == "delete":
            if op.get("prefix"):
                n = sum(self._delete(k) for k in sorted(k for k in self.kv if k.startswith(op["key"])))
            else:
                n = self._delete(op["key"])
            return {"deleted": n}
        if t == "get":
            if op.get("prefix"):
                return {"kvs": self._range(op["key"], op.get("end"))}
            kv = self.kv.get(op["key"])
            return {"kvs": [kv.to_wire(op["key"])] if kv else []}
        raise ValueError("unknown op %r" % t)

    # -- request handling ------------------------------------------------------------------
    def handle(self, req: dict, conn: "_Conn") -> dict:
        m = req.get("method")
        with self.lock:
            if m == "put":
                if req.get("if_not_exists") and req["key"] in self.kv:
                    return {"ok": True, "succeeded": False, "revision": self.rev}
                wire = self._put(req["key"], req.get("value", b""), int(req.get("lease", 0)))
                return {"ok": True, "succeeded": True, "revision": self.rev, "kv": wire}
            if m == "get":
                return {"ok": True, "revision": self.rev, **self._apply({"op": "get", **req})}
            if m == "delete":
                return {"ok": True, "revision": self.rev, **self._apply({"op": "delete", **req})}
            if m == "txn":
                ok = all(self._compare(c) for c in req.get("compare", []))
                results = [self._apply(op) for op in (req.get("success", []) if ok else req.get("failure", []))]
                return {"ok": True, "succeeded": ok, "revision": self.rev, "results": results}
            if m == "lease_grant":
                lid = int(req.get("lease_id") or 0) or self._next_lease
                self._next_lease = max(self._next_lease, lid) + 1
                self.leases[lid] = _Lease(lid, float(req["ttl"]))
                return {"ok": True, "lease": lid, "ttl": float(req["ttl"])}
            if m == "lease_keepalive":
                le = self.leases.get(int(req["lease"]))
                if le is None:
                    return {"ok": True, "ttl": 0}
                le.expiry = time.monotonic() + le.ttl
                return {"ok": True, "ttl": le.ttl}
            if m == "lease_revoke":
                self._revoke(int(req["lease"]))
                return {"ok": True}
            if m == "lease_ttl":
                le = self.leases.get(int(req["lease"]))
                return {"ok": True, "ttl": max(0.0, le.expiry - time.monotonic()) if le else -1,
                        "keys": sorted(le.keys) if le else []}
            if m == "watch":
                wid = int(req["watch_id"])
                w = _Watcher(conn, wid, req["key"], req.get("end"))
                start = int(req.get("start_revision") or 0)
                if start:
                    backlog = [e for e in self.events if e["revision"] >= start and
                               e["key"].startswith(w.prefix) and (w.end is None or e["key"] < w.end)]
                    if backlog:
                        w.push(backlog, self.rev)
                self.watchers[(conn.cid, wid)] = w
                return {"ok": True, "watch_id": wid, "revision": self.rev}
            if m == "cancel_watch":
                self.watchers.pop((conn.cid, int(req["watch_id"])), None)
                return {"ok": True}
            if m == "status":
                return {"ok": True, "revision": self.rev, "keys": len(self.kv), "leases": len(self.leases)}
        return {"ok": False, "error": "unknown method %r" % m}

    def _revoke(self, lid: int):
        le = self.leases.pop(lid, None)
        if le is None:
            return
        for k in sorted(le.keys):        # deterministic (key order): watchers of both server builds see the same stream
            self._delete(k)

    def expire(self):
        now = time.monotonic()
        with self.lock:
            for lid in [l.id for l in self.leases.values() if l.expiry <= now]:
                logger.debug("lease %d expired", lid)
                self._revoke(lid)

    def drop_conn(self, cid: int):
        with self.lock:
            for k in [k for k in self.watchers if k[0] == cid]:
                self.watchers.pop(k, None)


class _Watcher:
    def __init__(self, conn, wid, prefix, end):
        self.conn, self.wid, self.prefix, self.end = conn, wid, prefix, end

    def push(self, events, rev):
        self.conn.send({"watch_id": self.wid, "events": events, "revision": rev})


class _Conn:
    """One client connection.  Everything that goes out -- responses and watch events, in the order they were
    produced -- is queued and written by the connection's own writer thread: ``push`` is called under the state lock,
    and a client that stops reading (a paused process, a dead network path) must stall only itself, never the store.
    A connection whose backlog exceeds ``MAX_BACKLOG`` messages is dropped (its client reconnects and re-watches)."""

    _next = 0
    MAX_BACKLOG = 20000

    def __init__(self, sock):
        _Conn._next += 1
        self.cid = _Conn._next
        self.sock = sock
        self.closed = False
        self._q = queue.Queue()
        self._writer = threading.Thread(target=self._write_loop, daemon=True, name="kv-conn-writer")
        self._writer.start()

    def send(self, obj):
        if self.closed:
            return
        if self._q.qsize() > self.MAX_BACKLOG:
            logger.warning("connection %d does not read its messages (%d queued): dropping it", self.cid, self._q.qsize())
            self.close()
            return
        self._q.put(msgpack.packb(obj, use_bin_type=True))

    def _write_loop(self):
        while True:
            data = self._q.get()
            if data is None:
                return
            try:
                self.sock.sendall(_HDR.pack(len(data)) + data)
            except OSError:
                self.closed = True
                return

    def close(self):
        if not self.closed:
            self.closed = True
            try:
                self.sock.shutdown(socket.SHUT_RDWR)      # wakes the reader (and a writer blocked in sendall)
            except OSError:
                pass
        self._q.put(None)


class _Handler(socketserver.BaseRequestHandler):
    def handle(self):
        self.request.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        conn = _Conn(self.request)
        state: KVState = self.server.state
        try:
            while True:
                req = recv_msg(self.request)
                if req is None:
                    break
                try:
                    resp = state.handle(req, conn)
                except Exception as e:  # noqa: BLE001 - errors travel back to the client
                    resp = {"ok": False, "error": "%s: %s" % (type(e).__name__, e)}
                resp["id"] = req.get("id")
                conn.send(resp)
                if conn.closed:
                    break
        except OSError:
            pass
        finally:
            state.drop_conn(conn.cid)
            conn.close()


class _Server(socketserver.ThreadingTCPServer):
    allow_reuse_address = True
    daemon_threads = True
    request_queue_size = 256


class KVServer:
    """In-process handle: ``KVServer(port=0).start()``; ``.endpoint`` is ``"127.0.0.1:port"``."""

    def __init__(self, host: str = "127.0.0.1", port: int = 0, data_dir: Optional[str] = None,
                 snapshot_interval: float = 2.0):
        self.state = KVState()
        # optional durability (etcd keeps a data dir; a job whose coordination store restarts should not
        # lose its cluster / state / job-status keys): periodic atomic snapshots, reloaded on start
        self.data_dir, self.snapshot_interval = data_dir, snapshot_interval
        self._snap_rev = -1
        if data_dir:
            os.makedirs(data_dir, exist_ok=True)
            path = os.path.join(data_dir, "snapshot.bin")
            if os.path.exists(path):
                with open(path, "rb") as f:
                    self.state.restore(f.read(), lease_grace=10.0)
                self._snap_rev = self.state.rev
                logger.info("restored %d keys at revision %d from %s", len(self.state.kv), self.state.rev, path)
        self.server = _Server((host, port), _Handler)
        self.server.state = self.state
        self.host, self.port = self.server.server_address[:2]
        self._stop = threading.Event()
        self._threads: List[threading.Thread] = []

    @property
    def endpoint(self) -> str:
        return "%s:%d" % (self.host, self.port)

    def start(self) -> "KVServer":
        t = threading.Thread(target=self.server.serve_forever, kwargs={"poll_interval": 0.1},
                             name="kv-serve", daemon=True)
        t.start()
        e = threading.Thread(target=self._expire_loop, name="kv-expire", daemon=True)
        e.start()
        self._threads = [t, e]
        if self.data_dir:
            sn = threading.Thread(target=self._snapshot_loop, name="kv-snapshot", daemon=True)
            sn.start()
            self._threads.append(sn)
        return self

    def _expire_loop(self):
        while not self._stop.wait(0.1):
            self.state.expire()

    def save_snapshot(self):
        if not self.data_dir or self.state.rev == self._snap_rev:
            return
        blob, rev = self.state.snapshot(), self.state.rev
        tmp = os.path.join(self.data_dir, "snapshot.bin.tmp")
        with open(tmp, "wb") as f:
            f.write(blob)
            f.flush()
            os.fsync(f.fileno())
        os.replace(tmp, os.path.join(self.data_dir, "snapshot.bin"))      # atomic: readers never see a torn file
        self._snap_rev = rev

    def _snapshot_loop(self):
        while not self._stop.wait(self.snapshot_interval):
            try:
                self.save_snapshot()
            except OSError as e:
                logger.warning("snapshot failed: %s", e)

    def stop(self):
        self._stop.set()
        self.server.shutdown()
        self.server.server_close()
        try:
            self.save_snapshot()
        except OSError:
            pass

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()


def main(argv=None):
    ap = argparse.ArgumentParser(description="edl_b200 KV store (etcd-v3 semantics)")
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=2379)
    ap.add_argument("--log_level", type=int, default=20)
    ap.add_argument("--data_dir", default=None, help="keep periodic snapshots here and reload them on start")
    ap.add_argument("--snapshot_interval", type=float, default=2.0)
    ap.add_argument("--native", action="store_true",
                    help="exec the C++ / epoll build of this server (store/native/kv_server.cpp) instead")
    args = ap.parse_args(argv)
    logging.basicConfig(level=args.log_level)
    if args.native or os.environ.get("EDL_KV_NATIVE", "0") == "1":
        from . import native_server

        cmd = [native_server.build(), "--host", args.host, "--port", str(args.port),
               "--snapshot_interval", str(args.snapshot_interval)]
        if args.data_dir:
            cmd += ["--data_dir", args.data_dir]
        os.execv(cmd[0], cmd)
    srv = KVServer(args.host, args.port, args.data_dir, args.snapshot_interval).start()
    logger.info("kv store listening on %s", srv.endpoint)
    try:
        while True:
            time.sleep(3600)
    except KeyboardInterrupt:
        srv.stop()


if __name__ == "__main__":
    main()
