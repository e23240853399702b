"""KV store semantics (the substrate the reference gets from etcd v3) + EtcdClient surface.
Mirrors python/edl/tests/unittests/etcd_client_test.py with second-scale TTLs."""
import threading
import time

import pytest

from edl_b200.store import KVClient, KVServer


def test_put_get_txn_lease_watch(kv_server):
    c = KVClient([kv_server.endpoint])
    events = []
    c2 = KVClient(kv_server.endpoint)
    c2.add_watch_prefix_callback("/a/", lambda evs, rev: events.extend(evs))
    lease = c.lease(0.6)
    assert c.put_if_not_exists("/a/x", b"1", lease.id)
    assert not c.put_if_not_exists("/a/x", b"2")
    v, meta = c.get("/a/x")
    assert v == b"1" and meta["version"] == 1 and meta["lease"] == lease.id
    ok, _ = c.txn([{"key": "/a/x", "value": b"1"}], [{"op": "put", "key": "/a/y", "value": "ok"}])
    assert ok and c.get("/a/y")[0] == b"ok"
    ok, _ = c.txn([{"key": "/a/x", "value": b"nope"}], [{"op": "put", "key": "/a/y", "value": "bad"}],
                  [{"op": "put", "key": "/a/z", "value": "else"}])
    assert not ok and c.get("/a/y")[0] == b"ok" and c.get("/a/z")[0] == b"else"
    kvs, rev = c.get_prefix("/a/")
    assert [kv["key"] for kv in kvs] == ["/a/x", "/a/y", "/a/z"] and rev >= 4
    for _ in range(3):
        time.sleep(0.3)
        assert lease.refresh() > 0
    assert c.get("/a/x")[0] == b"1"
    time.sleep(1.0)
    assert c.get("/a/x")[0] is None  # lease expired -> key gone
    time.sleep(0.1)
    kinds = [(e["type"], e["key"]) for e in events]
    assert ("put", "/a/x") in kinds and ("delete", "/a/x") in kinds
    assert c.delete_prefix("/a/") == 2


def test_watch_from_revision_replays_history(kv_server):
    c = KVClient(kv_server.endpoint)
    c.put("/w/a", "1")
    _, rev = c.get_prefix("/w/")
    c.put("/w/b", "2")
    c.delete("/w/a")
    got = []
    c.add_watch_prefix_callback("/w/", lambda evs, r: got.extend(evs), start_revision=rev + 1)
    time.sleep(0.2)
    assert [(e["type"], e["key"]) for e in got] == [("put", "/w/b"), ("delete", "/w/a")]


def test_lock_is_exclusive(kv_server):
    c1, c2 = KVClient(kv_server.endpoint), KVClient(kv_server.endpoint)
    order = []

    def worker(c, name):
        with c.lock("/lock/k", ttl=5):
            order.append(name + "+")
            time.sleep(0.2)
            order.append(name + "-")

    ts = [threading.Thread(target=worker, args=(c, n)) for c, n in ((c1, "a"), (c2, "b"))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert order in (["a+", "a-", "b+", "b-"], ["b+", "b-", "a+", "a-"])


def test_client_reconnects(kv_server):
    c = KVClient(kv_server.endpoint)
    c.put("/r/a", "1")
    c._sock.close()  # simulate a dropped connection
    assert c.get("/r/a")[0] == b"1"


def test_etcd_client_surface(etcd):
    assert etcd.set_server_not_exists("svc", "127.0.0.1:1", "info1", ttl=1)
    assert etcd.set_server_not_exists("svc", "127.0.0.1:2", "info2", ttl=1)
    assert not etcd.set_server_not_exists("svc", "127.0.0.1:1", "other", ttl=1, timeout=0.2)
    servers, rev = etcd.get_service_with_revision("svc")
    assert sorted(s.server for s in servers) == ["127.0.0.1:1", "127.0.0.1:2"]
    seen = {"add": [], "rm": []}
    etcd.watch_service("svc", lambda add, rm: (seen["add"].extend(s.server for s in add),
                                               seen["rm"].extend(s.server for s in rm)), start_revision=rev + 1)
    # keep :1 alive by refreshing, let :2 expire
    for _ in range(6):
        time.sleep(0.3)
        etcd.refresh("svc", "127.0.0.1:1")
    assert [s.server for s in etcd.get_service("svc")] == ["127.0.0.1:1"]
    assert seen["rm"] == ["127.0.0.1:2"]
    etcd.set_server_permanent("svc", "perm", "x")
    time.sleep(1.5)
    assert "perm" in [s.server for s in etcd.get_service("svc")]
    etcd.remove_service("svc")
    assert etcd.get_service("svc") == []


def test_snapshot_restart_keeps_keys_revisions_and_leases(tmp_path):
    import time
    from edl_b200.store.client import KVClient
    from edl_b200.store.kv_server import KVServer

    d = str(tmp_path / "kvdata")
    srv = KVServer(port=0, data_dir=d, snapshot_interval=0.1).start()
    c = KVClient([srv.endpoint])
    c.connect()
    c.put("/job/status", b"RUNNING")
    lease = c.lease(30.0)
    c.put("/job/pods/a", b"pod-a", lease.id)
    c.put("/job/status", b"SUCCEED")
    _, meta = c.get("/job/status")
    rev, version = meta["mod_revision"], meta["version"]
    port = srv.port
    c.close()
    srv.stop()                                   # final snapshot on stop
    srv2 = KVServer(port=port, data_dir=d).start()
    try:
        c2 = KVClient([srv2.endpoint])
        c2.connect()
        v, meta = c2.get("/job/status")
        assert v == b"SUCCEED" and meta["mod_revision"] == rev and meta["version"] == version
        v, meta = c2.get("/job/pods/a")
        assert v == b"pod-a" and meta["lease"] == lease.id
        assert c2.lease_keepalive(lease.id) > 0            # the lease survived and can be refreshed
        r = c2.put("/job/new", b"x")
        assert r["kv"]["mod_revision"] > rev if "kv" in r else True
        c2.close()
    finally:
        srv2.stop()


@pytest.mark.parametrize("first,second", [("python", "native"), ("native", "python")])
def test_python_and_native_servers_share_the_snapshot_format(tmp_path, first, second):
    """Either server can take over the other's data directory (same msgpack snapshot, same lease grace)."""
    from edl_b200.store import NativeKVServer, native_server

    if not native_server.available():
        pytest.skip("no C++ compiler for the native store")
    impl = {"python": KVServer, "native": NativeKVServer}
    d = str(tmp_path / "kvdata")
    srv = impl[first](port=0, data_dir=d, snapshot_interval=0.1).start()
    c = KVClient([srv.endpoint])
    lease = c.lease(30.0)
    c.put("/job/pods/a", b"\x00binary\xff", lease.id)
    for i in range(3):
        c.put("/job/status", "v%d" % i)
    _, meta = c.get("/job/status")
    c.close()
    srv.stop()                                        # final snapshot on stop (SIGTERM for the native server)
    srv2 = impl[second](port=0, data_dir=d).start()
    try:
        c2 = KVClient([srv2.endpoint])
        v, m2 = c2.get("/job/status")
        assert v == b"v2" and m2 == meta
        v, m3 = c2.get("/job/pods/a")
        assert v == b"\x00binary\xff" and m3["lease"] == lease.id
        assert c2.lease_keepalive(lease.id) > 0
        assert c2.put("/job/new", b"x")["kv"]["mod_revision"] == meta["mod_revision"] + 1
        c2.lease_revoke(lease.id)
        assert c2.get("/job/pods/a")[0] is None
        c2.close()
    finally:
        srv2.stop()


def test_native_server_error_paths_and_many_clients(kv_server):
    """Unknown lease / unknown method come back as errors (not a dropped connection); 16 clients hammering one
    key through compare-and-swap transactions never lose an update."""
    from edl_b200.store import StoreError

    c = KVClient(kv_server.endpoint)
    with pytest.raises(StoreError):
        c.put("/e/x", b"1", lease=123456789)
    with pytest.raises(StoreError):
        c.call({"method": "no_such_method"})
    assert c.get("/e/x")[0] is None
    c.put("/cnt", "0")
    n_threads, n_incr = 16, 25

    def worker():
        cl = KVClient(kv_server.endpoint)
        for _ in range(n_incr):
            while True:
                v, _ = cl.get("/cnt")
                ok, _ = cl.txn([{"key": "/cnt", "value": v}], [{"op": "put", "key": "/cnt", "value": str(int(v) + 1)}])
                if ok:
                    break
        cl.close()

    ts = [threading.Thread(target=worker) for _ in range(n_threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert c.get("/cnt")[0] == str(n_threads * n_incr).encode()
    c.close()


def test_rejected_request_does_not_reconnect_or_duplicate_the_next_one(kv_server):
    """Regression (found by tests/test_store_differential.py): an application-level error used to be treated like a
    transport error -- the client closed its socket and resent, and the old reader thread then failed the NEXT
    request, which was in flight on the new socket, so that one was executed twice."""
    from edl_b200.store import StoreRequestError

    c = KVClient(kv_server.endpoint)
    c.put("/n", "0")
    sock = c._sock
    for i in range(20):
        with pytest.raises(StoreRequestError):
            c.put("/bad", b"x", lease=424242)                 # unknown lease
        ok, _ = c.txn([{"key": "/n", "value": str(i)}], [{"op": "put", "key": "/n", "value": str(i + 1)}])
        assert ok, "the compare-and-swap after a rejected request was applied twice"
    assert c._sock is sock, "a rejected request must not cost the connection"
    _, meta = c.get("/n")
    assert meta["version"] == 21
    c.close()


def test_a_watcher_that_never_reads_cannot_stall_the_store(kv_server):
    """Head-of-line blocking: a client subscribes to a prefix and then stops reading its socket (paused process, dead
    network path).  Other clients' writes must keep their latency; the store buffers for the slow one (and eventually
    drops it) instead of blocking under its lock."""
    import socket
    import struct

    import msgpack

    host, port = kv_server.endpoint.rsplit(":", 1)
    stuck = socket.create_connection((host, int(port)))
    stuck.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 4096)
    body = msgpack.packb({"method": "watch", "watch_id": 1, "key": "/hol/", "start_revision": 0, "id": 1}, use_bin_type=True)
    stuck.sendall(struct.pack("!I", len(body)) + body)          # ... and never recv()
    c = KVClient(kv_server.endpoint)
    payload = b"x" * 65536
    t0 = time.time()
    for i in range(400):                                        # 26 MB of events for the watcher that is not listening
        c.put("/hol/%d" % (i % 8), payload)
    dt = time.time() - t0
    assert dt < 20, "writes stalled behind a watcher that does not read (%.1fs)" % dt
    assert c.get("/hol/3")[0] == payload
    c.close()
    stuck.close()
