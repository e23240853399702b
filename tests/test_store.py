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
