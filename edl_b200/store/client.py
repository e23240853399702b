"""Client for the edl_b200 KV store (see kv_server.py): thread-safe request/response over one TCP
connection plus server-pushed watch events, with transparent reconnect to any listed endpoint
(reference behaviour: random endpoint pick + reconnect-once, discovery/etcd_client.py:39-48,67-83).
"""
from __future__ import annotations

import itertools
import queue
import logging
import random
import socket
import threading
import time
from typing import Callable, Dict, List, Optional

from .kv_server import recv_msg, send_msg

logger = logging.getLogger("edl.store")


class StoreError(Exception):
    pass


class NoValidEndpoint(StoreError):
    pass


class StoreRequestError(StoreError):
    """The server processed the request and rejected it (unknown lease, malformed transaction, ...): not a
    transport problem, so the client neither reconnects nor resends."""


class Lease:
    def __init__(self, client: "KVClient", lease_id: int, ttl: float):
        self.client, self.id, self.ttl = client, lease_id, ttl

    def refresh(self) -> float:
        return self.client.lease_keepalive(self.id)

    def revoke(self):
        self.client.lease_revoke(self.id)


class KVClient:
    def __init__(self, endpoints, timeout: float = 6.0):
        if isinstance(endpoints, str):
            endpoints = [e for e in endpoints.split(",") if e]
        self.endpoints = list(endpoints)
        self.timeout = timeout
        self._sock: Optional[socket.socket] = None
        self._wlock = threading.Lock()
        self._plock = threading.Lock()
        self._pending: Dict[int, dict] = {}
        self._ids = itertools.count(1)
        self._watch_ids = itertools.count(1)
        self._watches: Dict[int, dict] = {}
        self._reader: Optional[threading.Thread] = None
        self._cb_thread: Optional[threading.Thread] = None
        self._cb_queue: "queue.Queue" = queue.Queue()
        self._closed = False
        self._conn_lock = threading.RLock()

    # ------------------------------------------------------------------ connection
    def connect(self):
        with self._conn_lock:
            if self._sock is not None:
                return
            eps = list(self.endpoints)
            random.shuffle(eps)
            last = None
            for ep in eps:
                host, port = ep.rsplit(":", 1)
                try:
                    s = socket.create_connection((host, int(port)), timeout=self.timeout)
                    s.settimeout(None)
                    s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                    self._sock = s
                    self._reader = threading.Thread(target=self._read_loop, args=(s,), daemon=True,
                                                    name="kv-client-reader")
                    self._reader.start()
                    self._rewatch()
                    return
                except OSError as e:
                    last = e
            raise NoValidEndpoint("cannot reach any of %s: %s" % (self.endpoints, last))

    def close(self):
        self._closed = True
        with self._conn_lock:
            if self._sock is not None:
                try:
                    self._sock.shutdown(socket.SHUT_RDWR)
                except OSError:
                    pass
                self._sock.close()
                self._sock = None

    def _read_loop(self, sock):
        while True:
            try:
                msg = recv_msg(sock)
            except OSError:
                msg = None
            if msg is None:
                break
            if "watch_id" in msg and "events" in msg:
                w = self._watches.get(msg["watch_id"])
                if w is not None:
                    w["last_rev"] = max(w.get("last_rev", 0), msg.get("revision", 0))
                    # callbacks run on their own thread, in arrival order: a callback that issues a
                    # store request must not block the reader that has to deliver its response
                    self._dispatch(w["cb"], msg["events"], msg.get("revision", 0))
                continue
            with self._plock:
                slot = self._pending.get(msg.get("id"))
            if slot is not None:
                slot["resp"] = msg
                slot["ev"].set()
        # connection lost: fail everything that is waiting
        with self._conn_lock:
            if self._sock is sock:
                self._sock = None
        # ... that was sent on THIS socket (a reconnect may already have requests in flight on a new one)
        with self._plock:
            for slot in self._pending.values():
                if slot.get("sock") is sock:
                    slot["resp"] = None
                    slot["ev"].set()

    def _dispatch(self, cb, events, rev):
        with self._plock:
            if self._cb_thread is None or not self._cb_thread.is_alive():
                self._cb_thread = threading.Thread(target=self._cb_loop, daemon=True, name="kv-client-watch-cb")
                self._cb_thread.start()
        self._cb_queue.put((cb, events, rev))

    def _cb_loop(self):
        while not self._closed:
            try:
                cb, events, rev = self._cb_queue.get(timeout=1.0)
            except queue.Empty:
                continue
            try:
                cb(events, rev)
            except Exception:  # noqa: BLE001
                logger.exception("watch callback failed")

    def _rewatch(self):
        for wid, w in list(self._watches.items()):
            start = w.get("last_rev", 0) + 1 if w.get("last_rev") else w.get("start_revision", 0)
            try:
                self._call_once({"method": "watch", "watch_id": wid, "key": w["key"], "end": w.get("end"),
                                 "start_revision": start})
            except StoreError:
                pass

    def _call_once(self, req: dict) -> dict:
        sock = self._sock
        if sock is None:
            raise StoreError("not connected")
        rid = next(self._ids)
        req = dict(req, id=rid)
        slot = {"ev": threading.Event(), "resp": None, "sock": sock}
        with self._plock:
            self._pending[rid] = slot
        try:
            send_msg(sock, req, self._wlock)
            if not slot["ev"].wait(self.timeout):
                raise StoreError("store request timed out: %s" % req.get("method"))
        except OSError as e:
            raise StoreError(str(e))
        finally:
            with self._plock:
                self._pending.pop(rid, None)
        resp = slot["resp"]
        if resp is None:
            raise StoreError("connection lost")
        if not resp.get("ok"):
            raise StoreRequestError(resp.get("error", "store error"))
        return resp

    def call(self, req: dict) -> dict:
        """One request with reconnect-and-retry-once semantics."""
        if self._closed:
            raise StoreError("client closed")
        for attempt in (0, 1):
            try:
                self.connect()
                return self._call_once(req)
            except StoreRequestError:
                raise            # the server answered: the connection is fine and a resend would not change the answer
            except (StoreError, OSError):
                with self._conn_lock:
                    if self._sock is not None:
                        try:
                            self._sock.close()
                        except OSError:
                            pass
                        self._sock = None
                if attempt == 1:
                    raise
                time.sleep(0.05)
        raise StoreError("unreachable")

    # ------------------------------------------------------------------ KV API
    @staticmethod
    def _b(v) -> bytes:
        if isinstance(v, bytes):
            return v
        return str(v).encode("utf-8")

    def put(self, key: str, value, lease: int = 0) -> dict:
        return self.call({"method": "put", "key": key, "value": self._b(value), "lease": int(lease)})

    def put_if_not_exists(self, key: str, value, lease: int = 0) -> bool:
        r = self.call({"method": "put", "key": key, "value": self._b(value), "lease": int(lease),
                       "if_not_exists": True})
        return bool(r["succeeded"])

    def get(self, key: str):
        """-> (value bytes | None, meta dict | None)"""
        r = self.call({"method": "get", "key": key})
        if not r["kvs"]:
            return None, None
        kv = r["kvs"][0]
        return kv["value"], kv

    def get_prefix(self, prefix: str):
        """-> (list of kv dicts, header revision)"""
        r = self.call({"method": "get", "key": prefix, "prefix": True})
        return r["kvs"], r["revision"]

    def delete(self, key: str) -> int:
        return self.call({"method": "delete", "key": key})["deleted"]

    def delete_prefix(self, prefix: str) -> int:
        return self.call({"method": "delete", "key": prefix, "prefix": True})["deleted"]

    def txn(self, compare: List[dict], success: List[dict], failure: Optional[List[dict]] = None):
        """-> (succeeded, results).  compare items: {key, target: value|version|create|mod, op, value};
        ops: {op: put|delete|get, key, value, lease, prefix}."""
        def enc(ops):
            out = []
            for o in ops or []:
                o = dict(o)
                if "value" in o:
                    o["value"] = self._b(o["value"])
                out.append(o)
            return out
        cmp_ = []
        for c in compare:
            c = dict(c)
            if c.get("target", "value") == "value" and "value" in c:
                c["value"] = self._b(c["value"])
            cmp_.append(c)
        r = self.call({"method": "txn", "compare": cmp_, "success": enc(success), "failure": enc(failure)})
        return bool(r["succeeded"]), r["results"]

    # ------------------------------------------------------------------ leases
    def lease(self, ttl: float, lease_id: int = 0) -> Lease:
        r = self.call({"method": "lease_grant", "ttl": float(ttl), "lease_id": int(lease_id)})
        return Lease(self, r["lease"], r["ttl"])

    def lease_keepalive(self, lease_id: int) -> float:
        return self.call({"method": "lease_keepalive", "lease": int(lease_id)})["ttl"]

    def lease_revoke(self, lease_id: int):
        self.call({"method": "lease_revoke", "lease": int(lease_id)})

    def lease_ttl(self, lease_id: int) -> float:
        return self.call({"method": "lease_ttl", "lease": int(lease_id)})["ttl"]

    # ------------------------------------------------------------------ watches
    def add_watch_prefix_callback(self, prefix: str, callback: Callable[[list, int], None],
                                  start_revision: int = 0) -> int:
        """callback(events, header_revision); events: [{type: put|delete, key, kv, revision}]"""
        wid = next(self._watch_ids)
        self._watches[wid] = {"key": prefix, "end": None, "cb": callback, "start_revision": start_revision}
        self.call({"method": "watch", "watch_id": wid, "key": prefix, "start_revision": int(start_revision)})
        return wid

    def cancel_watch(self, watch_id: int):
        self._watches.pop(watch_id, None)
        try:
            self.call({"method": "cancel_watch", "watch_id": int(watch_id)})
        except StoreError:
            pass

    def status(self) -> dict:
        return self.call({"method": "status"})

    # ------------------------------------------------------------------ lock (lease-backed)
    def lock(self, key: str, ttl: float = 10.0) -> "StoreLock":
        return StoreLock(self, key, ttl)


class StoreLock:
    """Mutual exclusion on a key: put-if-absent under a lease (auto-released if the holder dies)."""

    def __init__(self, client: KVClient, key: str, ttl: float):
        self.client, self.key, self.ttl = client, key, ttl
        self.lease: Optional[Lease] = None

    def acquire(self, timeout: Optional[float] = 10.0) -> bool:
        deadline = None if timeout is None else time.time() + timeout
        while True:
            lease = self.client.lease(self.ttl)
            if self.client.put_if_not_exists(self.key, b"locked", lease.id):
                self.lease = lease
                return True
            lease.revoke()
            if deadline is not None and time.time() > deadline:
                return False
            time.sleep(0.05)

    def release(self):
        if self.lease is not None:
            self.lease.revoke()
            self.lease = None

    def __enter__(self):
        if not self.acquire():
            raise StoreError("could not acquire lock %s" % self.key)
        return self

    def __exit__(self, *exc):
        self.release()
