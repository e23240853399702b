"""``EtcdClient`` -- the registry client every EDL component talks to.

Same public surface and key layout (``/<root>/<service>/nodes/<server>``) as the reference's
python/edl/discovery/etcd_client.py:51-263, implemented on the in-repo store
(:mod:`edl_b200.store`) instead of the ``etcd3`` package.  Semantics kept: per-key lease cache,
``set_server_not_exists`` = put-if-absent under a TTL lease with retry-until-timeout, permanent
keys, ``refresh`` = lease keep-alive (or re-put when info changes), race-free
``get_service_with_revision`` + ``watch_service(start_revision=...)`` that coalesces put/delete
events into ``call_back(add_servers, rm_servers)``.
"""
from __future__ import annotations

import threading
import time
from typing import Callable, Dict, List, Optional

from ..store.client import KVClient, Lease, NoValidEndpoint, StoreError  # noqa: F401


class ServerMeta:
    def __init__(self, server, info, mod_revision, revision):
        self.server = server
        self.info = info  # bytes, like etcd values
        self.mod_revision = mod_revision
        self.revision = revision

    def __str__(self):
        return "server={}, info={}, mod_revision={}, revision={}".format(
            self.server, self.info, self.mod_revision, self.revision)

    __repr__ = __str__


class EtcdClient:
    def __init__(self, endpoints=("127.0.0.1:2379",), passwd=None, root="service", timeout=6):
        if isinstance(endpoints, str):
            endpoints = endpoints.split(",")
        assert isinstance(endpoints, (list, tuple, set)), "endpoints must be a list"
        self._endpoints = list(endpoints)
        self._passwd = passwd
        self._root = root
        self._timeout = timeout
        self._kv: Optional[KVClient] = None
        self._leases: Dict[str, Lease] = {}
        self._lock = threading.RLock()

    # ------------------------------------------------------------------ connection
    def init(self):
        self._kv = KVClient(self._endpoints, timeout=self._timeout)
        self._kv.connect()
        return self

    @property
    def kv(self) -> KVClient:
        if self._kv is None:
            self.init()
        return self._kv

    def close(self):
        if self._kv is not None:
            self._kv.close()
            self._kv = None

    # ------------------------------------------------------------------ paths
    def get_full_path(self, service_name, server):
        return "/{}/{}/nodes/{}".format(self._root, service_name, server)

    def _service_dir(self, service_name):
        return "/{}/{}/nodes/".format(self._root, service_name)

    def get_server_name_from_full_path(self, path, service_name):
        return path[len(self._service_dir(service_name)):]

    # ------------------------------------------------------------------ reads
    def get_service(self, service_name) -> List[ServerMeta]:
        return self.get_service_with_revision(service_name)[0]

    def get_service_with_revision(self, service_name):
        kvs, rev = self.kv.get_prefix(self._service_dir(service_name))
        servers = [ServerMeta(self.get_server_name_from_full_path(kv["key"], service_name), kv["value"],
                              kv["mod_revision"], rev) for kv in kvs]
        return servers, rev

    def get_value(self, service_name, server):
        return self.kv.get(self.get_full_path(service_name, server))[0]

    def get_key(self, key):
        return self.kv.get(key)

    def _get_server(self, service_name, server):
        value, meta = self.kv.get(self.get_full_path(service_name, server))
        if meta is None:
            return None, None, 0, 0, 0
        return value, meta["key"], meta["version"], meta["create_revision"], meta["mod_revision"]

    # ------------------------------------------------------------------ watches
    def watch_service(self, service_name, call_back: Callable, **kwargs):
        """call_back(add_servers, rm_servers) with lists of ServerMeta; ``start_revision=`` kwarg
        resumes from a revision returned by :meth:`get_service_with_revision`."""
        def services_change(events, header_rev):
            add_servers, rm_servers = {}, {}
            nodes = self._service_dir(service_name)
            for ev in events:
                if not ev["key"].startswith(nodes):
                    continue
                key = self.get_server_name_from_full_path(ev["key"], service_name)
                meta = ServerMeta(key, ev["kv"]["value"], ev["kv"]["mod_revision"], header_rev)
                if ev["type"] == "put":
                    rm_servers.pop(key, None)
                    add_servers[key] = meta
                elif ev["type"] == "delete":
                    add_servers.pop(key, None)
                    rm_servers[key] = meta
                else:
                    raise TypeError("store event type is not put or delete!")
            if not add_servers and not rm_servers:
                return
            call_back(list(add_servers.values()), list(rm_servers.values()))

        d = "/{}/{}/".format(self._root, service_name)
        return self.kv.add_watch_prefix_callback(d, services_change,
                                                 start_revision=kwargs.get("start_revision", 0))

    def cancel_watch(self, watch_id):
        return self.kv.cancel_watch(watch_id)

    # ------------------------------------------------------------------ writes
    def _get_lease(self, key, ttl=10) -> Lease:
        with self._lock:
            le = self._leases.get(key)
            if le is None or self.kv.lease_ttl(le.id) <= 0:
                le = self.kv.lease(ttl)
                self._leases[key] = le
            return le

    def set_server_not_exists(self, service_name, server, info, ttl=10, timeout=6):
        """put-if-absent under a TTL lease; retries (refreshing the lease) until ``timeout``.
        Returns True iff this caller created the key."""
        key = self.get_full_path(service_name, server)
        begin = time.time()
        while True:
            lease = self._get_lease(key, ttl)
            if self.kv.put_if_not_exists(key, info, lease.id):
                return True
            lease.refresh()
            if time.time() - begin > timeout:
                break
            time.sleep(min(1.0, max(0.05, timeout / 10.0)))
        return False

    def _set_server(self, service_name, server, info, ttl=10):
        key = self.get_full_path(service_name, server)
        lease = self._get_lease(key, ttl)
        return self.kv.put(key, info, lease.id)

    def set_server_permanent(self, service_name, server, info):
        key = self.get_full_path(service_name, server)
        self.kv.put(key, info)
        with self._lock:
            self._leases.pop(key, None)

    def remove_server(self, service_name, server):
        key = self.get_full_path(service_name, server)
        self.kv.delete(key)
        with self._lock:
            le = self._leases.pop(key, None)
        if le is not None:
            try:
                le.revoke()
            except StoreError:
                pass

    def remove_service(self, service_name):
        for s in self.get_service(service_name):
            self.remove_server(service_name, s.server)
        self.kv.delete_prefix("/{}/{}/".format(self._root, service_name))

    def refresh(self, service_name, server, info=None, ttl=10):
        if info is not None:
            self._set_server(service_name, server, info, ttl)
            return
        key = self.get_full_path(service_name, server)
        with self._lock:
            le = self._leases.get(key)
        if le is None or le.refresh() <= 0:
            raise StoreError("lease of %s is gone" % key)

    def lock(self, service_name, server, ttl=10):
        return self.kv.lock(self.get_full_path(service_name, server) + "/__lock__", ttl)

    # ------------------------------------------------------------------ extras used by the cluster layer
    def put_if_not_exists_with_lease(self, service_name, server, info, ttl):
        """Single attempt (no retry loop): returns (created, lease)."""
        key = self.get_full_path(service_name, server)
        lease = self.kv.lease(ttl)
        ok = self.kv.put_if_not_exists(key, info, lease.id)
        if not ok:
            lease.revoke()
            return False, None
        with self._lock:
            self._leases[key] = lease
        return True, lease

    def txn_put_if_value(self, guard_service, guard_server, guard_value, puts):
        """Atomically: if ``guard`` key holds ``guard_value`` then apply ``puts``
        (list of (service, server, info)).  The leader-guarded write of the reference
        (utils/cluster_generator.py:223-250, utils/state.py:186-200)."""
        compare = [{"key": self.get_full_path(guard_service, guard_server), "value": guard_value}]
        success = [{"op": "put", "key": self.get_full_path(s, k), "value": v} for s, k, v in puts]
        ok, _ = self.kv.txn(compare, success, [])
        return ok
