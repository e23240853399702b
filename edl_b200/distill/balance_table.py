"""Teacher <-> student balancing behind the discovery service.

Reference behaviour (python/edl/distill/balance_table.py:34-688):
* per service: ``max_conn_per_server = ceil(C/S)``, each client gets ``max(1, floor(S/C))`` servers
  capped by its ``require_num``; every client whose assignment changed gets a new ``version`` and
  learns about it on its next heartbeat;
* several discovery servers share the load: each registers itself under ``__balance__`` and a
  consistent-hash ring over the live discovery servers decides which one owns a service name
  (clients asking the wrong one get REDIRECT);
* clients that stop heart-beating for ``idle_seconds`` (7) are unregistered;
* teacher membership comes from the registry: get + watch-from-revision+1 per service.

The implementation below is original: explicit last-seen timestamps + a sweeper thread instead of
the weakref timing wheel, and a least-loaded greedy balancer.
"""
import logging
import math
import threading
import time

from ..discovery.consistent_hash import ConsistentHash
from ..discovery.etcd_client import EtcdClient
from ..protos.schema import Code

logger = logging.getLogger("edl.distill.balance")

BALANCE_SERVICE = "__balance__"


class _Client:
    __slots__ = ("name", "require_num", "version", "servers", "last_seen")

    def __init__(self, name, require_num):
        self.name, self.require_num, self.version, self.servers = name, max(1, int(require_num)), 0, []
        self.last_seen = time.time()


class Service:
    """Assignment state of one teacher service."""

    def __init__(self, name):
        self.name = name
        self.servers = set()
        self.clients = {}
        self.lock = threading.Lock()

    # ---- membership
    def update_servers(self, add=(), rm=()):
        with self.lock:
            before = set(self.servers)
            self.servers |= set(add)
            self.servers -= set(rm)
            if before != self.servers:
                self._rebalance()

    def add_client(self, name, require_num):
        with self.lock:
            c = self.clients.get(name)
            if c is None:
                c = self.clients[name] = _Client(name, require_num)
            else:
                c.require_num = max(1, int(require_num))
            c.last_seen = time.time()
            self._rebalance()
            return c

    def remove_client(self, name):
        with self.lock:
            if self.clients.pop(name, None) is not None:
                self._rebalance()

    def touch(self, name):
        with self.lock:
            c = self.clients.get(name)
            if c is not None:
                c.last_seen = time.time()
            return c

    def expire_clients(self, idle_seconds):
        now = time.time()
        with self.lock:
            dead = [n for n, c in self.clients.items() if now - c.last_seen > idle_seconds]
            for n in dead:
                self.clients.pop(n)
            if dead:
                self._rebalance()
        return dead

    # ---- the balancer (call with lock held)
    def _rebalance(self):
        S, C = len(self.servers), len(self.clients)
        old = {n: list(c.servers) for n, c in self.clients.items()}
        if S == 0 or C == 0:
            for c in self.clients.values():
                c.servers = []
        else:
            per_client = max(1, S // C)
            want = {n: min(c.require_num, per_client) for n, c in self.clients.items()}
            cap = max(math.ceil(C / S), math.ceil(sum(want.values()) / S))
            load = {s: 0 for s in self.servers}
            # keep still-valid links (stability), dropping the excess of over-served clients
            for n, c in sorted(self.clients.items()):
                kept = []
                for s in c.servers:
                    if s in load and len(kept) < want[n] and load[s] < cap:
                        kept.append(s)
                        load[s] += 1
                c.servers = kept
            # greedy fill from the least-loaded servers
            for n, c in sorted(self.clients.items(), key=lambda kv: len(kv[1].servers)):
                while len(c.servers) < want[n]:
                    cands = [s for s in self.servers if s not in c.servers]
                    if not cands:
                        break
                    s = min(cands, key=lambda x: (load[x], x))
                    c.servers.append(s)
                    load[s] += 1
        for n, c in self.clients.items():
            if sorted(old[n]) != sorted(c.servers):
                c.version += 1

    def snapshot(self, name):
        with self.lock:
            c = self.clients.get(name)
            return (c.version, list(c.servers)) if c is not None else None


class BalanceTable:
    def __init__(self, server, db_endpoints, idle_seconds=7, registry_root="service", ttl=6):
        self._server = server                     # my own "ip:port" as seen by clients
        self._db = EtcdClient(db_endpoints, root=registry_root)
        self._idle_seconds = idle_seconds
        self._ttl = ttl
        self._services = {}
        self._lock = threading.RLock()
        self._hash = ConsistentHash([])
        self._watches = {}
        self._stop = threading.Event()
        self._threads = []
        self._ready = False

    # ------------------------------------------------------------------ lifecycle
    def start(self):
        self._db.init()
        # register myself among the discovery servers and follow the peer set
        self._db.set_server_not_exists(BALANCE_SERVICE, self._server, "discovery", ttl=self._ttl, timeout=self._ttl)
        peers, rev = self._db.get_service_with_revision(BALANCE_SERVICE)
        for p in peers:
            self._hash.add_new_node(p.server)
        self._hash.add_new_node(self._server)
        self._db.watch_service(BALANCE_SERVICE, self._on_peers_change, start_revision=rev + 1)
        self._threads = [threading.Thread(target=self._keepalive, daemon=True, name="balance-keepalive"),
                         threading.Thread(target=self._sweeper, daemon=True, name="balance-sweeper")]
        for t in self._threads:
            t.start()
        self._ready = True
        return self

    def stop(self):
        self._stop.set()
        for t in self._threads:
            t.join(3)
        try:
            self._db.remove_server(BALANCE_SERVICE, self._server)
        except Exception:  # noqa: BLE001
            pass
        self._db.close()

    def _keepalive(self):
        while not self._stop.wait(min(2.0, self._ttl / 3.0)):
            try:
                self._db.refresh(BALANCE_SERVICE, self._server)
            except Exception:  # noqa: BLE001 - lease lost: re-register
                try:
                    self._db.set_server_not_exists(BALANCE_SERVICE, self._server, "discovery", ttl=self._ttl,
                                                   timeout=1)
                except Exception as e:  # noqa: BLE001
                    logger.warning("discovery self-registration failed: %s", e)

    def _sweeper(self):
        while not self._stop.wait(1.0):
            with self._lock:
                services = list(self._services.values())
            for svc in services:
                dead = svc.expire_clients(self._idle_seconds)
                if dead:
                    logger.info("service %s: clients timed out: %s", svc.name, dead)

    def _on_peers_change(self, add, rm):
        for s in add:
            self._hash.add_new_node(s.server)
        for s in rm:
            if s.server != self._server:
                self._hash.remove_node(s.server)

    # ------------------------------------------------------------------ teacher membership
    def _get_service(self, name, create=False):
        with self._lock:
            svc = self._services.get(name)
            if svc is None and create:
                svc = self._services[name] = Service(name)
                servers, rev = self._db.get_service_with_revision(name)
                svc.update_servers(add=[s.server for s in servers])
                self._watches[name] = self._db.watch_service(
                    name, lambda add, rm, _svc=svc: _svc.update_servers([s.server for s in add],
                                                                        [s.server for s in rm]),
                    start_revision=rev + 1)
            return svc

    def _owner(self, service_name):
        node, nodes, version = self._hash.get_node_nodes(service_name)
        return node, nodes, version

    # ------------------------------------------------------------------ RPC entry points
    def register_client(self, client, service_name, require_num, token=None):
        """-> (code, message, version, discovery_version, servers, discovery_servers)"""
        if not self._ready:
            return Code.NO_READY, "discovery server not ready", 0, 0, [], []
        if not client or not service_name:
            return Code.INVALID_ARGUMENT, "client and service_name are required", 0, 0, [], []
        owner, nodes, dver = self._owner(service_name)
        if owner is not None and owner != self._server:
            return Code.REDIRECT, owner, 0, dver, [], nodes
        with self._lock:
            for other in self._services.values():
                if other.name != service_name and other.snapshot(client) is not None:
                    return Code.REGISTER_OTHER_SERVICE, other.name, 0, dver, [], nodes
        svc = self._get_service(service_name, create=True)
        already = svc.snapshot(client) is not None
        svc.add_client(client, require_num)
        version, servers = svc.snapshot(client)
        code = Code.ALREADY_REGISTER if already else Code.OK
        return code, "", version, dver, servers, nodes

    def heartbeat(self, client, version, discovery_version):
        if not self._ready:
            return Code.NO_READY, "discovery server not ready", 0, 0, [], []
        with self._lock:
            services = list(self._services.values())
        for svc in services:
            c = svc.touch(client)
            if c is None:
                continue
            owner, nodes, dver = self._owner(svc.name)
            if owner is not None and owner != self._server:
                svc.remove_client(client)
                return Code.REDIRECT, owner, 0, dver, [], nodes
            cur_version, servers = svc.snapshot(client)
            new_nodes = nodes if dver != discovery_version else []
            if cur_version != version:
                return Code.OK, "", cur_version, dver, servers, new_nodes
            return Code.OK, "", cur_version, dver, [], new_nodes
        return Code.UNREGISTERED, "client {} is not registered here".format(client), 0, 0, [], []

    def unregister_client(self, client):
        with self._lock:
            services = list(self._services.values())
        for svc in services:
            svc.remove_client(client)
