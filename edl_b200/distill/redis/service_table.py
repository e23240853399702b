"""Balance state of the redis-flavoured balance server: same ceil/floor balancing as the etcd
flavour (shared ``Service`` balancer), keyed by the client's socket, teacher list refreshed from
redis every 3 s (reference: python/edl/distill/redis/service_table.py:27-273)."""
import logging
import threading

from ..balance_table import Service
from .redis_store import RedisStore

logger = logging.getLogger("edl.distill.redis")


class ServiceTable:
    def __init__(self, ip, port, passwd=None, refresh_s=3.0):
        self._store = RedisStore(ip, port, passwd)
        self._services = {}
        self._client_service = {}   # client key (fd) -> service name
        self._lock = threading.RLock()
        self._refresh_s = refresh_s
        self._stop = threading.Event()
        self._t = None

    def start(self):
        self._t = threading.Thread(target=self._refresh_loop, daemon=True, name="redis-service-refresh")
        self._t.start()
        return self

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(3)
        self._store.close()

    def _refresh(self, name):
        try:
            servers = {s["server"] for s in self._store.get_service(name)}
        except Exception as e:  # noqa: BLE001
            logger.warning("redis refresh of %s failed: %s", name, e)
            return
        svc = self._services[name]
        cur = set(svc.servers)
        svc.update_servers(add=servers - cur, rm=cur - servers)

    def _refresh_loop(self):
        while not self._stop.wait(self._refresh_s):
            with self._lock:
                names = list(self._services)
            for n in names:
                self._refresh(n)

    def add_client(self, key, service_name, require_num):
        with self._lock:
            svc = self._services.get(service_name)
            if svc is None:
                svc = self._services[service_name] = Service(service_name)
                self._refresh(service_name)
            self._client_service[key] = service_name
        svc.add_client(str(key), require_num)
        return svc.snapshot(str(key))

    def rm_client(self, key):
        with self._lock:
            name = self._client_service.pop(key, None)
            svc = self._services.get(name) if name else None
        if svc is not None:
            svc.remove_client(str(key))

    def get_servers(self, key):
        """-> (version, servers) or None"""
        with self._lock:
            name = self._client_service.get(key)
            svc = self._services.get(name) if name else None
        if svc is None:
            return None
        svc.touch(str(key))
        return svc.snapshot(str(key))
