"""Generic TTL registration: put-if-absent under a lease, keep the lease alive from a background
thread, report death through ``is_stopped()`` (reference: python/edl/utils/register.py:21-86)."""
import threading

from . import constants
from .exceptions import EdlRegisterError
from .log_utils import logger


class Register:
    def __init__(self, etcd, service, server, info, ttl=None):
        self._etcd, self._service, self._server, self._info = etcd, service, server, info
        self._ttl = float(ttl if ttl is not None else constants.ETCD_TTL)
        self._stop = threading.Event()
        self._dead = threading.Event()
        self._lock = threading.Lock()
        if not self._etcd.set_server_not_exists(service, server, info, ttl=self._ttl, timeout=self._ttl):
            raise EdlRegisterError("key {}/{} is already registered".format(service, server))
        self._t = threading.Thread(target=self._refresher, name="edl-register-%s" % service, daemon=True)
        self._t.start()

    def _refresher(self):
        period = max(0.05, self._ttl / 2.0)
        while not self._stop.wait(period):
            try:
                self._etcd.refresh(self._service, self._server, ttl=self._ttl)
            except Exception as e:  # noqa: BLE001 - lease lost / store unreachable => declare death
                logger.warning("register %s/%s lost: %s", self._service, self._server, e)
                self._dead.set()
                break

    def update(self, info):
        with self._lock:
            self._info = info
            self._etcd.refresh(self._service, self._server, info=info, ttl=self._ttl)

    def stop(self):
        self._stop.set()
        self._t.join(self._ttl)
        try:
            self._etcd.remove_server(self._service, self._server)
        except Exception:  # noqa: BLE001
            pass
        self._dead.set()

    def is_stopped(self):
        return self._dead.is_set() or self._stop.is_set()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.stop()
