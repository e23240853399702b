"""Per-pod watcher of the cluster record: flips ``changed`` when the stage or the ordered pod list
differs from the cluster this pod is running with (reference: python/edl/utils/cluster_watcher.py:23-120).
Besides the reference's polling it also subscribes to a store watch so a change is seen immediately."""
import threading

from . import cluster as edl_cluster
from . import constants
from .log_utils import logger


class Watcher:
    def __init__(self, job_env, cluster, pod=None, etcd=None):
        if etcd is None:
            from .etcd_db import get_global_etcd
            etcd = get_global_etcd(job_env.etcd_endpoints, job_env.job_id)
        # snapshot: the caller may keep mutating its own Cluster object
        self._etcd, self._cluster = etcd, edl_cluster.Cluster().from_dict(cluster.to_dict())
        self._new_cluster = self._cluster
        self._changed = threading.Event()
        self._stop = threading.Event()
        self._lock = threading.Lock()
        self._kick = threading.Event()
        self._watch_id = None
        try:
            self._watch_id = etcd.watch_service(constants.ETCD_CLUSTER, lambda add, rm: self._kick.set())
        except Exception:  # noqa: BLE001 - polling alone is sufficient
            self._watch_id = None
        self._t = threading.Thread(target=self._loop, name="edl-cluster-watcher", daemon=True)
        self._t.start()

    def _differs(self, new):
        old = self._cluster
        return new.stage != old.stage or new.get_pods_ids_list() != old.get_pods_ids_list()

    def _loop(self):
        while not self._stop.is_set():
            self._kick.wait(constants.POLL_INTERVAL)
            self._kick.clear()
            if self._stop.is_set():
                break
            try:
                new = edl_cluster.load_from_etcd(self._etcd, timeout=5)
            except Exception as e:  # noqa: BLE001
                logger.debug("watcher could not load cluster: %s", e)
                continue
            if new is None:
                continue
            with self._lock:
                self._new_cluster = new
                if self._differs(new):
                    self._changed.set()

    @property
    def changed(self):
        return self._changed.is_set()

    def is_changed(self):
        return self._changed.is_set()

    def get_cluster(self):
        with self._lock:
            return self._cluster

    def get_new_cluster(self):
        with self._lock:
            return self._new_cluster

    def stop(self):
        self._stop.set()
        self._kick.set()
        if self._watch_id is not None:
            try:
                self._etcd.cancel_watch(self._watch_id)
            except Exception:  # noqa: BLE001
                pass
        self._t.join(constants.POLL_INTERVAL + 1)

    def is_stopped(self):
        return self._stop.is_set()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.stop()
