"""Leader election on ``rank/0`` (reference: python/edl/utils/leader_pod.py:28-165).

Whoever wins ``put_if_not_exists(rank/0 = pod_id)`` under a TTL lease is the leader and runs the
cluster generator; followers retry the seize every poll interval, so a dead leader is replaced one
lease-TTL later."""
import threading

from . import constants
from .cluster_generator import Generator
from .error_utils import handle_errors_until_timeout
from .exceptions import EdlNotFoundLeader, EdlTableError
from .log_utils import logger
from .pod import Pod


class Register:
    def __init__(self, job_env, pod_id, cluster_generator=None, ttl=None, etcd=None):
        if etcd is None:
            from .etcd_db import get_global_etcd
            etcd = get_global_etcd(job_env.etcd_endpoints, job_env.job_id)
        self._etcd, self._job_env, self._pod_id = etcd, job_env, pod_id
        self._ttl = float(ttl if ttl is not None else constants.ETCD_TTL)
        self._generator = cluster_generator if cluster_generator is not None else Generator(job_env, pod_id, etcd=etcd)
        self._is_leader = False
        self._lease = None
        self._stop = threading.Event()
        self._dead = threading.Event()
        self._lock = threading.Lock()
        self._seize_leader()
        self._t = threading.Thread(target=self._refresher, name="edl-leader", daemon=True)
        self._t.start()

    def _seize_leader(self):
        ok, lease = self._etcd.put_if_not_exists_with_lease(constants.ETCD_POD_RANK, constants.ETCD_POD_LEADER,
                                                            self._pod_id, self._ttl)
        with self._lock:
            if ok:
                self._is_leader, self._lease = True, lease
                logger.info("pod %s is now the leader", self._pod_id)
                self._generator.start()
            else:
                self._is_leader = False
        return ok

    def _refresher(self):
        period = max(0.05, min(self._ttl / 2.0, constants.POLL_INTERVAL))
        while not self._stop.wait(period):
            try:
                if self._is_leader:
                    if self._lease.refresh() <= 0:
                        raise EdlTableError("leader lease expired")
                    if self._generator.is_stopped():
                        raise EdlTableError("cluster generator died")
                else:
                    self._seize_leader()
            except Exception as e:  # noqa: BLE001
                logger.warning("leader register of %s stopped: %s", self._pod_id, e)
                with self._lock:
                    self._is_leader = False     # the lease is gone: rank/0 may already belong to somebody else
                self._generator.stop()
                self._dead.set()
                break

    def is_leader(self):
        return self._is_leader

    def stop(self):
        self._stop.set()
        self._t.join(self._ttl)
        self._generator.stop()
        if self._is_leader:
            try:
                # delete rank/0 only while it still names THIS pod (one transaction): after a lost lease a new
                # leader may own the key, and removing it would take that pod down
                key = self._etcd.get_full_path(constants.ETCD_POD_RANK, constants.ETCD_POD_LEADER)
                self._etcd.kv.txn([{"key": key, "value": self._pod_id.encode() if isinstance(self._pod_id, str)
                                    else self._pod_id}], [{"op": "delete", "key": key}], [])
                if self._lease is not None:
                    self._lease.revoke()
            except Exception:  # noqa: BLE001
                pass
            self._is_leader = False
        self._dead.set()

    def is_stopped(self):
        return self._dead.is_set() or self._stop.is_set()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.stop()


@handle_errors_until_timeout
def get_pod_leader_id(etcd, timeout=15):
    value = etcd.get_value(constants.ETCD_POD_RANK, constants.ETCD_POD_LEADER)
    if value is None:
        raise EdlNotFoundLeader("no leader registered yet")
    return value.decode("utf-8") if isinstance(value, (bytes, bytearray)) else value


@handle_errors_until_timeout
def load_from_etcd(etcd, timeout=15):
    """-> the leader's Pod (from the resource table)."""
    leader_id = get_pod_leader_id(etcd, timeout=timeout)
    value = etcd.get_value(constants.ETCD_POD_RESOURCE, leader_id)
    if value is None:
        raise EdlTableError("leader {} has no resource record".format(leader_id))
    return Pod().from_json(value)
