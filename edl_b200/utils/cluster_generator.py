"""Leader-only rank-table builder (reference: python/edl/utils/cluster_generator.py:32-272).

Every poll interval the leader reconciles the stored cluster with the live ``resource`` table and
the pod status table:

* pods that disappeared (lease expired) or FAILED  -> rebuild from scratch, leader first, new stage;
* INITIAL pods waiting and room below ``max_nodes`` and no pod NEAR-THE-END -> append them
  (scale-out), new stage  -- implemented for real here; the reference's append path never worked
  (SURVEY App. E);
* fewer than ``min_nodes`` live pods -> nothing is written (everybody keeps waiting in the barrier).

The write is a store transaction guarded by "``rank/0`` still holds my pod id"."""
import threading

from . import cluster as edl_cluster
from . import constants, resource_pods, status as edl_status, train_status as edl_train_status
from .exceptions import EdlGenerateClusterError, EdlTableError
from .log_utils import logger


class Generator:
    def __init__(self, job_env, pod_id, etcd=None):
        if etcd is None:
            from .etcd_db import get_global_etcd
            etcd = get_global_etcd(job_env.etcd_endpoints, job_env.job_id)
        self._etcd, self._job_env, self._pod_id = etcd, job_env, pod_id
        self._stop = threading.Event()
        self._dead = threading.Event()
        self._t = None
        self._lock = threading.Lock()

    def start(self):
        with self._lock:
            if self._t is not None and self._t.is_alive():
                return
            self._stop.clear()
            self._dead.clear()
            try:
                self._generate_cluster_once()
            except EdlGenerateClusterError as e:
                logger.info("initial cluster not ready: %s", e)
            self._t = threading.Thread(target=self._loop, name="edl-cluster-generator", daemon=True)
            self._t.start()

    def _loop(self):
        while not self._stop.wait(constants.POLL_INTERVAL):
            try:
                self._generate_cluster_once()
            except EdlGenerateClusterError as e:
                logger.debug("cluster not (re)generated: %s", e)
            except Exception as e:  # noqa: BLE001
                logger.warning("cluster generator failed: %s", e)
                self._dead.set()
                break

    def stop(self):
        self._stop.set()
        t = self._t
        if t is not None and t is not threading.current_thread():
            t.join(constants.POLL_INTERVAL * 2 + 1)
        self._dead.set()

    def is_stopped(self):
        return self._dead.is_set()

    # ------------------------------------------------------------------ core
    def _limit(self):
        """Pod count the job may use right now: ``max_nodes`` unless a scheduler asked for fewer (ScaleIn)."""
        try:
            v = self._etcd.get_value(constants.ETCD_SCALE, "target")
            if v:
                return max(self._job_env.min_nodes, min(self._job_env.max_nodes, int(v)))
        except Exception:  # noqa: BLE001
            pass
        return self._job_env.max_nodes

    def _build_from_scratch(self, resource, failed):
        live = {pid: p for pid, p in resource.items() if pid not in failed}
        if self._pod_id not in live:
            raise EdlGenerateClusterError("leader {} has no resource record".format(self._pod_id))
        ordered = [live[self._pod_id]] + [p for pid, p in sorted(live.items()) if pid != self._pod_id]
        ordered = ordered[:self._limit()]
        c = edl_cluster.Cluster()
        c._pods = ordered
        c.new_stage()
        c.assign_ranks()
        return c

    def _append_inited(self, current, resource, inited):
        room = self._limit() - len(current.pods)
        have = current.get_pods_ids_set()
        new = [resource[pid] for pid in sorted(inited) if pid in resource and pid not in have][:max(0, room)]
        if not new:
            return None
        c = edl_cluster.Cluster().from_dict(current.to_dict())
        c._pods.extend(new)
        c.new_stage()
        c.assign_ranks()
        return c

    def _generate_cluster_once(self):
        etcd = self._etcd
        current = edl_cluster.load_from_etcd(etcd, timeout=5)
        resource = resource_pods.load_from_etcd(etcd, timeout=5)
        inited, running, succeed, failed = edl_status.load_pods_status_from_etcd(etcd, timeout=5)
        new_cluster = None
        if current is None:
            new_cluster = self._build_from_scratch(resource, failed)
        else:
            cur_ids = current.get_pods_ids_set()
            disappeared = cur_ids - set(resource) - succeed
            bad = cur_ids & failed
            if disappeared or bad:
                logger.info("pods disappeared:%s failed:%s -> rebuilding the cluster", sorted(disappeared), sorted(bad))
                new_cluster = self._build_from_scratch(resource, failed)
            elif len(current.pods) > self._limit():
                logger.info("scale-in requested: %d pods -> %d", len(current.pods), self._limit())
                keep = {p.id: resource[p.id] for p in current.pods if p.id in resource}     # current members only
                new_cluster = self._build_from_scratch(keep, failed)
            elif len(current.pods) < self._limit():
                waiting = (set(resource) - cur_ids) & (inited | (set(resource) - running - succeed - failed))
                if waiting and not edl_train_status.any_near_the_end(etcd, cur_ids, timeout=5):
                    new_cluster = self._append_inited(current, resource, waiting)
        if new_cluster is None:
            return current
        if len(new_cluster.pods) < self._job_env.min_nodes:
            raise EdlGenerateClusterError("only {} pods, need at least {}".format(
                len(new_cluster.pods), self._job_env.min_nodes))
        new_cluster.status = edl_status.Status.RUNNING
        ok = etcd.txn_put_if_value(constants.ETCD_POD_RANK, constants.ETCD_POD_LEADER, self._pod_id,
                                   [(constants.ETCD_CLUSTER, constants.ETCD_CLUSTER, new_cluster.to_json())])
        if not ok:
            raise EdlTableError("pod {} is no longer the leader".format(self._pod_id))
        logger.info("new cluster stage %s with %d pods", new_cluster.stage, len(new_cluster.pods))
        return new_cluster
