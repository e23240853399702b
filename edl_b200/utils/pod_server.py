"""Per-pod gRPC server; the leader's instance implements the stage barrier
(reference: python/edl/utils/pod_server.py:32-163)."""
import threading

from ..protos import rpc, schema
from . import cluster as edl_cluster
from . import exceptions, leader_pod
from .log_utils import logger


class PodServerServicer:
    def __init__(self, job_env, pod_id, etcd):
        self._job_env, self._pod_id, self._etcd = job_env, pod_id, etcd
        self._lock = threading.Lock()
        self._barrier_in = {}   # stage -> set of arrived pod ids

    def _check_leader(self):
        leader_id = leader_pod.get_pod_leader_id(self._etcd, timeout=3)
        if leader_id != self._pod_id:
            raise exceptions.EdlLeaderError("this pod {} is not the leader {}".format(self._pod_id, leader_id))

    # ScaleOut / ScaleIn: the hooks an external scheduler uses to resize a running job.  The reference
    # declares them but leaves them empty (python/edl/utils/pod_server.py:47-67).  Here the leader records
    # the requested pod count under ``scale/target``; the cluster generator (same pod) honours it on its
    # next pass: surplus pods drop out of the rank table (they idle as stand-bys), a raised target lets
    # waiting pods back in.  ``nodes_range`` still bounds the target.
    def _set_target(self, target):
        from . import constants

        target = max(self._job_env.min_nodes, min(self._job_env.max_nodes, int(target)))
        self._etcd.set_server_permanent(constants.ETCD_SCALE, "target", str(target))
        return target

    def ScaleOut(self, request, context):
        status = schema.common.Status()
        try:
            self._check_leader()
            self._set_target(self._job_env.max_nodes)
        except exceptions.EdlException as e:
            exceptions.serialize(status, e)
        return status

    def ScaleIn(self, request, context):
        status = schema.common.Status()
        try:
            self._check_leader()
            cluster = edl_cluster.load_from_etcd(self._etcd, timeout=3)
            cur = len(cluster.pods) if cluster is not None else self._job_env.max_nodes
            self._set_target(cur - max(0, int(request.num)))
        except exceptions.EdlException as e:
            exceptions.serialize(status, e)
        return status

    def Barrier(self, request, context):
        res = schema.pod_server.BarrierResponse()
        try:
            self._check_leader()
            cluster = edl_cluster.load_from_etcd(self._etcd, timeout=3)
            if cluster is None:
                raise exceptions.EdlBarrierError("no cluster generated yet")
            ids = cluster.get_pods_ids_set()
            if request.pod_id not in ids:
                raise exceptions.EdlBarrierError("pod {} is not in cluster stage {}".format(
                    request.pod_id, cluster.stage))
            with self._lock:
                arrived = self._barrier_in.setdefault(cluster.stage, set())
                arrived.add(request.pod_id)
                done = arrived >= ids
                for st in [s for s in self._barrier_in if s != cluster.stage]:
                    self._barrier_in.pop(st, None)
            if not done:
                raise exceptions.EdlBarrierError("stage {}: arrived {} of {}".format(
                    cluster.stage, len(arrived), len(ids)))
            res.cluster_json = cluster.to_json()
        except exceptions.EdlException as e:
            exceptions.serialize(res.status, e)
        except Exception as e:  # noqa: BLE001
            exceptions.serialize(res.status, exceptions.EdlInternalError(str(e)))
        return res


class PodServer:
    def __init__(self, job_env, pod_id, etcd=None):
        if etcd is None:
            from .etcd_db import get_global_etcd
            etcd = get_global_etcd(job_env.etcd_endpoints, job_env.job_id)
        self._job_env, self._pod_id, self._etcd = job_env, pod_id, etcd
        self._server = None
        self._port = None

    def start(self, concurrency=20, host="0.0.0.0"):
        self._server = rpc.make_server(concurrency)
        sv = PodServerServicer(self._job_env, self._pod_id, self._etcd)
        rpc.add_service(self._server, "pod_server.PodServer",
                        {"Barrier": sv.Barrier, "ScaleOut": sv.ScaleOut, "ScaleIn": sv.ScaleIn})
        self._port = self._server.add_insecure_port("{}:0".format(host))
        assert self._port > 0, "cannot bind pod server"
        self._server.start()
        logger.info("pod server of %s listening on %d", self._pod_id, self._port)
        return self

    @property
    def port(self):
        return self._port

    def stop(self, grace=0):
        if self._server is not None:
            self._server.stop(grace)
            self._server = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.stop()
