"""Client of the leader's PodServer (reference: python/edl/utils/pod_server_client.py:24-60)."""
import time

import grpc

from ..protos import schema
from . import cluster as edl_cluster
from . import exceptions
from .client import Client as _Base


class Client(_Base):
    def __init__(self, endpoint):
        super().__init__(endpoint, "pod_server.PodServer")

    def barrier(self, job_id, pod_id, timeout=15, interval=None):
        """Poll ``Barrier`` until every pod of the current stage arrived; returns the Cluster."""
        interval = interval if interval is not None else 0.2
        req = schema.pod_server.BarrierRequest(job_id=job_id, pod_id=pod_id)
        begin = time.time()
        last = None
        while True:
            try:
                res = self._stub.Barrier(req, timeout=max(1.0, min(5.0, timeout)))
                if not res.status.type:
                    return edl_cluster.Cluster().from_json(res.cluster_json)
                last = res.status
            except grpc.RpcError as e:
                last = None
                err = exceptions.EdlBarrierError("rpc to {} failed: {}".format(self._endpoint, e.code()))
                if time.time() - begin > timeout:
                    raise err
            if time.time() - begin > timeout:
                if last is not None:
                    exceptions.deserialize(last)
                raise exceptions.EdlBarrierError("barrier timed out after {}s".format(timeout))
            time.sleep(interval)

    def scale_out(self, timeout=5):
        st = self._stub.ScaleOut(schema.pod_server.ScaleOutRequest(), timeout=timeout)
        exceptions.deserialize(st)

    def scale_in(self, num, timeout=5):
        st = self._stub.ScaleIn(schema.pod_server.ScaleInRequest(num=num), timeout=timeout)
        exceptions.deserialize(st)
