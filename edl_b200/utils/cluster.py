"""The rank table: ordered pods + a ``stage`` id that changes on every membership change
(reference: python/edl/utils/cluster.py:29-175)."""
import json
import time
import uuid

from . import constants
from .error_utils import handle_errors_until_timeout
from .exceptions import EdlRankError, EdlTableError
from .pod import Pod
from .status import Status


class Cluster:
    def __init__(self):
        self._pods = []
        self._stage = None
        self._status = Status.INITIAL

    def new_stage(self):
        self._stage = str(uuid.uuid1())
        return self._stage

    def assign_ranks(self):
        """Pod rank = position; trainer global ranks = running sum over lower pods."""
        nxt = 0
        for i, pod in enumerate(self._pods):
            pod.rank = i
            nxt = pod.set_global_ranks(nxt)
        return nxt

    # ------------------------------------------------------------------ views
    @property
    def pods(self): return self._pods
    @property
    def stage(self): return self._stage
    @property
    def status(self): return self._status
    @status.setter
    def status(self, s): self._status = Status(int(s))

    def get_pods_nranks(self): return len(self._pods)
    def get_trainers_nranks(self): return sum(p.trainers_num for p in self._pods)
    def get_pods_ids_list(self): return [p.id for p in self._pods]
    def get_pods_ids_set(self): return set(p.id for p in self._pods)
    def get_pods_endpoints(self): return [p.endpoint for p in self._pods]

    def get_trainers_endpoints(self):
        eps = []
        for p in self._pods:
            eps.extend(p.get_trainers_endpoints())
        return eps

    def get_pod_by_id(self, pod_id):
        for p in self._pods:
            if p.id == pod_id:
                return p
        return None

    def get_leader_endpoint(self):
        assert self._pods, "empty cluster"
        return self._pods[0].endpoint

    def get_leader_id(self):
        assert self._pods, "empty cluster"
        return self._pods[0].id

    def __str__(self):
        return "stage:{} status:{} pods:[{}]".format(self._stage, Status(int(self._status)).name,
                                                     "; ".join(str(p) for p in self._pods))

    def details(self):
        return "stage:{} pods:[{}]".format(self._stage, "; ".join(p.details() for p in self._pods))

    def __eq__(self, other):
        return isinstance(other, Cluster) and self.to_dict() == other.to_dict()

    def __ne__(self, other):
        return not self == other

    # ------------------------------------------------------------------ JSON
    def to_dict(self):
        return {"pods": {str(i): p.to_dict() for i, p in enumerate(self._pods)}, "stage": self._stage,
                "status": int(self._status)}

    def to_json(self):
        return json.dumps(self.to_dict())

    def from_dict(self, d):
        self._stage, self._status = d["stage"], Status(int(d["status"]))
        pods = []
        for i, (k, v) in enumerate(sorted(d["pods"].items(), key=lambda kv: int(kv[0]))):
            if i != int(k):
                raise EdlRankError("rank {} is missing in {}".format(i, sorted(d["pods"])))
            pods.append(Pod().from_json(v) if isinstance(v, str) else Pod().from_dict(v))
        self._pods = pods
        return self

    def from_json(self, s):
        if isinstance(s, (bytes, bytearray)):
            s = s.decode("utf-8")
        return self.from_dict(json.loads(s))


@handle_errors_until_timeout
def load_from_etcd(etcd, timeout=60):
    value = etcd.get_value(constants.ETCD_CLUSTER, constants.ETCD_CLUSTER)
    if value is None:
        return None
    return Cluster().from_json(value)


def wait_to_load_from_etcd(etcd, timeout=60):
    """Block until a cluster record exists."""
    begin = time.time()
    while True:
        c = load_from_etcd(etcd, timeout=timeout)
        if c is not None:
            return c
        if time.time() - begin > timeout:
            raise EdlTableError("no cluster record after {}s".format(timeout))
        time.sleep(min(1.0, constants.POLL_INTERVAL))
