"""A pod = one launcher process + its trainers on one node (reference: python/edl/utils/pod.py:26-181).

Fixes relative to the reference: GPUs are split evenly over ``nproc_per_node``
trainers for any nproc <= #gpus; global ranks are assigned by the *cluster* as the running sum of
the lower pods' trainer counts (the reference's ``pod.rank + rank_in_pod`` collides for multi-trainer
pods); the JSON carries the pod status."""
import json
import uuid

from . import network_utils
from .status import Status
from .trainer import Trainer


class Pod:
    def __init__(self):
        self._id = None
        self._rank = None
        self._trainer_ports = None
        self._addr = None
        self._gpus = None
        self._trainers = []
        self._port = None
        self._status = Status.INITIAL

    # ------------------------------------------------------------------ (de)serialisation
    def to_dict(self):
        return {"id": self._id, "rank": self._rank, "port": self._port,
                "trainer_ports": self._trainer_ports, "addr": self._addr, "gpus": self._gpus,
                "status": int(self._status),
                "trainers": {str(i): t.to_dict() for i, t in enumerate(self._trainers)}}

    def to_json(self):
        return json.dumps(self.to_dict())

    def from_dict(self, d):
        self._id, self._rank, self._addr, self._port = d["id"], d["rank"], d["addr"], d["port"]
        self._trainer_ports, self._gpus = d["trainer_ports"], d["gpus"]
        self._status = Status(int(d.get("status", 0)))
        self._trainers = []
        for _, v in sorted(d["trainers"].items(), key=lambda kv: int(kv[0])):
            t = Trainer()
            t.from_json(v) if isinstance(v, str) else t.from_dict(v)
            self._trainers.append(t)
        return self

    def from_json(self, s):
        if isinstance(s, (bytes, bytearray)):
            s = s.decode("utf-8")
        return self.from_dict(json.loads(s))

    def from_env(self, job_env):
        self._id = str(uuid.uuid1())
        self._trainer_ports = list(job_env.trainer_ports)
        self._gpus = list(job_env.gpus)
        _, self._addr = network_utils.get_host_name_ip()
        n = job_env.nproc_per_node
        assert 1 <= n, "nproc_per_node must be >= 1"
        per, extra = (len(self._gpus) // n, len(self._gpus) % n) if self._gpus else (0, 0)
        self._trainers, b = [], 0
        for i in range(n):
            e = b + per + (1 if i < extra else 0)
            t = Trainer().from_pod(endpoint="{}:{}".format(self._addr, self._trainer_ports[i]),
                                   rank_in_pod=i, gpus=self._gpus[b:e])
            self._trainers.append(t)
            b = e
        return self

    # ------------------------------------------------------------------ accessors
    def __eq__(self, other):
        return isinstance(other, Pod) and self.to_dict() == other.to_dict()

    def __ne__(self, other):
        return not self == other

    def __str__(self):
        return "rank:{} id:{} addr:{} port:{} gpus:{} status:{} trainers_num:{}".format(
            self._rank, self._id, self._addr, self._port, self._gpus, self._status.name,
            len(self._trainers))

    def details(self):
        return str(self) + " trainers:[" + "; ".join(str(t) for t in self._trainers) + "]"

    @property
    def id(self): return self._id
    def get_id(self): return self._id
    @property
    def rank(self): return self._rank
    @rank.setter
    def rank(self, v): self._rank = v
    @property
    def port(self): return self._port
    @port.setter
    def port(self, v): self._port = v
    @property
    def addr(self): return self._addr
    @property
    def endpoint(self): return "{}:{}".format(self._addr, self._port)
    @property
    def trainers(self): return self._trainers
    @property
    def trainers_num(self): return len(self._trainers)
    @property
    def gpus(self): return self._gpus
    @property
    def status(self): return self._status
    @status.setter
    def status(self, s): self._status = Status(int(s))

    def set_global_ranks(self, first_rank: int) -> int:
        """Assign consecutive global ranks starting at ``first_rank``; returns the next free rank."""
        for i, t in enumerate(self._trainers):
            t.global_rank = first_rank + i
        return first_rank + len(self._trainers)

    def get_trainers_endpoints(self):
        return [t.endpoint for t in self._trainers]
