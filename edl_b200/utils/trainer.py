"""One training process (one GPU) inside a pod (reference: python/edl/utils/trainer.py:19-55)."""
import json
import uuid


class Trainer:
    def __init__(self):
        self._id = None
        self._rank_in_pod = None
        self._gpus = []
        self._endpoint = None
        self._global_rank = None

    def to_dict(self):
        return {"id": self._id, "rank_in_pod": self._rank_in_pod, "gpus": self._gpus,
                "endpoint": self._endpoint, "global_rank": self._global_rank}

    def to_json(self):
        return json.dumps(self.to_dict())

    def from_dict(self, d):
        self._id, self._rank_in_pod, self._gpus = d["id"], d["rank_in_pod"], d["gpus"]
        self._endpoint, self._global_rank = d["endpoint"], d["global_rank"]
        return self

    def from_json(self, s):
        return self.from_dict(json.loads(s) if isinstance(s, (str, bytes)) else s)

    def from_pod(self, endpoint, rank_in_pod, gpus):
        self._id = str(uuid.uuid1())
        self._global_rank = None
        self._rank_in_pod = rank_in_pod
        self._endpoint = endpoint
        self._gpus = list(gpus)
        return self

    def __eq__(self, other):
        return isinstance(other, Trainer) and self.to_dict() == other.to_dict()

    def __ne__(self, other):
        return not self == other

    def __str__(self):
        return "id:{} rank_in_pod:{} gpus:{} endpoint:{} global_rank:{}".format(
            self._id, self._rank_in_pod, self._gpus, self._endpoint, self._global_rank)

    @property
    def id(self): return self._id
    @property
    def global_rank(self): return self._global_rank
    @global_rank.setter
    def global_rank(self, v): self._global_rank = v
    @property
    def rank_in_pod(self): return self._rank_in_pod
    @property
    def gpus(self): return self._gpus
    @property
    def endpoint(self): return self._endpoint
