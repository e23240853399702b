"""Teacher transports behind one interface (the reference's ``PredictServer`` abstraction,
distill_worker.py:187-195, whose only real implementation wraps ``paddle_serving_client``):

* ``GrpcPredictClient``  -- off-box teachers: tensors over gRPC to ``teacher_server.TeacherServer``
* ``NopPredictClient``   -- fake teacher for tests (reference ``_TestNopPaddlePredictServer``)
* the same-box NVSwitch path does not go through this interface at all: see ``device_feed.py``.

``predict(feed_batch)`` takes a list of per-sample dicts ``{feed_name: ndarray}`` and returns a
list of per-sample dicts ``{fetch_name: ndarray}``."""
import time

import numpy as np

from ..protos import rpc, schema
from ..utils import exceptions


class PredictClient:
    def __init__(self, server, feeds, fetchs, conf=None):
        self.server, self.feeds, self.fetchs, self.conf = server, feeds, fetchs, conf

    def connect(self):
        raise NotImplementedError

    def predict(self, feed_batch):
        raise NotImplementedError

    def close(self):
        pass


def _to_pb(name, arr):
    arr = np.ascontiguousarray(arr)
    return schema.predict.Tensor(name=name, dtype=str(arr.dtype), shape=list(arr.shape), data=arr.tobytes())


def _from_pb(t):
    return np.frombuffer(t.data, dtype=np.dtype(t.dtype)).reshape(list(t.shape))


class GrpcPredictClient(PredictClient):
    def __init__(self, server, feeds, fetchs, conf=None, timeout=30.0, retries=3):
        super().__init__(server, feeds, fetchs, conf)
        self._timeout, self._retries = timeout, retries
        self._channel = self._stub = None
        self.teacher_feeds = None

    def connect(self):
        self._channel = rpc.insecure_channel(self.server)
        self._stub = rpc.Stub(self._channel, "edl.predict.PredictService")
        conf = self._stub.GetConf(schema.predict.ConfRequest(), timeout=self._timeout)
        # only the slots the teacher actually consumes are shipped (e.g. images, not labels)
        self.teacher_feeds = [f for f in self.feeds if f is not None and f in set(conf.feed_names)]
        missing = [f for f in self.fetchs if f not in set(conf.fetch_names)]
        if missing:
            raise exceptions.EdlInternalError("teacher {} cannot fetch {}".format(self.server, missing))
        return True

    def predict(self, feed_batch):
        req = schema.predict.PredictRequest(fetch=list(self.fetchs))
        for name in self.teacher_feeds:
            req.feeds.append(_to_pb(name, np.stack([np.asarray(s[name]) for s in feed_batch])))
        last = None
        for _ in range(self._retries):
            try:
                res = self._stub.Predict(req, timeout=self._timeout)
                exceptions.deserialize(res.status)
                outs = {t.name: _from_pb(t) for t in res.outputs}
                return [{k: outs[k][i] for k in self.fetchs} for i in range(len(feed_batch))]
            except Exception as e:  # noqa: BLE001 - rpc error or teacher-side failure
                last = e
                time.sleep(0.05)
        raise exceptions.EdlInternalError("predict on {} failed: {}".format(self.server, last))

    def close(self):
        if self._channel is not None:
            self._channel.close()
            self._channel = None


class NopPredictClient(PredictClient):
    """Returns zeros of the configured per-sample shape for every fetch; never touches the network."""

    def __init__(self, server, feeds, fetchs, conf=None, out_shape=(1,), delay=0.0):
        super().__init__(server, feeds, fetchs, conf)
        self._out_shape, self._delay = tuple(out_shape), delay
        self.teacher_feeds = [f for f in feeds if f is not None]

    def connect(self):
        return True

    def predict(self, feed_batch):
        if self._delay:
            time.sleep(self._delay)
        return [{k: np.zeros(self._out_shape, dtype=np.float32) for k in self.fetchs} for _ in feed_batch]
