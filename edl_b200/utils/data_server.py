"""DataServer: every pod serves the batches it produced (``GetBatchData``); the leader's instance
additionally balances batch ids across pods so that all pods finish an epoch together even when
their file slices are uneven or the world size changed mid-epoch
(reference: python/edl/utils/data_server.py:31-431 -- the reference re-balances only after every pod
has reported; here ``GetBatchDataMeta`` steals incrementally from the richest producer).

RPCs (wire-compatible with python/edl/protos/data_server.proto):
  GetFileList(pod_id, reader_name, file_list)        -> this pod's slice of the (verified) file list
  ReportBatchDataMeta(reader, pod, endpoint, ids)    -> leader records ids produced by ``pod``
  ReachDataEnd(reader, pod)                          -> ``pod`` will produce no more ids
  GetBatchDataMeta(reader, pod)                      -> ids (with producer endpoint) ``pod`` should consume;
                                                        EdlDataEndError once everything was handed out
  GetBatchData(BatchDataMeta)                        -> the records of those ids (any pod)
"""
import threading
from collections import OrderedDict, deque

from ..protos import rpc, schema
from . import exceptions
from .log_utils import logger

pb = schema.data_server


class _PodQueue:
    def __init__(self):
        self.endpoint = None
        self.ids = deque()
        self.ended = False


class PodsData:
    """Leader-side balancer state for one reader."""

    def __init__(self, reader_name, file_list, pod_ids, grant=4):
        self.last = {}           # consumer pod -> (seq, answer): re-sent until the consumer acknowledges it
        self.reader_name = reader_name
        self.file_list = list(file_list)
        self.pod_ids = sorted(pod_ids)
        self.grant = grant
        self.lock = threading.Lock()
        self.queues = {p: _PodQueue() for p in self.pod_ids}
        self.file_slices = {p: [] for p in self.pod_ids}
        for i, path in enumerate(self.file_list):           # round-robin file assignment
            self.file_slices[self.pod_ids[i % len(self.pod_ids)]].append((i, path))
        self.handed_out = 0

    def get_pod_file_list(self, pod_id):
        if pod_id not in self.file_slices:
            raise exceptions.EdlPodIDNotExistError(pod_id)
        return self.file_slices[pod_id]

    def put(self, pod_id, endpoint, batch_ids):
        with self.lock:
            q = self.queues.get(pod_id)
            if q is None:
                raise exceptions.EdlPodIDNotExistError(pod_id)
            q.endpoint = endpoint
            q.ids.extend(batch_ids)

    def set_data_end(self, pod_id):
        with self.lock:
            if pod_id not in self.queues:
                raise exceptions.EdlPodIDNotExistError(pod_id)
            self.queues[pod_id].ended = True

    def pop_acked(self, pod_id, ack_seq):
        """Idempotent hand-out: ``ack_seq`` is the sequence number of the last answer the consumer received.  If
        it is not the one we sent last, that answer was lost on the way: send it again (its ids are already off
        the queues) instead of handing out new ones.  -> (seq, answer)"""
        with self.lock:
            seq, ans = self.last.get(pod_id, (0, None))
            if ans and ack_seq != seq:
                return seq, ans
        ans = self.pop(pod_id)
        with self.lock:
            if ans:
                seq += 1
                self.last[pod_id] = (seq, ans)
            return seq, ans

    def pop(self, pod_id):
        """-> list of (producer_pod_id, endpoint, [ids]); [] = nothing available *yet*.
        Raises EdlDataEndError when every producer ended and all ids were handed out."""
        with self.lock:
            if pod_id not in self.queues:
                raise exceptions.EdlPodIDNotExistError(pod_id)
            mine = self.queues[pod_id]
            out = []
            if mine.ids:
                ids = [mine.ids.popleft() for _ in range(min(self.grant, len(mine.ids)))]
                out.append((pod_id, mine.endpoint, ids))
            else:
                # steal from the richest other producer (keeps pods in lock-step at epoch end)
                rich = max((q for p, q in self.queues.items() if p != pod_id), key=lambda q: len(q.ids), default=None)
                if rich is not None and len(rich.ids) > 0:
                    n = max(1, min(self.grant, len(rich.ids) // 2))
                    src_pod = [p for p, q in self.queues.items() if q is rich][0]
                    ids = [rich.ids.pop() for _ in range(n)]
                    out.append((src_pod, rich.endpoint, ids))
            if out:
                self.handed_out += sum(len(x[2]) for x in out)
                return out
            if all(q.ended and not q.ids for q in self.queues.values()):
                raise exceptions.EdlDataEndError("reader {} is drained".format(self.reader_name))
            return []


class DataServerServicer:
    """``capacity`` bounds the producer: ``put_batch`` BLOCKS while that many batches wait to be fetched (nothing
    is ever evicted: every cached id may already have been reported to the leader and promised to a consumer).
    Fetched batches move to a small ring of recently served ones so that a consumer whose ``GetBatchData`` answer
    was lost gets the same batches again when it retries."""

    def __init__(self, pod_id, is_leader_fn=None, capacity=256, served_keep=64):
        self._pod_id = pod_id
        self._is_leader_fn = is_leader_fn
        self._lock = threading.Lock()
        self._space = threading.Condition(self._lock)
        self._pods_data = {}                      # reader_name -> PodsData (leader only)
        self._batches = OrderedDict()             # batch_data_id -> BatchData pb (every pod)
        self._served = OrderedDict()              # recently fetched batches (retry safety net)
        self._capacity = max(32, int(capacity))
        self._served_keep = served_keep
        self._closed = False

    # ---- local batch cache (producer side)
    def put_batch(self, batch, timeout=None):
        """Blocks while the cache is full (back-pressure on the generator).  False if the server was closed."""
        with self._space:
            while len(self._batches) >= self._capacity and not self._closed:
                if not self._space.wait(timeout=timeout if timeout is not None else 1.0) and timeout is not None:
                    raise exceptions.EdlAccessDataError("batch cache of pod {} stayed full for {} s".format(
                        self._pod_id, timeout))
            if self._closed:
                return False
            self._batches[batch.batch_data_id] = batch
            return True

    def pop_batch(self, batch_id):
        with self._space:
            b = self._batches.pop(batch_id, None)
            if b is not None:
                self._served[batch_id] = b
                while len(self._served) > self._served_keep:
                    self._served.popitem(last=False)
                self._space.notify_all()
                return b
            return self._served.get(batch_id)       # a retried fetch of something already handed out

    def close(self):
        with self._space:
            self._closed = True
            self._space.notify_all()

    # ---- leader registration of a reader
    def create_reader(self, reader_name, file_list, pod_ids):
        with self._lock:
            if reader_name not in self._pods_data:
                self._pods_data[reader_name] = PodsData(reader_name, file_list, pod_ids)
            return self._pods_data[reader_name]

    def _reader(self, name):
        with self._lock:
            pd = self._pods_data.get(name)
        if pd is None:
            raise exceptions.EdlReaderNameError("reader {} is unknown on this server".format(name))
        return pd

    # ---- RPCs
    def GetFileList(self, request, context):
        res = pb.FileListResponse()
        try:
            pd = self._reader(request.reader_name)
            sent = [(e.idx, e.path) for e in request.file_list]
            if sent and sent != list(enumerate(pd.file_list)):
                raise exceptions.EdlFileListNotMatchError("file list of pod {} differs from the leader's".format(
                    request.pod_id))
            for idx, path in pd.get_pod_file_list(request.pod_id):
                res.file_list.append(pb.FileListElement(idx=idx, path=path))
        except exceptions.EdlException as e:
            exceptions.serialize(res.status, e)
        return res

    def ReportBatchDataMeta(self, request, context):
        res = schema.common.EmptyRet()
        try:
            self._reader(request.reader_name).put(request.pod_id, request.data_server_endpoint,
                                                  list(request.batch_data_ids))
        except exceptions.EdlException as e:
            exceptions.serialize(res.status, e)
        return res

    def ReachDataEnd(self, request, context):
        res = schema.common.EmptyRet()
        try:
            self._reader(request.reader_name).set_data_end(request.pod_id)
        except exceptions.EdlException as e:
            exceptions.serialize(res.status, e)
        return res

    def GetBatchDataMeta(self, request, context):
        res = pb.BatchDataMetaResponse()
        try:
            seq, answer = self._reader(request.reader_name).pop_acked(request.pod_id, int(request.ack_seq))
            res.seq = seq
            for producer, endpoint, ids in answer:
                m = res.data.add()
                m.reader_name, m.producer_pod_id, m.consumer_pod_id = request.reader_name, producer, request.pod_id
                m.data_server_endpoint = endpoint or ""
                m.batch_data_ids.extend(ids)
        except exceptions.EdlException as e:
            exceptions.serialize(res.status, e)
        return res

    def GetBatchData(self, request, context):
        res = pb.BatchDataResponse()
        try:
            for bid in request.batch_data_ids:
                b = self.pop_batch(bid)
                if b is None:
                    raise exceptions.EdlAccessDataError("batch {} is not on pod {}".format(bid, self._pod_id))
                res.data.append(b)
        except exceptions.EdlException as e:
            exceptions.serialize(res.status, e)
        return res


class DataServer:
    """``host``: interface to bind -- the pod's own address by default (``start(addr=...)``), NOT 0.0.0.0: the
    server hands out training records to whoever asks."""

    def __init__(self, pod_id, host=None, capacity=256):
        self.servicer = DataServerServicer(pod_id, capacity=capacity)
        self._server = None
        self._host = host
        self.port = None

    def start(self, addr="127.0.0.1", concurrency=20):
        self._server = rpc.make_server(concurrency)
        sv = self.servicer
        rpc.add_service(self._server, "data_server.DataServer", {
            "GetFileList": sv.GetFileList, "ReportBatchDataMeta": sv.ReportBatchDataMeta,
            "ReachDataEnd": sv.ReachDataEnd, "GetBatchDataMeta": sv.GetBatchDataMeta,
            "GetBatchData": sv.GetBatchData})
        self.port = self._server.add_insecure_port("{}:0".format(self._host or addr))
        assert self.port > 0
        self._server.start()
        self.endpoint = "{}:{}".format(addr, self.port)
        logger.debug("data server of %s on %s", sv._pod_id, self.endpoint)
        return self

    def stop(self):
        self.servicer.close()
        if self._server is not None:
            self._server.stop(0)
            self._server = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.stop()
