"""Client side of the DataServer RPCs with a per-endpoint channel cache
(reference: python/edl/utils/data_server_client.py:34-152)."""
import threading

from ..protos import rpc, schema
from . import exceptions
from .error_utils import handle_errors_until_timeout

pb = schema.data_server


class _Conn:
    def __init__(self, endpoint):
        self.channel = rpc.insecure_channel(endpoint)
        self.stub = rpc.Stub(self.channel, "data_server.DataServer")


class Client:
    def __init__(self):
        self._conns = {}
        self._acks = {}
        self._lock = threading.Lock()

    def _stub(self, endpoint):
        with self._lock:
            c = self._conns.get(endpoint)
            if c is None:
                c = self._conns[endpoint] = _Conn(endpoint)
            return c.stub

    def close(self):
        with self._lock:
            for c in self._conns.values():
                c.channel.close()
            self._conns.clear()

    @handle_errors_until_timeout
    def get_file_list(self, leader_endpoint, reader_name, pod_id, file_list, timeout=60):
        req = pb.FileListRequest(pod_id=pod_id, reader_name=reader_name)
        for i, p in enumerate(file_list):
            req.file_list.append(pb.FileListElement(idx=i, path=p))
        res = self._stub(leader_endpoint).GetFileList(req, timeout=10)
        exceptions.deserialize(res.status)
        return [(e.idx, e.path) for e in res.file_list]

    @handle_errors_until_timeout
    def report_batch_data_meta(self, leader_endpoint, reader_name, pod_id, dataserver_endpoint, batch_data_ids,
                               timeout=60):
        req = pb.ReportBatchDataMetaRequest(reader_name=reader_name, pod_id=pod_id,
                                            data_server_endpoint=dataserver_endpoint)
        req.batch_data_ids.extend(batch_data_ids)
        res = self._stub(leader_endpoint).ReportBatchDataMeta(req, timeout=10)
        exceptions.deserialize(res.status)

    @handle_errors_until_timeout
    def reach_data_end(self, leader_endpoint, reader_name, pod_id, timeout=60):
        res = self._stub(leader_endpoint).ReachDataEnd(
            pb.ReachDataEndRequest(reader_name=reader_name, pod_id=pod_id), timeout=10)
        exceptions.deserialize(res.status)

    def _get_batch_data_meta(self, leader_endpoint, reader_name, pod_id, ack_seq, timeout=60):
        """Retries transport / transient errors only: "the epoch is drained" is an answer, not a failure."""
        import time

        deadline = time.time() + timeout
        while True:
            try:
                res = self._stub(leader_endpoint).GetBatchDataMeta(
                    pb.GetBatchDataMetaRequest(reader_name=reader_name, pod_id=pod_id, ack_seq=ack_seq), timeout=10)
                exceptions.deserialize(res.status)
                return res
            except exceptions.EdlDataEndError:
                raise
            except exceptions.EdlException:
                if time.time() >= deadline:
                    raise
                time.sleep(0.05)

    def get_batch_data_meta(self, leader_endpoint, reader_name, pod_id, timeout=60):
        """-> list of BatchDataMeta; raises EdlDataEndError when the epoch is drained.  Safe to retry: the request
        carries the sequence number of the last answer received, the leader re-sends an unacknowledged answer."""
        key = (leader_endpoint, reader_name, pod_id)
        with self._lock:
            ack = self._acks.get(key, 0)
        res = self._get_batch_data_meta(leader_endpoint, reader_name, pod_id, ack, timeout=timeout)
        with self._lock:
            fresh = int(res.seq) != ack
            self._acks[key] = int(res.seq)
        return list(res.data) if fresh else []

    @handle_errors_until_timeout
    def get_batch_data(self, meta, timeout=60):
        res = self._stub(meta.data_server_endpoint).GetBatchData(meta, timeout=30)
        exceptions.deserialize(res.status)
        return list(res.data)
