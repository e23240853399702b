"""Elastic data reader (reference: python/edl/collective/distribute_reader.py:33-391 -- a design
sketch there; implemented for real here).

Every pod runs three cooperating pieces:
* a **generator** thread that takes this pod's slice of the file list from the leader, splits files
  into records (``FileSplitter``), packs ``batch_size`` records into a ``BatchData`` kept in the local
  ``DataServer`` and reports the produced batch ids to the leader;
* the local **DataServer** that serves those batches to whoever is told to consume them;
* an **accesser** thread that asks the leader which batch ids this pod should consume next (its own
  first, then ids stolen from slower / richer pods) and fetches them locally or over gRPC.

``Reader`` yields ``{"meta": {...}, "data": [record, ...]}``; ``meta`` carries
``file_idx / begin / end`` so the trainer can record consumed ranges into ``State.DataCheckpoint``
(``edl.notify_end_one_batch``) and a resumed epoch skips what was already trained on.
"""
import queue
import threading
import uuid

from ..protos import schema
from ..utils import data_server as edl_data_server
from ..utils import data_server_client, exceptions
from ..utils import reader as edl_reader
from ..utils.log_utils import logger

pb = schema.data_server


# Record fields travel between pods over plain gRPC: the wire format is DATA ONLY (bytes, str, int, float, bool, None,
# numpy arrays as dtype + shape + raw buffer, lists / tuples of those).  No pickle: whoever can reach a pod's data
# server could otherwise execute code in every trainer that fetches from it.
_NP_OK = frozenset("?bhilqBHILQefd")          # numpy dtype kinds allowed on the wire (no object arrays)


def _encode_field(v):
    import json
    import struct

    if isinstance(v, (bytes, bytearray, memoryview)):
        return b"b" + bytes(v)
    if isinstance(v, str):
        return b"s" + v.encode("utf-8")
    if isinstance(v, (bool, int, float)) or v is None:
        return b"j" + json.dumps(v).encode()
    try:
        import numpy as np
        if isinstance(v, np.generic):
            v = np.asarray(v)
        if isinstance(v, np.ndarray):
            if v.dtype.char not in _NP_OK:
                raise TypeError("numpy dtype %s is not allowed in a record field" % v.dtype)
            head = json.dumps({"d": v.dtype.str, "s": list(v.shape)}).encode()
            return b"n" + struct.pack("<I", len(head)) + head + np.ascontiguousarray(v).tobytes()
    except ImportError:
        pass
    if isinstance(v, (list, tuple)):
        parts = [_encode_field(x) for x in v]
        out = [b"l" if isinstance(v, list) else b"t", struct.pack("<I", len(parts))]
        for part in parts:
            out += [struct.pack("<I", len(part)), part]
        return b"".join(out)
    raise TypeError("record field of type %s cannot be shipped between pods (bytes, str, numbers, numpy arrays and "
                    "lists / tuples of those are)" % type(v).__name__)


def _decode_field(b):
    import json
    import struct

    tag, body = bytes(b[:1]), bytes(b[1:])
    if tag == b"b":
        return body
    if tag == b"s":
        return body.decode("utf-8")
    if tag == b"j":
        return json.loads(body.decode())
    if tag == b"n":
        import numpy as np
        (n,) = struct.unpack_from("<I", body, 0)
        head = json.loads(body[4:4 + n].decode())
        dt = np.dtype(head["d"])
        if dt.char not in _NP_OK:
            raise ValueError("refusing numpy dtype %s from the wire" % dt)
        return np.frombuffer(body[4 + n:], dtype=dt).reshape(head["s"]).copy()
    if tag in (b"l", b"t"):
        (count,) = struct.unpack_from("<I", body, 0)
        off, items = 4, []
        for _ in range(count):
            (ln,) = struct.unpack_from("<I", body, off)
            items.append(_decode_field(body[off + 4:off + 4 + ln]))
            off += 4 + ln
        return items if tag == b"l" else tuple(items)
    raise ValueError("unknown record field tag %r" % tag)


class DataGenerator(threading.Thread):
    def __init__(self, reader):
        super().__init__(daemon=True, name="edl-data-generator")
        self.r = reader
        self.error = None

    def run(self):
        r = self.r
        try:
            files = r._client.get_file_list(r._leader_endpoint, r._name, r._pod_id, r._file_list, timeout=60)
            pending, n_batches = [], 0
            for file_idx, path in files:
                cur, begin, last = [], None, None
                for rec in r._splitter(path):
                    rec_no = int(rec[0])
                    if r._data_checkpoint is not None and r._data_checkpoint.is_processed(file_idx, rec_no):
                        continue
                    if begin is None:
                        begin = rec_no
                    last = rec_no
                    cur.append(rec)
                    if len(cur) == r._batch_size:
                        pending.append(self._emit(file_idx, begin, last, cur))
                        cur, begin = [], None
                        n_batches += 1
                        if len(pending) >= 8:
                            self._report(pending)
                            pending = []
                if cur:
                    pending.append(self._emit(file_idx, begin, last, cur))
                    n_batches += 1
            if pending:
                self._report(pending)
            r._client.reach_data_end(r._leader_endpoint, r._name, r._pod_id, timeout=60)
            logger.debug("pod %s produced %d batches", r._pod_id, n_batches)
        except Exception as e:  # noqa: BLE001
            logger.exception("data generator failed")
            self.error = e
            try:
                r._client.reach_data_end(r._leader_endpoint, r._name, r._pod_id, timeout=5)
            except Exception:  # noqa: BLE001
                pass

    def _emit(self, file_idx, begin, end, records):
        r = self.r
        bid = "{}:{}:{}:{}:{}".format(r._pod_id[:8], file_idx, begin, end, uuid.uuid4().hex[:6])
        b = pb.BatchData(batch_data_id=bid)
        for rec in records:
            rr = b.records.add()
            rr.record_no = int(rec[0])
            rr.field_data.extend(_encode_field(f) for f in rec[1:])
        r._server.servicer.put_batch(b)
        return bid

    def _report(self, ids):
        r = self.r
        r._client.report_batch_data_meta(r._leader_endpoint, r._name, r._pod_id, r._server.endpoint, ids, timeout=60)


class DataAccesser(threading.Thread):
    def __init__(self, reader):
        super().__init__(daemon=True, name="edl-data-accesser")
        self.r = reader

    def run(self):
        r = self.r
        try:
            while not r._stop.is_set():
                try:
                    metas = r._client.get_batch_data_meta(r._leader_endpoint, r._name, r._pod_id)
                except exceptions.EdlDataEndError:
                    break
                if not metas:
                    r._stop.wait(0.02)
                    continue
                for m in metas:
                    if m.producer_pod_id == r._pod_id:
                        batches = [r._server.servicer.pop_batch(b) for b in m.batch_data_ids]
                        if any(b is None for b in batches):
                            raise exceptions.EdlAccessDataError("local batch missing")
                    else:
                        batches = r._client.get_batch_data(m, timeout=60)
                    for b in batches:
                        r._out.put(b)
            r._out.put(None)
        except Exception as e:  # noqa: BLE001
            logger.exception("data accesser failed")
            r._out.put(e)


class Reader:
    """``Reader(file_list, file_splitter_cls, batch_size, cache_capcity=100)`` -- iterate to get
    ``{"meta", "data"}`` dicts.  ``pod_id / leader_endpoint / etcd / pod_ids`` default to the
    launcher-provided trainer environment (one reader per pod: use it from the pod's rank-0
    trainer, or give each trainer its own ``name``).  ``cache_capcity`` is also the balancing granularity: the
    accesser prefetches that many batches, so a small value hands batches out at the pace they are consumed and the
    work stealing then keeps the pods finishing together; a large one favours throughput of the fetch path."""

    def __init__(self, file_list, file_splitter_cls, batch_size, cache_capcity=100, name=None, pod_id=None,
                 leader_endpoint=None, etcd=None, pod_ids=None, is_leader=None, data_checkpoint=None,
                 server_addr="127.0.0.1"):
        self._file_list = list(file_list)
        self._splitter = file_splitter_cls() if isinstance(file_splitter_cls, type) else file_splitter_cls
        self._batch_size = batch_size
        self._name = name or "reader"
        self._data_checkpoint = data_checkpoint
        self._out = queue.Queue(maxsize=max(2, cache_capcity))
        self._stop = threading.Event()
        self._client = data_server_client.Client()
        if pod_id is None:
            from ..utils.env import TrainerEnv
            env = TrainerEnv()
            pod_id, pod_ids = env.pod_id, env.pod_ids or [env.pod_id]
            is_leader = env.pod_leader_id == env.pod_id
            if etcd is None and env.etcd_endpoints:
                from ..discovery.etcd_client import EtcdClient
                etcd = EtcdClient(env.etcd_endpoints, root=env.job_id)
                etcd.init()
        self._pod_id, self._pod_ids = pod_id, list(pod_ids or [pod_id])
        # the generator may run at most this far ahead of the consumers (back-pressure; nothing is ever evicted)
        self._server = edl_data_server.DataServer(pod_id, capacity=max(64, 4 * cache_capcity)).start(addr=server_addr)
        if is_leader or (is_leader is None and self._pod_ids[0] == pod_id):
            self._server.servicer.create_reader(self._name, self._file_list, self._pod_ids)
        if etcd is not None:
            edl_reader.save_to_etcd(etcd, self._name, pod_id, self._server.endpoint, timeout=30)
            if leader_endpoint is None:
                metas = edl_reader.check_dist_readers(etcd, self._name, self._pod_ids, timeout=120)
                leader_endpoint = metas[self._pod_ids[0]].endpoint
        self._leader_endpoint = leader_endpoint or self._server.endpoint
        self._gen = self._acc = None

    @property
    def endpoint(self):
        return self._server.endpoint

    def __iter__(self):
        self._gen, self._acc = DataGenerator(self), DataAccesser(self)
        self._gen.start()
        self._acc.start()
        while True:
            item = self._out.get()
            if item is None:
                break
            if isinstance(item, Exception):
                raise item
            parts = item.batch_data_id.split(":")
            meta = {"batch_data_id": item.batch_data_id, "file_idx": int(parts[1]), "begin": int(parts[2]),
                    "end": int(parts[3])}
            data = [(rec.record_no,) + tuple(_decode_field(f) for f in rec.field_data) for rec in item.records]
            yield {"meta": meta, "data": data}
        if self._gen.error is not None:
            raise self._gen.error

    def stop(self):
        self._stop.set()
        self._server.stop()
        self._client.close()
