"""Reader metadata registered in the store at ``reader/<name>/nodes/<pod_id>``
(reference: python/edl/utils/reader.py:22-99)."""
import json
import time

from . import constants
from .error_utils import handle_errors_until_timeout
from .exceptions import EdlTableError


class ReaderMeta:
    def __init__(self, name, pod_id, data_server_endpoint):
        self._name, self._pod_id, self._endpoint = name, pod_id, data_server_endpoint

    def to_json(self):
        return json.dumps({"name": self._name, "pod_id": self._pod_id, "endpoint": self._endpoint})

    def from_json(self, s):
        if isinstance(s, (bytes, bytearray)):
            s = s.decode("utf-8")
        d = json.loads(s)
        self._name, self._pod_id, self._endpoint = d["name"], d["pod_id"], d["endpoint"]
        return self

    @property
    def name(self): return self._name
    @property
    def pod_id(self): return self._pod_id
    @property
    def endpoint(self): return self._endpoint

    def __eq__(self, o):
        return isinstance(o, ReaderMeta) and self.to_json() == o.to_json()


def _table(reader_name):
    return "{}/{}".format(constants.ETCD_READER, reader_name)


@handle_errors_until_timeout
def save_to_etcd(etcd, reader_name, pod_id, data_server_endpoint, timeout=30):
    meta = ReaderMeta(reader_name, pod_id, data_server_endpoint)
    etcd.set_server_permanent(_table(reader_name), pod_id, meta.to_json())
    return meta


@handle_errors_until_timeout
def load_from_etcd(etcd, reader_name, pod_id, timeout=30):
    value = etcd.get_value(_table(reader_name), pod_id)
    if value is None:
        raise EdlTableError("reader {} of pod {} is not registered".format(reader_name, pod_id))
    return ReaderMeta(None, None, None).from_json(value)


def check_dist_readers(etcd, reader_name, pod_ids, timeout=60):
    """Block until every pod of the cluster has registered its data server for ``reader_name``;
    returns {pod_id: ReaderMeta}."""
    begin = time.time()
    while True:
        got = {s.server: ReaderMeta(None, None, None).from_json(s.info) for s in etcd.get_service(_table(reader_name))}
        if set(pod_ids) <= set(got):
            return {p: got[p] for p in pod_ids}
        if time.time() - begin > timeout:
            raise EdlTableError("readers missing for pods {}".format(sorted(set(pod_ids) - set(got))))
        time.sleep(0.2)
