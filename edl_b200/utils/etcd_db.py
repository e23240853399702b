"""Process-global store handle (reference: python/edl/utils/etcd_db.py:19-27)."""
import threading

from ..discovery.etcd_client import EtcdClient
from . import constants

_lock = threading.Lock()
_etcd = None


def get_global_etcd(etcd_endpoints=None, job_id=None):
    global _etcd
    with _lock:
        if _etcd is None:
            assert etcd_endpoints is not None and job_id is not None, "first call needs endpoints and job id"
            _etcd = EtcdClient(endpoints=list(etcd_endpoints), root=job_id,
                               timeout=constants.ETCD_CONN_TIMEOUT)
            _etcd.init()
        return _etcd


def reset_global_etcd():
    global _etcd
    with _lock:
        if _etcd is not None:
            _etcd.close()
        _etcd = None
