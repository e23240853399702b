"""Convenience constructors (reference: python/edl/utils/etcd_utils.py)."""
from ..discovery.etcd_client import EtcdClient
from . import constants


def get_etcd(job_env):
    etcd = EtcdClient(endpoints=list(job_env.etcd_endpoints), root=job_env.job_id,
                      timeout=constants.ETCD_CONN_TIMEOUT)
    etcd.init()
    return etcd
