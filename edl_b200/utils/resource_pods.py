"""The ``resource`` table: every live pod keeps ``resource/<pod_id> = pod_json`` alive under a TTL
lease (reference: python/edl/utils/resource_pods.py:24-71)."""
import time

from . import constants
from . import register as edl_register
from .error_utils import handle_errors_until_timeout
from .exceptions import EdlWaitFollowersReleaseError
from .pod import Pod


class Register(edl_register.Register):
    def __init__(self, job_env, pod_id, pod_json, ttl=None, etcd=None):
        if etcd is None:
            from .etcd_db import get_global_etcd
            etcd = get_global_etcd(job_env.etcd_endpoints, job_env.job_id)
        super().__init__(etcd, constants.ETCD_POD_RESOURCE, pod_id, pod_json, ttl)


@handle_errors_until_timeout
def load_from_etcd(etcd, timeout=15):
    """-> {pod_id: Pod} of every currently registered pod."""
    pods = {}
    for s in etcd.get_service(constants.ETCD_POD_RESOURCE):
        pods[s.server] = Pod().from_json(s.info)
    return pods


@handle_errors_until_timeout
def wait_resource(etcd, pod_id, timeout=15):
    """Leader side: succeed once every *other* pod has released its resource key (so the leader can
    declare the final job status last)."""
    others = [p for p in load_from_etcd(etcd, timeout=timeout) if p != pod_id]
    if others:
        raise EdlWaitFollowersReleaseError("followers still registered: {}".format(others))
    return True


def wait_followers_release(etcd, pod_id, timeout=60):
    begin = time.time()
    while time.time() - begin < timeout:
        try:
            return wait_resource(etcd, pod_id, timeout=1)
        except EdlWaitFollowersReleaseError:
            time.sleep(0.2)
    return False
