"""Job / pod status tables (reference: python/edl/utils/status.py:22-110)."""
import json
from enum import IntEnum

from . import constants
from .error_utils import handle_errors_until_timeout
from .exceptions import EdlTableError
from .log_utils import logger


class Status(IntEnum):
    INITIAL = 0
    RUNNING = 1
    PENDING = 2
    SUCCEED = 3
    FAILED = 4

    @staticmethod
    def bool_to_status(b):
        return Status.SUCCEED if b else Status.FAILED


def _dump(status):
    return json.dumps({"status": int(status)})


def _parse(value):
    if value is None:
        return None
    if isinstance(value, (bytes, bytearray)):
        value = value.decode("utf-8")
    return Status(int(json.loads(value)["status"]))


@handle_errors_until_timeout
def load_job_status_from_etcd(etcd, timeout=30):
    value = etcd.get_value(constants.ETCD_JOB_STATUS, "status")
    return _parse(value)


@handle_errors_until_timeout
def save_job_status_to_etcd(etcd, status, timeout=30):
    etcd.set_server_permanent(constants.ETCD_JOB_STATUS, "status", _dump(status))


def save_job_flag_to_etcd(etcd, pod_id, flag, timeout=30):
    """Persist the final job state.  (The reference only logs the failure case,
    status.py:60-66 -- here FAILED is written too so a relaunch can see it.)"""
    save_job_status_to_etcd(etcd, Status.bool_to_status(flag), timeout=timeout)
    logger.info("pod %s set job status %s", pod_id, "SUCCEED" if flag else "FAILED")


@handle_errors_until_timeout
def save_pod_status_to_etcd(etcd, pod_id, status, timeout=30):
    etcd.set_server_permanent(constants.ETCD_POD_STATUS, pod_id, _dump(status))


@handle_errors_until_timeout
def load_pod_status_from_etcd(etcd, pod_id, timeout=30):
    return _parse(etcd.get_value(constants.ETCD_POD_STATUS, pod_id))


def save_pod_flag_to_etcd(etcd, pod_id, flag, timeout=30):
    save_pod_status_to_etcd(etcd, pod_id, Status.bool_to_status(flag), timeout=timeout)


@handle_errors_until_timeout
def load_pods_status_from_etcd(etcd, timeout=30):
    """-> (inited, running, succeeded, failed) sets of pod ids."""
    sets = {Status.INITIAL: set(), Status.RUNNING: set(), Status.SUCCEED: set(), Status.FAILED: set(),
            Status.PENDING: set()}
    for s in etcd.get_service(constants.ETCD_POD_STATUS):
        try:
            sets[_parse(s.info)].add(s.server)
        except (ValueError, KeyError, TypeError) as e:
            raise EdlTableError("bad pod status record %s: %s" % (s, e))
    return sets[Status.INITIAL], sets[Status.RUNNING], sets[Status.SUCCEED], sets[Status.FAILED]
