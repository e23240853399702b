"""Per-pod training progress flag (reference: python/edl/utils/train_status.py:21-45): lets the
cluster generator refuse scale-out for a job that is about to finish
(doc/edl_collective_design_doc.md:26-29).  NEARTHEEND and SUCCEED are distinct values here (the
reference aliases both to 3)."""
import json
from enum import IntEnum

from . import constants
from .error_utils import handle_errors_until_timeout


class TrainStatus(IntEnum):
    INITIAL = 0
    RUNNING = 1
    NEARTHEEND = 2
    SUCCEED = 3
    FAILED = 4


@handle_errors_until_timeout
def save_to_etcd(etcd, pod_id, status, timeout=30):
    etcd.set_server_permanent(constants.ETCD_TRAIN_STATUS, pod_id, json.dumps({"status": int(status)}))


@handle_errors_until_timeout
def load_from_etcd(etcd, pod_id, timeout=30):
    value = etcd.get_value(constants.ETCD_TRAIN_STATUS, pod_id)
    if value is None:
        return None
    if isinstance(value, (bytes, bytearray)):
        value = value.decode("utf-8")
    return TrainStatus(int(json.loads(value)["status"]))


def any_near_the_end(etcd, pod_ids, timeout=30):
    for pid in pod_ids:
        st = load_from_etcd(etcd, pid, timeout=timeout)
        if st in (TrainStatus.NEARTHEEND, TrainStatus.SUCCEED):
            return True
    return False
