"""Store table names and timing constants (reference: python/edl/utils/constants.py:15-39).

The reference's values (lease TTL 15 s, 3 s polls) make a membership change take ~25-35 s to be
noticed and acted on.  The same knobs exist here but every one of them can be shortened
through ``EDL_*`` environment variables -- the B200 fast path (in-process re-planning, no NCCL
re-bootstrap) makes short intervals worthwhile."""
import os


def _f(name, default):
    return float(os.environ.get(name, default))


ETCD_POD_RESOURCE = "resource"
ETCD_POD_RANK = "rank"
ETCD_POD_STATUS = "pod_status"
ETCD_JOB_STATUS = "job_status"
ETCD_TRAIN_STATUS = "train_status"
ETCD_CLUSTER = "cluster"
ETCD_READER = "reader"
ETCD_STATE = "state"
ETCD_SCALE = "scale"        # scale/target = pod count an external scheduler asked for (ScaleIn / ScaleOut RPCs)
ETCD_POD_LEADER = "0"   # key of the leader record inside the rank table

ETCD_CONN_TIMEOUT = _f("EDL_ETCD_CONN_TIMEOUT", 6)
ETCD_TTL = _f("EDL_ETCD_TTL", 15)
ETCD_OPERATION_TIMEOUT = _f("EDL_ETCD_OPERATION_TIMEOUT", 60)
POLL_INTERVAL = _f("EDL_POLL_INTERVAL", 3)          # generator / watcher / leader re-seize / supervision
BARRIER_TIMEOUT = _f("EDL_BARRIER_TIMEOUT", 600)
RESCALE_BARRIER_TIMEOUT = _f("EDL_RESCALE_BARRIER_TIMEOUT", 60)
KILL_GRACE = _f("EDL_KILL_GRACE", 3)
# seconds a launcher waits, after one of its trainers died, for the membership to change before it declares the job failed
COLLATERAL_GRACE = _f("EDL_COLLATERAL_GRACE", ETCD_TTL + 3 * POLL_INTERVAL + 1)
# seconds a SIGTERMed pod waits for the job to re-plan without it (k8s default grace period: 30 s)
LEAVE_GRACE = _f("EDL_LEAVE_GRACE", 25)
INPLACE_ACK_TIMEOUT = _f("EDL_INPLACE_ACK_TIMEOUT", 60)   # s a launcher waits for its trainers to enter the new stage in place

ALL_TABLES = [ETCD_POD_RESOURCE, ETCD_POD_RANK, ETCD_POD_STATUS, ETCD_JOB_STATUS, ETCD_TRAIN_STATUS,
              ETCD_CLUSTER, ETCD_READER, ETCD_STATE, ETCD_SCALE]


def clean_etcd(etcd):
    """Remove every table of this job (test helper, reference constants.py:30-39)."""
    for table in ALL_TABLES:
        etcd.remove_service(table)
