"""Launcher / trainer environment resolution: CLI flag -> env var -> default
(reference: python/edl/utils/env.py:22-229)."""
import os

from . import network_utils
from .log_utils import logger


def get_gpus():
    """Visible GPU ids as strings.  ``CUDA_VISIBLE_DEVICES`` wins; otherwise every device torch
    sees; on a GPU-less host a single pseudo device "0" keeps the plumbing testable.  (The reference
    dereferences ``None`` when the variable is unset, env.py:23-27.)"""
    cvd = os.getenv("CUDA_VISIBLE_DEVICES")
    if cvd is not None and cvd.strip() != "":
        return [x.strip() for x in cvd.split(",") if x.strip() != ""]
    try:
        import torch

        n = torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        n = 0
    return [str(i) for i in range(n)] if n > 0 else ["0"]


def get_from_dict_or_env(args, name, key, default=""):
    if args and name in args and args[name] is not None:
        return args[name]
    return os.getenv(key, default)


class JobEnv:
    def __init__(self, args=None):
        args = args or {}
        self._platform = os.getenv("PADDLE_RUNNING_PLATFORM", "")
        self._job_id = get_from_dict_or_env(args, "job_id", "PADDLE_JOB_ID")
        assert self._job_id, "job_id must be set (--job_id or PADDLE_JOB_ID)"
        eps = get_from_dict_or_env(args, "etcd_endpoints", "PADDLE_ETCD_ENDPOINTS")
        assert eps, "etcd_endpoints must be set (--etcd_endpoints or PADDLE_ETCD_ENDPOINTS)"
        self._etcd_endpoints = [e for e in str(eps).split(",") if e]
        self._ce_test = int(os.getenv("PADDLE_EDL_ONLY_FOR_CE_TEST", "0"))
        # checkpoint file system (HDFS in the reference; any shared path here)
        self._hdfs_home = get_from_dict_or_env(args, "hdfs_home", "PADDLE_EDL_HDFS_HOME")
        self._hdfs_name = get_from_dict_or_env(args, "hdfs_name", "PADDLE_EDL_HDFS_NAME")
        self._hdfs_path = get_from_dict_or_env(args, "hdfs_path", "PADDLE_EDL_HDFS_PATH")
        self._hdfs_ugi = get_from_dict_or_env(args, "hdfs_ugi", "PADDLE_EDL_HDFS_UGI")
        # nodes range "min:max" (the env name keeps the reference's spelling)
        nr = get_from_dict_or_env(args, "nodes_range", "PADDLE_EDLNODES_RANAGE") or \
            os.getenv("PADDLE_EDL_NODES_RANGE", "")
        assert nr, "nodes_range must be set, e.g. 2:8"
        a = str(nr).split(":")
        assert len(a) == 2, "nodes_range is not min:max : {}".format(nr)
        self._min_nodes, self._max_nodes = int(a[0]), int(a[1])
        assert 1 <= self._min_nodes <= self._max_nodes
        self._gpus = get_gpus()
        nproc = get_from_dict_or_env(args, "nproc_per_node", "PADDLE_EDL_NPROC_PERNODE")
        self._nproc_per_node = int(nproc) if str(nproc) not in ("", "None") else len(self._gpus)
        assert self._nproc_per_node >= 1
        if self._platform == "PADDLE_CLOUD" and os.getenv("PADDLE_TRAINER_PORTS"):
            self._trainer_ports = [p for p in os.getenv("PADDLE_TRAINER_PORTS").split(",") if p]
            assert len(self._trainer_ports) >= self._nproc_per_node, "not enough PADDLE_TRAINER_PORTS"
        else:
            self._trainer_ports = [str(p) for p in network_utils.find_free_ports(self._nproc_per_node)]
        self._log_dir = get_from_dict_or_env(args, "log_dir", "PADDLE_EDL_LOG_DIR") or "./log"
        self._log_level = int(get_from_dict_or_env(args, "log_level", "PADDLE_EDL_LOG_LEVEL") or 20)
        logger.debug("job env: %s", self)

    @property
    def gpus(self): return self._gpus
    @property
    def nproc_per_node(self): return self._nproc_per_node
    @property
    def etcd_endpoints(self): return self._etcd_endpoints
    @property
    def job_id(self): return self._job_id
    @property
    def trainer_ports(self): return self._trainer_ports
    @property
    def min_nodes(self): return self._min_nodes
    @property
    def max_nodes(self): return self._max_nodes
    @property
    def hdfs_home(self): return self._hdfs_home
    @property
    def hdfs_name(self): return self._hdfs_name
    @property
    def hdfs_path(self): return self._hdfs_path
    @property
    def hdfs_ugi(self): return self._hdfs_ugi
    @property
    def log_dir(self): return self._log_dir
    @property
    def log_level(self): return self._log_level
    @property
    def platform(self): return self._platform

    def __str__(self):
        return " ".join("{}:{}".format(k, v) for k, v in vars(self).items())


class TrainerEnv:
    """What the launcher exports to every trainer process (Appendix B of SURVEY.md; reference
    env.py:179-229).  ``EDL_POD_LEADER_ID`` / ``EDL_POD_IDS`` are exported here as well (the
    reference's TrainerEnv expects them but its launcher never sets them)."""

    def __init__(self, environ=None):
        e = environ if environ is not None else os.environ
        self._job_id = e.get("PADDLE_JOB_ID", "")
        self._pod_id = e.get("PADDLE_POD_ID", "")
        self._pod_leader_id = e.get("EDL_POD_LEADER_ID", "")
        self._etcd_endpoints = [x for x in e.get("PADDLE_ETCD_ENDPOINTS", "").split(",") if x]
        self._global_rank = int(e.get("PADDLE_TRAINER_ID", "0"))
        self._rank_in_pod = int(e.get("PADDLE_TRAINER_RANK_IN_POD", "0"))
        self._trainer_endpoints = [x for x in e.get("PADDLE_TRAINER_ENDPOINTS", "").split(",") if x]
        self._pod_ids = [x for x in e.get("EDL_POD_IDS", "").split(",") if x]
        self._size = int(e.get("PADDLE_TRAINERS_NUM", str(max(1, len(self._trainer_endpoints)))))
        self._gpus = [x for x in e.get("FLAGS_selected_gpus", "").split(",") if x]
        self._current_endpoint = e.get("PADDLE_CURRENT_ENDPOINT", "")
        self._stage = e.get("EDL_STAGE", "")
        self._ckpt_path = e.get("PADDLE_EDL_FLEET_CHECKPOINT_PATH", e.get("PADDLE_EDL_HDFS_PATH", ""))

    @property
    def job_id(self): return self._job_id
    @property
    def pod_id(self): return self._pod_id
    @property
    def pod_leader_id(self): return self._pod_leader_id
    @property
    def etcd_endpoints(self): return self._etcd_endpoints
    @property
    def global_rank(self): return self._global_rank
    @property
    def rank_in_pod(self): return self._rank_in_pod
    @property
    def trainer_endpoints(self): return self._trainer_endpoints
    @property
    def pod_ids(self): return self._pod_ids
    @property
    def size(self): return self._size
    @property
    def gpus(self): return self._gpus
    @property
    def current_endpoint(self): return self._current_endpoint
    @property
    def stage(self): return self._stage
    @property
    def checkpoint_path(self): return self._ckpt_path

    def torch_distributed_env(self, master_port_offset=0):
        """MASTER_ADDR/PORT, RANK, WORLD_SIZE, LOCAL_RANK derived from the Paddle-style contract, so a
        trainer can call ``torch.distributed.init_process_group`` directly."""
        master = self._trainer_endpoints[0] if self._trainer_endpoints else "127.0.0.1:29500"
        host, port = master.rsplit(":", 1)
        return {"MASTER_ADDR": host, "MASTER_PORT": str(int(port) + master_port_offset),
                "RANK": str(self._global_rank), "WORLD_SIZE": str(self._size),
                "LOCAL_RANK": str(self._rank_in_pod)}
