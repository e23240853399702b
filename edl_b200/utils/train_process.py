"""Trainer process supervision (reference: python/edl/utils/train_process.py:25-188): start one
process per trainer with the environment contract of SURVEY App. B, poll exit codes, tail the
rank-0 log, terminate whole process trees."""
import os
import signal
import subprocess
import sys
import time

import psutil

from . import constants
from .log_utils import logger


class TrainerProc:
    def __init__(self):
        self.proc = None
        self.log_fn = None
        self.rank = None
        self.cmd = None
        self.log_offset = None
        self.local_rank = None


def trainer_env(job_env, cluster, pod, trainer):
    """The launcher -> trainer interface: Paddle-compatible names plus torch.distributed names."""
    eps = cluster.get_trainers_endpoints()
    master_host, master_port = eps[0].rsplit(":", 1)
    return {
        "PADDLE_JOB_ID": str(job_env.job_id),
        "PADDLE_POD_ID": str(pod.id),
        "PADDLE_ETCD_ENDPOINTS": ",".join(job_env.etcd_endpoints),
        "PADDLE_TRAINER_ID": str(trainer.global_rank),
        "PADDLE_TRAINER_RANK_IN_POD": str(trainer.rank_in_pod),
        "FLAGS_selected_gpus": ",".join(str(g) for g in trainer.gpus),
        "PADDLE_CURRENT_ENDPOINT": str(trainer.endpoint),
        "PADDLE_TRAINERS_NUM": str(cluster.get_trainers_nranks()),
        "PADDLE_TRAINER_ENDPOINTS": ",".join(eps),
        "EDL_POD_LEADER_ID": str(cluster.get_leader_id()),
        "EDL_POD_IDS": ",".join(cluster.get_pods_ids_list()),
        "EDL_STAGE": str(cluster.stage),
        "PADDLE_EDL_HDFS_PATH": str(getattr(job_env, "hdfs_path", "") or ""),
        # torch.distributed view of the same rank table
        "MASTER_ADDR": master_host,
        "MASTER_PORT": master_port,
        "RANK": str(trainer.global_rank),
        "WORLD_SIZE": str(cluster.get_trainers_nranks()),
        "LOCAL_RANK": str(trainer.rank_in_pod),
    }


def _child_setup():
    """Runs in the forked child before exec: own session (so the whole trainer tree can be signalled as a group) and
    PR_SET_PDEATHSIG, so a launcher that is SIGKILLed or crashes never leaves orphan trainers holding GPUs and
    spinning in collectives (Linux only; silently skipped elsewhere)."""
    os.setsid()
    try:
        import ctypes

        libc = ctypes.CDLL("libc.so.6", use_errno=True)
        libc.prctl(1, int(signal.SIGKILL), 0, 0, 0)      # PR_SET_PDEATHSIG = 1
    except Exception:  # noqa: BLE001
        pass


def start(job_env, cluster, pod, training_script, training_script_args, log_dir=None):
    base_env = dict(os.environ)
    # proxies can make peers unreachable during communicator bootstrap
    base_env.pop("http_proxy", None)
    base_env.pop("https_proxy", None)
    procs = []
    for idx, t in enumerate(pod.trainers):
        env = dict(base_env)
        env.update(trainer_env(job_env, cluster, pod, t))
        cmd = [sys.executable, "-u", training_script] + list(training_script_args)
        fn = None
        if log_dir is not None:
            os.makedirs(log_dir, exist_ok=True)
            fn = open(os.path.join(log_dir, "workerlog.%d" % idx), "a")
            proc = subprocess.Popen(cmd, env=env, stdout=fn, stderr=fn, preexec_fn=_child_setup)
        else:
            proc = subprocess.Popen(cmd, env=env, preexec_fn=_child_setup)
        tp = TrainerProc()
        tp.proc, tp.rank, tp.log_fn, tp.local_rank, tp.cmd = proc, t.global_rank, fn, idx, cmd
        tp.log_offset = fn.tell() if fn else None
        procs.append(tp)
        logger.debug("started trainer rank %s pid %d", t.global_rank, proc.pid)
    logger.info("stage %s: started %d trainers of %d", cluster.stage, len(procs), cluster.get_trainers_nranks())
    return procs


def _descendants(procs):
    out = []
    for tp in procs:
        try:
            p = psutil.Process(tp.proc.pid)
            out.append(p)
            out.extend(p.children(recursive=True))
        except psutil.NoSuchProcess:
            pass
    return out


def terminate(procs, grace=None):
    """SIGTERM the trainers and all their descendants, SIGKILL what is left after ``grace`` s."""
    grace = constants.KILL_GRACE if grace is None else grace
    victims = _descendants(procs)
    for tp in procs:
        if tp.log_fn is not None:
            try:
                tp.log_fn.close()
            except Exception:  # noqa: BLE001
                pass
            tp.log_fn = None
    for p in victims:
        try:
            p.send_signal(signal.SIGTERM)
        except psutil.NoSuchProcess:
            pass
    gone, alive = psutil.wait_procs(victims, timeout=grace)
    for p in alive:
        try:
            p.kill()
        except psutil.NoSuchProcess:
            pass
    gone, alive = psutil.wait_procs(alive, timeout=1)
    for tp in procs:
        try:
            tp.proc.wait(timeout=1)
        except Exception:  # noqa: BLE001
            pass
    if alive:
        logger.error("could not kill %s", alive)
        return False
    logger.info("terminated %d trainer process trees", len(procs))
    return True


def pull_worker_log(tp, out=sys.stdout):
    if tp.log_fn is None:
        return
    try:
        with open(tp.log_fn.name, "r", errors="replace") as fin:
            fin.seek(tp.log_offset, 0)
            for line in fin:
                out.write(line)
            tp.log_offset = fin.tell()
    except OSError:
        pass


def watch(procs):
    """One supervision poll: returns (alive, failed_exit_code).

    alive=True while any trainer is still running; when all exited, failed_exit_code is None on
    success or the first non-zero exit code."""
    alive = False
    failed = None
    for tp in procs:
        if tp.local_rank == 0:
            pull_worker_log(tp)
        ret = tp.proc.poll()
        if ret is None:
            alive = True
        elif ret != 0 and failed is None:
            failed = ret
    if failed is not None:
        return False, failed
    return alive, None


def wait_all(procs, timeout=None, poll=0.2):
    begin = time.time()
    while True:
        alive, failed = watch(procs)
        if failed is not None or not alive:
            return failed
        if timeout is not None and time.time() - begin > timeout:
            return -1
        time.sleep(poll)
