#!/usr/bin/env python
"""Rescale recovery time through the real launcher (BASELINE.json "rescale recovery time after +-1 pod"), CPU / gloo,
fit_a_line: pod A runs alone, pod B joins, the leader's ScaleIn RPC evicts B.  Measured for the reference's
stop-resume mode and for the in-place mode (edl_b200/elastic.py):

  join   = first epoch reported at world 2  -  launch of pod B          (includes B's interpreter + torch import)
  stall  = longest gap between two consecutive epoch reports around the change (what training actually lost)
  leave  = first epoch reported at world 1  -  ScaleIn RPC

    python tools/bench_elastic_launch.py [--native-store] [--out profiles/elastic_launch_cpu.json]
    python tools/bench_elastic_launch.py --trainer resnet --gpus-per-pod 4 --out gpurun_out/elastic_launch_8gpu.json
        # on an 8-GPU box: pod A = GPUs 0-3, pod B = GPUs 4-7, ResNet50_vd, 4 -> 8 -> 4 trainers
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
import uuid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRAIN = os.path.join(ROOT, "examples", "fit_a_line", "train.py")

from edl_b200.discovery.etcd_client import EtcdClient  # noqa: E402
from edl_b200.store import KVServer, NativeKVServer  # noqa: E402
from edl_b200.utils import leader_pod, pod_server_client  # noqa: E402


CFG = {"trainer": "fit", "gpus_per_pod": 0, "leave": "scale_in"}
RESNET = os.path.join(ROOT, "examples", "collective", "resnet50", "train.py")


def launch(endpoint, job, tmp, name, mode):
    g = CFG["gpus_per_pod"]
    first = 0 if name == "A" else g
    env = dict(os.environ, PYTHONPATH=ROOT, PADDLE_RUNNING_PLATFORM="", EDL_POD_IP="127.0.0.1",
               CUDA_VISIBLE_DEVICES=",".join(str(first + i) for i in range(g)),
               FIT_REPORT_DIR=os.path.join(tmp, "report"), EDL_PROGRESS_FILE=os.path.join(tmp, "report", "epochs.jsonl"),
               EDL_INPLACE_CHECK_EVERY="3" if CFG["trainer"] == "fit" else "10",
               EDL_ETCD_TTL="1.5", EDL_POLL_INTERVAL="0.3", EDL_KILL_GRACE="1",
               FIT_INJECT_FAULT_FILE=os.path.join(tmp, "fault.now"), RESNET_INJECT_FAULT_FILE=os.path.join(tmp, "fault.now"))
    os.makedirs(os.path.join(tmp, "report"), exist_ok=True)
    if CFG["trainer"] == "fit":
        script = [TRAIN, "--epochs", "100000", "--epoch_sleep", "0.02", "--ckpt", os.path.join(tmp, "ckpt")]
    else:
        script = [RESNET, "--model", "ResNet50_vd", "--epochs", "1000", "--steps_per_epoch", "400",
                  "--ckpt", os.path.join(tmp, "ckpt")]
    cmd = [sys.executable, "-u", "-m", "edl_b200.collective.launch", "--nodes_range", "1:2",
           "--nproc_per_node", str(max(1, g)),
           "--etcd_endpoints", endpoint, "--job_id", job, "--log_dir", os.path.join(tmp, "log" + name),
           "--hdfs_path", os.path.join(tmp, "ckpt"), "--rescale_mode", mode] + script
    return subprocess.Popen(cmd, env=env, stdout=open(os.path.join(tmp, name + ".log"), "w"), stderr=subprocess.STDOUT,
                            start_new_session=True)


def run(mode, server_cls):
    tmp = tempfile.mkdtemp(prefix="edl_elastic_")
    job = "bench_" + uuid.uuid4().hex[:6]

    def epochs():
        p = os.path.join(tmp, "report", "epochs.jsonl")
        return [json.loads(l) for l in open(p)] if os.path.exists(p) else []

    def wait_world(w, timeout=120, min_new=5):
        n0 = len(epochs())
        deadline = time.time() + timeout
        while time.time() < deadline:
            e = epochs()
            if len(e) >= n0 + min_new and all(x["world"] == w for x in e[-min_new:]):
                return e
            time.sleep(0.05)
        raise RuntimeError("world never became %d (%s)" % (w, mode))

    def stall(e, t_event):
        ts = [x["t"] for x in e if x["t"] > t_event - 2.0]
        return max(b - a for a, b in zip(ts[:-1], ts[1:]))

    w1 = max(1, CFG["gpus_per_pod"])
    w2 = 2 * w1
    with server_cls() as srv:
        a = launch(srv.endpoint, job, tmp, "A", mode)
        b = None
        try:
            wait_world(w1)
            t_join = time.time()
            b = launch(srv.endpoint, job, tmp, "B", mode)
            e = wait_world(w2, timeout=300)
            join = min(x["t"] for x in e if x["world"] == w2 and x["t"] > t_join) - t_join
            join_stall = stall(e, t_join)
            pids_before = {x["pid"] for x in e if x["world"] == w1 and x["t"] < t_join}
            survivor_kept = e[-1]["pid"] in pids_before
            if CFG["leave"] == "false_alarm":
                # nobody leaves: rank 1 reports ONE failed collective; both trainers soft-reset (same stage, new generation)
                t_ev = time.time()
                open(os.path.join(tmp, "fault.now"), "w").close()
                time.sleep(1.0)
                e = wait_world(w2, timeout=120, min_new=8)
                return {"mode": mode, "leave": "false_alarm", "store": server_cls.__name__, "join_s": join,
                        "join_stall_s": join_stall, "false_alarm_stall_s": stall(e, t_ev),
                        "survivor_process_kept": e[-1]["pid"] in {x["pid"] for x in e if x["t"] < t_ev},
                        "steady_epoch_s": sorted(y["t"] - x["t"] for x, y in zip(e[-5:-1], e[-4:]))[1]}
            etcd = EtcdClient([srv.endpoint], root=job)
            etcd.init()
            t_leave = time.time()
            if CFG["leave"] == "scale_in":
                cli = pod_server_client.Client(leader_pod.load_from_etcd(etcd, timeout=5).endpoint)
                cli.scale_in(1)
                cli.close()
            elif CFG["leave"] == "sigterm":
                b.terminate()
            else:
                import psutil

                for q in [psutil.Process(b.pid)] + psutil.Process(b.pid).children(recursive=True):
                    try:
                        q.kill()
                    except psutil.NoSuchProcess:
                        pass
            e = wait_world(w1, timeout=300)
            leave = min(x["t"] for x in e if x["world"] == w1 and x["t"] > t_leave) - t_leave
            leave_stall = stall(e, t_leave)
            etcd.close()
            return {"mode": mode, "leave": CFG["leave"], "store": server_cls.__name__, "join_s": join, "join_stall_s": join_stall,
                    "leave_s": leave, "leave_stall_s": leave_stall, "survivor_process_kept": survivor_kept,
                    "steady_epoch_s": sorted(y["t"] - x["t"] for x, y in zip(e[-5:-1], e[-4:]))[1]}
        except Exception:
            # keep the evidence: launcher and trainer logs of both pods (gpurun_out/ travels back from the GPU box)
            import shutil
            dst = os.path.join(ROOT, "gpurun_out", "elastic_fail_%s_%s" % (mode, CFG["leave"]))
            shutil.rmtree(dst, ignore_errors=True)
            shutil.copytree(tmp, dst, ignore=shutil.ignore_patterns("ckpt*", "*.pt"))
            raise
        finally:
            import psutil

            for p in (a, b):
                if p is None:
                    continue
                try:
                    tree = [psutil.Process(p.pid)] + psutil.Process(p.pid).children(recursive=True)
                except psutil.NoSuchProcess:
                    continue
                for q in tree:            # launchers AND their trainers (own sessions): nothing may outlive the bench
                    try:
                        q.kill()
                    except psutil.NoSuchProcess:
                        pass


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--native-store", action="store_true")
    ap.add_argument("--trainer", default="fit", choices=["fit", "resnet"])
    ap.add_argument("--gpus-per-pod", type=int, default=0, help="GPUs (= trainers) per pod; 0 = CPU / gloo")
    ap.add_argument("--leave", default="scale_in", choices=["scale_in", "sigterm", "kill", "false_alarm"],
                    help="how pod B leaves: the leader's ScaleIn RPC, SIGTERM to its launcher (graceful leave), or SIGKILL "
                         "of launcher and trainers (hot recovery in place / restart of everybody in restart mode)")
    ap.add_argument("--out", default="")
    ap.add_argument("--modes", default="restart,inplace", help="comma-separated subset of restart,inplace")
    args = ap.parse_args()
    CFG.update(trainer=args.trainer, gpus_per_pod=args.gpus_per_pod, leave=args.leave)
    cls = NativeKVServer if args.native_store else KVServer
    res = []
    for mode in [m for m in args.modes.split(",") if m]:
        try:
            res.append(run(mode, cls))
        except Exception as e:  # noqa: BLE001 - the other mode is still worth measuring
            res.append({"mode": mode, "leave": CFG["leave"], "error": repr(e)[:300]})
    for r in res:
        print(json.dumps(r))
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"trainer": CFG["trainer"], "gpus_per_pod": CFG["gpus_per_pod"],
                       "note": "fit_a_line runs are CPU / gloo; test timing constants (TTL 1.5 s, polls 0.3 s, kill grace 1 s); "
                               "the reference's constants are 15 s / 3 s / 3 s (BASELINE.md)", "runs": res}, f, indent=1)
