"""JobClient demo: runs on every node, polls the JobServer and starts / stops one elastic launcher
per pod the server wants running on this node.

Documented surface of the (absent) reference module kept: flags ``--log_level --package_sh
--pod_path <train script>``; environment ``PADDLE_RUNING_ENV=PADDLE_EDL``, ``PADDLE_JOBSERVER``,
``PADDLE_JOB_ID``, ``PADDLE_POD_ID`` (README.md:149-153, example/demo/collective/
start_job_client.sh:21-37); ``package.sh -pod_id X`` is executed to materialise the pod's working
directory before its launcher starts (example/demo/collective/resnet50/package.sh:19-52).
"""
import argparse
import json
import logging
import os
import signal
import subprocess
import sys
import threading
import time
import urllib.request

from ...utils.network_utils import get_extern_ip

logger = logging.getLogger("edl.jobclient")


class JobClient:
    def __init__(self, job_server, job_id, pod_path, package_sh=None, node_ip=None, etcd_endpoints=None,
                 nodes_range="1:8", poll_s=3.0, extra_launch_args=None, log_dir="./log", python=sys.executable):
        self.job_server = job_server.rstrip("/")
        self.job_id, self.pod_path, self.package_sh = job_id, pod_path, package_sh
        self.node_ip = node_ip or get_extern_ip()
        self.etcd_endpoints = etcd_endpoints or os.environ.get("PADDLE_ETCD_ENDPOINTS", "127.0.0.1:2379")
        self.nodes_range, self.poll_s = nodes_range, poll_s
        self.extra = list(extra_launch_args or [])
        self.log_dir, self.python = log_dir, python
        self.procs = {}
        self.version = -1
        self._stop = threading.Event()
        self.events = []   # (time, "start"/"stop", pod_id) -- for tests / reports

    def _get(self, path):
        with urllib.request.urlopen(self.job_server + path, timeout=5) as r:
            return json.loads(r.read())

    def _package(self, pod_id):
        if self.package_sh:
            subprocess.run(["bash", self.package_sh, "-pod_id", pod_id], check=False)

    def _start_pod(self, pod_id, gpus):
        self._package(pod_id)
        env = dict(os.environ)
        env.update({"PADDLE_RUNING_ENV": "PADDLE_EDL", "PADDLE_JOB_ID": self.job_id, "PADDLE_POD_ID": pod_id,
                    "PADDLE_JOBSERVER": self.job_server})
        if gpus:
            env["CUDA_VISIBLE_DEVICES"] = ",".join(str(g) for g in gpus)
        cmd = [self.python, "-u", "-m", "edl_b200.collective.launch", "--nodes_range", self.nodes_range,
               "--etcd_endpoints", self.etcd_endpoints, "--job_id", self.job_id,
               "--log_dir", os.path.join(self.log_dir, pod_id)] + self.extra + [self.pod_path]
        os.makedirs(self.log_dir, exist_ok=True)
        out = open(os.path.join(self.log_dir, "launcher_%s.log" % pod_id), "a")
        p = subprocess.Popen(cmd, env=env, stdout=out, stderr=subprocess.STDOUT, start_new_session=True)
        self.procs[pod_id] = p
        self.events.append((time.time(), "start", pod_id))
        logger.info("started pod %s (pid %d)", pod_id, p.pid)

    def _stop_pod(self, pod_id, sig=signal.SIGTERM):
        p = self.procs.pop(pod_id, None)
        if p is None:
            return
        try:
            os.killpg(os.getpgid(p.pid), sig)
        except ProcessLookupError:
            pass
        try:
            p.wait(timeout=10)
        except subprocess.TimeoutExpired:
            os.killpg(os.getpgid(p.pid), signal.SIGKILL)
        self.events.append((time.time(), "stop", pod_id))
        logger.info("stopped pod %s", pod_id)

    def reconcile_once(self):
        snap = self._get("/job/%s/node/%s" % (self.job_id, self.node_ip))
        want = {pid: p for pid, p in snap["pods"].items() if p["running"]}
        for pid in list(self.procs):
            if self.procs[pid].poll() is not None:     # launcher exited by itself (job done / failed)
                self.procs.pop(pid)
        for pid in list(self.procs):
            if pid not in want:
                self._stop_pod(pid)
        if snap["version"] != self.version or True:
            for pid, p in want.items():
                if pid not in self.procs and not self._finished(pid):
                    self._start_pod(pid, p.get("gpus"))
        self.version = snap["version"]

    def _finished(self, pod_id):
        return False

    def run(self):
        while not self._stop.is_set():
            try:
                self.reconcile_once()
            except Exception as e:  # noqa: BLE001
                logger.warning("job server poll failed: %s", e)
            self._stop.wait(self.poll_s)
        for pid in list(self.procs):
            self._stop_pod(pid)

    def stop(self):
        self._stop.set()


def main(argv=None):
    ap = argparse.ArgumentParser(description="EDL JobClient demo")
    ap.add_argument("--log_level", type=int, default=20)
    ap.add_argument("--package_sh", type=str, default=None)
    ap.add_argument("--pod_path", type=str, required=True, help="the training entry script of a pod")
    ap.add_argument("--nodes_range", type=str, default=os.environ.get("PADDLE_EDLNODES_RANAGE", "1:8"))
    args = ap.parse_args(argv)
    logging.basicConfig(level=args.log_level)
    server = os.environ.get("PADDLE_JOBSERVER", "http://127.0.0.1:8180")
    job_id = os.environ.get("PADDLE_JOB_ID", "edl_demo_job")
    cli = JobClient(server, job_id, args.pod_path, args.package_sh, nodes_range=args.nodes_range)
    signal.signal(signal.SIGTERM, lambda *a: cli.stop())
    cli.run()


if __name__ == "__main__":
    main()
