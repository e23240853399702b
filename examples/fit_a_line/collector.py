#!/usr/bin/env python
"""Cluster utilisation monitor for elastic jobs on Kubernetes (reference: example/fit_a_line/collector.py,
used to plot how EDL fills idle resources).  Polls the pod list through ``k8s/k8s_tools`` (kubernetes
client if installed, else ``kubectl -o json``) and prints one line per interval:

    time  submitted  pending  running-trainers  gpu-requested/allocatable  per-job phase

    python examples/fit_a_line/collector.py --label edl-job --interval 10
"""
import argparse
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "k8s")))

NOT_EXISTS, PENDING, RUNNING, FINISHED, KILLED = "N/A", "PENDING", "RUNNING", "FINISH", "KILLED"


class JobInfo:
    def __init__(self, name):
        self.name, self.status = name, NOT_EXISTS
        self.submit_time = self.start_time = self.end_time = -1.0
        self.parallelism = 0


def _kubectl_json(args):
    out = subprocess.run(["kubectl"] + args + ["-o", "json"], capture_output=True, text=True, check=True).stdout
    return json.loads(out)


class Collector:
    def __init__(self, label_key="edl-job", namespace=None, lister=None, nodes=None):
        self.label_key = label_key
        self.namespace = namespace or os.getenv("NAMESPACE", "default")
        self._lister = lister or self._list_pods
        self._nodes = nodes or self._list_nodes
        self.jobs = {}
        self.t0 = time.time()

    def _list_pods(self):
        items = _kubectl_json(["get", "pods", "-n", self.namespace, "-l", self.label_key])["items"]
        pods = []
        for p in items:
            req = 0
            for c in p["spec"]["containers"]:
                req += int(c.get("resources", {}).get("limits", {}).get("nvidia.com/gpu", 0))
            pods.append({"name": p["metadata"]["name"], "job": p["metadata"]["labels"].get(self.label_key, ""),
                         "phase": p["status"].get("phase", "Unknown"), "gpus": req})
        return pods

    def _list_nodes(self):
        items = _kubectl_json(["get", "nodes"])["items"]
        return sum(int(n["status"].get("allocatable", {}).get("nvidia.com/gpu", 0)) for n in items)

    def run_once(self):
        pods = self._lister()
        now = time.time() - self.t0
        seen = set()
        for p in pods:
            j = self.jobs.setdefault(p["job"], JobInfo(p["job"]))
            seen.add(p["job"])
            if j.submit_time < 0:
                j.submit_time = now
        for name, j in self.jobs.items():
            mine = [p for p in pods if p["job"] == name]
            phases = [p["phase"] for p in mine]
            j.parallelism = sum(1 for ph in phases if ph == "Running")
            if not mine:
                j.status = FINISHED if j.status in (RUNNING, FINISHED) else (KILLED if j.status == PENDING else j.status)
                if j.end_time < 0 and j.status in (FINISHED, KILLED):
                    j.end_time = now
            elif "Running" in phases:
                if j.start_time < 0:
                    j.start_time = now
                j.status = RUNNING
            elif all(ph == "Succeeded" for ph in phases):
                j.status, j.end_time = FINISHED, (now if j.end_time < 0 else j.end_time)
            elif "Failed" in phases and "Pending" not in phases:
                j.status = KILLED
            else:
                j.status = PENDING
        alloc = self._nodes()
        used = sum(p["gpus"] for p in pods if p["phase"] == "Running")
        return {"t": round(now, 1), "submitted": len(self.jobs),
                "pending": sum(1 for j in self.jobs.values() if j.status == PENDING),
                "running_trainers": sum(j.parallelism for j in self.jobs.values()),
                "gpu_util": "%d/%d" % (used, alloc),
                "jobs": {n: "%s:%d" % (j.status, j.parallelism) for n, j in self.jobs.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--label", default="edl-job")
    ap.add_argument("--interval", type=float, default=10.0)
    ap.add_argument("--count", type=int, default=0, help="0 = forever")
    args = ap.parse_args()
    c = Collector(args.label)
    n = 0
    while True:
        print(json.dumps(c.run_once()), flush=True)
        n += 1
        if args.count and n >= args.count:
            return
        time.sleep(args.interval)


if __name__ == "__main__":
    main()
