#!/usr/bin/env python
"""Pod discovery helpers for Kubernetes deployments (reference: k8s/k8s_tools.py:29-184, py2):
``fetch_ips / fetch_endpoints / fetch_id / wait_pods_running / count_pods_by_phase``.
Uses the ``kubernetes`` client when it is installed, else ``kubectl``.

    python k8s/k8s_tools.py fetch_endpoints edl-job=myjob 6170
"""
import json
import os
import subprocess
import sys
import time

NAMESPACE = os.getenv("NAMESPACE", "default")


def _list_pods(label_selector):
    try:
        from kubernetes import client, config

        if os.getenv("KUBERNETES_SERVICE_HOST"):
            config.load_incluster_config()
        else:
            config.load_kube_config()
        items = client.CoreV1Api().list_namespaced_pod(namespace=NAMESPACE, label_selector=label_selector).items
        return [{"name": p.metadata.name, "ip": p.status.pod_ip, "phase": p.status.phase,
                 "start": str(p.status.start_time)} for p in items]
    except ImportError:
        out = subprocess.run(["kubectl", "get", "pods", "-n", NAMESPACE, "-l", label_selector, "-o", "json"],
                             capture_output=True, text=True, check=True).stdout
        return [{"name": p["metadata"]["name"], "ip": p["status"].get("podIP"), "phase": p["status"].get("phase"),
                 "start": p["status"].get("startTime", "")} for p in json.loads(out)["items"]]


def count_pods_by_phase(label_selector, phase):
    return sum(1 for p in _list_pods(label_selector) if p["phase"] == phase)


def fetch_pods_info(label_selector, phase=None):
    return sorted(((p["start"], p["ip"], p["name"]) for p in _list_pods(label_selector)
                   if phase is None or p["phase"] == phase))


def wait_pods_running(label_selector, desired, interval=5):
    while True:
        n = count_pods_by_phase(label_selector, "Running")
        print("label selector: %s, desired: %s, running: %d" % (label_selector, desired, n), flush=True)
        if n >= int(desired):
            return
        time.sleep(interval)


def fetch_ips(label_selector):
    return ",".join(ip for _, ip, _ in fetch_pods_info(label_selector, "Running") if ip)


def fetch_endpoints(label_selector, port):
    return ",".join("%s:%s" % (ip, port) for _, ip, _ in fetch_pods_info(label_selector, "Running") if ip)


def fetch_id(label_selector, my_ip=None):
    """Rank of this pod among the running pods of the job (ordered by start time, then ip)."""
    my_ip = my_ip or os.getenv("POD_IP")
    ips = [ip for _, ip, _ in fetch_pods_info(label_selector, "Running")]
    return ips.index(my_ip) if my_ip in ips else -1


if __name__ == "__main__":
    cmd, args = sys.argv[1], sys.argv[2:]
    fn = {"fetch_ips": fetch_ips, "fetch_endpoints": fetch_endpoints, "fetch_id": fetch_id,
          "wait_pods_running": wait_pods_running, "count_pods_by_phase": count_pods_by_phase}[cmd]
    r = fn(*args)
    if r is not None:
        print(r)
