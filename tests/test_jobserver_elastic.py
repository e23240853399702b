"""JobServer/JobClient driven elasticity on CPU/gloo with the fit_a_line workload: pods are added and
removed by the JobServer schedule while training keeps resuming from atomic checkpoints with the LR
rescaled to the world size (BASELINE.json config 0; reference demo: README.md:121-160)."""
import json
import os
import threading
import time
import urllib.request
import uuid

import pytest

from edl_b200.demo.collective.job_client_demo import JobClient
from edl_b200.demo.collective.job_server_demo import JobServer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRAIN = os.path.join(ROOT, "examples", "fit_a_line", "train.py")


def test_jobserver_http_api():
    with JobServer(["10.0.0.1", "10.0.0.2"], pod_num_of_node=2, gpu_num_of_node=8, time_interval_to_change=0,
                   port=0, host="127.0.0.1") as js:
        base = "http://127.0.0.1:%d" % js.port
        snap = json.loads(urllib.request.urlopen(base + "/job/j/pods").read())
        assert len(snap["pods"]) == 4 and all(p["running"] for p in snap["pods"].values())
        node = json.loads(urllib.request.urlopen(base + "/job/j/node/10.0.0.2").read())
        assert len(node["pods"]) == 2 and sorted(p["gpus"] for p in node["pods"].values()) == [[0, 1, 2, 3], [4, 5, 6, 7]]
        keep = sorted(snap["pods"])[:1]
        req = urllib.request.Request(base + "/job/j/schedule", data=json.dumps({"running": keep}).encode(), method="POST")
        v = json.loads(urllib.request.urlopen(req).read())["version"]
        snap = json.loads(urllib.request.urlopen(base + "/job/j/pods").read())
        assert snap["version"] == v and [k for k, p in snap["pods"].items() if p["running"]] == keep
        for _ in range(5):
            js.state.flip()
            assert sum(p["running"] for p in js.state.snapshot()["pods"].values()) >= 1


@pytest.mark.slow
def test_elastic_fit_a_line_2_1_2_pods(kv_server, tmp_path):
    job = "fit_" + uuid.uuid4().hex[:6]
    ckpt, report = str(tmp_path / "ckpt"), str(tmp_path / "report")
    with JobServer(["127.0.0.1"], pod_num_of_node=2, gpu_num_of_node=0, time_interval_to_change=0, port=0,
                   host="127.0.0.1") as js:
        os.environ["FIT_REPORT_DIR"] = report
        cli = JobClient("http://127.0.0.1:%d" % js.port, job, TRAIN, node_ip="127.0.0.1",
                        etcd_endpoints=kv_server.endpoint, nodes_range="1:2", poll_s=0.3,
                        extra_launch_args=["--nproc_per_node", "1", "--hdfs_path", ckpt],
                        log_dir=str(tmp_path / "log"))
        # the trainer reads its arguments from the pod script's argv: wrap with a tiny pod script
        pod = tmp_path / "pod.py"
        pod.write_text("import sys, runpy\nsys.argv = [%r, '--epochs', '400', '--epoch_sleep', '0.05', '--ckpt', %r]\n"
                       "runpy.run_path(%r, run_name='__main__')\n" % (TRAIN, ckpt, TRAIN))
        cli.pod_path = str(pod)
        t = threading.Thread(target=cli.run, daemon=True)
        t.start()

        def epochs():
            p = os.path.join(report, "epochs.jsonl")
            return [json.loads(l) for l in open(p)] if os.path.exists(p) else []

        def wait_world(w, timeout, min_new=3):
            n0 = len(epochs())
            deadline = time.time() + timeout
            while time.time() < deadline:
                e = epochs()
                if len(e) >= n0 + min_new and all(x["world"] == w for x in e[-min_new:]):
                    return e
                time.sleep(0.3)
            raise AssertionError("world never became %d: %s" % (w, epochs()[-5:]))

        pods = sorted(js.state.snapshot()["pods"])
        e2 = wait_world(2, 90)
        js.state.set_running({pods[0]})                      # scale in: 2 -> 1
        e1 = wait_world(1, 90)
        js.state.set_running(set(pods))                      # scale out: 1 -> 2
        e2b = wait_world(2, 90)
        cli.stop()
        t.join(30)
        os.environ.pop("FIT_REPORT_DIR", None)
    ep = [x["epoch"] for x in e2b]
    assert ep == sorted(set(ep)), "epochs must resume exactly where the checkpoint left off: %s" % ep
    lr2, lr1 = e2[-1]["lr"], e1[-1]["lr"]
    assert abs(lr1 * 2 - lr2) < 1e-9 and abs(e2b[-1]["lr"] - lr2) < 1e-9   # linear LR rescale both ways
    assert e2b[-1]["loss"] < e2[0]["loss"]                                  # and it keeps learning
