"""End-to-end launcher tests with several launcher processes on one host against one store
(reference: tests/unittests/test_launch.sh:40-88): static 2-pod job, job-status idempotence,
trainer failure, and elastic join / leave with trainer restart."""
import glob
import json
import os
import signal
import subprocess
import sys
import time
import uuid


from edl_b200.discovery.etcd_client import EtcdClient
from edl_b200.utils import cluster as edl_cluster, status as edl_status

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "tests", "launch_demo.py")


def _launch(endpoint, job_id, nodes_range, log_dir, extra_env=None, gpus="0"):
    env = dict(os.environ)
    env.update({"PYTHONPATH": ROOT, "CUDA_VISIBLE_DEVICES": gpus, "PADDLE_RUNNING_PLATFORM": "",
                "EDL_POD_IP": "127.0.0.1"})
    env.update(extra_env or {})
    cmd = [sys.executable, "-u", "-m", "edl_b200.collective.launch", "--nodes_range", nodes_range,
           "--nproc_per_node", "1", "--etcd_endpoints", endpoint, "--job_id", job_id, "--log_dir", log_dir,
           "--log_level", "10", DEMO]
    return subprocess.Popen(cmd, env=env, stdout=open(log_dir + ".launcher.log", "w"), stderr=subprocess.STDOUT,
                            start_new_session=True)


def _wait(procs, timeout):
    deadline = time.time() + timeout
    for p in procs:
        p.wait(timeout=max(1, deadline - time.time()))
    return [p.returncode for p in procs]


def test_two_pods_complete_and_relaunch_is_noop(kv_server, tmp_path):
    job = "job_" + uuid.uuid4().hex[:6]
    rec = str(tmp_path / "rec")
    procs = [_launch(kv_server.endpoint, job, "2:2", str(tmp_path / ("log%d" % i)), {"DEMO_RECORD_DIR": rec},
                     gpus=str(i)) for i in range(2)]
    assert _wait(procs, 90) == [0, 0], open(str(tmp_path / "log0.launcher.log")).read()[-3000:]
    starts = [json.load(open(f)) for f in glob.glob(rec + "/start_*.json")]
    assert len(starts) == 2
    assert sorted(s["PADDLE_TRAINER_ID"] for s in starts) == ["0", "1"]
    assert all(s["PADDLE_TRAINERS_NUM"] == "2" and len(s["PADDLE_TRAINER_ENDPOINTS"].split(",")) == 2 for s in starts)
    assert all(s["WORLD_SIZE"] == "2" and s["EDL_POD_LEADER_ID"] in s["EDL_POD_IDS"] for s in starts)
    etcd = EtcdClient([kv_server.endpoint], root=job)
    etcd.init()
    assert edl_status.load_job_status_from_etcd(etcd) == edl_status.Status.SUCCEED
    # job-level idempotence: a relaunched pod exits immediately with 0
    again = _launch(kv_server.endpoint, job, "2:2", str(tmp_path / "log_again"))
    assert _wait([again], 30) == [0]


def test_trainer_failure_fails_the_job(kv_server, tmp_path):
    job = "job_" + uuid.uuid4().hex[:6]
    procs = [_launch(kv_server.endpoint, job, "1:1", str(tmp_path / "log"), {"PADDLE_DEMO_EXIT_CODE": "7"})]
    assert _wait(procs, 60) == [1]
    etcd = EtcdClient([kv_server.endpoint], root=job)
    etcd.init()
    assert edl_status.load_job_status_from_etcd(etcd) == edl_status.Status.FAILED


def test_elastic_join_and_leave(kv_server, tmp_path):
    """1 pod running -> a 2nd joins (trainers restart with world 2) -> the 2nd is killed (world 1
    again) -> training finishes.  This is the reference's stop-resume elasticity."""
    job = "job_" + uuid.uuid4().hex[:6]
    rec = str(tmp_path / "rec")
    done = str(tmp_path / "done.flag")
    env = {"DEMO_RECORD_DIR": rec, "DEMO_RUN_SECONDS": "120", "DEMO_DONE_FLAG": done}
    from edl_b200.utils.network_utils import find_free_ports
    mport = find_free_ports(1)[0]
    a = _launch(kv_server.endpoint, job, "1:2", str(tmp_path / "logA"), dict(env, EDL_METRICS_PORT=str(mport)), gpus="0")

    def metrics():
        import urllib.request
        body = urllib.request.urlopen("http://127.0.0.1:%d/metrics" % mport, timeout=5).read().decode()
        return {ln.split("{")[0]: float(ln.rsplit(" ", 1)[1]) for ln in body.splitlines() if ln}

    def worlds():
        return sorted((json.load(open(f))["t"], json.load(open(f))["WORLD_SIZE"]) for f in glob.glob(rec + "/start_*.json"))

    def wait_for(pred, timeout, what):
        deadline = time.time() + timeout
        while time.time() < deadline:
            if pred():
                return
            time.sleep(0.2)
        raise AssertionError("timeout waiting for %s; starts=%s\n%s" % (
            what, worlds(), open(str(tmp_path / "logA.launcher.log")).read()[-4000:]))

    wait_for(lambda: [w for _, w in worlds()] == ["1"], 40, "first start with world 1")
    t_join = time.time()
    b = _launch(kv_server.endpoint, job, "1:2", str(tmp_path / "logB"), env, gpus="1")
    wait_for(lambda: [w for _, w in worlds()].count("2") == 2, 60, "both pods restarted with world 2")
    join_latency = max(t for t, w in worlds() if w == "2") - t_join
    wait_for(lambda: metrics().get("edl_launcher_world_size") == 2, 10, "launcher metrics to show world 2")
    m = metrics()                                  # the launcher's Prometheus endpoint saw the elastic event
    assert m["edl_launcher_world_size"] == 2 and m["edl_launcher_rescales_total"] >= 1 and m["edl_launcher_is_leader"] == 1
    assert m["edl_launcher_trainers_alive"] == 1 and m["edl_launcher_last_rescale_seconds"] > 0
    os.killpg(os.getpgid(b.pid), signal.SIGKILL)   # pod B dies hard (no clean deregistration)
    t_leave = time.time()
    wait_for(lambda: [w for _, w in worlds()][-1] == "1" and len(worlds()) == 4, 60, "A back to world 1")
    leave_latency = worlds()[-1][0] - t_leave
    open(done, "w").write("x")
    assert _wait([a], 60) == [0]
    etcd = EtcdClient([kv_server.endpoint], root=job)
    etcd.init()
    c = edl_cluster.load_from_etcd(etcd)
    assert len(c.pods) == 1
    print("join latency %.2fs leave latency %.2fs" % (join_latency, leave_latency))
