"""liveft np-matching rendezvous: 2 nodes needed; jobs start only when both are there; a failing
child maps to exit code 101 (level 1); rank env is exported."""
import glob
import json
import os
import subprocess
import sys
import time
import uuid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "tests", "launch_demo.py")


def _node(endpoint, job, np_, env_extra, out):
    env = dict(os.environ)
    env.update({"PYTHONPATH": ROOT, "POD_IP": "127.0.0.1", "EDL_POLL_INTERVAL": "0.3", "EDL_ETCD_TTL": "1.5"})
    env.update(env_extra)
    return subprocess.Popen([sys.executable, "-m", "edl_b200.liveft.launch", "--elastic_server", endpoint,
                             "--job_id", job, "--np", str(np_), DEMO], env=env, stdout=open(out, "w"),
                            stderr=subprocess.STDOUT)


def test_np_match_and_completion(kv_server, tmp_path):
    job = "lf_" + uuid.uuid4().hex[:6]
    rec = str(tmp_path / "rec")
    env = {"DEMO_RECORD_DIR": rec, "PADDLE_POD_ID": "x", "DEMO_RUN_SECONDS": "1.5"}
    a = _node(kv_server.endpoint, job, 2, env, str(tmp_path / "a.log"))
    time.sleep(2.0)
    assert a.poll() is None and not glob.glob(rec + "/start_*")   # one node of two: must keep waiting
    b = _node(kv_server.endpoint, job, 2, env, str(tmp_path / "b.log"))
    assert a.wait(30) == 0 and b.wait(30) == 0, open(str(tmp_path / "a.log")).read()[-2000:]
    starts = [json.load(open(f)) for f in glob.glob(rec + "/start_*")]
    assert len(starts) == 2
    assert sorted(s["PADDLE_TRAINER_ID"] for s in starts) == ["0", "1"]


def test_child_failure_maps_to_restart_code(kv_server, tmp_path):
    job = "lf_" + uuid.uuid4().hex[:6]
    p = _node(kv_server.endpoint, job, 1, {"PADDLE_DEMO_EXIT_CODE": "5"}, str(tmp_path / "c.log"))
    assert p.wait(30) == 101
    p = _node(kv_server.endpoint, job + "b", 1, {"PADDLE_DEMO_EXIT_CODE": "5", "PADDLE_ELASTIC_FAULT_TOLERANC_LEVEL": "2"},
              str(tmp_path / "d.log"))
    assert p.wait(30) == 3


def test_level2_shrinks_to_the_surviving_node(kv_server, tmp_path):
    """Fault-tolerance level 2: when a node is lost for good the job continues on the survivors (np is lowered in
    the store) instead of holding for a replacement."""
    import signal

    job = "lf_" + uuid.uuid4().hex[:6]
    rec = str(tmp_path / "rec")
    env = {"DEMO_RECORD_DIR": rec, "PADDLE_POD_ID": "x", "DEMO_RUN_SECONDS": "4", "PADDLE_ELASTIC_FAULT_TOLERANC_LEVEL": "2"}
    a = _node(kv_server.endpoint, job, 2, env, str(tmp_path / "a.log"))
    b = _node(kv_server.endpoint, job, 2, env, str(tmp_path / "b.log"))
    deadline = time.time() + 20
    while time.time() < deadline and len(glob.glob(rec + "/start_*")) < 2:
        time.sleep(0.1)
    assert len(glob.glob(rec + "/start_*")) == 2
    b.send_signal(signal.SIGKILL)                 # the node disappears without cleaning up; its lease expires
    assert a.wait(60) == 0, open(str(tmp_path / "a.log")).read()[-2000:]
    starts = sorted((json.load(open(f)) for f in glob.glob(rec + "/start_*")), key=lambda s: s["t"])
    assert len(starts) == 3                        # two initial trainers + the survivor's relaunch
    assert starts[-1]["PADDLE_TRAINERS_NUM"] == "1" and starts[-1]["PADDLE_TRAINER_ID"] == "0"
