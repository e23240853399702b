"""In-place elastic rescale end to end (edl_b200/elastic.py): real launchers, real gloo collectives, one store.

pod A alone -> pod B joins: A's trainer KEEPS ITS PROCESS, re-rendezvouses through the store, hands its state to
B's trainer by broadcast and doubles the LR -> the leader's ScaleIn RPC evicts B: B's trainer leaves quietly, A's
trainer continues alone (same process, LR halved) -> the job finishes and is recorded SUCCEED."""
import json
import os
import subprocess
import sys
import time
import uuid

import pytest
import torch.distributed as dist

from edl_b200.discovery.etcd_client import EtcdClient
from edl_b200.utils import leader_pod, pod_server_client, status as edl_status

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRAIN = os.path.join(ROOT, "examples", "fit_a_line", "train.py")


def _launch(endpoint, job, log_dir, report, ckpt, epochs, mode="inplace"):
    env = dict(os.environ)
    env.update({"PYTHONPATH": ROOT, "CUDA_VISIBLE_DEVICES": "", "PADDLE_RUNNING_PLATFORM": "", "EDL_POD_IP": "127.0.0.1",
                "FIT_REPORT_DIR": report, "EDL_INPLACE_CHECK_EVERY": "3", "EDL_INPLACE_ACK_TIMEOUT": "40",
                # `epochs` is only an upper bound: the test ends the job (_finish) once it has seen what it wanted,
                # so a slow pod start on a loaded box cannot make the job finish under the test's feet
                "FIT_FINISH_FILE": report + ".finish"})
    cmd = [sys.executable, "-u", "-m", "edl_b200.collective.launch", "--nodes_range", "1:2", "--nproc_per_node", "1",
           "--etcd_endpoints", endpoint, "--job_id", job, "--log_dir", log_dir, "--log_level", "10",
           "--hdfs_path", ckpt, "--rescale_mode", mode,
           TRAIN, "--epochs", str(epochs), "--epoch_sleep", "0.05", "--ckpt", ckpt]
    return subprocess.Popen(cmd, env=env, stdout=open(log_dir + ".launcher.log", "w"), stderr=subprocess.STDOUT,
                            start_new_session=True)


def _finish(report):
    open(report + ".finish", "w").close()


@pytest.mark.slow
def test_join_and_scale_in_without_restarting_the_survivor(kv_server, tmp_path):
    job = "inplace_" + uuid.uuid4().hex[:6]
    report, ckpt = str(tmp_path / "report"), str(tmp_path / "ckpt")

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
            time.sleep(0.2)
        logs = "".join(open(f).read()[-3000:] for f in (str(tmp_path / "logA.launcher.log"), str(tmp_path / "logB.launcher.log"))
                       if os.path.exists(f))
        raise AssertionError("world never became %d: %s\n%s" % (w, epochs()[-4:], logs))

    a = _launch(kv_server.endpoint, job, str(tmp_path / "logA"), report, ckpt, 2000)
    b = None
    try:
        e1 = wait_world(1, 60)
        pid_a = e1[-1]["pid"]
        b = _launch(kv_server.endpoint, job, str(tmp_path / "logB"), report, ckpt, 2000)
        e2 = wait_world(2, 90)
        assert e2[-1]["pid"] == pid_a, "the surviving trainer was restarted on scale-out"
        assert abs(e2[-1]["lr"] - 2 * e1[-1]["lr"]) < 1e-9                       # linear LR rescale, in place
        # scheduler-driven scale-in through the leader's RPC: the evicted pod leaves quietly
        etcd = EtcdClient([kv_server.endpoint], root=job)
        etcd.init()
        leader = leader_pod.load_from_etcd(etcd, timeout=5)
        cli = pod_server_client.Client(leader.endpoint)
        cli.scale_in(1)
        cli.close()
        e1b = wait_world(1, 90)
        assert e1b[-1]["pid"] == pid_a, "the surviving trainer was restarted on scale-in"
        assert abs(e1b[-1]["lr"] - e1[-1]["lr"]) < 1e-9
        _finish(report)
        assert b.wait(timeout=60) == 0                                           # evicted pod: clean exit
        assert a.wait(timeout=120) == 0
        assert edl_status.load_job_status_from_etcd(etcd) == edl_status.Status.SUCCEED
        ep = [x["epoch"] for x in epochs()]
        assert ep == sorted(set(ep)), "an epoch was reported twice: %s" % ep
        assert epochs()[-1]["loss"] < e1[0]["loss"]
        log_a = open(str(tmp_path / "logA.launcher.log")).read()
        assert log_a.count("rescaled IN PLACE") >= 2 and "falling back to stop-resume" not in log_a
        worker = open(str(tmp_path / "logA" / "workerlog.0")).read()
        assert worker.count("rescaled in place") >= 2
        etcd.close()
    finally:
        for p in (a, b):
            if p is not None and p.poll() is None:
                os.killpg(os.getpgid(p.pid), 9)


def test_rendezvous_store_and_commit_protocol(kv_server):
    """KVRendezvousStore semantics torch relies on + the commit / withdraw exclusion of the stage rendezvous."""
    from edl_b200 import elastic
    from edl_b200.store import KVClient

    kv = KVClient(kv_server.endpoint)
    st = elastic.KVRendezvousStore(kv, "/t/pg/", timeout_s=0.3)
    assert isinstance(st, dist.Store)
    st.set("a", b"1")
    assert st.get("a") == b"1" and st.check(["a"]) and not st.check(["a", "zz"])
    assert [st.add("cnt", 2), st.add("cnt", 3)] == [2, 5] and st.get("cnt") == b"5"
    assert st.compare_set("cas", b"", b"x") == b"x" and st.compare_set("cas", b"nope", b"y") == b"x"
    assert st.compare_set("cas", b"x", b"y") == b"y"
    with pytest.raises(RuntimeError):
        st.get("never")
    assert st.num_keys() == 3 and st.delete_key("a") and st.num_keys() == 2
    assert st.has_extended_api()
    st.append("log", b"ab")
    st.append("log", b"cd")
    st.multi_set(["m1", "m2"], [b"x", b"yy"])
    assert st.multi_get(["log", "m1", "m2"]) == [b"abcd", b"x", b"yy"]
    # commit needs every ready key; a withdrawal is only possible while the commit record is absent
    job, stage = "j", "S"
    keys = [elastic.ready_key(job, stage, r) for r in range(2)]
    ckey = elastic.commit_key(job, stage)
    kv.put(keys[0], b"survivor")
    commit = lambda: kv.txn([{"key": k, "target": "version", "op": ">", "value": 0} for k in keys] +  # noqa: E731
                            [{"key": ckey, "target": "version", "op": "==", "value": 0}],
                            [{"op": "put", "key": ckey, "value": b"{}"}])[0]
    withdraw = lambda k: kv.txn([{"key": ckey, "target": "version", "op": "==", "value": 0}],  # noqa: E731
                                [{"op": "delete", "key": k}])[0]
    assert not commit()                       # rank 1 has not arrived
    kv.put(keys[1], b"joiner")
    assert withdraw(keys[1]) and not commit()  # withdrew in time: the stage cannot be committed without it
    kv.put(keys[1], b"joiner")
    assert commit() and not withdraw(keys[1]) and not commit()   # committed: nobody can leave, nobody commits twice
    kv.close()


@pytest.mark.slow
def test_resnet_trainer_rescales_in_place(kv_server, tmp_path):
    """The flagship trainer path (StudentTrainer.rebuild + sync_from) under the launcher, CPU / gloo, tiny ResNet_vd:
    pod B joins mid-epoch, A's trainer stays alive, a false alarm is survived by a soft reset, both finish the job."""
    job = "inplace_rn_" + uuid.uuid4().hex[:6]
    ckpt = str(tmp_path / "ckpt")
    script = os.path.join(ROOT, "examples", "collective", "resnet50", "train.py")
    fault = str(tmp_path / "fault.now")

    def launch(name):
        env = dict(os.environ)
        env.update({"PYTHONPATH": ROOT, "CUDA_VISIBLE_DEVICES": "", "PADDLE_RUNNING_PLATFORM": "", "EDL_POD_IP": "127.0.0.1",
                    "EDL_INPLACE_CHECK_EVERY": "4", "EDL_INPLACE_ACK_TIMEOUT": "60", "OMP_NUM_THREADS": "2",
                    "RESNET_INJECT_FAULT_FILE": fault})
        cmd = [sys.executable, "-u", "-m", "edl_b200.collective.launch", "--nodes_range", "1:2", "--nproc_per_node", "1",
               "--etcd_endpoints", kv_server.endpoint, "--job_id", job, "--log_dir", str(tmp_path / ("log" + name)),
               "--hdfs_path", ckpt, "--rescale_mode", "inplace", script, "--model", "ResNet18_vd", "--width_mult", "0.125",
               "--image_size", "32", "--class_dim", "10", "--batch_size", "4", "--epochs", "2", "--steps_per_epoch", "240",
               "--solo_step_sleep", "0.4", "--ckpt", ckpt]     # alone, epoch 0 lasts >= 96 s: B joins whatever the load
        return subprocess.Popen(cmd, env=env, stdout=open(str(tmp_path / (name + ".launcher.log")), "w"),
                                stderr=subprocess.STDOUT, start_new_session=True)

    def worker_log(name):
        p = tmp_path / ("log" + name) / "workerlog.0"
        return p.read_text() if p.exists() else ""

    a = launch("A")
    b = None
    try:
        deadline = time.time() + 120
        while "trainbatch 10 " not in worker_log("A"):
            assert time.time() < deadline and a.poll() is None, worker_log("A")[-2000:]
            time.sleep(0.2)
        b = launch("B")
        deadline = time.time() + 200
        while "rescaled in place: world 1 -> 2" not in worker_log("A"):
            assert time.time() < deadline and a.poll() is None, worker_log("A")[-2000:]
            time.sleep(0.2)
        # a false alarm on top: rank 1 reports a failed collective although both pods are alive; both trainers drop the
        # group, see that nobody left and re-form the same stage (soft reset) -- StudentTrainer.rebuild with an unchanged
        # world size, state from rank 0
        open(fault, "w").close()
        assert a.wait(timeout=400) == 0, worker_log("A")[-3000:]
        assert b.wait(timeout=120) == 0, worker_log("B")[-3000:]
        la, lb = worker_log("A"), worker_log("B")
        assert "rescaled in place: world 1 -> 2" in la, la[-3000:]
        assert "rescaled in place: world 2 -> 2" in la and "rescaled in place: world 2 -> 2" in lb, (la[-2000:], lb[-2000:])
        assert "injected collective fault" in lb
        assert "Traceback" not in la and "Traceback" not in lb
        assert "falling back to stop-resume" not in (tmp_path / "A.launcher.log").read_text()
        etcd = EtcdClient([kv_server.endpoint], root=job)
        etcd.init()
        assert edl_status.load_job_status_from_etcd(etcd) == edl_status.Status.SUCCEED
        etcd.close()
        from edl_b200.checkpoint import load_check_point
        tensors, ts, state_json = load_check_point(ckpt)
        assert tensors is not None and ts.epoch_no == 1 and json.loads(state_json)["world"] == 2
    finally:
        for p in (a, b):
            if p is not None and p.poll() is None:
                os.killpg(os.getpgid(p.pid), 9)


@pytest.mark.slow
def test_hot_recovery_when_a_pod_dies_hard(kv_server, tmp_path):
    """Pod B is SIGKILLed while both pods train: A's trainer sees its collective fail, drops the broken group, waits
    for the store to publish the smaller stage and carries on ALONE IN THE SAME PROCESS (the reference -- and this
    launcher's restart mode -- kill and restart every trainer of the job)."""
    job = "hot_" + uuid.uuid4().hex[:6]
    report, ckpt = str(tmp_path / "report"), str(tmp_path / "ckpt")

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
            time.sleep(0.2)
        raise AssertionError("world never became %d: %s\n%s" % (w, epochs()[-4:],
                                                                  open(str(tmp_path / "logA.launcher.log")).read()[-3000:]))

    a = _launch(kv_server.endpoint, job, str(tmp_path / "logA"), report, ckpt, 2000)
    b = None
    try:
        e1 = wait_world(1, 60)
        pid_a = e1[-1]["pid"]
        b = _launch(kv_server.endpoint, job, str(tmp_path / "logB"), report, ckpt, 2000)
        wait_world(2, 90)
        import psutil

        victims = [psutil.Process(b.pid)] + psutil.Process(b.pid).children(recursive=True)
        for v in victims:                                     # launcher B AND its trainer (own session) die without a word
            try:
                v.kill()
            except psutil.NoSuchProcess:
                pass
        e1b = wait_world(1, 90)
        assert e1b[-1]["pid"] == pid_a, "the survivor was restarted instead of recovering in place"
        assert abs(e1b[-1]["lr"] - e1[-1]["lr"]) < 1e-9
        _finish(report)
        assert a.wait(timeout=120) == 0
        worker = open(str(tmp_path / "logA" / "workerlog.0")).read()
        assert "recovered in place: world 2 -> 1" in worker, worker[-2000:]
        ep = [x["epoch"] for x in epochs()]
        assert ep == sorted(set(ep))
    finally:
        for p in (a, b):
            if p is not None and p.poll() is None:
                os.killpg(os.getpgid(p.pid), 9)


@pytest.mark.slow
def test_false_alarm_is_survived_by_a_soft_reset(kv_server, tmp_path):
    """A collective fails on one rank although BOTH pods are alive (injected; in production: a rank stalled longer than the
    communication time-out).  Both trainers drop the group, see that the membership does not change, re-form the same
    stage under a fresh namespace (generation 1) and continue from rank 0's state -- same processes, same world size."""
    job = "soft_" + uuid.uuid4().hex[:6]
    report, ckpt = str(tmp_path / "report"), str(tmp_path / "ckpt")
    fault = str(tmp_path / "fault.now")
    os.environ["FIT_INJECT_FAULT_FILE"] = fault

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
            time.sleep(0.2)
        raise AssertionError("world never became %d: %s" % (w, epochs()[-4:]))

    a = b = None
    try:
        a = _launch(kv_server.endpoint, job, str(tmp_path / "logA"), report, ckpt, 4000)
        wait_world(1, 60)
        b = _launch(kv_server.endpoint, job, str(tmp_path / "logB"), report, ckpt, 4000)
        e2 = wait_world(2, 90)
        pid_a, lr2 = e2[-1]["pid"], e2[-1]["lr"]
        open(fault, "w").close()                              # rank 1 raises once
        deadline = time.time() + 90
        wa = wb = ""
        while time.time() < deadline:
            wa = open(str(tmp_path / "logA" / "workerlog.0")).read()
            wb = open(str(tmp_path / "logB" / "workerlog.0")).read()
            if "recovered in place: world 2 -> 2" in wa and "recovered in place: world 2 -> 2" in wb:
                break
            time.sleep(0.3)
        assert "recovered in place: world 2 -> 2" in wa, wa[-2000:]
        assert "recovered in place: world 2 -> 2" in wb, wb[-2000:]
        assert "injected collective fault" in wb
        e3 = wait_world(2, 60)                                # training goes on with both pods ...
        assert e3[-1]["pid"] == pid_a                         # ... in the same processes, LR untouched
        assert abs(e3[-1]["lr"] - lr2) < 1e-9
        _finish(report)
        assert a.wait(timeout=120) == 0
        ep = [x["epoch"] for x in epochs()]
        assert ep == sorted(set(ep))
    finally:
        os.environ.pop("FIT_INJECT_FAULT_FILE", None)
        for p in (a, b):
            if p is not None and p.poll() is None:
                os.killpg(os.getpgid(p.pid), 9)


@pytest.mark.slow
def test_sigterm_is_a_graceful_leave(kv_server, tmp_path):
    """SIGTERM to a launcher (scheduler eviction, k8s pod deletion) = announce the departure first: the pod gives up its
    registrations, the job re-plans without it, its trainers leave at the agreed step -- the survivor never sees a
    broken collective (no hot recovery needed) and keeps its process."""
    import signal

    job = "leave_" + uuid.uuid4().hex[:6]
    report, ckpt = str(tmp_path / "report"), str(tmp_path / "ckpt")

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
            time.sleep(0.2)
        raise AssertionError("world never became %d: %s" % (w, epochs()[-4:]))

    a = _launch(kv_server.endpoint, job, str(tmp_path / "logA"), report, ckpt, 2000)
    b = None
    try:
        pid_a = wait_world(1, 60)[-1]["pid"]
        b = _launch(kv_server.endpoint, job, str(tmp_path / "logB"), report, ckpt, 2000)
        wait_world(2, 90)
        b.send_signal(signal.SIGTERM)                             # the launcher only; its trainer is left alone
        e1 = wait_world(1, 60)
        assert e1[-1]["pid"] == pid_a
        _finish(report)
        assert b.wait(timeout=60) == 0
        assert a.wait(timeout=120) == 0
        worker = open(str(tmp_path / "logA" / "workerlog.0")).read()
        assert "rescaled in place: world 2 -> 1" in worker and "collective failed" not in worker, worker[-2000:]
        assert "is leaving" in open(str(tmp_path / "logB.launcher.log")).read()
    finally:
        for p in (a, b):
            if p is not None and p.poll() is None:
                os.killpg(os.getpgid(p.pid), 9)


@pytest.mark.slow
def test_elastic_reader_consumes_every_record_once_across_an_inplace_join(kv_server, tmp_path):
    """Elastic data plane + in-place rescale: pod A starts reading alone, pod B joins mid-epoch; the pods exchange their
    consumed ranges at the stage rendezvous and re-create the balanced reader -- no record is lost, none is read twice,
    and A's process survives."""
    job = "reader_" + uuid.uuid4().hex[:6]
    out = tmp_path / "out"
    out.mkdir()
    files = []
    for i, n in enumerate([150, 90, 120, 60]):
        p = tmp_path / ("f%d.txt" % i)
        p.write_text("".join("file%d-line%d\n" % (i, j) for j in range(n)))
        files.append(str(p))
    demo = os.path.join(ROOT, "tests", "elastic_reader_demo.py")

    def launch(name):
        env = dict(os.environ)
        env.update({"PYTHONPATH": ROOT, "CUDA_VISIBLE_DEVICES": "", "PADDLE_RUNNING_PLATFORM": "", "EDL_POD_IP": "127.0.0.1",
                    "READER_DEMO_OUT": str(out), "EDL_INPLACE_ACK_TIMEOUT": "60", "READER_DEMO_STEP": "0.15"})
        cmd = [sys.executable, "-u", "-m", "edl_b200.collective.launch", "--nodes_range", "1:2", "--nproc_per_node", "1",
               "--etcd_endpoints", kv_server.endpoint, "--job_id", job, "--log_dir", str(tmp_path / ("log" + name)),
               "--rescale_mode", "inplace", demo, ",".join(files)]
        return subprocess.Popen(cmd, env=env, stdout=open(str(tmp_path / (name + ".launcher.log")), "w"),
                                stderr=subprocess.STDOUT, start_new_session=True)

    def records():
        rows = []
        for f in out.glob("consumed_*.jsonl"):
            rows += [json.loads(l) for l in open(f)]
        return rows

    a = launch("A")
    b = None
    try:
        deadline = time.time() + 90
        while len(records()) < 24:
            assert time.time() < deadline and a.poll() is None, (tmp_path / "A.launcher.log").read_text()[-3000:]
            time.sleep(0.1)
        b = launch("B")
        assert a.wait(timeout=240) == 0, (tmp_path / "logA" / "workerlog.0").read_text()[-3000:]
        assert b.wait(timeout=120) == 0, (tmp_path / "logB" / "workerlog.0").read_text()[-3000:]
        rows = records()
        seen = [(r["file"], r["rec"]) for r in rows]
        want = {(i, j) for i, n in enumerate([150, 90, 120, 60]) for j in range(n)}
        assert len(seen) == len(set(seen)), "records consumed twice: %d" % (len(seen) - len(set(seen)))
        assert set(seen) == want, "records lost: %d" % len(want - set(seen))
        by_pid = {}
        for r in rows:
            by_pid.setdefault(r["pid"], set()).add(r["world"])
        assert len(by_pid) == 2 and any(w == {1, 2} for w in by_pid.values()), by_pid   # A read at world 1 AND 2: same process
        assert "reader demo rescaled in place: 1 -> 2" in (tmp_path / "logA" / "workerlog.0").read_text()
    finally:
        for p in (a, b):
            if p is not None and p.poll() is None:
                os.killpg(os.getpgid(p.pid), 9)


@pytest.mark.slow
def test_restart_mode_survives_a_hard_pod_death(kv_server, tmp_path):
    """Stop-resume mode, pod B SIGKILLed: A's trainer dies of the broken collective (gloo) BEFORE the store has noticed
    the dead pod.  The launcher treats a trainer exit that coincides with a membership change as collateral damage,
    restarts its trainers for the smaller stage and the job still succeeds (it used to be declared FAILED)."""
    import psutil

    job = "collateral_" + uuid.uuid4().hex[:6]
    report, ckpt = str(tmp_path / "report"), str(tmp_path / "ckpt")

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
            time.sleep(0.2)
        raise AssertionError("world never became %d: %s\n%s" % (w, epochs()[-4:],
                                                                  open(str(tmp_path / "logA.launcher.log")).read()[-3000:]))

    a = _launch(kv_server.endpoint, job, str(tmp_path / "logA"), report, ckpt, 2000, mode="restart")
    b = None
    try:
        pid1 = wait_world(1, 60)[-1]["pid"]
        b = _launch(kv_server.endpoint, job, str(tmp_path / "logB"), report, ckpt, 2000, mode="restart")
        wait_world(2, 90)
        for v in [psutil.Process(b.pid)] + psutil.Process(b.pid).children(recursive=True):
            try:
                v.kill()
            except psutil.NoSuchProcess:
                pass
        e = wait_world(1, 90)
        assert e[-1]["pid"] != pid1                              # restart mode: a NEW trainer process carries on
        _finish(report)
        assert a.wait(timeout=120) == 0
        log_a = open(str(tmp_path / "logA.launcher.log")).read()
        assert "treating the exit as collateral" in log_a
        etcd = EtcdClient([kv_server.endpoint], root=job)
        etcd.init()
        assert edl_status.load_job_status_from_etcd(etcd) == edl_status.Status.SUCCEED
        etcd.close()
        ep = [x["epoch"] for x in epochs()]
        assert ep == sorted(ep), ep
    finally:
        for p in (a, b):
            if p is not None and p.poll() is None:
                os.killpg(os.getpgid(p.pid), 9)


@pytest.mark.slow
def test_distill_example_rescales_in_place_with_a_live_teacher(kv_server, tmp_path):
    """The elastic DISTILL example end to end (examples/distill/resnet/train.py, the reference's
    example/distill/resnet/train_with_fleet.py under the launcher): students pull soft labels from a teacher server
    through the DistillReader, pod B joins in place while pod A trains, an injected false alarm is survived by a soft reset
    (ElasticContext.recover() with an unchanged membership), both finish the job."""
    from edl_b200.distill.teacher_server import TeacherServer
    from edl_b200.models.teacher_zoo import build

    job = "inplace_distill_" + uuid.uuid4().hex[:6]
    ckpt = str(tmp_path / "ckpt")
    script = os.path.join(ROOT, "examples", "distill", "resnet", "train.py")
    fault = str(tmp_path / "fault.now")
    model, feeds, fetches, shapes = build("resnext_tiny")
    srv = TeacherServer(model, feeds, fetches, shapes).start()

    def launch(name):
        env = dict(os.environ)
        env.update({"PYTHONPATH": ROOT, "CUDA_VISIBLE_DEVICES": "", "PADDLE_RUNNING_PLATFORM": "", "EDL_POD_IP": "127.0.0.1",
                    "EDL_INPLACE_CHECK_EVERY": "4", "EDL_INPLACE_ACK_TIMEOUT": "60", "OMP_NUM_THREADS": "2",
                    "DISTILL_INJECT_FAULT_FILE": fault})
        cmd = [sys.executable, "-u", "-m", "edl_b200.collective.launch", "--nodes_range", "1:2", "--nproc_per_node", "1",
               "--etcd_endpoints", kv_server.endpoint, "--job_id", job, "--log_dir", str(tmp_path / ("log" + name)),
               "--hdfs_path", ckpt, "--rescale_mode", "inplace", script, "--model", "ResNet18_vd", "--width_mult", "0.125",
               "--image_shape", "3,64,64", "--class_dim", "16", "--batch_size", "4", "--total_images", "960",
               "--num_epochs", "2", "--use_distill_service", "1", "--distill_teachers", srv.endpoint,
               "--teacher_batch_size", "4", "--fetch_steps", "5", "--solo_step_sleep", "0.4", "--checkpoint", ckpt]
        return subprocess.Popen(cmd, env=env, stdout=open(str(tmp_path / (name + ".launcher.log")), "w"),
                                stderr=subprocess.STDOUT, start_new_session=True)

    def worker_log(name):
        p = tmp_path / ("log" + name) / "workerlog.0"
        return p.read_text() if p.exists() else ""

    a = launch("A")
    b = None
    try:
        deadline = time.time() + 180
        while "batch 10," not in worker_log("A"):
            assert time.time() < deadline and a.poll() is None, worker_log("A")[-2000:]
            time.sleep(0.2)
        b = launch("B")
        deadline = time.time() + 300
        while "rescaled in place: world 1 -> 2" not in worker_log("A"):
            assert time.time() < deadline and a.poll() is None, worker_log("A")[-2000:]
            time.sleep(0.2)
        open(fault, "w").close()                              # rank 1 reports one failed collective; nobody left
        assert a.wait(timeout=500) == 0, worker_log("A")[-3000:]
        assert b.wait(timeout=120) == 0, worker_log("B")[-3000:]
        la, lb = worker_log("A"), worker_log("B")
        assert "rescaled in place: world 1 -> 2" in la, la[-3000:]
        assert "rescaled in place: world 2 -> 2" in la and "rescaled in place: world 2 -> 2" in lb, (la[-2000:], lb[-2000:])
        assert "Traceback" not in la and "Traceback" not in lb
        etcd = EtcdClient([kv_server.endpoint], root=job)
        etcd.init()
        assert edl_status.load_job_status_from_etcd(etcd) == edl_status.Status.SUCCEED
        etcd.close()
    finally:
        srv.stop()
        for p in (a, b):
            if p is not None and p.poll() is None:
                os.killpg(os.getpgid(p.pid), 9)
