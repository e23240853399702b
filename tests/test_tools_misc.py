"""Small host-side utilities: k8s pod discovery helpers (fake kubectl), the hadoop-CLI file system (fake hadoop),
the step-window profiler."""
import json
import os
import stat
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _script(path, body):
    path.write_text("#!/bin/bash\n" + body)
    path.chmod(path.stat().st_mode | stat.S_IEXEC)


def test_k8s_tools_with_fake_kubectl(tmp_path, monkeypatch):
    pods = {"items": [
        {"metadata": {"name": "t-1"}, "status": {"podIP": "10.0.0.2", "phase": "Running", "startTime": "2026-01-01T00:00:02Z"}},
        {"metadata": {"name": "t-0"}, "status": {"podIP": "10.0.0.1", "phase": "Running", "startTime": "2026-01-01T00:00:01Z"}},
        {"metadata": {"name": "t-2"}, "status": {"phase": "Pending"}}]}
    (tmp_path / "pods.json").write_text(json.dumps(pods))
    _script(tmp_path / "kubectl", "cat %s\n" % (tmp_path / "pods.json"))
    env = dict(os.environ, PATH="%s:%s" % (tmp_path, os.environ["PATH"]), POD_IP="10.0.0.2")
    code = ("import sys; sys.modules['kubernetes'] = None\n"          # force the kubectl fallback
            "sys.path.insert(0, %r); import k8s_tools as k\n"
            "print(k.fetch_ips('edl=1')); print(k.fetch_endpoints('edl=1', 7164)); print(k.fetch_id('edl=1'));"
            "print(k.count_pods_by_phase('edl=1', 'Pending'))" % os.path.join(ROOT, "k8s"))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["10.0.0.1,10.0.0.2", "10.0.0.1:7164,10.0.0.2:7164", "1", "1"]


def test_hdfs_client_drives_the_hadoop_cli(tmp_path, monkeypatch):
    log = tmp_path / "calls.log"
    _script(tmp_path / "hadoop", 'echo "$@" >> %s\nif [ "$4" == "-test" ] || [ "$6" == "-test" ]; then exit 1; fi\n'
                                 'if [[ "$*" == *"-ls"* ]]; then echo "drwxr-xr-x - u g 0 2026-01-01 00:00 /ckpt/__paddle_checkpoint__.3"; '
                                 'echo "-rw-r--r-- 3 u g 12 2026-01-01 00:00 /ckpt/readme"; fi\n' % log)
    monkeypatch.setenv("PATH", "%s:%s" % (tmp_path, os.environ["PATH"]))
    from edl_b200.checkpoint.fs import HDFSClient, get_fs

    fs = HDFSClient("hdfs://nn:9000", "user,pw", time_out=2000, sleep_inter=100)
    assert fs.available and fs.need_upload_download()
    assert fs.ls_dir("/ckpt") == (["__paddle_checkpoint__.3"], ["readme"])
    fs.mkdirs("/ckpt/a")
    fs.upload("/tmp/x", "/ckpt/a/x")
    fs.mv("/ckpt/a", "/ckpt/b")
    assert not fs.is_exist("/ckpt/missing")
    calls = log.read_text()
    assert "fs.default.name=hdfs://nn:9000" in calls and "hadoop.job.ugi=user,pw" in calls
    assert "-mkdir -p /ckpt/a" in calls and "-put -f /tmp/x /ckpt/a/x" in calls and "-mv /ckpt/a /ckpt/b" in calls
    assert isinstance(get_fs("hdfs://nn:9000", "user,pw"), HDFSClient)
    assert type(get_fs(None, None)).__name__ == "LocalFS"


def test_step_profiler_window(tmp_path):
    from edl_b200.utils.profiler import StepProfiler, nvtx_range

    out = str(tmp_path / "profile_pass_0")
    with StepProfiler(start=2, stop=4, out=out) as prof:
        for _ in range(6):
            with nvtx_range("step"):
                torch.randn(64, 64) @ torch.randn(64, 64)
            prof.step()
    assert os.path.exists(os.path.join(out, "kernels.txt"))
    assert "mm" in open(os.path.join(out, "kernels.txt")).read()
    off = StepProfiler(enabled=False)
    off.step()
    assert off.prof is None


def test_rescale_bench_runs_on_gloo(tmp_path):
    """tools/bench_rescale.py end to end on CPU: 3 ranks -> 2 -> 3 in place (rebuild + sync_from + LR rescale),
    then the stop-resume path through a versioned checkpoint."""
    out = tmp_path / "rescale.json"
    port = 29000 + os.getpid() % 900
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "3", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "bench_rescale.py"), "--cpu", "--model", "ResNet18_vd",
           "--width", "0.125", "--image", "32", "--classes", "10", "--batch-per-gpu", "4", "--drop", "1", "--steps", "2",
           "--ckpt-dir", str(tmp_path / "ck"), "--out", str(out)]
    p = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, CUDA_VISIBLE_DEVICES="", MASTER_PORT=str(port)),
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    res = json.loads(out.read_text())
    assert res["replicas_identical_after_grow"] is True and res["comm_error"] == 0
    for k in ("shrink_inplace_s", "grow_inplace_s", "shrink_stop_resume_s", "grow_stop_resume_s", "checkpoint_save_s"):
        assert res[k] > 0
    assert any(d.startswith("__edl_checkpoint__.") for d in os.listdir(tmp_path / "ck"))


def test_lint_gate_is_clean():
    """tools/lint.py (docstrings, line length, unused imports, bare excepts, mutable defaults) over the whole tree."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "lint.py")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout[-3000:]
