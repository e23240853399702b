"""State JSON round trip through the store with the leader guard (reference test_state.py:64-99),
adjust callbacks, and atomic versioned checkpoints."""
import json
import os

import pytest
import torch

from edl_b200.checkpoint import (LocalFS, TrainStatus, latest_version, list_versions,
                                 load_check_point, save_check_point)
from edl_b200.collective import serializable
from edl_b200.utils import constants, exceptions, state as edl_state


class UserDefined(serializable.SerializableBase):
    def __init__(self):
        self.learning_rate = 1.11

    def from_json(self, s):
        self.learning_rate = json.loads(s)["learning_rate"]

    def to_json(self):
        return json.dumps({"learning_rate": self.learning_rate})


def _make_state():
    st = edl_state.State(total_batch_size=1000, user_defined=UserDefined())
    st.model_path = "model_path"
    dp = st.data_checkpoint
    dp.reader_name, dp.file_list = "reader", ["0", "1"]
    dp.processed_data = {"0": [[0, 1], [2, 3]], "1": [[4, 5], [6, 7]]}
    ts = st.train_status
    ts.epoch_no, ts.global_step_no = 1, 2
    a = edl_state.EpochAttr()
    a.epoch_no, a.world_size, a.step_num, a.avg_step_time, a.step_no_of_epoch = 1, 1, 10, 100, 5
    ts.update_epoch_attr(1, a)
    return st


def test_state_roundtrip_and_leader_guard(etcd):
    st = _make_state()
    etcd.set_server_permanent(constants.ETCD_POD_RANK, constants.ETCD_POD_LEADER, "0")
    edl_state.save_to_etcd(etcd, "0", st, timeout=5)
    ud = UserDefined()
    ud.learning_rate = 0.0
    st2 = edl_state.load_from_etcd(etcd, st.name, user_defined=ud, timeout=5)
    assert st2 == st and ud.learning_rate == 1.11
    assert st2.train_status.get_epoch_attr(1).avg_step_time == 100 and st2.train_status.next() == 2
    assert st2.data_checkpoint.is_processed(0, 3) and not st2.data_checkpoint.is_processed(1, 8)
    etcd.set_server_permanent(constants.ETCD_POD_RANK, constants.ETCD_POD_LEADER, "1")
    with pytest.raises(exceptions.EdlEtcdIOError):
        edl_state.save_to_etcd(etcd, "0", st, timeout=0.5)


def test_adjust_functions_and_bookkeeping():
    lr = {"v": 0.1}
    bs = {"v": 32}
    st = edl_state.State(total_batch_size=256)
    st.register_adjust_function([edl_state.linear_scale_lr(lambda: lr["v"], lambda v: lr.__setitem__("v", v))])
    st.adjust(8, 6)
    assert abs(lr["v"] - 0.075) < 1e-9 and st.total_batch_size == 192
    st2 = edl_state.State(total_batch_size=256)
    st2.register_adjust_function(edl_state.keep_total_batch(lambda v: bs.__setitem__("v", v)))
    st2.adjust(8, 4)
    assert bs["v"] == 64 and st2.total_batch_size == 256
    for _ in range(3):
        st.end_one_batch(world_size=6, step_time=0.1)
    assert st.global_step_no == 3 and st.step_no_of_epoch == 3
    st.data_checkpoint.mark(0, 0, 9)
    st.data_checkpoint.mark(0, 10, 19)
    assert st.data_checkpoint.processed_data["0"] == [[0, 19]]
    st.end_one_epoch()
    assert st.epoch_no == 1 and st.data_checkpoint.processed_data == {}


def test_checkpoint_versions_are_atomic(tmp_path):
    path = str(tmp_path / "ckpt")
    fs = LocalFS()
    assert load_check_point(path, fs)[0] is None and load_check_point(path, fs)[1].next() == 0
    m = torch.nn.Linear(4, 3)
    for epoch in range(4):
        with torch.no_grad():
            m.weight.fill_(float(epoch))
        v = save_check_point(path, {"model": m.state_dict()}, TrainStatus(epoch, epoch * 10), fs, trainer_id=0,
                             state_json='{"e": %d}' % epoch)
        assert v == epoch
    assert save_check_point(path, {}, TrainStatus(9), fs, trainer_id=1) == -1   # only rank 0 writes
    assert list_versions(path, fs) == [2, 3]                                   # keep=2
    # a crashed writer leaves only a temp dir: it must be ignored
    os.makedirs(os.path.join(path, "__edl_checkpoint__.4.tmp.dead"))
    os.makedirs(os.path.join(path, "__edl_checkpoint__.5"))                    # no meta.json => incomplete
    assert latest_version(path, fs) == 3
    tensors, ts, sj = load_check_point(path, fs)
    assert ts.next() == 4 and ts.global_step == 30 and sj == '{"e": 3}'
    assert float(tensors["model"]["weight"][0, 0]) == 3.0


# ---------------------------------------------------------------------------------------------------------------
# Remote file systems: a fake ``hadoop`` CLI (maps "hadoop fs -<cmd>" onto a local directory tree) stands in for
# HDFS, so HDFSClient / BDFS, the checkpoint upload / download path and the distill conf fetch run for real.
_FAKE_HADOOP = r'''#!/usr/bin/env python3
import os, shutil, sys
root = os.environ["FAKE_HDFS_ROOT"]
a = sys.argv[1:]
assert a[0] == "fs", a
a = a[1:]
while a and a[0] == "-D":
    a = a[2:]
m = lambda p: os.path.join(root, p.lstrip("/"))
cmd, rest = a[0], [x for x in a[1:] if not x.startswith("-")]
if cmd == "-ls":
    p = m(rest[0])
    if not os.path.isdir(p):
        sys.exit(1)
    for f in sorted(os.listdir(p)):
        full = os.path.join(p, f)
        print("%s   - u g 0 2020-01-01 00:00 %s" % ("drwxr-xr-x" if os.path.isdir(full) else "-rw-r--r--",
                                                    os.path.join(rest[0], f)))
elif cmd == "-test":
    sys.exit(0 if os.path.exists(m(rest[0])) else 1)
elif cmd == "-mkdir":
    os.makedirs(m(rest[0]), exist_ok=True)
elif cmd == "-rm":
    p = m(rest[0])
    shutil.rmtree(p, ignore_errors=True) if os.path.isdir(p) else (os.path.exists(p) and os.remove(p))
elif cmd == "-mv":
    os.replace(m(rest[0]), m(rest[1]))
elif cmd == "-put":
    src, dst = rest[0], m(rest[1])
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copytree(src, dst, dirs_exist_ok=True) if os.path.isdir(src) else shutil.copyfile(src, dst)
elif cmd == "-get":
    src, dst = m(rest[0]), rest[1]
    if os.path.isdir(src):
        shutil.copytree(src, os.path.join(dst, os.path.basename(src)) if os.path.isdir(dst) else dst, dirs_exist_ok=True)
    else:
        shutil.copyfile(src, os.path.join(dst, os.path.basename(src)) if os.path.isdir(dst) else dst)
else:
    sys.exit(2)
'''


@pytest.fixture
def fake_hdfs(tmp_path, monkeypatch):
    bin_dir, root = tmp_path / "bin", tmp_path / "hdfs"
    bin_dir.mkdir()
    root.mkdir()
    exe = bin_dir / "hadoop"
    exe.write_text(_FAKE_HADOOP)
    exe.chmod(0o755)
    monkeypatch.setenv("PATH", str(bin_dir) + os.pathsep + os.environ["PATH"])
    monkeypatch.setenv("FAKE_HDFS_ROOT", str(root))
    return root


def test_checkpoint_through_a_remote_file_system(fake_hdfs):
    """save / load with HDFSClient: tensors are written locally, uploaded under the temporary name and committed by
    the REMOTE rename; nothing of the checkpoint stays on the local disk (reference: fleet.save_check_point with
    BDFS, example/collective/resnet50/train_with_fleet.py:422-434)."""
    from edl_b200.checkpoint import TrainStatus, load_check_point, save_check_point
    from edl_b200.checkpoint.fs import BDFS, HDFSClient

    fs = HDFSClient("hdfs://fake", "user,pass", time_out=3000, sleep_inter=100)
    assert BDFS is HDFSClient and fs.available and fs.need_upload_download()
    sd = {"w": torch.arange(6.0).view(2, 3), "step": 7}
    assert save_check_point("/job/ckpt", sd, TrainStatus(3, 70), fs) == 0
    assert save_check_point("/job/ckpt", {"w": sd["w"] * 2, "step": 8}, TrainStatus(4, 80), fs, state_json='{"lr": 0.1}') == 1
    assert not os.path.exists("/job/ckpt")                                    # nothing leaked onto the local disk
    names = sorted(os.listdir(fake_hdfs / "job" / "ckpt"))
    assert names == ["__edl_checkpoint__.0", "__edl_checkpoint__.1"], names
    tensors, ts, sj = load_check_point("/job/ckpt", fs)
    assert torch.equal(tensors["w"], sd["w"] * 2) and tensors["step"] == 8 and ts.epoch_no == 4 and sj == '{"lr": 0.1}'
    old, ts0, _ = load_check_point("/job/ckpt", fs, version=0)
    assert torch.equal(old["w"], sd["w"]) and ts0.global_step == 70
    for _ in range(3):                                                        # keep-N on the remote side
        save_check_point("/job/ckpt", sd, TrainStatus(9), fs, keep=2)
    assert len(os.listdir(fake_hdfs / "job" / "ckpt")) == 2


def test_distill_conf_file_from_hdfs(fake_hdfs, tmp_path, monkeypatch):
    """python/edl/distill/utils.py:19-34: the teacher's serving conf is downloaded from HDFS when the
    PADDLE_DISTILL_HDFS_* variables are set."""
    from edl_b200.distill.utils import get_conf_file

    (fake_hdfs / "conf").mkdir()
    (fake_hdfs / "conf" / "serving_client_conf.prototxt").write_text('feed_var { name: "image" }\n')
    monkeypatch.delenv("PADDLE_DISTILL_CONF_FILE", raising=False)
    monkeypatch.setenv("PADDLE_DISTILL_HDFS_NAME", "hdfs://fake")
    monkeypatch.setenv("PADDLE_DISTILL_HDFS_UGI", "user,pass")
    monkeypatch.setenv("PADDLE_DISTILL_HDFS_PATH", "/conf/serving_client_conf.prototxt")
    dst = str(tmp_path / "serving_conf" / "serving_client_conf.prototxt")
    assert get_conf_file(dst) == dst and "feed_var" in open(dst).read()


def test_leader_key_is_only_removed_by_its_owner(etcd):
    """A pod that lost its lease must not delete a NEW leader's rank/0 key when it shuts down."""
    from edl_b200.utils import constants, leader_pod

    class _Gen:
        def start(self): pass
        def stop(self): pass
        def is_stopped(self): return False

    class _Env:
        etcd_endpoints, job_id = None, "j"

    reg = leader_pod.Register(_Env(), "pod-old", cluster_generator=_Gen(), ttl=5, etcd=etcd)
    assert reg.is_leader()
    # somebody else took over (the old leader's lease expired while it was partitioned)
    etcd.remove_server(constants.ETCD_POD_RANK, constants.ETCD_POD_LEADER)
    assert etcd.set_server_not_exists(constants.ETCD_POD_RANK, constants.ETCD_POD_LEADER, "pod-new", ttl=30)
    reg._is_leader = True                      # stale belief of the old leader
    reg.stop()
    assert etcd.get_value(constants.ETCD_POD_RANK, constants.ETCD_POD_LEADER) in ("pod-new", b"pod-new")
