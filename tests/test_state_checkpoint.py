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
