"""DistillReader pipeline: ordering, batch reassembly for the 3 reader kinds, teacher failure /
retire / join mid-stream, real gRPC teacher (reference: distill_reader_test.py:21-52 with a NOP
teacher; here additionally a live TeacherServer on CPU)."""
import time

import numpy as np
import torch

from edl_b200.distill import distill_worker
from edl_b200.distill.distill_reader import DistillReader
from edl_b200.distill.predict_client import PredictClient
from edl_b200.distill.teacher_server import TeacherServer


def _samples(n):
    for i in range(n):
        yield (np.full((2, 3), i, dtype=np.float32), np.array([i], dtype=np.int64))


def _batches(n, bs):
    buf = []
    for s in _samples(n):
        buf.append(s)
        if len(buf) == bs:
            yield buf
            buf = []
    if buf:
        yield buf


class _EchoClient(PredictClient):
    """score = mean of the image slot (lets the test check sample <-> prediction pairing)."""
    fail_servers = set()
    calls = {}

    def connect(self):
        if self.server in self.fail_servers:
            raise RuntimeError("down")
        self.teacher_feeds = ["image"]
        return True

    def predict(self, feed_batch):
        if self.server in self.fail_servers:
            raise RuntimeError("teacher died")
        _EchoClient.calls[self.server] = _EchoClient.calls.get(self.server, 0) + 1
        time.sleep(0.002)
        return [{"score": np.array([f["image"].mean()], dtype=np.float32)} for f in feed_batch]


def _make(teachers="t1:1,t2:2", tbs=4):
    dr = DistillReader(ins=["image", "label"], predicts=["score"])
    dr.set_teacher_batch_size(tbs)
    dr.set_fixed_teacher(teachers)
    dr.set_predict_client_factory(lambda s, f, o, c: _EchoClient(s, f, o, c))
    return dr


def test_sample_list_order_and_boundaries_over_epochs():
    _EchoClient.fail_servers = set()
    n, bs = 24 * 8 + 2, 8
    dr = _make()
    reader = dr.set_sample_list_generator(lambda: _batches(n, bs))
    for epoch in range(5):
        seen = 0
        for batch in reader():
            assert len(batch) == (bs if seen + bs <= n else n - seen)
            for img, label, score in batch:
                assert int(label[0]) == seen and float(score[0]) == float(seen) and img.shape == (2, 3)
                seen += 1
        assert seen == n
    dr.stop()


def test_sample_and_batch_generators():
    _EchoClient.fail_servers = set()
    dr = _make(tbs=3)
    r = dr.set_sample_generator(lambda: _samples(10))
    out = list(r())
    assert [int(s[1][0]) for s in out] == list(range(10)) and all(float(s[2][0]) == i for i, s in enumerate(out))
    dr.stop()

    def batch_gen():
        for b in _batches(10, 4):
            yield (np.stack([s[0] for s in b]), np.stack([s[1] for s in b]))
    dr2 = _make(tbs=3)
    r2 = dr2.set_batch_generator(batch_gen)
    outs = list(r2())
    assert [o[0].shape[0] for o in outs] == [4, 4, 2]
    assert np.allclose(np.concatenate([o[2][:, 0] for o in outs]), np.arange(10))
    dr2.stop()


def test_teacher_failure_requeues_and_consumer_can_break_early():
    _EchoClient.fail_servers = set()
    _EchoClient.calls = {}
    dr = _make("a:1,b:2,c:3")
    reader = dr.set_sample_list_generator(lambda: _batches(400, 8))
    seen = 0
    for batch in reader():
        for _, label, score in batch:
            assert int(label[0]) == seen == int(score[0])
            seen += 1
        if seen == 80:
            _EchoClient.fail_servers = {"b:2"}   # one teacher dies mid-epoch
    assert seen == 400
    # early break then a fresh full epoch
    for i, batch in enumerate(reader()):
        if i == 3:
            break
    _EchoClient.fail_servers = set()
    assert sum(len(b) for b in reader()) == 400
    dr.stop()


def test_nop_teacher_flag():
    distill_worker._NOP_PREDICT_TEST = True
    try:
        dr = DistillReader(ins=["image", None], predicts=["score"])
        dr.set_teacher_batch_size(4)
        dr.set_fixed_teacher("127.0.0.1:1")
        r = dr.set_sample_list_generator(lambda: _batches(50, 8))
        assert sum(len(b) for b in r()) == 50
        dr.stop()
    finally:
        distill_worker._NOP_PREDICT_TEST = False


def test_real_grpc_teacher():
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(6, 5), torch.nn.Softmax(-1))
    with TeacherServer(model, ["image"], ["score"], {"image": [2, 3]}, max_wait_ms=1.0) as srv:
        dr = DistillReader(ins=["image", "label"], predicts=["score"])
        dr.set_teacher_batch_size(4)
        dr.set_fixed_teacher([srv.endpoint])
        r = dr.set_sample_list_generator(lambda: _batches(30, 6))
        n = 0
        for batch in r():
            for img, label, score in batch:
                ref = model(torch.from_numpy(img)[None]).detach().numpy()[0]
                assert np.allclose(score, ref, atol=1e-5)
                n += 1
        assert n == 30 and srv.served == 30
        dr.stop()


def test_reader_in_a_forked_process_gives_the_same_stream():
    """set_reader_process(): the user's generator runs in a forked child (the reference forks its reader worker), tasks
    cross a bounded pipe, ordering / batch boundaries / early break behave exactly like the thread version, a reader
    exception surfaces in the consumer."""
    import os

    _EchoClient.fail_servers = set()
    n, bs = 50, 8
    pids = []

    def gen():
        for b in _batches(n, bs):
            yield [(img, np.array([int(lab[0]), os.getpid()], dtype=np.int64)) for img, lab in b]

    dr = _make().set_reader_process(True)
    reader = dr.set_sample_list_generator(gen)
    for epoch in range(3):
        seen = 0
        for batch in reader():
            assert len(batch) == (bs if seen + bs <= n else n - seen)
            for img, label, score in batch:
                assert int(label[0]) == seen and float(score[0]) == float(seen)
                pids.append(int(label[1]))
                seen += 1
        assert seen == n
    assert os.getpid() not in set(pids)                    # the generator really ran elsewhere
    it = reader()                                          # early break: the child is stopped, the next epoch is complete
    next(it)
    it.close()
    assert sum(len(b) for b in reader()) == n
    dr.stop()

    def bad():
        yield from _batches(8, 4)
        raise ValueError("boom in the reader")

    dr2 = _make().set_reader_process(True)
    r2 = dr2.set_sample_list_generator(bad)
    try:
        for _ in r2():
            pass
        raise AssertionError("the reader's exception was swallowed")
    except RuntimeError as e:
        assert "boom in the reader" in str(e)
    dr2.stop()
