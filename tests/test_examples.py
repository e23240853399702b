"""Smoke-run the example workloads on CPU with tiny shapes (the reference has no automated example
tests; these keep the ported workloads of SURVEY 2.10 runnable)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, timeout=240, env=None):
    e = dict(os.environ, CUDA_VISIBLE_DEVICES="", **(env or {}))
    p = subprocess.run([sys.executable] + args, cwd=ROOT, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=timeout, text=True)
    assert p.returncode == 0, p.stdout[-3000:]
    return p.stdout


def test_distill_resnet_example_pure_mixup_eval(tmp_path):
    out = run(["examples/distill/resnet/train.py", "--model", "ResNet18_vd", "--width_mult", "0.125", "--image_shape", "3,32,32",
               "--class_dim", "10", "--batch_size", "4", "--total_images", "32", "--num_epochs", "1", "--use_mixup", "1",
               "--use_label_smoothing", "1", "--do_test", "1", "--fetch_steps", "2", "--checkpoint", str(tmp_path / "ck")])
    assert "Pass 0, batch 0" in out and "test acc1" in out
    assert any(d.startswith("__paddle_checkpoint__") or d.startswith("ckpt") or d for d in os.listdir(tmp_path / "ck"))


def test_distill_resnet_example_with_service(tmp_path):
    from edl_b200.distill.teacher_server import TeacherServer
    from edl_b200.models.teacher_zoo import build

    model, feeds, fetches, shapes = build("resnext_tiny")
    srv = TeacherServer(model, feeds, fetches, shapes).start()
    try:
        out = run(["examples/distill/resnet/train.py", "--model", "ResNet18_vd", "--width_mult", "0.125", "--image_shape", "3,64,64",
                   "--class_dim", "16", "--batch_size", "4", "--total_images", "16", "--num_epochs", "1",
                   "--use_distill_service", "1", "--distill_teachers", srv.endpoint, "--teacher_batch_size", "4",
                   "--fetch_steps", "1", "--checkpoint", str(tmp_path / "ck")])
    finally:
        srv.stop()
    assert "Pass 0, batch 3" in out


def test_dgc_and_recompute_flags(tmp_path):
    out = run(["examples/distill/resnet/train.py", "--model", "ResNet18_vd", "--width_mult", "0.125", "--image_shape", "3,32,32",
               "--class_dim", "10", "--batch_size", "4", "--total_images", "24", "--num_epochs", "1", "--use_dgc", "1",
               "--rampup_begin_step", "2", "--use_recompute", "1", "--fetch_steps", "1", "--checkpoint", str(tmp_path / "ck")])
    assert "Pass 0, batch 5" in out


@pytest.mark.parametrize("model", ["ResNet50", "VGG11", "ResNet18_vd"])
def test_collective_resnet_example_models(tmp_path, model):
    out = run(["examples/collective/resnet50/train.py", "--model", model, "--width_mult", "0.125", "--image_size", "32",
               "--class_dim", "10", "--batch_size", "4", "--epochs", "1", "--steps_per_epoch", "3", "--ckpt", str(tmp_path / "ck")])
    assert "Pass 0 trainbatch 0" in out


def test_collective_resnet_example_reads_a_jpeg_file_list(tmp_path):
    """--data_dir: train_list.txt + JPEGs through ImageBatchLoader (cv2 threads on a box without a GPU; --use_dali
    switches to nvJPEG + the fused augmentation kernel on one)."""
    cv2 = pytest.importorskip("cv2")
    import numpy as np

    rng = np.random.RandomState(0)
    data = tmp_path / "data"
    data.mkdir()
    lines = []
    for i in range(12):
        cv2.imwrite(str(data / ("im%02d.jpg" % i)), rng.randint(0, 256, (40 + i, 50, 3)).astype(np.uint8))
        lines.append("im%02d.jpg %d" % (i, i % 10))
    (data / "train_list.txt").write_text("\n".join(lines) + "\n")
    out = run(["examples/collective/resnet50/train.py", "--model", "ResNet18_vd", "--width_mult", "0.125", "--image_size", "32",
               "--class_dim", "10", "--batch_size", "4", "--epochs", "2", "--data_dir", str(data), "--use_dali", "true",
               "--reader_threads", "2", "--ckpt", str(tmp_path / "ck")])
    assert "Pass 0 trainbatch 0" in out and "Pass 1 trainbatch 0" in out


def test_seqfile_roundtrip_and_ctr_dump(tmp_path):
    import io
    import struct
    sys.path.insert(0, os.path.join(ROOT, "examples", "ctr"))
    from seqfile import SequenceFileReader, SequenceFileWriter

    buf = io.BytesIO()
    w = SequenceFileWriter(buf)
    recs = [(struct.pack(">Q", i), os.urandom(40)) for i in range(300)]     # > SYNC_INTERVAL: sync escapes are exercised
    for k, v in recs:
        w.write(k, v)
    buf.seek(0)
    r = SequenceFileReader(buf)
    assert r.key_class == b"org.apache.hadoop.io.BytesWritable"
    assert list(r) == recs

    import torch
    from edl_b200.models.ctr_dnn import CtrDnn
    m = CtrDnn(sparse_feature_dim=50, embedding_size=4, num_sparse=3, hidden=(8,))
    torch.save({"model": m.state_dict()}, tmp_path / "ctr.pt")
    out = run(["examples/ctr/dumper.py", "--model_path", str(tmp_path / "ctr.pt"), "--output_dir", str(tmp_path / "dump"), "--shards", "2"])
    assert "dumped 150 rows" in out
    rows = []
    for i in range(2):
        rows += list(SequenceFileReader(open(tmp_path / "dump" / ("part-%05d" % i), "rb")))
    assert len(rows) == 150 and all(len(v) == 16 for _, v in rows)


def test_collector_tracks_job_phases():
    sys.path.insert(0, os.path.join(ROOT, "examples", "fit_a_line"))
    from collector import Collector

    state = {"pods": [{"name": "a-0", "job": "a", "phase": "Pending", "gpus": 8}]}
    c = Collector(lister=lambda: state["pods"], nodes=lambda: 16)
    assert c.run_once()["jobs"] == {"a": "PENDING:0"}
    state["pods"] = [{"name": "a-0", "job": "a", "phase": "Running", "gpus": 8}, {"name": "a-1", "job": "a", "phase": "Running", "gpus": 8}]
    r = c.run_once()
    assert r["jobs"] == {"a": "RUNNING:2"} and r["gpu_util"] == "16/16" and r["running_trainers"] == 2
    state["pods"] = []
    assert c.run_once()["jobs"] == {"a": "FINISH:0"}


@pytest.mark.parametrize("nn_type", ["mlp", "conv"])
def test_recognize_digits_example(tmp_path, nn_type):
    out = run(["examples/fit_a_line/recognize_digits.py", "--nn_type", nn_type, "--epochs", "2", "--samples", "256",
               "--ckpt", str(tmp_path / "ck")])
    assert "epoch 1 loss" in out


def test_nlp_student_and_teacher_examples(tmp_path):
    out = run(["examples/distill/nlp/train.py", "--epochs", "2", "--samples", "256", "--vocab", "400"])
    assert "epoch 1 loss" in out
    tsv = tmp_path / "train.tsv"
    tsv.write_text("text_a\tlabel\n" + "".join("good nice fine great %d\t1\nbad awful poor sad %d\t0\n" % (i, i) for i in range(40)))
    out = run(["examples/distill/nlp/train.py", "--epochs", "3", "--train_tsv", str(tsv)])
    assert float(out.strip().splitlines()[-1].split()[-1]) > 0.9          # dev acc on the separable toy corpus
    out = run(["examples/distill/nlp/fine_tune.py", "--epochs", "1", "--samples", "128", "--vocab", "400", "--save", str(tmp_path / "t.pt")])
    assert "epoch 0 loss" in out and (tmp_path / "t.pt").exists()


@pytest.mark.parametrize("model", ["ctr_dnn", "deepfm"])
def test_ctr_example_trains(model):
    out = run(["examples/ctr/train.py", "--model", model, "--steps", "12", "--batch", "64", "--vocab", "101"])
    assert "step 10 loss" in out and "examples_per_s" in out


def test_deepfm_second_order_term_matches_pairwise_sum():
    """The (sum^2 - sum of squares)/2 identity against the explicit sum over slot pairs."""
    import torch

    from edl_b200.models.ctr_dnn import DeepFM

    torch.manual_seed(0)
    m = DeepFM(sparse_feature_dim=31, embedding_size=4, num_sparse=5, num_dense=3, hidden=(8,))
    for t in m.first_order:
        torch.nn.init.normal_(t.weight)
    dense = torch.rand(6, 3)
    ids = torch.randint(0, 31, (6, 5, 1))
    logit = m(dense, ids)
    assert logit.shape == (6, 2) and torch.all(logit[:, 0] == 0)
    embs = m.slot_embeddings(ids)
    pair = sum((embs[i] * embs[j]).sum(1) for i in range(5) for j in range(i + 1, 5))
    first = m.dense_first(dense)[:, 0] + sum(m.first_order[s](ids[:, s])[:, 0] for s in range(5))
    x = torch.cat(embs + [dense], 1)
    for fc in m.fcs:
        x = torch.relu(fc(x))
    want = first + pair + m.out(x)[:, 0]
    assert torch.allclose(logit[:, 1], want, atol=1e-5)
    logit[:, 1].sum().backward()
    assert all(t.weight.grad is not None for t in m.tables)
