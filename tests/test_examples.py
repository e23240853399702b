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
