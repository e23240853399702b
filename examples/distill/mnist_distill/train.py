#!/usr/bin/env python
"""MNIST distillation: small CNN student, served CNN teacher, soft-label loss (reference:
example/distill/mnist_distill/train_with_fleet.py:134-145 -- fetch name ``fc_0.tmp_2``).

    python -m paddle_edl.distill.teacher_server --model mnist_cnn --port 9292 &
    python examples/distill/mnist_distill/train.py --use_distill_service 1 --distill_teachers 127.0.0.1:9292
(with no teacher flags an in-process teacher is started; data is synthetic digits).
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from paddle_edl import ops  # noqa: E402
from paddle_edl.distill.distill_reader import DistillReader  # noqa: E402
from paddle_edl.distill.teacher_server import TeacherServer  # noqa: E402
from paddle_edl.models.small import MnistStudent  # noqa: E402
from paddle_edl.models.teacher_zoo import build  # noqa: E402


def digits(n, seed):
    rng = np.random.RandomState(seed)
    protos = np.random.RandomState(0).rand(10, 1, 28, 28).astype("float32")
    for _ in range(n):
        y = rng.randint(0, 10)
        yield (protos[y] + 0.3 * rng.randn(1, 28, 28)).astype("float32"), np.array([y], dtype="int64")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--use_distill_service", type=int, default=1)
    ap.add_argument("--distill_teachers", default="")
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--samples", type=int, default=1024)
    args = ap.parse_args()
    srv = None
    if args.use_distill_service and not args.distill_teachers:
        model, feeds, fetches, shapes = build("mnist_cnn")
        srv = TeacherServer(model, feeds, fetches, shapes).start()
        args.distill_teachers = srv.endpoint
    student = MnistStudent()
    opt = torch.optim.Adam(student.parameters(), 1e-3)

    def batches(epoch):
        buf = []
        for s in digits(args.samples, epoch):
            buf.append(s)
            if len(buf) == args.batch:
                yield buf
                buf = []

    reader = None
    if args.use_distill_service:
        dr = DistillReader(ins=["img", "label"], predicts=["fc_0.tmp_2"])
        dr.set_teacher_batch_size(16)
        dr.set_fixed_teacher(args.distill_teachers)
    for epoch in range(args.epochs):
        if args.use_distill_service:
            if reader is None:
                reader = dr.set_sample_list_generator(lambda: batches(epoch))
            it = reader()
        else:
            it = batches(epoch)
        for i, batch in enumerate(it):
            x = torch.from_numpy(np.stack([s[0] for s in batch]))
            y = torch.from_numpy(np.concatenate([s[1] for s in batch]))
            logits = student(x)
            if args.use_distill_service:
                soft = torch.from_numpy(np.stack([s[2] for s in batch]))
                loss = ops.soft_cross_entropy(logits, soft, "probs")
            else:
                loss = ops.soft_cross_entropy(logits, y, "labels")
            opt.zero_grad()
            loss.backward()
            opt.step()
            if i % 5 == 0:
                acc = (logits.argmax(-1) == y).float().mean().item()
                print("epoch %d batch %d loss %.4f acc %.3f" % (epoch, i, float(loss), acc), flush=True)
    if args.use_distill_service:
        dr.stop()
    if srv:
        srv.stop()


if __name__ == "__main__":
    main()
