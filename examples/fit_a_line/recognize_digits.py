#!/usr/bin/env python
"""Digit recognition with elastic data parallelism (reference: example/fit_a_line/fluid/recognize_digits.py
-- MLP or conv net on MNIST with the parameter-server transpiler; here the same nets train under the
elastic launcher with all-reduced gradients, see SURVEY 2.8 "PS -> elastic DP").

    python -m paddle_edl.collective.launch --nodes_range 1:2 --nproc_per_node 1 --etcd_endpoints 127.0.0.1:2379 \
        --job_id digits examples/fit_a_line/recognize_digits.py --nn_type conv --epochs 3
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")))
import edl_b200 as edl  # noqa: E402
from edl_b200 import ops  # noqa: E402
from edl_b200.checkpoint import LocalFS, TrainStatus, load_check_point, save_check_point  # noqa: E402
from edl_b200.parallel import ElasticDataParallel  # noqa: E402


class MLP(nn.Module):
    def __init__(self):
        super().__init__()
        self.f1, self.f2, self.out = nn.Linear(784, 200), nn.Linear(200, 200), nn.Linear(200, 10)

    def forward(self, x):
        return self.out(torch.tanh(self.f2(torch.tanh(self.f1(x.flatten(1))))))


class ConvNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.c1, self.c2 = nn.Conv2d(1, 20, 5), nn.Conv2d(20, 50, 5)
        self.out = nn.Linear(50 * 4 * 4, 10)

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.c1(x)), 2)
        x = F.max_pool2d(F.relu(self.c2(x)), 2)
        return self.out(x.flatten(1))


def synthetic_digits(n, seed):
    protos = np.random.RandomState(0).rand(10, 1, 28, 28).astype("float32")
    rng = np.random.RandomState(seed)
    y = rng.randint(0, 10, n)
    x = protos[y] + 0.3 * rng.randn(n, 1, 28, 28).astype("float32")
    return torch.from_numpy(x), torch.from_numpy(y)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nn_type", default="mlp", choices=["mlp", "conv"])
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--batch_size", type=int, default=64)
    ap.add_argument("--samples", type=int, default=2048)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--ckpt", default=os.environ.get("PADDLE_EDL_HDFS_PATH") or "./digits_ckpt")
    args = ap.parse_args()
    env = edl.init_distributed()
    world, rank = env.size, env.global_rank
    torch.manual_seed(0)
    model = MLP() if args.nn_type == "mlp" else ConvNet()
    dp = ElasticDataParallel(model)
    opt = ops.FlatAdam(dp.flat, lr=args.lr)
    fs = LocalFS()
    tensors, ts, _ = load_check_point(args.ckpt, fs, trainer_id=rank)
    if tensors is not None:
        model.load_state_dict(tensors["model"])
        dp.flat.sync_master_from_params()
        opt.load_state_dict(tensors["optim"])
    for epoch in range(ts.next(), args.epochs):
        x, y = synthetic_digits(args.samples // world, 1000 * epoch + rank)
        correct = 0
        for i in range(0, len(x) - args.batch_size + 1, args.batch_size):
            xb, yb = x[i:i + args.batch_size], y[i:i + args.batch_size]
            dp.zero_grad()
            logits = dp(xb)
            loss = F.cross_entropy(logits, yb)
            loss.backward()
            dp.finish()
            opt.step()
            correct += int((logits.argmax(-1) == yb).sum())
        if rank == 0:
            print("epoch %d loss %.4f acc %.3f world %d" % (epoch, float(loss), correct / max(1, len(x)), world), flush=True)
            save_check_point(args.ckpt, {"model": model.state_dict(), "optim": opt.state_dict()}, TrainStatus(epoch, 0), fs)


if __name__ == "__main__":
    main()
