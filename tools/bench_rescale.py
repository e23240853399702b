#!/usr/bin/env python
"""Rescale recovery time after -1 / +1 pod (BASELINE.json metric "rescale recovery time after +-1 pod";
config "ResNet50_vd elastic 8->6->8 pods mid-run with checkpoint reload and LR rescale").

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_rescale.py --drop 2
    torchrun --nproc-per-node 2 tools/bench_rescale.py --cpu --model ResNet18_vd --width 0.125 --image 32 --drop 1

Two recovery paths are timed, both from "the new membership is known" to "the first optimizer step at the new
world size has completed on every surviving rank" (host wall clock, max over ranks -- recovery IS host work):

* in-place  : ``StudentTrainer.rebuild(new_group)`` -- new NVSwitch-symmetric gradient slab + bucket plan,
              parameters / fp32 masters / momentum stay in HBM, the step graph is re-captured; on the way
              back up the joiners take parameters and optimizer state from rank 0 over NVLink
              (``sync_from``), and the learning rate is rescaled linearly with the global batch
              (reference: ``lr * batch * num_trainers / 256``, example/collective/resnet50/train_with_fleet.py:129-141);
* stop-resume: what the reference does (utils/launcher.py:221-244) minus process start-up: rank 0 writes a
              versioned checkpoint, every member of the new world builds a fresh trainer and loads it.

The reference's own floor for this path is set by its polling constants (15 s lease TTL, 3 s
polls, trainer restart + NCCL bootstrap + checkpoint reload).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="ResNet50_vd")
    ap.add_argument("--width", type=float, default=1.0)
    ap.add_argument("--image", type=int, default=224)
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--batch-per-gpu", type=int, default=32)
    ap.add_argument("--drop", type=int, default=2, help="ranks that leave (the highest ones) and later re-join")
    ap.add_argument("--steps", type=int, default=20, help="steady-state steps timed at each world size")
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--ckpt-dir", default="")
    ap.add_argument("--out", default="")
    return ap.parse_args()


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    cuda = torch.cuda.is_available() and not args.cpu
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    if cuda:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    assert 0 < args.drop < world, "--drop must leave at least one rank"

    from edl_b200 import models
    from edl_b200.checkpoint import LocalFS, TrainStatus, load_check_point, save_check_point
    from edl_b200.ops.optim import scaled_lr
    from edl_b200.trainer import StudentTrainer

    B = args.batch_per_gpu
    dtype = torch.bfloat16 if cuda else torch.float32
    shape = (3, args.image, args.image)

    def new_model():
        torch.manual_seed(0)
        m = getattr(models, args.model)(class_dim=args.classes, width_mult=args.width)
        return models.to_train_dtype(m, dtype, dev).train()

    def new_trainer(group, n):
        return StudentTrainer(new_model(), B, image_shape=shape, num_classes=args.classes,
                              lr=scaled_lr(0.1, B, n), group=group, use_graph=cuda, dtype=dtype)

    def sync(group=None):
        if cuda:
            torch.cuda.synchronize(dev)
        dist.barrier(group)

    def wall_max(t0, group=None):
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item())

    torch.manual_seed(100 + rank)
    x = torch.randn(B, *shape).to(dtype).contiguous(memory_format=torch.channels_last)
    t = torch.softmax(torch.randn(B, args.classes), -1).to(dtype)
    if cuda:
        x, t = x.pin_memory(), t.pin_memory()

    def steady(tr, group, n_steps):
        for _ in range(3):
            tr.step(x, t)
        sync(group)
        t0 = time.perf_counter()
        for _ in range(n_steps):
            tr.step(x, t)
        if cuda:
            torch.cuda.synchronize(dev)
        return wall_max(t0, group) / n_steps * 1e3

    survivors = list(range(world - args.drop))
    small = dist.new_group(ranks=survivors)                       # collective over the full world
    solo = [dist.new_group(ranks=[r]) for r in range(world)][rank]
    alive = rank in survivors
    res = {"world": world, "drop": args.drop, "model": args.model, "batch_per_gpu": B,
           "device": "cuda" if cuda else "cpu"}

    tr = new_trainer(None, world)
    res["ms_per_step_full"] = steady(tr, None, args.steps)

    # ---- shrink in place: world -> world - drop ------------------------------------------------
    sync()
    t0 = time.perf_counter()
    tr.prepare_rescale()        # fused optimizer: the sharded master / momentum slices are made complete (old stage)
    res["consolidate_s"] = wall_max(t0)
    tr.rebuild(small if alive else solo)
    if alive:
        tr.set_lr(scaled_lr(0.1, B, len(survivors)))
        tr.step(x, t)
        if cuda:
            torch.cuda.synchronize(dev)
        res["shrink_inplace_s"] = wall_max(t0, small)
        res["ms_per_step_small"] = steady(tr, small, args.steps)
    sync()

    # ---- grow in place: world - drop -> world (joiners sync from rank 0 over the fabric) --------
    t0 = time.perf_counter()
    if alive:
        tr.prepare_rescale()
    tr.rebuild(None)
    tr.sync_from(0)
    tr.set_lr(scaled_lr(0.1, B, world))
    tr.step(x, t)
    if cuda:
        torch.cuda.synchronize(dev)
    res["grow_inplace_s"] = wall_max(t0)
    res["ms_per_step_full_again"] = steady(tr, None, args.steps)
    flat = torch.cat([g.param.flatten().float() for g in tr.dp.flat.groups.values()])
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    res["replicas_identical_after_grow"] = bool(torch.equal(flat, ref))
    res["comm_error"] = tr.dp.check_comm_error()

    # ---- stop-resume (the reference's path, process start-up excluded) --------------------------
    ckpt_dir = args.ckpt_dir or os.path.join(tempfile.gettempdir(), "edl_rescale_ckpt_%s" % os.environ.get("MASTER_PORT"))
    fs = LocalFS()
    sync()
    t0 = time.perf_counter()
    tr.consolidate()
    if rank == 0:
        save_check_point(ckpt_dir, tr.state_dict(), TrainStatus(0), fs)
    dist.barrier()
    res["checkpoint_save_s"] = wall_max(t0)
    del tr
    sync()
    t0 = time.perf_counter()
    if alive:
        tr2 = new_trainer(small, len(survivors))
        sd, _, _ = load_check_point(ckpt_dir, fs, map_location=dev)
        tr2.load_state_dict(sd)
        tr2.step(x, t)
        if cuda:
            torch.cuda.synchronize(dev)
        res["shrink_stop_resume_s"] = wall_max(t0, small)
        del tr2
    sync()
    t0 = time.perf_counter()
    tr3 = new_trainer(None, world)
    sd, _, _ = load_check_point(ckpt_dir, fs, map_location=dev)
    tr3.load_state_dict(sd)
    tr3.step(x, t)
    if cuda:
        torch.cuda.synchronize(dev)
    res["grow_stop_resume_s"] = wall_max(t0)
    if rank == 0:
        print(json.dumps(res))
        if args.out:
            with open(args.out, "w") as fh:
                json.dump(res, fh, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
