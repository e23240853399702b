#!/usr/bin/env python
"""ResNet50_vd student <- ResNeXt101_32x16d teacher: the published-benchmark workload of the
reference (example/distill/resnet/train_with_fleet.py; flags :60-109, distill wiring :236-259,
440-465, hot loop :468-524, eval :537-575).

Modes
  (default)                 pure data-parallel training on hard labels
  --use_distill_service 1   samples flow through ``DistillReader``: teachers are found through the
                            discovery service (``--discovery H:P --service_name NAME``) or given with
                            ``--distill_teachers ip:port,...``; every sample comes back with the teacher
                            ``score`` appended and the student trains on soft-label cross-entropy
  same-box teacher GPUs     see ``bench.py --mode distill`` / ``paddle_edl.distill.device_trainer``: the
                            teacher rank writes logits straight into the student's HBM over NVSwitch

Elastic: run it through ``python -m paddle_edl.collective.launch``; every restart reloads the newest
checkpoint and rescales the LR to the new world size.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
sys.path.insert(0, ROOT)

import edl_b200 as edl  # noqa: E402
from edl_b200 import elastic, ops  # noqa: E402
from edl_b200.checkpoint import LocalFS, TrainStatus, load_check_point, save_check_point  # noqa: E402
from edl_b200.models import ResNetVd, to_train_dtype  # noqa: E402
from edl_b200.ops.optim import cosine_decay_with_warmup, piecewise_decay_with_warmup, scaled_lr  # noqa: E402
from edl_b200.parallel import DGCMomentum  # noqa: E402
from edl_b200.trainer import StudentTrainer  # noqa: E402
from edl_b200.utils.profiler import StepProfiler  # noqa: E402


def str2bool(v):
    return str(v).lower() in ("1", "true", "yes", "y")


def parse():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    a = ap.add_argument
    a("--batch_size", type=int, default=32, help="per trainer")
    a("--total_images", type=int, default=1281167)
    a("--num_epochs", type=int, default=120)
    a("--class_dim", type=int, default=1000)
    a("--image_shape", default="3,224,224")
    a("--model", default="ResNet50_vd")
    a("--lr", type=float, default=0.1)
    a("--lr_strategy", default="cosine_warmup_decay", choices=["cosine_warmup_decay", "piecewise_decay"])
    a("--l2_decay", type=float, default=1e-4)
    a("--momentum_rate", type=float, default=0.9)
    a("--fp16", type=str2bool, default=False, help="fp16 + loss scaling instead of bf16")
    a("--scale_loss", type=float, default=128.0)
    a("--use_dynamic_loss_scaling", type=str2bool, default=True)
    a("--use_label_smoothing", type=str2bool, default=False)
    a("--label_smoothing_epsilon", type=float, default=0.1)
    a("--use_mixup", type=str2bool, default=False)
    a("--mixup_alpha", type=float, default=0.2)
    a("--do_test", type=str2bool, default=False)
    a("--profile", type=str2bool, default=False)
    a("--fetch_steps", type=int, default=10)
    a("--fuse", type=str2bool, default=True, help="bucketed fused all-reduce (False: one bucket per tensor group)")
    a("--fuse_mb", type=float, default=16.0, help="bucket size (FLAGS_fuse_parameter_memory_size)")
    a("--nccl_comm_num", type=int, default=1, help="accepted for CLI parity: the P2P kernels use one comm stream")
    a("--allreduce", default="auto", choices=["auto", "oneshot", "twoshot", "multimem", "nccl"])
    a("--use_dgc", type=str2bool, default=False)
    a("--rampup_begin_step", type=int, default=5008)
    a("--use_recompute", type=str2bool, default=False)
    a("--use_distill_service", type=str2bool, default=False)
    a("--distill_teachers", default=None, help="fixed teachers ip:port,ip:port")
    a("--discovery", default=None, help="discovery / balance servers for dynamic teachers")
    a("--service_name", default="ResNeXt101_32x16d")
    a("--teacher_batch_size", type=int, default=16)
    a("--checkpoint", default=os.environ.get("PADDLE_EDL_HDFS_PATH") or "./distill_resnet_ckpt")
    a("--data_dir", default=None, help="ImageNet-style directory with train_list.txt ('path label' lines, JPEGs decoded by "
      "paddle_edl.utils.image_pipeline) or a directory of .pt shards {'images': uint8 NHWC, 'labels': int64}; synthetic if unset")
    a("--max_steps", type=int, default=0)
    a("--width_mult", type=float, default=1.0)
    a("--inject_fault_file", default=os.environ.get("DISTILL_INJECT_FAULT_FILE", ""),
      help="testing: once this file exists, rank 1 reports ONE failed collective although every pod is alive (a false "
           "alarm: exercises the soft reset of ElasticContext.recover())")
    a("--solo_step_sleep", type=float, default=0.0,
      help="elastic demos / tests: seconds to sleep per step while the job has ONE trainer, so that a second pod has "
           "time to join whatever the speed of the box")
    return ap.parse_args()


def sample_stream(args, rank, world, epoch):
    """Yields per-sample (image float32 CHW, label int64[1]) -- the 'sample list' reader format."""
    c, h, w = (int(v) for v in args.image_shape.split(","))
    if args.data_dir and os.path.exists(os.path.join(args.data_dir, "train_list.txt")):
        # JPEG files: threaded decode + random-resized crop into uint8 NHWC, normalised per sample here because the
        # DistillReader protocol is per sample (the pure-DP path below the reader ships uint8 batches instead)
        from edl_b200.utils import image_pipeline as ip
        samples = ip.read_file_list(os.path.join(args.data_dir, "train_list.txt"))
        ld = ip.ImageBatchLoader(samples, 64, size=h, train=True, rank=rank, world=world, seed=epoch, pin=False, drop_last=False)
        mean = np.array([0.485, 0.456, 0.406], dtype="float32").reshape(3, 1, 1) * 255
        std = np.array([0.229, 0.224, 0.225], dtype="float32").reshape(3, 1, 1) * 255
        for img, lab, flip in ld:
            arr = img.numpy()
            for i in range(arr.shape[0]):
                x = arr[i][:, ::-1] if int(flip[i]) else arr[i]
                yield ((x.transpose(2, 0, 1).astype("float32") - mean) / std), np.array([int(lab[i])], dtype="int64")
        return
    if args.data_dir:
        shards = sorted(f for f in os.listdir(args.data_dir) if f.endswith(".pt"))[rank::world]
        mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1) * 255
        std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1) * 255
        for f in shards:
            blob = torch.load(os.path.join(args.data_dir, f))
            for img, lab in zip(blob["images"], blob["labels"]):
                x = (img.permute(2, 0, 1).float() - mean) / std
                yield x.numpy(), np.array([int(lab)], dtype="int64")
        return
    rng = np.random.RandomState(1000 * epoch + rank)
    n = args.total_images // world
    for _ in range(n):
        yield rng.randn(c, h, w).astype("float32"), rng.randint(0, args.class_dim, (1,)).astype("int64")


def batches_of(stream, bs):
    buf = []
    for s in stream:
        buf.append(s)
        if len(buf) == bs:
            yield buf
            buf = []


def main():
    args = parse()
    ctx = None
    if elastic.inplace_requested():       # launcher --rescale_mode inplace: survive membership changes (edl_b200/elastic.py)
        ctx = elastic.ElasticContext(check_every=int(os.environ.get("EDL_INPLACE_CHECK_EVERY", "20")))
        info = ctx.start()
        world, rank = info.size, info.rank
    else:
        env = edl.init_distributed()
        world, rank = env.size, env.global_rank
    cuda = torch.cuda.is_available()
    dev = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
    c, h, w = (int(v) for v in args.image_shape.split(","))
    bs = args.batch_size
    dtype = (torch.float16 if args.fp16 else torch.bfloat16) if cuda else torch.float32
    layers = int("".join(ch for ch in args.model.split("_")[0] if ch.isdigit()) or 50)
    torch.manual_seed(0)
    model = to_train_dtype(ResNetVd(layers, args.class_dim, width_mult=args.width_mult, recompute=args.use_recompute),
                           dtype, dev).train()
    base_lr = scaled_lr(args.lr, bs, world)
    soft = args.use_distill_service or args.use_mixup
    eps = args.label_smoothing_epsilon if args.use_label_smoothing else 0.0
    opt_factory = None
    if args.use_dgc:
        def opt_factory(flat):
            return DGCMomentum(flat, lr=base_lr, momentum=args.momentum_rate, weight_decay=args.l2_decay,
                               rampup_begin_step=args.rampup_begin_step, rampup_step=max(1, args.rampup_begin_step // 5))
    tr = StudentTrainer(
        model, bs, image_shape=(c, h, w), num_classes=args.class_dim, lr=base_lr, momentum=args.momentum_rate,
        weight_decay=args.l2_decay, target_kind="probs" if soft else "labels",
        use_graph=cuda and not args.use_dgc and not args.use_recompute, dtype=dtype,
        fabric=ctx.fabric if ctx is not None else None,
        bucket_cap_mb=args.fuse_mb if args.fuse else 1e9, algo=args.allreduce,
        loss_scaling=args.scale_loss if args.fp16 else None, dynamic_loss_scaling=args.use_dynamic_loss_scaling,
        optimizer=opt_factory,
        loss_fn=None if soft else (lambda z, t: ops.soft_cross_entropy(z, t, "labels", label_smoothing=eps)))
    if args.use_dgc:
        tr.opt.dp = tr.dp
    fs = LocalFS()

    def take_cursor_from(root, cursor):
        """In-place state handoff: parameters, optimizer state and the (epoch, step) cursor of ``root``."""
        if world <= 1:
            return cursor
        tr.sync_from(root)
        box = [cursor]
        if ctx is not None:
            return tuple(ctx.broadcast_object(cursor, root))
        dist.broadcast_object_list(box, src=root)
        return box[0]

    if ctx is not None and info.root is not None:
        first_epoch, step = take_cursor_from(info.root, None)       # joined a running job: nothing read from disk
    else:
        tensors, ts, _ = load_check_point(args.checkpoint, fs, trainer_id=rank, map_location=dev)
        if tensors is not None:
            tr.load_state_dict(tensors)
        first_epoch, step = ts.next(), ts.global_step
    steps_per_epoch = max(1, args.total_images // (bs * world))
    shard = {"rank": rank, "world": world}                         # the data stream follows the current stage

    dr = None
    if args.use_distill_service:
        from edl_b200.distill.distill_reader import DistillReader
        dr = DistillReader(ins=["image", "label"], predicts=["score"])
        dr.set_teacher_batch_size(args.teacher_batch_size)
        if args.distill_teachers:
            dr.set_fixed_teacher(args.distill_teachers)
        else:
            dr.set_dynamic_teacher(args.discovery.split(","), args.service_name, require_max_teacher=4)
    epoch_box = [0]
    reader = None
    if dr is not None:
        reader = dr.set_sample_list_generator(
            lambda: batches_of(sample_stream(args, shard["rank"], shard["world"], epoch_box[0]), bs))

    onehot_eye = None
    prof = StepProfiler(100, 105, "./profile_pass_0", enabled=args.profile, rank=rank)
    epoch = first_epoch
    injected = False
    while epoch < args.num_epochs:
        epoch_box[0] = epoch
        it = reader() if reader is not None else batches_of(sample_stream(args, rank, world, epoch), bs)
        t0, seen = time.time(), 0
        switch, failed = False, ""
        for bi, batch in enumerate(it):
            if args.max_steps and bi >= args.max_steps:
                break
            if (ctx is not None and world > 1 and rank == 1 and not injected and args.inject_fault_file
                    and os.path.exists(args.inject_fault_file)):
                injected = True
                failed = "injected collective fault (testing the soft reset)"
                break
            try:
                if ctx is not None and bi > 0 and ctx.poll(agree=tr.dp.agree):
                    switch = True
                    break
            except RuntimeError as e:                           # a collective timed out: a peer is gone (or a false alarm)
                failed = str(e).splitlines()[0][:200]
                break
            lr = (cosine_decay_with_warmup(step, base_lr, steps_per_epoch, args.num_epochs)
                  if args.lr_strategy.startswith("cosine")
                  else piecewise_decay_with_warmup(step, base_lr, steps_per_epoch, [30, 60, 80]))
            tr.set_lr(lr)
            x = torch.from_numpy(np.stack([s[0] for s in batch])).to(tr.static_x.dtype)
            y = torch.from_numpy(np.concatenate([s[1] for s in batch]))
            if soft:
                if args.use_distill_service:
                    t = torch.from_numpy(np.stack([s[2] for s in batch])).float()
                else:
                    if onehot_eye is None:
                        onehot_eye = torch.eye(args.class_dim)
                    t = onehot_eye[y] * (1 - eps) + eps / args.class_dim
                if args.use_mixup:
                    lam = float(np.random.beta(args.mixup_alpha, args.mixup_alpha))
                    perm = torch.randperm(x.shape[0])
                    x = lam * x.float() + (1 - lam) * x.float()[perm]
                    t = lam * t + (1 - lam) * t[perm]
                    x = x.to(tr.static_x.dtype)
                tgt = t.to(tr.static_t.dtype)
            else:
                tgt = y
            x = x.contiguous(memory_format=torch.channels_last)
            try:
                loss = tr.step(x.pin_memory() if cuda else x, tgt.pin_memory() if cuda else tgt)
            except RuntimeError as e:
                # library collectives raise when a peer is gone; the fabric kernels time out into an error word that the
                # next poll reports
                if ctx is None:
                    raise
                failed = str(e).splitlines()[0][:200]
                break
            prof.step()
            step += 1
            seen += bs
            if args.solo_step_sleep and world == 1:
                time.sleep(args.solo_step_sleep)
            if bi % args.fetch_steps == 0 and rank == 0:
                print("Pass %d, batch %d, loss %.5f, lr %.5f, speed %.1f img/s" % (
                    epoch, bi, float(loss), lr, seen * world / max(1e-6, time.time() - t0)), flush=True)
        if ctx is not None and not switch and not failed:
            try:
                switch = ctx.poll(force=True, agree=tr.dp.agree)
            except RuntimeError as e:
                failed = str(e).splitlines()[0][:200]
        if switch or failed:
            if hasattr(it, "close"):
                it.close()                                      # stop the (distill) reader of the old shard
            try:
                if failed:
                    # hot recovery (a pod died) or soft reset (false alarm): same process, state from the root rank
                    print("rank %d: %s -- recovering in place" % (rank, failed), flush=True)
                    info = ctx.recover()
                else:
                    tr.prepare_rescale()                        # fused optimizer: sharded state made complete (old stage)
                    info = ctx.rescale()
            except elastic.EdlEvicted:
                print("rank %d: pod left the job (scale-in); exiting" % rank, flush=True)
                break
            old_world, world, rank = world, info.size, info.rank
            shard.update(rank=rank, world=world)
            tr.rebuild(None, fabric=ctx.fabric, failed=bool(failed))
            epoch, step = take_cursor_from(info.root, (epoch, step))
            base_lr = scaled_lr(args.lr, bs, world)
            steps_per_epoch = max(1, args.total_images // (bs * world))
            print("rescaled in place: world %d -> %d, rank %d, pid %d" % (old_world, world, rank, os.getpid()), flush=True)
            continue                                            # this epoch again, re-sharded for the new world
        if args.do_test:
            val = batches_of(sample_stream(args, rank, world, 10 ** 6), bs)
            ev = tr.evaluate((torch.from_numpy(np.stack([s[0] for s in b])), torch.from_numpy(np.concatenate([s[1] for s in b])))
                             for _, b in zip(range(8), val))
            if rank == 0:
                print("Pass %d test acc1 %.4f acc5 %.4f (n=%d)" % (epoch, ev["acc1"], ev["acc5"], ev["n"]), flush=True)
        tr.consolidate()                 # fused optimizer: every rank's slices of master / momentum -> complete state
        if rank == 0:
            save_check_point(args.checkpoint, tr.state_dict(), TrainStatus(epoch, step), fs, trainer_id=0,
                             state_json=json.dumps({"world": world, "lr": base_lr}))
        if ctx is not None:
            ctx.barrier()
        elif world > 1:
            dist.barrier()
        epoch += 1
    if dr is not None:
        dr.stop()
    if ctx is not None:
        ctx.close()
    elif world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
