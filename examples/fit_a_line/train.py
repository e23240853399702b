#!/usr/bin/env python
"""fit_a_line: elastic data-parallel linear regression (``fc(13 -> 1)``, SGD 1e-3, batch 20) -- the
CPU / gloo plumbing config (BASELINE.json configs[0]; reference workload:
example/fit_a_line/fluid/fit_a_line.py:26-44, which runs it in parameter-server mode).

With ``EDL_RESCALE_MODE=inplace`` (launcher flag ``--rescale_mode inplace``) the trainer owns an
``ElasticContext``: on a membership change it stays alive, re-rendezvouses through the store, rebuilds the
data-parallel engine for the new world, hands its state to joiners by broadcast and rescales the LR -- no
process restart, no checkpoint reload (edl_b200/elastic.py).  Otherwise (the reference's stop-resume):

Started by the elastic launcher; every (re)start it
  1. joins the process group described by the launcher's environment,
  2. reloads the newest checkpoint (params + optimizer + epoch cursor + State JSON),
  3. applies the hyper-parameter rescale policy for the new world size (linear LR scaling),
  4. trains the remaining epochs, rank 0 checkpointing atomically at every epoch end.

    python -m paddle_edl.collective.launch --nodes_range 1:4 --nproc_per_node 1 --etcd_endpoints H:P \
        --job_id fit --hdfs_path /tmp/fit_ckpt examples/fit_a_line/train.py --epochs 20
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import edl_b200 as edl  # noqa: E402
from edl_b200 import ops  # noqa: E402
from edl_b200 import elastic  # noqa: E402
from edl_b200.checkpoint import LocalFS, TrainStatus, load_check_point, save_check_point  # noqa: E402
from edl_b200.models.small import FitALine  # noqa: E402
from edl_b200.parallel import ElasticDataParallel  # noqa: E402
from edl_b200.utils import state as edl_state  # noqa: E402


def synthetic_housing(n=512, seed=7):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 13, generator=g)
    w = torch.randn(13, 1, generator=g)
    y = x @ w + 0.5 + 0.01 * torch.randn(n, 1, generator=g)
    return x, y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--batch", type=int, default=20)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--ckpt", type=str, default=os.environ.get("PADDLE_EDL_HDFS_PATH") or "./fit_a_line_ckpt")
    ap.add_argument("--epoch_sleep", type=float, default=0.0, help="slow epochs down (elastic demos)")
    ap.add_argument("--report", type=str, default=os.environ.get("FIT_REPORT_DIR", ""))
    ap.add_argument("--inject_fault_file", type=str, default=os.environ.get("FIT_INJECT_FAULT_FILE", ""),
                    help="testing: once this file exists, rank --inject_fault_rank raises ONE collective error although "
                         "every pod is alive (a false alarm: exercises the soft reset of ElasticContext.recover())")
    ap.add_argument("--inject_fault_rank", type=int, default=int(os.environ.get("FIT_INJECT_FAULT_RANK", "1")))
    ap.add_argument("--finish_file", type=str, default=os.environ.get("FIT_FINISH_FILE", ""),
                    help="stop at the first epoch end at which this file exists (--epochs stays the upper bound)")
    args = ap.parse_args()

    inplace = elastic.inplace_requested()
    ctx = info = None
    if inplace:
        ctx = elastic.ElasticContext("gloo", check_every=int(os.environ.get("EDL_INPLACE_CHECK_EVERY", "5")))
        info = ctx.start()
        world, rank = info.size, info.rank
    else:
        env = edl.init_distributed("gloo")
        world, rank = env.size, env.global_rank
    torch.manual_seed(0)
    model = FitALine()
    dp = ElasticDataParallel(model)
    opt = ops.FlatSGDMomentum(dp.flat, lr=args.lr, momentum=0.0, weight_decay=0.0)

    # ---- elastic state: LR follows the world size ("linear scale" policy) ----
    state = edl_state.TorchState(total_batch_size=args.batch * world, model=model)
    state.register_adjust_function([edl_state.linear_scale_lr(lambda: opt.lr, opt.set_lr)])

    fs = LocalFS()

    def take_state_from(root, cursor, prev_world):
        """Collective over the (new) world: parameters, optimizer state, LR, the epoch cursor of ``root`` and the
        world size that LR belongs to.  Returns (cursor, prev_world) as ``root`` sees them."""
        if world <= 1:
            return cursor, prev_world
        dp.broadcast_parameters(root)
        for st in opt.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    dp.broadcast_tensor(v, root)
        box = [cursor, opt.lr, prev_world]
        dist.broadcast_object_list(box, src=root)
        opt.set_lr(float(box[1]))
        return box[0], int(box[2])

    if inplace and info.root is not None:
        # joining a RUNNING job: the survivors hand their state over, the checkpoint is not read
        start_epoch, prev_world = take_state_from(info.root, None, world)
        state.adjust(prev_world, world)
    else:
        tensors, train_status, state_json = load_check_point(args.ckpt, fs, trainer_id=rank)
        prev_world = world
        if tensors is not None:
            model.load_state_dict(tensors["model"])
            dp.flat.sync_master_from_params()
            opt.load_state_dict(tensors["optim"])
            if state_json:
                saved = json.loads(state_json)
                prev_world = int(saved.get("world", world))
                opt.set_lr(float(saved.get("lr", opt.lr)))
        state.adjust(prev_world, world)          # LR <- LR * world / prev_world
        start_epoch = train_status.next()

    x, y = synthetic_housing()
    n = x.shape[0]
    loss = torch.zeros(())
    epoch = start_epoch
    injected = False
    while epoch < args.epochs:
        g = torch.Generator().manual_seed(epoch)          # shuffle seed = epoch: reproducible after resume
        perm = torch.randperm(n, generator=g)
        shard = perm[rank::world]
        switch = broken = False
        stop = torch.zeros(1)
        try:
            if (inplace and not injected and world > 1 and rank == args.inject_fault_rank and args.inject_fault_file
                    and os.path.exists(args.inject_fault_file)):
                injected = True
                raise RuntimeError("injected collective fault (testing the soft reset)")
            for i in range(0, len(shard) - args.batch + 1, args.batch):
                idx = shard[i:i + args.batch]
                dp.zero_grad()
                loss = torch.nn.functional.mse_loss(dp(x[idx]), y[idx])
                loss.backward()
                dp.finish()
                opt.step()
                edl.notify_end_one_batch(None, state)
                if inplace and ctx.poll():
                    switch = True
                    break
            if inplace and not switch:
                switch = ctx.poll(force=True)              # epoch boundary: every rank asks, every rank agrees
            if not switch:
                edl.notify_end_one_epoch(state)
                if world > 1:
                    dist.all_reduce(loss)
                    loss /= world
                    dist.barrier()                         # everybody finished the epoch before it is checkpointed
                if args.finish_file:                       # every rank must take the same decision
                    stop.fill_(1.0 if os.path.exists(args.finish_file) else 0.0)
                    if world > 1:
                        dist.all_reduce(stop, op=dist.ReduceOp.MAX)
        except RuntimeError as e:                          # a peer died inside a collective (gloo raises)
            if not inplace:
                raise
            print("rank %d: collective failed (%s); recovering in place" % (rank, str(e).splitlines()[0][:120]), flush=True)
            switch = broken = True
        if switch:
            try:
                info = ctx.recover() if broken else ctx.rescale()
            except elastic.EdlEvicted:
                print("rank %d: pod left the job (scale-in); exiting" % rank, flush=True)
                ctx.close()
                return 0
            old_world, world, rank = world, info.size, info.rank
            dp.rebuild(None)                               # new process group => new gradient buffers / bucket plan
            epoch, prev_world = take_state_from(info.root, epoch, old_world)
            state.adjust(prev_world, world)                # LR follows the world size
            print("%s in place: world %d -> %d, rank %d, pid %d, %.2fs" % (
                "recovered" if broken else "rescaled", old_world, world, rank, os.getpid(), info.rendezvous_s), flush=True)
            continue                                       # redo this epoch with the new sharding
        if rank == 0:
            save_check_point(args.ckpt, {"model": model.state_dict(), "optim": opt.state_dict()},
                             TrainStatus(epoch, state.global_step_no), fs, trainer_id=0,
                             state_json=json.dumps({"world": world, "lr": opt.lr, "state": state.to_dict()}))
            print("epoch %d world %d lr %.5f loss %.5f" % (epoch, world, opt.lr, float(loss)), flush=True)
            if args.report:
                os.makedirs(args.report, exist_ok=True)
                with open(os.path.join(args.report, "epochs.jsonl"), "a") as f:
                    f.write(json.dumps({"epoch": epoch, "world": world, "lr": opt.lr, "loss": float(loss),
                                        "t": time.time(), "pid": os.getpid()}) + "\n")
        if args.epoch_sleep:
            time.sleep(args.epoch_sleep)
        epoch += 1
        if float(stop) > 0:
            break
    if ctx is not None:
        ctx.close()
        return 0
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
