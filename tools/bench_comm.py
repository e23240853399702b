#!/usr/bin/env python
"""Communication numbers of the data-parallel step (BASELINE.json: "exposed comm ms/step", all-reduce roofline).

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_comm.py --sweep --exposed

``--sweep``   device-timed (CUDA events, max over ranks) all-reduce of 64 KB ... 256 MB messages, bf16 and fp32,
              for our kernels (P2P two-shot, NVLS multimem, one-shot for small messages) next to NCCL on the same
              box.  Reported: ms, algorithm bandwidth (bytes / t), bus bandwidth (2 (W-1)/W bytes / t) and the
              fraction of the NVLink 5 per-direction roofline (900 GB/s per GPU; NVLS moves bytes/W * (W-1) less
              through the links, its ceiling is therefore higher than the P2P one -- both are printed).
``--exposed`` the ResNet50_vd step at world W with the fused bucket all-reduces enabled (overlapped with
              backward) minus the same captured step with communication disabled (gradients stay local):
              exposed comm ms/step.  The all-reduce-only time of the same buckets, back to back on an idle GPU,
              is printed next to it: hidden = isolated - exposed.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

NVLINK_DIR_GBPS = 900.0     # NVLink 5, one direction, per GPU (B200_PROFILING.md)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--exposed", action="store_true")
    ap.add_argument("--max-mb", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch-per-gpu", type=int, default=32)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--blocks", type=int, nargs="*", default=[16, 32, 64])
    ap.add_argument("--out", default="")
    return ap.parse_args()


def timed(fn, iters, dev, group=None):
    """ms per call: 3 warm-up calls, then ``iters`` calls between two events; max over ranks."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    dist.barrier(group)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sweep(args, dev, world, rank):
    from edl_b200.ops import native
    from edl_b200.parallel.symm import SymmetricPool

    C = native()
    pool = SymmetricPool((args.max_mb + 8) << 20, device=dev)
    rows = []
    sizes = [64 << 10, 256 << 10, 1 << 20, 4 << 20, 16 << 20, 51 << 20, 128 << 20, 256 << 20]
    sizes = [s for s in sizes if s <= (args.max_mb << 20)]
    for dtype in (torch.bfloat16, torch.float32):
        esz = 2 if dtype == torch.bfloat16 else 4
        for nbytes in sizes:
            n = nbytes // esz // 8 * 8
            pool._off = pool._sig_total                 # bump allocator: reuse the payload area per size
            sl = pool.alloc(n, dtype)
            sl.tensor.normal_()
            out = torch.empty(n, dtype=dtype, device=dev)
            nccl_buf = torch.randn(n, device=dev).to(dtype)
            cands = [("nccl", 0, lambda: dist.all_reduce(nccl_buf))]
            for nb in args.blocks:
                cands.append(("twoshot", nb, lambda nb=nb: C.allreduce_twoshot(
                    sl.data_ptrs, sl.sig_ptrs, 0, rank, sl.tensor, n, 1.0 / world, None, None, False, nb, 30.0)))
                if pool.has_multicast:
                    cands.append(("multimem", nb, lambda nb=nb: C.allreduce_twoshot(
                        sl.data_ptrs, sl.sig_ptrs, sl.mc_ptr, rank, sl.tensor, n, 1.0 / world, None, None, True, nb, 30.0)))
            if nbytes <= (4 << 20):
                cands.append(("oneshot", 16, lambda: C.allreduce_oneshot(
                    sl.data_ptrs, sl.sig_ptrs, rank, out, n, 1.0 / world, None, None, 16, 30.0)))
            for name, nb, fn in cands:
                ms = timed(fn, args.iters, dev)
                sl.tensor.normal_()                      # repeated averaging underflows nothing, but keep values sane
                algbw = nbytes / (ms * 1e-3) / 1e9
                busbw = algbw * 2 * (world - 1) / world
                row = {"dtype": str(dtype).split(".")[-1], "bytes": nbytes, "algo": name, "blocks": nb, "ms": ms,
                       "algbw_GBps": algbw, "busbw_GBps": busbw,
                       # P2P two-shot sends and receives (W-1)/W of the buffer twice per direction: busbw/2 per direction
                       "frac_nvlink_dir": busbw / 2 / NVLINK_DIR_GBPS}
                rows.append(row)
                if rank == 0:
                    print("%-8s %9d B  %-8s blocks %2d  %8.3f ms  algbw %7.1f  busbw %7.1f GB/s  (%.0f%% of %d GB/s/dir)" % (
                        row["dtype"], nbytes, name, nb, ms, algbw, busbw, 100 * row["frac_nvlink_dir"], NVLINK_DIR_GBPS),
                        flush=True)
    err = pool.check_error()
    return {"rows": rows, "multicast": pool.has_multicast, "comm_error": err}


def exposed(args, dev, world, rank):
    from edl_b200 import ops
    from edl_b200.models import ResNet50_vd, to_train_dtype
    from edl_b200.trainer import StudentTrainer

    B = args.batch_per_gpu
    torch.manual_seed(0)
    model = to_train_dtype(ResNet50_vd(), torch.bfloat16, dev).train()
    tr = StudentTrainer(model, B, lr=0.1 * B * world / 256.0, use_graph=True)
    x = torch.randn(B, 3, 224, 224).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).pin_memory()
    t = torch.softmax(torch.randn(B, 1000), -1).to(torch.bfloat16).pin_memory()
    tr.step(x, t)
    res = {"world": world, "batch_per_gpu": B, "buckets": [
        {"bytes": b.numel * (2 if b.dtype == torch.bfloat16 else 4), "algo": b.algo} for b in tr.dp.buckets]}
    res["ms_per_step_with_comm"] = timed(tr.step_device, args.steps, dev)
    # the same step with the reductions switched off: re-capture (the graph bakes the launches in)
    res["fused_optimizer"] = bool(tr.dp.bucket_opt)
    res["launched"] = [list(a) for a in tr.dp.last_algos]
    tr.dp.consolidate_optimizer_state()
    tr.dp.enabled = False
    tr.graph = None
    tr.step(x, t)
    res["ms_per_step_no_comm"] = timed(tr.step_device, args.steps, dev)
    res["exposed_comm_ms_per_step"] = res["ms_per_step_with_comm"] - res["ms_per_step_no_comm"]
    # the bucket all-reduces alone, back to back on an otherwise idle GPU
    tr.dp.enabled = True
    C = ops.native()

    def all_buckets():
        for b in tr.dp.buckets:
            g = tr.dp.flat.groups[b.dtype]
            sl = tr.dp.slices[b.dtype]
            off = b.start * g.grad.element_size()
            C.allreduce_twoshot([p + off for p in sl.data_ptrs], sl.sig_ptrs, (sl.mc_ptr + off) if sl.mc_ptr else 0,
                                rank, g.grad, b.numel, 1.0 / world, None, None, b.algo == "multimem",
                                tr.dp.comm_blocks, 30.0)

    res["isolated_allreduce_ms_per_step"] = timed(all_buckets, args.steps, dev)
    res["hidden_comm_ms_per_step"] = res["isolated_allreduce_ms_per_step"] - max(res["exposed_comm_ms_per_step"], 0.0)
    grad_bytes = sum(b["bytes"] for b in res["buckets"])
    res["grad_bytes_per_step"] = grad_bytes
    res["isolated_busbw_GBps"] = grad_bytes * 2 * (world - 1) / world / (res["isolated_allreduce_ms_per_step"] * 1e-3) / 1e9
    res["comm_error"] = tr.dp.check_comm_error()
    if rank == 0:
        print(json.dumps(res), flush=True)
    return res


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    assert torch.cuda.is_available() and world > 1, "needs >= 2 GPUs (torchrun)"
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    if args.sweep:
        out["sweep"] = sweep(args, dev, world, rank)
    if args.exposed:
        out["exposed"] = exposed(args, dev, world, rank)
    if rank == 0 and args.out:
        with open(args.out, "w") as fh:
            json.dump(out, fh, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
