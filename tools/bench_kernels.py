#!/usr/bin/env python
"""Per-kernel microbenchmarks on one B200: every distinct ResNet50_vd 1x1-conv GEMM (fwd / dgrad /
wgrad) and the BN / pool / optimizer / loss kernels, timed with CUDA events (warm-up, L2 flushed
between iterations) and reported as achieved bytes/s and FLOP/s against MEASURED_PEAKS.json.

    python tools/bench_kernels.py [--out profiles/kernels_rNN.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edl_b200 import ops  # noqa: E402

DEV = "cuda"


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return d["hbm_gbs"] * 1e9, d["bf16_tflops"] * 1e12, "measured"
    except Exception:
        return 6.65e12, 1.59e15, "fallback"


_flush = None


def timeit(fn, iters=20, warmup=3):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        _flush.zero_()  # > L2 (126 MB): evict the working set
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernels.json"))
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    bw, fl, src = peaks()
    B = args.batch
    rows = []

    def rec(name, t, bytes_, flops=0.0, ref_t=None):
        r = {"kernel": name, "us": t * 1e6, "GBps": bytes_ / t / 1e9, "hbm_frac": bytes_ / t / bw,
             "TFLOPs": flops / t / 1e12, "tc_frac": flops / t / fl if flops else 0.0}
        if ref_t is not None:
            r["lib_us"] = ref_t * 1e6
            r["speedup_vs_lib"] = ref_t / t
        rows.append(r)
        print("%-46s %8.1f us  %7.0f GB/s (%.2f of %s HBM)  %6.1f TF/s%s" % (
            name, r["us"], r["GBps"], r["hbm_frac"], src, r["TFLOPs"],
            ("  lib %.1f us (x%.2f)" % (r["lib_us"], r["speedup_vs_lib"])) if ref_t else ""))

    # ---- 1x1 conv GEMMs of ResNet50_vd, per-GPU batch B
    shapes = [(56, 64, 64), (56, 64, 256), (56, 256, 64), (56, 256, 128), (28, 128, 512), (28, 256, 512),
              (28, 512, 128), (28, 512, 256), (14, 256, 1024), (14, 512, 1024), (14, 1024, 256),
              (14, 1024, 512), (7, 512, 2048), (7, 1024, 2048), (7, 2048, 512)]
    for hw, cin, cout in shapes:
        M = B * hw * hw
        x = torch.randn(M, cin, device=DEV).bfloat16()
        w = (torch.randn(cout, cin, device=DEV) * 0.05).bfloat16()
        dy = torch.randn(M, cout, device=DEV).bfloat16()
        y = torch.empty(M, cout, device=DEV, dtype=torch.bfloat16)
        dx = torch.empty(M, cin, device=DEV, dtype=torch.bfloat16)
        stats = torch.zeros(2 * cout, device=DEV)
        flops = 2.0 * M * cin * cout
        t = timeit(lambda: ops.gemm_bf16(x, w, out=y, col_stats=stats))
        tl = timeit(lambda: torch.matmul(x, w.t(), out=y))
        rec("fwd+stats  M=%d K=%d N=%d" % (M, cin, cout), t, 2 * (M * cin + cout * cin + M * cout), flops, tl)
        t = timeit(lambda: ops.gemm_bf16(dy, w, out=dx, b_mn_major=True))
        tl = timeit(lambda: torch.matmul(dy, w, out=dx))
        rec("dgrad      M=%d K=%d N=%d" % (M, cout, cin), t, 2 * (M * cout + cout * cin + M * cin), flops, tl)
        sink = torch.zeros(cout, cin, device=DEV, dtype=torch.bfloat16)
        from edl_b200.ops.gemm import _wgrad
        t = timeit(lambda: _wgrad(dy, x, (cout, cin), sink, None))
        dwl = torch.empty(cout, cin, device=DEV, dtype=torch.bfloat16)
        tl = timeit(lambda: torch.matmul(dy.t(), x, out=dwl))
        rec("wgrad      M=%d K=%d N=%d" % (cout, M, cin), t, 2 * (M * cout + M * cin + 2 * cout * cin), flops, tl)

    # ---- BN kernels on the big activations
    for hw, c in [(112, 64), (56, 256), (56, 64), (28, 512), (14, 1024), (7, 2048)]:
        x = torch.randn(B, c, hw, hw, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
        n = x.numel()
        g = torch.ones(c, device=DEV)
        b = torch.zeros(c, device=DEV)
        rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
        for res in (False, True):
            r = torch.randn_like(x) if res else None
            xr = x.detach().requires_grad_(True)
            rr = r.detach().requires_grad_(True) if res else None
            holder = {}

            def fwd():
                holder["y"] = ops.batch_norm_act(xr, g, b, rm, rv, residual=rr, relu=True, training=True)
            t = timeit(fwd)
            rec("bn fwd(stats+apply) %dx%dx%d%s" % (c, hw, hw, "+res" if res else ""), t, 2 * n * (3 + (1 if res else 0)))
            dy = torch.randn_like(x)

            def bwd():
                fwd()
                holder["y"].backward(dy)
            tb = timeit(bwd) - t
            rec("bn bwd(reduce+apply) %dx%dx%d%s" % (c, hw, hw, "+res" if res else ""), tb,
                2 * n * ((7 if res else 5) + (1 if res else 0)))

    # ---- pools / optimizer / loss
    x = torch.randn(B, 64, 112, 112, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    t = timeit(lambda: ops.max_pool_3x3_s2(x))
    rec("maxpool3x3s2 fwd 64x112x112", t, 2 * x.numel() * 1.25 + x.numel() // 4)
    from edl_b200.parallel import FlatParams
    lin = torch.nn.Linear(5000, 5000, bias=False).to(DEV)
    lin.weight.data = lin.weight.data.bfloat16()
    flat = FlatParams(lin)
    opt = ops.FlatSGDMomentum(flat, lr=0.1)
    nparam = flat.total_numel()
    t = timeit(lambda: opt.step())
    rec("fused SGD-momentum 25M bf16 params", t, nparam * (2 + 4 + 4 + 4 + 4 + 2))
    z = torch.randn(B, 1000, device=DEV).bfloat16().requires_grad_(True)
    tt = torch.softmax(torch.randn(B, 1000, device=DEV), -1).bfloat16()
    t = timeit(lambda: ops.soft_cross_entropy(z, tt))
    rec("soft-CE fwd 32x1000", t, B * 1000 * 4)

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"peaks": {"hbm_Bps": bw, "bf16_flops": fl, "source": src}, "batch": B, "rows": rows},
              open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
