#!/usr/bin/env python
"""In-graph time of every BatchNorm kernel for every BN shape of ResNet50_vd (batch 32): the kernels
are captured back to back in a CUDA graph over a ROTATING set of buffers (so that L2 does not hold the
operands unless the working set is tiny) and timed with events.  Prints us / kernel, the bytes it must
move and the achieved fraction of the measured copy bandwidth, plus the per-step total.

    python tools/bn_shapes.py [--hot]      (--hot: one buffer set, operands L2 resident)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edl_b200 import ops  # noqa: E402

dev = "cuda"
C = ops.native()
hot = "--hot" in sys.argv
peak = 6.5e12
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
except Exception:  # noqa: BLE001
    pass
N = 32
# (C, H, W, residual, relu, count per step)
shapes = [(32, 112, 112, 0, 1, 2), (64, 112, 112, 0, 1, 1),
          (64, 56, 56, 0, 1, 6), (256, 56, 56, 1, 1, 3), (256, 56, 56, 0, 0, 1),
          (128, 56, 56, 0, 1, 1), (128, 28, 28, 0, 1, 7), (512, 28, 28, 1, 1, 4), (512, 28, 28, 0, 0, 1),
          (256, 28, 28, 0, 1, 1), (256, 14, 14, 0, 1, 11), (1024, 14, 14, 1, 1, 6), (1024, 14, 14, 0, 0, 1),
          (512, 14, 14, 0, 1, 1), (512, 7, 7, 0, 1, 5), (2048, 7, 7, 1, 1, 3), (2048, 7, 7, 0, 0, 1)]
REPS = 24


def timed(fn_list):
    for f in fn_list:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fn_list:
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * len(fn_list))


tot = {"stats": 0.0, "apply": 0.0, "bwd_reduce": 0.0, "bwd_apply": 0.0}
print("%-22s %10s %10s %12s %12s   (us/kernel, fraction of copy bw)" % ("shape", "stats", "apply", "bwd_reduce", "bwd_apply"))
for (c, h, w, res, relu, cnt) in shapes:
    m = N * h * w
    nbytes = m * c * 2
    nset = 1 if hot else max(2, min(12, int(400e6 // (nbytes * (4 + res))) + 1))
    S = []
    for _ in range(nset):
        d = {"x": torch.randn(m, c, device=dev).bfloat16(), "y": torch.empty(m, c, device=dev, dtype=torch.bfloat16),
             "dy": torch.randn(m, c, device=dev).bfloat16(), "dx": torch.empty(m, c, device=dev, dtype=torch.bfloat16),
             "res": torch.randn(m, c, device=dev).bfloat16() if res else None,
             "dres": torch.empty(m, c, device=dev, dtype=torch.bfloat16) if res else None}
        S.append(d)
    sums, dsums = torch.zeros(2 * c, device=dev), torch.zeros(2 * c, device=dev)
    g_, b_ = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    mean, rstd = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    C.bn_stats(S[0]["x"], sums)
    C.bn_apply(S[0]["x"], S[0]["res"], S[0]["y"], sums, g_, b_, None, None, mean, rstd, 1e-5, 0.1, bool(relu))
    f_stats = [(lambda d=S[i % nset]: C.bn_stats(d["x"], sums)) for i in range(REPS)]
    f_apply = [(lambda d=S[i % nset]: C.bn_apply(d["x"], d["res"], d["y"], sums, g_, b_, None, None, mean, rstd, 1e-5, 0.1, bool(relu)))
               for i in range(REPS)]
    need_y = bool(res)
    f_red = [(lambda d=S[i % nset]: C.bn_bwd_reduce(d["dy"], d["x"], d["y"] if need_y else None, g_, b_, mean, rstd, dsums, bool(relu)))
             for i in range(REPS)]
    f_bap = [(lambda d=S[i % nset]: C.bn_bwd_apply(d["dy"], d["x"], d["y"] if need_y else None, g_, b_, mean, rstd, dsums, d["dx"],
                                                   d["dres"], dg, db, bool(relu), False)) for i in range(REPS)]
    t = {"stats": timed(f_stats), "apply": timed(f_apply), "bwd_reduce": timed(f_red), "bwd_apply": timed(f_bap)}
    moved = {"stats": nbytes, "apply": nbytes * (2 + res), "bwd_reduce": nbytes * (2 + need_y), "bwd_apply": nbytes * (3 + need_y + res)}
    for k in tot:
        tot[k] += t[k] * cnt
    print("%-22s %s" % ("%dx%dx%d%s x%d" % (c, h, w, "+res" if res else "", cnt),
                        " ".join("%6.1f(%.2f)" % (t[k], moved[k] / (t[k] * 1e-6) / peak) for k in ("stats", "apply", "bwd_reduce", "bwd_apply"))), flush=True)
    del S
print("per-step totals (us): " + ", ".join("%s %.0f" % kv for kv in tot.items()) + "  | sum %.0f (stats only counted for all layers; GEMM-fused layers skip it)" % sum(tot.values()))
