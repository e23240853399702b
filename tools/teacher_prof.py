#!/usr/bin/env python
"""Teacher (ResNeXt101_32x16d) forward time and per-kernel breakdown: python tools/teacher_prof.py [--fp8] [--batch 32]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edl_b200.models.resnext import ResNeXt101_32x16d, to_inference_dtype  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--fp8", action="store_true")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--no-graph", action="store_true")
args = ap.parse_args()
dev = "cuda"
m = to_inference_dtype(ResNeXt101_32x16d(), torch.bfloat16, dev).eval()
x = torch.randn(args.batch, 3, 224, 224, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
if args.fp8:
    print("fp8 layers:", m.enable_fp8(x))
with torch.no_grad():
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    g = None
    if not args.no_graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = m(x)
    run = (lambda: g.replay()) if g is not None else (lambda: m(x))
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("forward %.2f ms / batch %d = %.0f img/s, %.0f TFLOP/s" % (ms, args.batch, args.batch / ms * 1e3, 72.3e9 * args.batch / ms / 1e9))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            run()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=100))
