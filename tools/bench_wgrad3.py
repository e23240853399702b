#!/usr/bin/env python
"""3x3 weight-gradient microbenchmark: own tcgen05 kernel (csrc/conv3x3_wgrad.cu) vs the library kernel, per ResNet50_vd
layer shape and per split-K factor; CUDA events, L2 flushed between iterations.

    python tools/bench_wgrad3.py [--batch 32] [--out gpurun_out/wgrad3.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edl_b200 import ops  # noqa: E402
from edl_b200.ops import gemm as G  # noqa: E402
from tools.bench_kernels import timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "wgrad3.json"))
    args = ap.parse_args()
    B, dev = args.batch, "cuda"
    rows = []
    for c, hw in [(64, 56), (128, 28), (256, 14), (512, 7)]:
        x = torch.randn(B, c, hw, hw, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        dy = (torch.randn(B, c, hw, hw, device=dev) * 0.1).bfloat16().contiguous(memory_format=torch.channels_last)
        w = torch.zeros(c, 3, 3, c, device=dev, dtype=torch.bfloat16)
        wv = w.permute(0, 3, 1, 2)
        sink = torch.zeros(w.numel(), device=dev, dtype=torch.bfloat16)
        flops = 2.0 * B * hw * hw * c * c * 9
        t_lib = timeit(lambda: torch.ops.aten.convolution_backward(
            dy, x, wv, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [False, True, False]))
        tiles = ops.native().conv3x3_wgrad_tiles(c, c)
        kblocks = ops.native().conv3x3_wgrad_kblocks(B, hw, hw)
        auto = max(1, min(max(1, 148 // tiles), max(1, kblocks // 2)))
        row = {"C": c, "HW": hw, "lib_us": t_lib * 1e6, "tiles": tiles, "kblocks": kblocks, "auto_split": auto, "own_us": {}}
        for split in sorted({1, 2, 4, 8, 16, 24, 49, auto, max(1, 296 // tiles), max(1, 444 // tiles)}):
            if split > kblocks:
                continue
            t = timeit(lambda: G.conv3x3_wgrad(x, dy, w.shape, sink, split_k=split))
            row["own_us"][str(split)] = t * 1e6
        best = min(row["own_us"].items(), key=lambda kv: kv[1])
        row["best_split"], row["best_us"] = int(best[0]), best[1]
        row["best_tflops"] = flops / (best[1] * 1e-6) / 1e12
        row["lib_tflops"] = flops / t_lib / 1e12
        rows.append(row)
        print("C=%d %dx%d  lib %.1f us (%.0f TF/s)  own auto(split %d) %.1f us  best split %d: %.1f us (%.0f TF/s)  all %s" % (
            c, hw, hw, row["lib_us"], row["lib_tflops"], auto, row["own_us"][str(auto)], row["best_split"], row["best_us"],
            row["best_tflops"], {k: round(v, 1) for k, v in row["own_us"].items()}), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
