#!/usr/bin/env python
"""Phase timeline of CTA 0 of one persistent tcgen05 kernel launch (clock64 stamps written by the kernel itself when
``set_persist_trace`` is given a buffer; csrc/gemm_persist.cu EDL_TRACE):

  python tools/trace_persist.py fwd M K N | dgradbn M K N [y] | conv3 N C K H W

Roles: producer (0: first slot of the tile is free), MMA (0: accumulator free, 1: first operands landed, 2: last ones
landed), epilogue (0: tile loop top, 1: staging buffer free, 2: accumulator complete, 3: tile packed into the staging
buffer, 4: store issued, 5: BN tiles landed, 6: reduction / statistics done, 7: all warps done).  Printed relative to the
first stamp, in cycles."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edl_b200 import ops  # noqa: E402

op, a = sys.argv[1], sys.argv[2:]
dev = "cuda"
torch.manual_seed(0)
nat = ops.native()
buf = torch.zeros(3 * 16 * 8, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

if op == "fwd":
    m, k, n = [int(v) for v in a[:3]]
    x = torch.randn(m, k, device=dev).bfloat16()
    w = torch.randn(n, k, device=dev).bfloat16()
    y = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(2 * n, device=dev)
    fn = lambda: ops.gemm_bf16(x, w, out=y, col_stats=st)  # noqa: E731
elif op == "dgradbn":
    from edl_b200.ops.bn import BNBackwardHook
    m, k, n = [int(v) for v in a[:3]]
    dyv = torch.randn(m, k, device=dev).bfloat16()
    w = (torch.randn(k, n, device=dev) * 0.1).bfloat16()
    h = BNBackwardHook()
    h.x = torch.randn(m, n, device=dev).bfloat16()
    h.y = torch.randn(m, n, device=dev).bfloat16() if len(a) > 3 else None
    h.relu = True
    h.mean, h.rstd = torch.randn(n, device=dev) * 0.1, torch.rand(n, device=dev) + 0.5
    h.gamma, h.beta = torch.rand(n, device=dev) + 0.5, torch.randn(n, device=dev) * 0.2
    h.dsums = torch.zeros(2 * n, device=dev)
    out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm_bf16(dyv, w, out=out, b_mn_major=True, bn=h)  # noqa: E731
elif op == "conv3":
    nb, c, k, hh, ww = [int(v) for v in a[:5]]
    x = torch.randn(nb, c, hh, ww, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(k, 3, 3, c, device=dev) * 0.05).bfloat16()
    st = torch.zeros(2 * k, device=dev)
    nat.set_conv_halo(False)            # the generic ring is the instrumented path
    fn = lambda: ops.conv3x3(x, wt, st)  # noqa: E731
else:
    raise SystemExit("unknown op " + op)

for _ in range(3):
    fn()
torch.cuda.synchronize()
for mode in ("warm", "cold"):
    if mode == "cold":
        flush.zero_()
    buf.zero_()
    nat.set_persist_trace(buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    nat.set_persist_trace(None)
    t = buf.view(3, 16, 8).cpu()
    nz = t[t > 0]
    base = int(nz.min()) if nz.numel() else 0
    print("== %s %s (%s): kernel %.1f us" % (op, " ".join(a), mode, e0.elapsed_time(e1) * 1e3))
    for role, name in enumerate(("producer", "mma", "epilogue")):
        for tile in range(12):
            row = [int(v) - base if int(v) > 0 else -1 for v in t[role, tile]]
            if any(v >= 0 for v in row):
                print("  %-8s tile %2d  %s" % (name, tile, " ".join("%7d" % v for v in row)))
