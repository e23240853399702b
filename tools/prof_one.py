#!/usr/bin/env python
"""Run ONE op a few times so that `ncu --set full -k regex:<kernel> -s 3 -c 1 python tools/prof_one.py <op> ...`
captures exactly the kernel of interest.

  wgrad  COUT CIN M      split-K wgrad GEMM (dy[M,COUT]^T @ x[M,CIN]) with fused finalize into a bf16 sink
  fwd    M K N           1x1-conv forward GEMM with BN-statistics epilogue
  dgrad  M K N
  bnbwd  N C H W [fused|stream|regs] [res]
  bnfwd  N C H W [fused|stream|regs] [res]
  conv3 / conv3dgrad  N CIN COUT H W   tcgen05 implicit-GEMM 3x3 convolution
  wgrad3 N CIN COUT H W [SPLIT]        tcgen05 3x3 weight gradient
  conv3g N C COUT H W GROUPS           inference 3x3 (grouped) convolution with folded-BN epilogue (the teacher's layers)
  gemmi  M K N                         inference 1x1 convolution: GEMM + scale / shift / ReLU epilogue (wide tiles)
  dgradbn M K N [y]                    1x1 dgrad with the fused BatchNorm-backward reduction in its epilogue
  sgd | softce B | rope T H D          fused optimizer / loss / rotary kernels
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edl_b200 import ops  # noqa: E402
from edl_b200.ops.gemm import _wgrad  # noqa: E402

op = sys.argv[1]
a = sys.argv[2:]
dev = "cuda"
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(fn, iters=6):
    for _ in range(iters):
        flush.zero_()
        fn()
    torch.cuda.synchronize()


if op == "wgrad":
    cout, cin, m = int(a[0]), int(a[1]), int(a[2])
    dy = torch.randn(m, cout, device=dev).bfloat16()
    x = torch.randn(m, cin, device=dev).bfloat16()
    sink = torch.zeros(cout, cin, device=dev, dtype=torch.bfloat16)
    run(lambda: _wgrad(dy, x, (cout, cin), sink, None))
elif op in ("fwd", "dgrad"):
    m, k, n = int(a[0]), int(a[1]), int(a[2])
    x = torch.randn(m, k, device=dev).bfloat16()
    y = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    if op == "fwd":
        w = torch.randn(n, k, device=dev).bfloat16()
        st = torch.zeros(2 * n, device=dev)
        run(lambda: ops.gemm_bf16(x, w, out=y, col_stats=st))
    else:
        w = torch.randn(k, n, device=dev).bfloat16()
        run(lambda: ops.gemm_bf16(x, w, out=y, b_mn_major=True))
elif op in ("bnbwd", "bnfwd"):
    n, c, h, w = [int(v) for v in a[:4]]
    path = a[4] if len(a) > 4 else "stream"
    res = len(a) > 5 and a[5] == "res"
    ops.set_fused_bn(path == "fused")
    ops.native().bn_set_stream_kernels(path != "regs")
    x = torch.randn(n, c, h, w, device=dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    r = torch.randn_like(x).requires_grad_(True) if res else None
    g, b = torch.ones(c, device=dev, requires_grad=True), torch.zeros(c, device=dev, requires_grad=True)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    dyv = torch.randn_like(x)

    def f():
        y = ops.batch_norm_act(x, g, b, rm, rv, residual=r, relu=True, training=True)
        if op == "bnbwd":
            y.backward(dyv)
    run(f)
elif op in ("conv3", "conv3dgrad"):
    n, c, k, h, w = [int(v) for v in a[:5]]          # batch, Cin, Cout, H, W
    x = torch.randn(n, c if op == "conv3" else k, h, w, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(k, 3, 3, c, device=dev) * 0.05).bfloat16()
    y = torch.empty(n, k if op == "conv3" else c, h, w, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    st = torch.zeros(2 * k, device=dev)
    run(lambda: ops.native().conv3x3(x, wt, y, op != "conv3", st if op == "conv3" else None, None, False))
elif op == "conv3g":
    n, c, k, h, w, gr = [int(v) for v in a[:6]]
    x = torch.randn(n, c, h, w, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(k, 3, 3, c // gr, device=dev) * 0.05).bfloat16()
    sc, sh = torch.rand(k, device=dev) + 0.5, torch.randn(k, device=dev) * 0.1
    run(lambda: ops.conv3x3_infer(x, wt, sc, sh, relu=True, groups=gr))
elif op == "gemmi":
    m, k, n = int(a[0]), int(a[1]), int(a[2])
    x = torch.randn(m, k, device=dev).bfloat16()
    w = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
    y = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    sc, sh = torch.rand(n, device=dev) + 0.5, torch.randn(n, device=dev) * 0.1
    run(lambda: ops.gemm_bf16(x, w, out=y, col_scale=sc, col_shift=sh, relu=True))
elif op == "dgradbn":
    from edl_b200.ops.bn import BNBackwardHook
    m, k, n = int(a[0]), int(a[1]), int(a[2])
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
    run(lambda: ops.gemm_bf16(dyv, w, out=out, b_mn_major=True, bn=h))
elif op == "wgrad3":
    n, c, k, h, w = [int(v) for v in a[:5]]          # batch, Cin, Cout, H, W
    from edl_b200.ops import gemm as G
    x = torch.randn(n, c, h, w, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    dy = (torch.randn(n, k, h, w, device=dev) * 0.1).bfloat16().contiguous(memory_format=torch.channels_last)
    sink = torch.zeros(k * 9 * c, device=dev, dtype=torch.bfloat16)
    split = int(a[5]) if len(a) > 5 else None
    run(lambda: G.conv3x3_wgrad(x, dy, (k, 3, 3, c), sink, split_k=split))
elif op == "sgd":
    from edl_b200.parallel import FlatParams
    lin = torch.nn.Linear(5000, 5000, bias=False).to(dev)
    lin.weight.data = lin.weight.data.bfloat16()
    flat = FlatParams(lin)
    opt = ops.FlatSGDMomentum(flat, lr=0.1)
    for g in flat.groups.values():
        g.grad.normal_()
    run(lambda: opt.step())
elif op == "softce":
    z = torch.randn(int(a[0]), 1000, device=dev).bfloat16().requires_grad_(True)
    tt = torch.softmax(torch.randn(int(a[0]), 1000, device=dev), -1).bfloat16()
    run(lambda: ops.soft_cross_entropy(z, tt).backward())
elif op == "rope":
    T, H, D = int(a[0]), int(a[1]), int(a[2])
    xq = torch.randn(T, H, D, device=dev).bfloat16()
    from edl_b200.ops.misc import rope, rope_tables
    cos, sin = rope_tables(T, D, device=dev)
    run(lambda: rope(xq, cos, sin))
else:
    raise SystemExit("unknown op " + op)
