import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edl_b200 import ops
from edl_b200.ops.bn import BNBackwardHook
dev = "cuda"
torch.manual_seed(0)


def ref_sums(dy, x, y, mean, rstd, gamma, beta, relu):
    dyf, xf = dy.float(), x.float()
    if relu:
        mask = (y.float() > 0) if y is not None else (torch.addcmul(beta - mean * gamma * rstd, xf, gamma * rstd) > 0)
        dyf = dyf * mask
    xh = (xf - mean) * rstd
    return torch.cat([dyf.sum(0), (dyf * xh).sum(0)])


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


for (m, k, n) in [(2048, 64, 32), (2048, 128, 64), (5000, 64, 128), (6272, 512, 256), (300, 64, 1024)]:
    for relu, has_y in [(True, False), (True, True), (False, False)]:
        a = torch.randn(m, k, device=dev).bfloat16()
        w = (torch.randn(k, n, device=dev) * 0.1).bfloat16()
        x = torch.randn(m, n, device=dev).bfloat16()
        y = torch.randn(m, n, device=dev).bfloat16() if has_y else None
        mean, rstd = torch.randn(n, device=dev) * 0.1, torch.rand(n, device=dev) + 0.5
        gamma, beta = torch.rand(n, device=dev) + 0.5, torch.randn(n, device=dev) * 0.2
        h = BNBackwardHook()
        h.x, h.y, h.mean, h.rstd, h.gamma, h.beta, h.relu = x, y, mean, rstd, gamma, beta, relu
        h.dsums = torch.zeros(2 * n, device=dev)
        d = ops.gemm_bf16(a, w, b_mn_major=True, bn=h)
        torch.cuda.synchronize()
        ref = ref_sums(d, x, y, mean, rstd, gamma, beta, relu)
        dref = a.float() @ w.float()
        print("gemm m%d k%d n%d relu%d y%d  D rel %.2e  dsums rel %.3e  done=%s" % (m, k, n, relu, has_y, rel(d.float(), dref), rel(h.dsums, ref), h.done), flush=True)

for (nb, c, hh, ww) in [(8, 64, 16, 16), (8, 128, 8, 8), (32, 256, 14, 14), (5, 64, 11, 20)]:
    for relu in (True, False):
        dy = torch.randn(nb, c, hh, ww, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        wt = (torch.randn(c, 3, 3, c, device=dev) * 0.05).bfloat16()
        x = torch.randn(nb, c, hh, ww, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        mean, rstd = torch.randn(c, device=dev) * 0.1, torch.rand(c, device=dev) + 0.5
        gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.2
        h = BNBackwardHook()
        h.x, h.y, h.mean, h.rstd, h.gamma, h.beta, h.relu = x, None, mean, rstd, gamma, beta, relu
        h.dsums = torch.zeros(2 * c, device=dev)
        dx = torch.empty_like(x)
        ops.native().conv3x3(dy, wt, dx, True, None, h.as_list(nb * hh * ww, c), relu)
        torch.cuda.synchronize()
        d2 = dx.permute(0, 2, 3, 1).reshape(-1, c)
        x2 = x.permute(0, 2, 3, 1).reshape(-1, c)
        ref = ref_sums(d2, x2, None, mean, rstd, gamma, beta, relu)
        print("conv3 n%d c%d %dx%d relu%d dsums rel %.3e" % (nb, c, hh, ww, relu, rel(h.dsums, ref)), flush=True)
