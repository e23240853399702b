"""tcgen05 implicit-GEMM 3x3 convolution (csrc/conv3x3.cu): fprop + fused BN statistics + dgrad vs the
fp32 PyTorch convolution; autograd wrapper incl. the side-stream library wgrad."""
import pytest
import torch
import torch.nn.functional as F

from edl_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


SHAPES = [  # n, cin, cout, h, w
    (2, 64, 64, 56, 56), (3, 128, 128, 28, 28), (4, 256, 256, 14, 14), (5, 512, 512, 7, 7),
    (2, 64, 128, 10, 12), (1, 128, 64, 9, 30), (3, 64, 192, 5, 5), (2, 64, 64, 33, 128), (2, 64, 64, 11, 20)]


@pytest.mark.parametrize("n,cin,cout,h,w", SHAPES)
def test_conv3x3_fprop_and_stats(n, cin, cout, h, w):
    torch.manual_seed(0)
    x = torch.randn(n, cin, h, w, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, 3, 3, cin, device=DEV) * 0.05).bfloat16()
    assert ops.conv3x3_supported(x, wt)
    stats = torch.zeros(2 * cout, device=DEV)
    y = ops.conv3x3(x, wt, stats)
    ref = F.conv2d(x.float(), wt.permute(0, 3, 1, 2).float(), None, 1, 1)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert _rel(y, ref) < 1e-2
    yf = y.float()
    assert _rel(stats[:cout], yf.sum((0, 2, 3))) < 2e-3
    assert _rel(stats[cout:], (yf * yf).sum((0, 2, 3))) < 2e-3


@pytest.mark.parametrize("n,cin,cout,h,w", [s for s in SHAPES if s[1] % 64 == 0 and (s[1] <= 64 or s[1] % 128 == 0)])
def test_conv3x3_dgrad_and_wgrad(n, cin, cout, h, w):
    torch.manual_seed(1)
    x = torch.randn(n, cin, h, w, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wt = (torch.randn(cout, 3, 3, cin, device=DEV) * 0.05).bfloat16().requires_grad_(True)
    y = ops.conv3x3(x, wt)
    dy = torch.randn_like(y)
    y.backward(dy)
    xr = x.detach().float().requires_grad_(True)
    wr = wt.detach().float().requires_grad_(True)
    F.conv2d(xr, wr.permute(0, 3, 1, 2), None, 1, 1).backward(dy.float())
    assert _rel(x.grad, xr.grad) < 1e-2
    assert _rel(wt.grad, wr.grad) < 1e-2


def test_conv3x3_unsupported_shapes_are_reported():
    x = torch.randn(2, 32, 8, 8, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    assert not ops.conv3x3_supported(x, torch.randn(64, 3, 3, 32, device=DEV).bfloat16())      # Cin % 64
    x = torch.randn(2, 64, 8, 200, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    assert not ops.conv3x3_supported(x, torch.randn(64, 3, 3, 64, device=DEV).bfloat16())      # W > 128
