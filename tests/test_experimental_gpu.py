"""Kernels written after the round-1 GPU budget was spent: compiled and wired in behind environment switches,
NOT yet validated on hardware.  They are skipped unless ``EDL_TEST_EXPERIMENTAL=1`` so that an unvalidated kernel
cannot take the regular GPU suite down; the first GPU call of the next round runs

    EDL_TEST_EXPERIMENTAL=1 python -m pytest tests/test_experimental_gpu.py -q

and whatever passes gets promoted (switch default flipped, test moved to its permanent file)."""
import os

import pytest
import torch
import torch.nn.functional as F

from edl_b200 import ops

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("EDL_TEST_EXPERIMENTAL", "0") != "1",
                                 reason="experimental kernels: set EDL_TEST_EXPERIMENTAL=1")]
DEV = "cuda"


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


WGRAD_SHAPES = [  # n, cin, cout, h, w      (ResNet50_vd stages first, then odd geometries)
    (8, 64, 64, 56, 56), (8, 128, 128, 28, 28), (8, 256, 256, 14, 14), (32, 512, 512, 7, 7),
    (2, 64, 128, 8, 16), (4, 128, 64, 14, 14), (3, 64, 192, 6, 8), (16, 64, 64, 7, 7), (2, 192, 64, 4, 28)]


@pytest.mark.parametrize("n,cin,cout,h,w", WGRAD_SHAPES)
@pytest.mark.parametrize("split", [1, 4, None])
def test_conv3x3_wgrad_tcgen05(n, cin, cout, h, w, split):
    torch.manual_seed(0)
    x = torch.randn(n, cin, h, w, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    dy = (torch.randn(n, cout, h, w, device=DEV) * 0.1).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = torch.zeros(cout, 3, 3, cin, device=DEV, dtype=torch.bfloat16)
    if not ops.conv3x3_wgrad_supported(x, wt):
        pytest.skip("geometry not supported by the pixel-box planner")
    xr = x.float()
    wr = torch.zeros(cout, cin, 3, 3, device=DEV, requires_grad=True)
    F.conv2d(xr, wr, None, 1, 1).backward(dy.float())
    ref = wr.grad.permute(0, 2, 3, 1)                       # KRSC
    got = ops.conv3x3_wgrad(x, dy, wt.shape, None, split_k=split)
    assert got.shape == ref.shape
    assert _rel(got, ref) < 1e-2, (split, _rel(got, ref))
    # accumulate into a sink (the flat gradient bucket): sink += dW, twice
    sink = torch.full((cout * 9 * cin,), 0.25, device=DEV, dtype=torch.bfloat16)
    assert ops.conv3x3_wgrad(x, dy, wt.shape, sink, split_k=split) is None
    assert _rel(sink.view(ref.shape), ref + 0.25) < 1e-2
    # the shared split-K workspace and the tile counters are left all-zero
    from edl_b200.ops import gemm as G
    ws, counters = G._splitk_workspace(x.device, cout * 9 * cin, 1)
    torch.cuda.synchronize()
    assert float(ws.abs().max()) == 0.0 and int(counters.abs().max()) == 0


def test_conv3x3_autograd_with_own_wgrad(monkeypatch):
    from edl_b200.ops import gemm as G

    monkeypatch.setattr(G, "OWN_WGRAD3", True)
    torch.manual_seed(1)
    x = torch.randn(4, 64, 28, 28, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wt = (torch.randn(128, 3, 3, 64, device=DEV) * 0.05).bfloat16().requires_grad_(True)
    y = ops.conv3x3(x, wt)
    dy = torch.randn_like(y)
    y.backward(dy)
    xr = x.detach().float().requires_grad_(True)
    wr = wt.detach().float().requires_grad_(True)
    F.conv2d(xr, wr.permute(0, 3, 1, 2), None, 1, 1).backward(dy.float())
    assert _rel(x.grad, xr.grad) < 1e-2
    assert _rel(wt.grad, wr.grad) < 1e-2
