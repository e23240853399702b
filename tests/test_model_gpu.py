"""Flagship model on the native kernels vs the same weights run through plain fp32 PyTorch ops."""
import pytest
import torch
import torch.nn.functional as F

from edl_b200 import ops
from edl_b200.models import ResNetVd, to_train_dtype
from edl_b200.models.resnet_vd import Bottleneck
from edl_b200.trainer import StudentTrainer

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref_unit(u, x, residual=None):
    """Same rounding points as the fused path (bf16 conv output, fp32 BN math, bf16 activation),
    but built from stock PyTorch ops."""
    w = u.weight.detach().permute(0, 3, 1, 2)
    y = F.conv2d(x, w, None, u.stride, (u.k - 1) // 2).float()
    y = F.batch_norm(y, None, None, u.bn.weight.detach().float(), u.bn.bias.detach().float(), True, 0.1, u.bn.eps)
    if residual is not None:
        y = y + residual.float()
    return (torch.relu(y) if u.bn.relu else y).bfloat16()


def _ref_forward(m, x):
    for u in m.stem:
        x = _ref_unit(u, x)
    x = F.max_pool2d(x, 3, 2, 1)
    for b in m.blocks:
        s = x
        if b.short is not None:
            s = F.avg_pool2d(x.float(), 2, 2, 0, ceil_mode=True, count_include_pad=False).bfloat16() if b.pool else x
            s = _ref_unit(b.short, s)
        if isinstance(b, Bottleneck):
            x = _ref_unit(b.c, _ref_unit(b.b, _ref_unit(b.a, x)), s)
        else:
            x = _ref_unit(b.b, _ref_unit(b.a, x), s)
    x = x.float().mean((2, 3)).bfloat16()
    return F.linear(x.float(), m.fc_weight.detach().float(), m.fc_bias.detach().float())


@pytest.mark.parametrize("layers,impl", [(18, "auto"), (50, "auto"), (50, "cudnn")])
def test_model_matches_fp32_reference(layers, impl):
    torch.manual_seed(0)
    m = to_train_dtype(ResNetVd(layers, class_dim=104, impl=impl, width_mult=0.5), torch.bfloat16, DEV).train()
    x = torch.randn(16, 3, 96, 96, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    y = m(x)
    ref = _ref_forward(m, x)
    rel = ((y.float() - ref).norm() / ref.norm()).item()
    assert rel < 0.15, rel  # bf16 rounding through 50+ BN layers
    t = torch.softmax(torch.randn(16, 104, device=DEV), -1).bfloat16()
    loss = ops.soft_cross_entropy(y, t)
    loss.backward()
    g = m.fc_weight.grad
    assert g is not None and torch.isfinite(g.float()).all()
    first = m.stem[0].weight.grad
    assert first is not None and torch.isfinite(first.float()).all() and first.float().abs().sum() > 0


@pytest.mark.parametrize("graph", [False, True])
def test_trainer_learns(graph):
    torch.manual_seed(0)
    m = to_train_dtype(ResNetVd(18, class_dim=16, width_mult=0.25), torch.bfloat16, DEV).train()
    tr = StudentTrainer(m, batch_size=16, image_shape=(3, 32, 32), num_classes=16, lr=0.05, use_graph=graph)
    x = torch.randn(16, 3, 32, 32).bfloat16().contiguous(memory_format=torch.channels_last).pin_memory()
    t = torch.zeros(16, 16)
    t[torch.arange(16), torch.arange(16)] = 1.0
    t = t.bfloat16().pin_memory()
    losses = [float(tr.step(x, t).item()) for _ in range(30)]
    assert losses[-1] < 0.5 * losses[0], losses
    assert ops.launches() > 0
