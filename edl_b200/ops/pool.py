"""NHWC pooling ops (csrc/pool.cu): max 3x3/s2/p1, avg 2x2/s2 (ceil, exclusive), global average.
Reference: Paddle ``pool2d`` call sites in example/distill/resnet/models/resnet_vd.py:97-102,
183-189,131-132."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .bn import _cl


def _nhwc(x):
    return x.permute(0, 2, 3, 1)  # contiguous NHWC view of a channels_last tensor


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        from . import native, count_launch

        x = _cl(x)
        n, c, h, w = x.shape
        ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        y = torch.empty((n, c, ho, wo), device=x.device, dtype=x.dtype,
                        memory_format=torch.channels_last)
        idx = torch.empty((n, ho, wo, c), device=x.device, dtype=torch.uint8)
        native().maxpool3x3s2_fwd(_nhwc(x), _nhwc(y), idx)
        count_launch()
        ctx.save_for_backward(idx)
        ctx.shape = (n, c, h, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import native, count_launch

        (idx,) = ctx.saved_tensors
        n, c, h, w = ctx.shape
        dy = _cl(dy)
        dx = torch.empty((n, c, h, w), device=dy.device, dtype=dy.dtype,
                         memory_format=torch.channels_last)
        native().maxpool3x3s2_bwd(_nhwc(dy), idx, _nhwc(dx))
        count_launch()
        return dx


class _AvgPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        from . import native, count_launch

        x = _cl(x)
        n, c, h, w = x.shape
        y = torch.empty((n, c, (h + 1) // 2, (w + 1) // 2), device=x.device, dtype=x.dtype,
                        memory_format=torch.channels_last)
        native().avgpool2x2_fwd(_nhwc(x), _nhwc(y))
        count_launch()
        ctx.shape = (n, c, h, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import native, count_launch

        n, c, h, w = ctx.shape
        dy = _cl(dy)
        dx = torch.empty((n, c, h, w), device=dy.device, dtype=dy.dtype,
                         memory_format=torch.channels_last)
        native().avgpool2x2_bwd(_nhwc(dy), _nhwc(dx))
        count_launch()
        return dx


class _GapFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        from . import native, count_launch

        x = _cl(x)
        n, c, h, w = x.shape
        y = torch.empty((n, c), device=x.device, dtype=x.dtype)
        native().gap_fwd(_nhwc(x), y)
        count_launch()
        ctx.shape = (n, c, h, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import native, count_launch

        n, c, h, w = ctx.shape
        dx = torch.empty((n, c, h, w), device=dy.device, dtype=dy.dtype,
                         memory_format=torch.channels_last)
        native().gap_bwd(dy.contiguous(), _nhwc(dx))
        count_launch()
        return dx


def max_pool_3x3_s2(x):
    if x.is_cuda:
        return _MaxPoolFn.apply(x)
    return F.max_pool2d(x, 3, 2, 1)


def avg_pool_2x2(x):
    if x.is_cuda:
        return _AvgPoolFn.apply(x)
    return F.avg_pool2d(x, 2, 2, 0, ceil_mode=True, count_include_pad=False)


def global_avg_pool(x):
    """[N, C, H, W] -> [N, C]"""
    if x.is_cuda:
        return _GapFn.apply(x)
    return x.float().mean((2, 3)).to(x.dtype)
