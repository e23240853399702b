"""Plain ResNet (v1.5: stride on the 3x3) -- the model of the reference's elastic ImageNet example
(example/collective/resnet50/models/resnet.py: 7x7/2 stem + 3x3/2 max-pool, bottleneck stride in the
3x3 conv, 1x1 strided projection shortcut).  Built from the same fused ``ConvBNAct`` units as the
``_vd`` student, so the 1x1 convolutions run on the tcgen05 GEMM and every BN is the fused op."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .resnet_vd import ConvBNAct, FCHead, _DEPTHS


class _Bottleneck(nn.Module):
    def __init__(self, cin, width, stride, impl):
        super().__init__()
        cout = width * 4
        self.a = ConvBNAct(cin, width, 1, 1, True, impl=impl)
        self.b = ConvBNAct(width, width, 3, stride, True, impl=impl)
        self.c = ConvBNAct(width, cout, 1, 1, True, impl=impl)
        self.short = ConvBNAct(cin, cout, 1, stride, False, impl=impl) if (cin != cout or stride != 1) else None

    def forward(self, x):
        if self.a.can_fork(x):           # both gradients of x are summed in conv a's dgrad epilogue
            y, xs = self.a(x, fork=True)
        else:
            ops.drop_bn_hook(x)
            y, xs = self.a(x), x
        s = self.short(xs) if self.short is not None else xs
        return self.c(self.b(y), residual=s)


class _Basic(nn.Module):
    def __init__(self, cin, width, stride, impl):
        super().__init__()
        self.a = ConvBNAct(cin, width, 3, stride, True, impl=impl)
        self.b = ConvBNAct(width, width, 3, 1, True, impl=impl)
        self.short = ConvBNAct(cin, width, 1, stride, False, impl=impl) if (cin != width or stride != 1) else None

    def forward(self, x):
        ops.drop_bn_hook(x)
        s = self.short(x) if self.short is not None else x
        return self.b(self.a(x), residual=s)


class ResNet(nn.Module):
    def __init__(self, layers=50, class_dim=1000, impl="auto", width_mult=1.0):
        super().__init__()
        depth = _DEPTHS[layers]
        bottleneck = layers >= 50
        w = [max(8, int(c * width_mult)) for c in (64, 128, 256, 512)]
        self.stem = ConvBNAct(3, w[0], 7, 2, True, impl=impl)
        blocks, cin = [], w[0]
        for stage, n in enumerate(depth):
            for i in range(n):
                stride = 2 if i == 0 and stage != 0 else 1
                if bottleneck:
                    blocks.append(_Bottleneck(cin, w[stage], stride, impl))
                    cin = w[stage] * 4
                else:
                    blocks.append(_Basic(cin, w[stage], stride, impl))
                    cin = w[stage]
        self.blocks = nn.Sequential(*blocks)
        self.fc = FCHead(cin, class_dim)     # last: gradient-ready order == reverse registration order
        self.impl = impl

    def forward(self, x):
        x = ops.max_pool_3x3_s2(self.stem(x))
        x = ops.global_avg_pool(self.blocks(x))
        if x.is_cuda and x.dtype == torch.bfloat16 and self.impl != "cudnn":
            return ops.linear_bf16(x, self.fc.weight, self.fc.bias)
        return F.linear(x, self.fc.weight, self.fc.bias.to(x.dtype))


def ResNet18(**kw): return ResNet(18, **kw)
def ResNet34(**kw): return ResNet(34, **kw)
def ResNet50(**kw): return ResNet(50, **kw)
def ResNet101(**kw): return ResNet(101, **kw)
def ResNet152(**kw): return ResNet(152, **kw)
