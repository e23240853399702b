"""VGG 11/13/16/19 (reference: example/collective/resnet50/models/vgg.py -- conv3x3+ReLU groups,
2x2 max-pool, fc 4096-4096-class_dim with dropout 0.5).  No BN, so there is nothing to fuse beyond
what cuDNN's conv+bias+ReLU already does; the two 4096-wide FC layers run on the tcgen05 GEMM."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops

_CFG = {11: [1, 1, 2, 2, 2], 13: [2, 2, 2, 2, 2], 16: [2, 2, 3, 3, 3], 19: [2, 2, 4, 4, 4]}


class _Linear(nn.Module):
    def __init__(self, cin, cout, relu):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))
        self.bias = nn.Parameter(torch.zeros(cout, dtype=torch.float32))
        nn.init.normal_(self.weight, 0.0, 0.01)
        self.relu = relu

    def forward(self, x):
        if x.is_cuda and x.dtype == torch.bfloat16 and self.weight.shape[0] % 8 == 0:
            y = ops.linear_bf16(x, self.weight, self.bias)
        else:
            y = F.linear(x, self.weight, self.bias.to(x.dtype))
        return F.relu(y) if self.relu else y


class VGG(nn.Module):
    def __init__(self, layers=16, class_dim=1000, width_mult=1.0, image_size=224, dropout=0.5):
        super().__init__()
        chans = [max(8, int(c * width_mult)) for c in (64, 128, 256, 512, 512)]
        convs, cin = [], 3
        for n, c in zip(_CFG[layers], chans):
            for _ in range(n):
                convs += [nn.Conv2d(cin, c, 3, 1, 1), nn.ReLU(inplace=True)]
                cin = c
            convs.append(nn.MaxPool2d(2, 2))
        self.features = nn.Sequential(*convs)
        hidden = max(64, int(4096 * width_mult))
        side = image_size // 32
        self.fc1 = _Linear(cin * side * side, hidden, True)
        self.fc2 = _Linear(hidden, hidden, True)
        self.fc3 = _Linear(hidden, class_dim, False)
        self.dropout = dropout

    def forward(self, x):
        x = self.features(x)
        x = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1) if x.is_contiguous(memory_format=torch.channels_last) \
            else x.flatten(1)
        x = F.dropout(self.fc1(x), self.dropout, self.training)
        x = F.dropout(self.fc2(x), self.dropout, self.training)
        return self.fc3(x)


def VGG11(**kw): return VGG(11, **kw)
def VGG13(**kw): return VGG(13, **kw)
def VGG16(**kw): return VGG(16, **kw)
def VGG19(**kw): return VGG(19, **kw)
