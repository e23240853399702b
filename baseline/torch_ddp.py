"""Comparator: the same ResNet50_vd / batch / optimizer written with stock PyTorch modules --
cuDNN convolutions + cuDNN BatchNorm (channels_last, bf16), torch.optim.SGD(momentum) and
DistributedDataParallel over NCCL.  This is the "library path on the same box" stand-in for the
reference's Paddle + cuDNN + NCCL build (which cannot be installed offline); it is
NOT the `--impl reference` arm.  `bench.py --impl torch` runs it through the same timing harness.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F


def conv_bn(cin, cout, k, stride=1, relu=True):
    layers = [nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, bias=False), nn.BatchNorm2d(cout)]
    if relu:
        layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*layers)


class Bottleneck(nn.Module):
    def __init__(self, cin, width, stride, if_first):
        super().__init__()
        cout = width * 4
        self.a = conv_bn(cin, width, 1)
        self.b = conv_bn(width, width, 3, stride)
        self.c = conv_bn(width, cout, 1, relu=False)
        self.short = None
        if cin != cout or stride != 1 or if_first:
            if if_first or stride == 1:
                self.short = conv_bn(cin, cout, 1, stride, relu=False)
            else:
                self.short = nn.Sequential(nn.AvgPool2d(2, 2, 0, ceil_mode=True, count_include_pad=False),
                                           conv_bn(cin, cout, 1, 1, relu=False))

    def forward(self, x):
        s = x if self.short is None else self.short(x)
        return F.relu(self.c(self.b(self.a(x))) + s)


class TorchResNet50vd(nn.Module):
    def __init__(self, class_dim=1000, depth=(3, 4, 6, 3)):
        super().__init__()
        self.stem = nn.Sequential(conv_bn(3, 32, 3, 2), conv_bn(32, 32, 3), conv_bn(32, 64, 3),
                                  nn.MaxPool2d(3, 2, 1))
        blocks, cin = [], 64
        for stage, n in enumerate(depth):
            for i in range(n):
                blocks.append(Bottleneck(cin, 64 << stage, 2 if i == 0 and stage else 1, stage == 0 and i == 0))
                cin = (64 << stage) * 4
        self.blocks = nn.Sequential(*blocks)
        self.fc = nn.Linear(cin, class_dim)

    def forward(self, x):
        x = self.blocks(self.stem(x))
        return self.fc(x.mean((2, 3)))


class TorchDDPTrainer:
    """Same step()/step_device() surface as edl_b200.trainer.StudentTrainer."""

    def __init__(self, batch_size, device, layers=50, use_graph=True, lr=0.1):
        depth = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}[layers]
        self.device = device
        self.model = TorchResNet50vd(depth=depth).to(device=device, dtype=torch.bfloat16)
        self.model = self.model.to(memory_format=torch.channels_last).train()
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        if self.world > 1:
            # DDP + whole-step CUDA-graph capture recipe: build DDP under a side stream
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.net = nn.parallel.DistributedDataParallel(self.model, device_ids=[device.index],
                                                               bucket_cap_mb=16, gradient_as_bucket_view=True)
            torch.cuda.current_stream().wait_stream(side)
        else:
            self.net = self.model
        self.opt = torch.optim.SGD(self.model.parameters(), lr=lr, momentum=0.9, weight_decay=1e-4)
        self.static_x = torch.zeros(batch_size, 3, 224, 224, dtype=torch.bfloat16, device=device).contiguous(
            memory_format=torch.channels_last)
        self.static_t = torch.zeros(batch_size, 1000, dtype=torch.bfloat16, device=device)
        self.static_loss = torch.zeros((), dtype=torch.float32, device=device)
        self.use_graph = use_graph
        self.graph = None

    def _body(self):
        self.opt.zero_grad(set_to_none=False)
        logits = self.net(self.static_x)
        loss = -(self.static_t.float() * F.log_softmax(logits.float(), -1)).sum(-1).mean()
        loss.backward()
        self.opt.step()
        self.static_loss.copy_(loss.detach())

    def capture(self):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(11 if self.world > 1 else 3):  # DDP needs ~11 warm-up iters before capture
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()

    def step_device(self):
        if self.use_graph:
            if self.graph is None:
                try:
                    self.capture()
                except Exception as e:  # noqa: BLE001 - stock DDP is not always capturable: run eagerly
                    import sys
                    print("torch baseline: CUDA-graph capture failed (%s); running eagerly" % type(e).__name__,
                          file=sys.stderr)
                    torch.cuda.synchronize()
                    self.use_graph = False
                    self.graph = None
                    self._body()
                    return self.static_loss
            self.graph.replay()
        else:
            self._body()
        return self.static_loss

    def step(self, images, targets):
        self.static_x.copy_(images, non_blocking=True)
        self.static_t.copy_(targets, non_blocking=True)
        return self.step_device()
