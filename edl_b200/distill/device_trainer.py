"""Same-box distillation service: student ranks train ResNet50_vd while teacher ranks run
ResNeXt101_32x16d and feed soft labels through the NVSwitch-direct link (``device_feed.py``).

BASELINE.json config 2 ("student on GPUs 0-3, teacher on GPUs 4-7, logit ship over NVSwitch"); the
reference's equivalent is DistillReader + 40 Paddle-Serving P4 teachers (README.md:85, 1514 img/s).
Each student rank ``s`` is paired with teacher rank ``n_students + s``.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .. import ops
from ..trainer import StudentTrainer
from .device_feed import DeviceDistillLink


class DistillStudentTrainer(StudentTrainer):
    """StudentTrainer whose targets arrive from the paired teacher GPU instead of the host."""

    def __init__(self, model, batch_size, link: DeviceDistillLink, **kw):
        kw.setdefault("target_kind", "logits")
        super().__init__(model, batch_size, **kw)
        self.link = link

    def _step_body(self):
        self.link.seq.add_(1)                       # device-side step counter (graph replay safe)
        self.dp.zero_grad()
        if self.arena is not None:
            self.arena.zero()
        x = self.static_x
        self.link.send_images(x, 0, seq=self.link.seq)      # -> teacher HBM, flag release
        logits = self.model(x if x.dtype == self.dtype else x.to(self.dtype))
        loss = self.link.loss(logits, 0, seq=self.link.seq)  # acquires the teacher's logits, fused soft-CE
        loss.backward()
        self.dp.finish()
        self.opt.step()
        self.static_loss.copy_(loss.detach())

    def step(self, images, targets=None):
        self.static_x.copy_(images, non_blocking=True)
        return self.step_device()


class TeacherWorker:
    """Teacher side of the link: wait for the student's batch, forward, ship logits."""

    def __init__(self, model, link: DeviceDistillLink, use_graph: bool = True, dtype=torch.bfloat16):
        self.model, self.link, self.dtype = model, link, dtype
        self.use_graph = use_graph
        self.graph = None
        self.device = link.pool.device
        self.steps_done = 0

    def _body(self):
        self.link.seq.add_(1)
        img = self.link.wait_images(0, seq=self.link.seq)
        img = img if img.dtype == self.dtype else img.to(self.dtype)
        if self.link.fused_fc and hasattr(self.model, "forward_features") and self.dtype == torch.bfloat16:
            feats = self.model.forward_features(img)
            self.link.ship_linear(feats, self.model.fc_weight, self.model.fc_bias, 0, seq=self.link.seq)
        else:
            assert not self.link.fused_fc, "link was built with fused_fc=True but the model cannot use it"
            self.link.send_logits(self.model(img).to(torch.bfloat16), 0, seq=self.link.seq)

    def capture(self, warmup: int = 3):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        torch.cuda.synchronize(self.device)

    def step(self):
        if self.use_graph:
            if self.graph is None:
                self.capture()
            self.graph.replay()
        else:
            self._body()
        self.steps_done += 1


def split_roles(world: int):
    """-> (n_students, students, teachers).  world 1 is not a service configuration."""
    assert world >= 2 and world % 2 == 0, "the distill service needs an even number (>= 2) of GPUs"
    n = world // 2
    return n, list(range(n)), list(range(n, world))
