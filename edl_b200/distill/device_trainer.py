"""Same-box distillation service: student ranks train ResNet50_vd while teacher ranks run
ResNeXt101_32x16d and feed soft labels through the NVSwitch-direct link (``device_feed.py``).

BASELINE.json config 2 ("student on GPUs 0-3, teacher on GPUs 4-7, logit ship over NVSwitch"); the
reference's equivalent is DistillReader + 40 Paddle-Serving P4 teachers (README.md:85, 1514 img/s).
Each student rank ``s`` is paired with teacher rank ``n_students + s``.
"""
from __future__ import annotations

import torch

from .. import ops
from ..trainer import StudentTrainer
from .device_feed import DeviceDistillLink


_EAGER_STEPS = 4     # protocol steps run eagerly on both sides before the per-slot CUDA graphs are captured


class DistillStudentTrainer(StudentTrainer):
    """StudentTrainer whose targets arrive from the paired teacher GPU instead of the host.

    Software-pipelined by one batch: protocol step ``k`` ships batch ``k`` to the teacher (ring slot
    ``k % 2``) and trains on batch ``k - 1``, whose logits the teacher produced while the student was busy
    with step ``k - 1``.  Student and teacher therefore overlap completely and the pair runs at
    max(student step, teacher forward) instead of their sum.  Step 0 only ships.  One CUDA graph per slot."""

    def __init__(self, model, batch_size, link: DeviceDistillLink, **kw):
        kw.setdefault("target_kind", "logits")
        super().__init__(model, batch_size, **kw)
        assert link.slots >= 2, "the pipelined student needs a 2-slot link"
        self.link = link
        self.static_xs = [self.static_x, torch.zeros_like(self.static_x)]
        self.seq_prev = torch.zeros_like(link.seq)           # sequence number of the batch being trained on
        self.k = 0
        self.graphs = [None, None]
        self.launches_per_slot = [0, 0]
        # eager warm-up steps and the captures share ONE stream: autograd binds every parameter's gradient
        # accumulator to the stream it first ran on, and a capture may not depend on work of another stream
        self._stream = self._cap_stream                     # the module's step stream (see StudentTrainer)

    def _body(self, p: int, train: bool):
        self.seq_prev.copy_(self.link.seq)
        self.link.seq.add_(1)                                # device-side counters (graph replay safe)
        self.dp.zero_grad()
        if self.arena is not None:
            self.arena.zero()
        self.link.send_images(self.static_xs[p], p, seq=self.link.seq)      # batch k -> teacher HBM
        if not train:
            return
        x = self.static_xs[1 - p]
        logits = self.model(x if x.dtype == self.dtype else x.to(self.dtype))
        loss = self.link.loss(logits, 1 - p, seq=self.seq_prev)  # acquires the teacher's logits of batch k-1
        loss.backward()
        self.dp.finish()
        self.opt.step()
        self.static_loss.copy_(loss.detach())

    def step_device(self):
        p = self.k % 2
        if not self.use_graph or self.k < _EAGER_STEPS:
            cur = torch.cuda.current_stream(self.device)
            self._stream.wait_stream(cur)                    # the H2D copy of step() ran on the caller's stream
            with torch.cuda.stream(self._stream):
                self._body(p, train=self.k > 0)
            cur.wait_stream(self._stream)
        else:
            if self.graphs[p] is None:
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                before = ops.launches()
                with torch.cuda.graph(g, stream=self._stream):
                    self._body(p, train=True)
                self.launches_per_slot[p] = ops.launches() - before
                self.graphs[p] = g
            self.graphs[p].replay()
            ops.count_launch(self.launches_per_slot[p])
        self.k += 1
        self.steps_done += 1
        return self.static_loss

    def step(self, images, targets=None):
        self.static_xs[self.k % 2].copy_(images, non_blocking=True)
        return self.step_device()


class TeacherWorker:
    """Teacher side of the link: wait for the student's batch in slot ``k % 2``, forward, ship the logits
    (classifier GEMM epilogue -> student's HBM).  Same eager-then-graph schedule as the student."""

    def __init__(self, model, link: DeviceDistillLink, use_graph: bool = True, dtype=torch.bfloat16):
        self.model, self.link, self.dtype = model, link, dtype
        self.use_graph = use_graph
        self.graphs = [None, None]
        self.device = link.pool.device
        self.steps_done = 0

    def _body(self, slot: int):
        self.link.seq.add_(1)
        img = self.link.wait_images(slot, seq=self.link.seq)
        img = img if img.dtype == self.dtype else img.to(self.dtype)
        if self.link.fused_fc and hasattr(self.model, "forward_features") and self.dtype == torch.bfloat16:
            feats = self.model.forward_features(img)
            self.link.ship_linear(feats, self.model.fc_weight, self.model.fc_bias, slot, seq=self.link.seq)
        else:
            assert not self.link.fused_fc, "link was built with fused_fc=True but the model cannot use it"
            self.link.send_logits(self.model(img).to(torch.bfloat16), slot, seq=self.link.seq)

    def step(self):
        slot = self.steps_done % 2
        if not self.use_graph or self.steps_done < _EAGER_STEPS:
            self._body(slot)
        else:
            if self.graphs[slot] is None:
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._body(slot)
                self.graphs[slot] = g
            self.graphs[slot].replay()
        self.steps_done += 1


def split_roles(world: int):
    """-> (n_students, students, teachers).  world 1 is not a service configuration."""
    assert world >= 2 and world % 2 == 0, "the distill service needs an even number (>= 2) of GPUs"
    n = world // 2
    return n, list(range(n)), list(range(n, world))
