"""Device-resident distillation link between a student rank and a teacher rank on the same NVSwitch
domain -- the B200 replacement of the reference's DistillReader -> Paddle Serving round trip when
teacher and student share a box (BASELINE.json config 2: "teacher on GPUs 4-7,
student on 0-3, logit ship over NVSwitch").

Data path per step ``n`` (ring slot ``n % slots``), no host involvement, no NCCL, no RPC:

  student : ``send_images(x)``     peer_ship: x -> TEACHER's HBM slot, release-flag img_ready = n+1
  teacher : ``wait_images()``      device-side acquire of img_ready, returns the slot view
            forward(...)           (ResNeXt101_32x16d)
            ``ship_linear(f,W,b)`` fused path (default): the classifier GEMM's epilogue TMA-stores its
                                   logit tiles straight into the STUDENT's HBM slot and the last CTA
                                   releases logit_ready = n+1 (csrc/gemm.cu, "GEMM -> peer ship")
            ``send_logits(z)``     unfused path: logit_ship copies z + per-row softmax stats
  student : ``loss(logits)``       soft_ce_recv: acquires logit_ready, fused soft-label CE on the slot;
                                   backward = soft_ce_bwd on the same slot

Slot reuse needs no extra acknowledgement as long as the student sends the images of step ``n`` only
after it has consumed the loss of step ``n - slots`` (program order in ``DistillStudentStep``).

All kernels take their sequence number from a device counter, so both sides are CUDA-graph capturable.
"""
from __future__ import annotations

from typing import Optional

import torch

from ..parallel.symm import SymmetricPool


class _Regions:
    """Identical carve-up of every rank's slab (symmetric allocation)."""

    def __init__(self, pool: SymmetricPool, batch: int, image_shape, num_classes: int, slots: int, dtype):
        self.slots, self.batch, self.num_classes = slots, batch, num_classes
        c, h, w = image_shape
        self.img = [pool.alloc(batch * c * h * w, dtype) for _ in range(slots)]
        self.logit = [pool.alloc(batch * num_classes, torch.bfloat16) for _ in range(slots)]
        self.stats = [pool.alloc(batch * 2, torch.float32) for _ in range(slots)]
        self.flags = pool.alloc(64, torch.int32)   # [0:slots] img_ready, [16:16+slots] logit_ready
        self.image_shape = (c, h, w)

    def img_flag_ptr(self, rank, slot):
        return self.flags.data_ptrs[rank] + 4 * slot

    def logit_flag_ptr(self, rank, slot):
        return self.flags.data_ptrs[rank] + 4 * (16 + slot)


class DeviceDistillLink:
    def __init__(self, pool: SymmetricPool, peer_rank: int, role: str, batch: int, image_shape=(3, 224, 224),
                 num_classes: int = 1000, slots: int = 2, dtype=torch.bfloat16, temperature: float = 1.0,
                 timeout_s: float = 60.0, fused_fc: bool = True):
        assert role in ("student", "teacher")
        self.pool, self.peer, self.role = pool, peer_rank, role
        self.rank = pool.rank
        self.r = _Regions(pool, batch, image_shape, num_classes, slots, dtype)
        self.slots, self.batch, self.num_classes = slots, batch, num_classes
        self.temperature, self.timeout_s = temperature, timeout_s
        # fused GEMM->ship leaves the row statistics to the receiving loss kernel; both ends of a link
        # must agree (same constructor arguments on both ranks).  TMA needs a 16-byte row pitch.
        self.fused_fc = fused_fc and num_classes % 8 == 0
        dev = pool.device
        self.seq = torch.zeros(1, dtype=torch.int32, device=dev)      # device step counter (graph-safe)
        self.done = torch.zeros(4, dtype=torch.int32, device=dev)     # per-kernel block counters
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self._step = 0   # host mirror: selects the slot (slot choice must be static per captured graph)

    # ------------------------------------------------------------------ student side
    def send_images(self, x: torch.Tensor, slot: int, seq: Optional[torch.Tensor] = None, seq_imm: int = 0):
        from ..ops import native, count_launch

        assert self.role == "student"
        if x.dim() == 4:      # ship the NHWC bytes; the flat view is what the copy kernel wants
            x = x.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)
        x = x.contiguous().view(-1)
        native().peer_ship(x, self.r.img[slot].data_ptrs[self.peer], x.numel() * x.element_size(),
                           self.r.img_flag_ptr(self.peer, slot), seq, seq_imm, self.done[0:1])
        count_launch()

    def loss(self, logits: torch.Tensor, slot: int, seq: Optional[torch.Tensor] = None, seq_imm: int = 0,
             student_temperature: float = 1.0, kl: bool = False, loss_scale: float = 1.0):
        assert self.role == "student"
        return _RecvLossFn.apply(logits, self, slot, seq, seq_imm, student_temperature, kl, loss_scale)

    # ------------------------------------------------------------------ teacher side
    def wait_images(self, slot: int, seq: Optional[torch.Tensor] = None, seq_imm: int = 0) -> torch.Tensor:
        from ..ops import native, count_launch

        assert self.role == "teacher"
        flag = self.r.flags.tensor[slot:slot + 1]
        native().wait_flag_async(flag, seq, seq_imm, self.timeout_s, self.err)
        count_launch()
        c, h, w = self.r.image_shape
        t = self.r.img[slot].tensor.view(self.batch, h, w, c).permute(0, 3, 1, 2)   # NHWC memory, NCHW view
        return t

    def send_logits(self, logits: torch.Tensor, slot: int, seq: Optional[torch.Tensor] = None, seq_imm: int = 0):
        from ..ops import native, count_launch

        assert self.role == "teacher"
        native().logit_ship(logits.contiguous(), self.r.logit[slot].data_ptrs[self.peer],
                            self.r.stats[slot].data_ptrs[self.peer], self.temperature,
                            self.r.logit_flag_ptr(self.peer, slot), seq, seq_imm, self.done[1:2])
        count_launch()

    def ship_linear(self, feats: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], slot: int,
                    seq: Optional[torch.Tensor] = None, seq_imm: int = 0):
        """logits = feats @ weight.T + bias, written by the GEMM epilogue directly into the student's
        slot (one kernel: tcgen05 GEMM + NVLink store + flag release)."""
        from ..ops import native, count_launch

        assert self.role == "teacher" and self.fused_fc
        feats = feats.contiguous()
        assert feats.shape[0] == self.batch and weight.shape[0] == self.num_classes
        native().gemm_bf16_ship(feats, weight, self.r.logit[slot].data_ptrs[self.peer], self.num_classes, bias,
                                self.r.logit_flag_ptr(self.peer, slot), seq, seq_imm, self.done[2:3])
        count_launch()

    def check_error(self) -> int:
        return int(self.err.item())


class _RecvLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, link, slot, seq, seq_imm, s_temp, kl, loss_scale):
        from ..ops import native, count_launch

        logits = logits.contiguous()
        n, c = logits.shape
        loss = torch.zeros((), device=logits.device, dtype=torch.float32)
        row_stats = torch.empty(n, 4, device=logits.device, dtype=torch.float32)
        slot_t = link.r.logit[slot].tensor.view(n, c)
        stats_t = None if link.fused_fc else link.r.stats[slot].tensor.view(n, 2)
        flag = link.r.flags.tensor[16 + slot:17 + slot]
        native().soft_ce_recv(logits, slot_t, stats_t, flag, seq, seq_imm, loss, row_stats, s_temp,
                              link.temperature, kl, loss_scale, link.timeout_s, link.err)
        count_launch()
        ctx.save_for_backward(logits, slot_t, row_stats)
        ctx.cfg = (s_temp, link.temperature, loss_scale)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        from ..ops import native, count_launch

        logits, slot_t, row_stats = ctx.saved_tensors
        s_temp, t_temp, loss_scale = ctx.cfg
        dlogits = torch.empty_like(logits)
        native().soft_ce_bwd(logits, slot_t, None, row_stats, grad_out.float().contiguous(), dlogits, 1, s_temp,
                             t_temp, 0.0, loss_scale)
        count_launch()
        return dlogits, None, None, None, None, None, None, None


def pool_bytes_needed(batch: int, image_shape=(3, 224, 224), num_classes: int = 1000, slots: int = 2,
                      dtype=torch.bfloat16) -> int:
    c, h, w = image_shape
    esz = torch.empty((), dtype=dtype).element_size()
    per = batch * c * h * w * esz + batch * num_classes * 2 + batch * 8 + 1024
    return slots * per + (1 << 20)
