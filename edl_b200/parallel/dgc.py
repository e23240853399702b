"""Deep Gradient Compression on the flat gradient buffers (optional strategy of the reference:
``DGCMomentumOptimizer(rampup_begin_step, rampup_step, sparsity)`` behind ``--use_dgc``,
example/distill/resnet/train_with_fleet.py:96-97,106-122; scripts/train_gpu.sh:59-65).

Before ``rampup_begin_step`` the step is the ordinary dense path (fused all-reduce + fused momentum).
Afterwards each rank keeps two residual buffers per dtype group,

    u <- m * u + (g + wd * w)        momentum correction
    v <- v + u                       local gradient accumulation

ships only the top-k entries of ``|v|`` (index + value, ``k = numel * (1 - sparsity)``) with ONE
all-gather per group, zeroes the shipped positions in ``u`` and ``v`` (momentum factor masking) and
applies the averaged sparse sum to the fp32 master weights.  On an NVSwitch box the dense fused
all-reduce is faster than this for ResNet-sized models -- DGC is here for inventory parity and for
multi-node jobs behind slow fabrics; it is not on the benchmark path.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist

from ..ops.optim import FlatSGDMomentum


class DGCMomentum(FlatSGDMomentum):
    def __init__(self, flat, dp=None, lr=0.1, momentum=0.9, weight_decay=1e-4, rampup_begin_step: int = 0,
                 rampup_step: int = 1, sparsity: Sequence[float] = (0.75, 0.9375, 0.984375, 0.996, 0.999),
                 group: Optional[dist.ProcessGroup] = None):
        super().__init__(flat, lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.dp, self.group = dp, group
        self.rampup_begin_step, self.rampup_step = int(rampup_begin_step), max(1, int(rampup_step))
        self.sparsity = list(sparsity)
        self.t = 0
        self.u = {dt: torch.zeros(g.numel, dtype=torch.float32, device=flat.device) for dt, g in flat.groups.items()}
        self.v = {dt: torch.zeros(g.numel, dtype=torch.float32, device=flat.device) for dt, g in flat.groups.items()}
        self.sent_elems = 0
        if dp is not None and self.rampup_begin_step <= 0:
            dp.enabled = False

    # the warm-up of the sparsity: one entry of ``sparsity`` per rampup_step/len(sparsity) steps
    def current_sparsity(self) -> float:
        k = self.t - self.rampup_begin_step
        per = max(1, self.rampup_step // len(self.sparsity))
        return self.sparsity[min(len(self.sparsity) - 1, max(0, k // per))]

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    @torch.no_grad()
    def step(self):
        if self.t < self.rampup_begin_step:
            super().step()
            self.t += 1
            if self.t >= self.rampup_begin_step and self.dp is not None:
                self.dp.enabled = False          # from now on gradients stay local until compressed
            return
        world = self._world()
        sp = self.current_sparsity()
        lr = self.lr_t
        for dt, g in self.flat.groups.items():
            master = g.master if g.master is not None else g.param
            grad = g.grad.float()
            if self.grad_scale_t is not None:
                grad = grad * self.grad_scale_t
            grad = grad + self.weight_decay * master
            u, v = self.u[dt], self.v[dt]
            u.mul_(self.momentum).add_(grad)
            v.add_(u)
            k = max(1, int(g.numel * (1.0 - sp)))
            _, idx = torch.topk(v.abs(), k, sorted=False)
            val = v[idx]
            v[idx] = 0.0
            u[idx] = 0.0
            if world > 1:
                all_idx = torch.empty(world * k, dtype=idx.dtype, device=idx.device)
                all_val = torch.empty(world * k, dtype=val.dtype, device=val.device)
                dist.all_gather_into_tensor(all_idx, idx, group=self.group)
                dist.all_gather_into_tensor(all_val, val, group=self.group)
            else:
                all_idx, all_val = idx, val
            self.sent_elems += k
            upd = torch.zeros_like(master)
            upd.index_add_(0, all_idx, all_val)
            master.add_(upd * (-lr / world))
            if g.master is not None:
                g.param.copy_(master.to(g.param.dtype))
        self.t += 1

    def state_dict(self):
        sd = super().state_dict()
        sd["dgc"] = {"t": self.t, "u": {str(k): t for k, t in self.u.items()}, "v": {str(k): t for k, t in self.v.items()}}
        return sd

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        d = sd.get("dgc")
        if d:
            self.t = d["t"]
            for k in self.u:
                self.u[k].copy_(d["u"][str(k)])
                self.v[k].copy_(d["v"][str(k)])
            if self.dp is not None:
                self.dp.enabled = self.t < self.rampup_begin_step
