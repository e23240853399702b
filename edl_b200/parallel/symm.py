"""Symmetric (peer-mapped) memory for one elastic *stage*.

Every rank allocates one slab, the slab handles are exchanged through the process group's store
and mapped into every peer (CUDA VMM; NVLS multicast alias when the fabric supports it).  The slab
is carved into a signal pad (barrier flags / epochs / scalar scratch, see csrc/comm.cuh) and
payload sub-buffers (gradient buckets, logit-ship rings).  On an elastic stage change the whole
``SymmetricPool`` is dropped and re-created for the new group -- the reference instead restarts
every trainer process and re-bootstraps NCCL over TCP (utils/train_process.py:37-41,55).

``torch.distributed._symmetric_memory`` provides allocation + handle exchange (plumbing); every
byte that moves over NVLink is moved by our own kernels (csrc/allreduce.cu, csrc/logit_ship.cu).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist

SIG_BYTES = 16384  # >= comm_sig_words()*4, rounded up; payload stays 16 KiB aligned


@dataclass
class SymmSlice:
    """A typed window [offset, offset+nbytes) of the slab, with the peer/multicast addresses."""
    tensor: torch.Tensor           # local view
    data_ptrs: List[int]           # address of this window in every rank's slab (rank order)
    sig_ptrs: List[int]            # signal pad address in every rank's slab
    mc_ptr: int                    # multicast alias of this window, 0 if unavailable
    rank: int
    world: int


class SymmetricPool:
    """One peer-mapped slab per rank; bump allocation of windows inside it."""

    def __init__(self, nbytes: int, group: Optional[dist.ProcessGroup] = None,
                 device: Optional[torch.device] = None, channels: int = 2):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.channels = channels
        self._sig_total = SIG_BYTES * channels
        total = self._sig_total + _round_up(nbytes, 16384)
        self.slab = symm_mem.empty(total, dtype=torch.uint8, device=self.device)
        self.handle = symm_mem.rendezvous(self.slab, self.group)
        self.slab.zero_()
        self.base_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        mc = 0
        try:
            if os.environ.get("EDL_DISABLE_MULTICAST", "0") != "1" and self.handle.multicast_ptr:
                mc = int(self.handle.multicast_ptr)
        except Exception:  # pragma: no cover - older handle objects
            mc = 0
        self.mc_base = mc
        self._off = self._sig_total
        self.total = total
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)

    @property
    def has_multicast(self) -> bool:
        return self.mc_base != 0

    def sig_ptrs(self, channel: int = 0) -> List[int]:
        assert 0 <= channel < self.channels
        return [b + channel * SIG_BYTES for b in self.base_ptrs]

    def sig_tensor(self, channel: int = 0) -> torch.Tensor:
        return self.slab[channel * SIG_BYTES:(channel + 1) * SIG_BYTES].view(torch.int32)

    def alloc(self, numel: int, dtype: torch.dtype, channel: int = 0) -> SymmSlice:
        esz = torch.empty((), dtype=dtype).element_size()
        nbytes = _round_up(numel * esz, 256)
        if self._off + nbytes > self.total:
            raise MemoryError("symmetric pool exhausted: need %d more bytes" % (self._off + nbytes - self.total))
        off = self._off
        self._off += nbytes
        t = self.slab[off:off + numel * esz].view(dtype)
        return SymmSlice(tensor=t, data_ptrs=[b + off for b in self.base_ptrs],
                         sig_ptrs=self.sig_ptrs(channel), mc_ptr=(self.mc_base + off) if self.mc_base else 0,
                         rank=self.rank, world=self.world)

    def check_error(self, channel: int = 0) -> int:
        """Returns 0, or 1+peer if a barrier timed out waiting for ``peer`` (host sync)."""
        from ..ops import native

        w = native().comm_error_word_offset()
        return int(self.sig_tensor(channel)[w].item())


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def symmetric_memory_supported() -> bool:
    if not torch.cuda.is_available():
        return False
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
        return True
    except Exception:
        return False
