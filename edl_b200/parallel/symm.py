"""Symmetric (peer-mapped) memory for one elastic *stage*, bootstrapped through a key-value store.

Every rank creates one physical slab with the CUDA virtual-memory-management API (``csrc/vmm.cpp``:
``cuMemCreate`` + POSIX file-descriptor export), publishes the name of its handle server in the
rendezvous store, maps every peer's slab and -- when the NVSwitch supports it -- binds all slabs to
one multicast object whose alias address is the target of the ``multimem.*`` instructions
(NVLS in-switch reduction / broadcast).  The slab is carved into a signal pad (barrier flags / epochs /
scalar scratch / error word, see csrc/comm.cuh) and payload windows (gradient buckets, parameter
shadows, logit-ship rings).  On an elastic stage change the whole ``SymmetricPool`` is dropped and
re-created for the new membership.

No NCCL communicator and no ``torch.distributed._symmetric_memory`` is involved: the reference restarts
every trainer and re-bootstraps NCCL over TCP on each stage change (python/edl/utils/train_process.py:
37-41,55); here survivors only swap a pointer table.  The store is anything with the
``torch.distributed.Store`` calls ``set`` / ``get``: the job's own KV store (``elastic.KVRendezvousStore``,
in-place elastic mode -- no process group needed at all) or the store behind an existing process
group (torchrun).  Every byte that moves over NVLink is moved by our own kernels
(csrc/allreduce.cu, csrc/logit_ship.cu).
"""
from __future__ import annotations

import hashlib
import os
import time
import uuid
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


@dataclass
class Fabric:
    """Membership of one stage without a process group: who am I, how many are we, and the store we meet in."""
    store: object                  # set(key, bytes) / get(key) -> bytes (blocking)
    rank: int
    world: int
    tag: str                       # unique per stage (key prefix)


_GROUP_SEQ = {}


def _group_fabric(group) -> Fabric:
    """Fabric description of a torch process group: its ranks meet in the default store under a prefix made of
    the member list and a per-group sequence number (pools are created in the same order by every member)."""
    from torch.distributed import distributed_c10d as c10d

    g = group if group is not None else dist.group.WORLD
    rank, world = dist.get_rank(g), dist.get_world_size(g)
    members = [dist.get_global_rank(g, r) for r in range(world)] if group is not None else list(range(world))
    key = hashlib.sha1((",".join(map(str, members))).encode()).hexdigest()[:12]
    store = c10d._get_default_store()
    gen = (id(store), key)
    seq = _GROUP_SEQ.get(gen, 0)
    _GROUP_SEQ[gen] = seq + 1
    return Fabric(store=store, rank=rank, world=world, tag="pg-%s-%d" % (key, seq))


class SymmetricPool:
    """One peer-mapped slab per rank; bump allocation of windows inside it."""

    def __init__(self, nbytes: int, group: Optional[dist.ProcessGroup] = None,
                 device: Optional[torch.device] = None, channels: int = 2, fabric: Optional[Fabric] = None,
                 timeout_s: float = 120.0):
        from ..ops import native

        C = native()
        err = C.vmm_driver_error()
        if err:
            raise RuntimeError("symmetric memory needs the CUDA driver: " + err)
        self.fabric = fabric if fabric is not None else _group_fabric(group)
        self.group = group
        self.rank, self.world = self.fabric.rank, self.fabric.world
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.channels = channels
        self.timeout_s = timeout_s
        self._sig_total = SIG_BYTES * channels
        self._p = "edl_symm/%s/" % self.fabric.tag
        self._bar = 0
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)                        # the primary context exists before the driver calls
        t0 = time.time()
        total = self._sig_total + _round_up(nbytes, 16384)
        name = "edl-symm-%s-%d" % (uuid.uuid4().hex[:16], self.rank)
        self._slab = C.SymmSlab(self.device.index, total, self.world, self.rank, name)
        self._set("srv/%d" % self.rank, name)
        names = [name if r == self.rank else self._get("srv/%d" % r) for r in range(self.world)]
        for r in range(self.world):
            if r != self.rank:
                self._slab.map_peer(r, names[r])
        self.mc_base = self._setup_multicast(names[0]) if self.world > 1 else 0
        self.slab = self._slab.tensor()
        self.total = int(self.slab.numel())
        self.slab.zero_()
        torch.cuda.synchronize(self.device)
        self.barrier()                                            # everybody zeroed its flags before anyone signals
        self._slab.stop_server()
        self.base_ptrs = [int(p) for p in self._slab.ptrs()]
        self._off = self._sig_total
        self.setup_s = time.time() - t0

    # ------------------------------------------------------------------ store helpers
    def _set(self, key: str, value):
        self.fabric.store.set(self._p + key, value if isinstance(value, (bytes, bytearray)) else str(value).encode())

    def _get(self, key: str) -> str:
        deadline = time.time() + self.timeout_s
        while True:
            try:
                v = self.fabric.store.get(self._p + key)
                return v.decode() if isinstance(v, (bytes, bytearray)) else str(v)
            except Exception:                                     # noqa: BLE001 - a store's own (shorter) timeout
                if time.time() > deadline:
                    raise RuntimeError("symmetric pool %s: rank %d timed out waiting for %r" % (
                        self.fabric.tag, self.rank, key))

    def _all_ok(self, what: str, ok: bool) -> bool:
        """Store barrier that also carries one bit: True iff every rank reports ``ok``."""
        self._set("%s/%d" % (what, self.rank), "1" if ok else "0")
        return all(self._get("%s/%d" % (what, r)) == "1" for r in range(self.world))

    def barrier(self):
        """Host-side barrier of the pool's members through the store (no collective library involved)."""
        self._bar += 1
        self._all_ok("bar%d" % self._bar, True)

    # ------------------------------------------------------------------ NVLS alias
    def _setup_multicast(self, root_name: str) -> int:
        # In-place elastic mode keeps the NVLS alias OFF unless asked for (EDL_INPLACE_MULTICAST=1): when a member of
        # a multicast team dies (SIGKILL), a surviving GPU that still issues multimem operations takes a CONTAINED
        # NVLink error that poisons its CUDA context -- measured on 2 x B200, profiles/elastic_launch_gpu.json -- while
        # plain peer loads / stores of the dead rank's (still referenced) memory merely make our barriers time out,
        # which is the event hot recovery is built on.
        inplace = os.environ.get("EDL_RESCALE_MODE", "").lower() == "inplace"
        want = (self._slab.mc_supported() and os.environ.get("EDL_DISABLE_MULTICAST", "0") != "1"
                and (not inplace or os.environ.get("EDL_INPLACE_MULTICAST", "0") == "1"))
        if not self._all_ok("mc_want", want):
            return 0
        ok = True
        try:
            if self.rank == 0:
                self._slab.mc_create()
        except RuntimeError:
            ok = False
        if not self._all_ok("mc_created", ok):
            return 0
        try:
            if self.rank != 0:
                self._slab.mc_import(root_name)
            self._slab.mc_add_device()
        except RuntimeError:
            ok = False
        if not self._all_ok("mc_added", ok):                      # every device is in the team before the first bind
            return 0
        try:
            self._slab.mc_bind()
        except RuntimeError:
            ok = False
        if not self._all_ok("mc_bound", ok):
            return 0
        return int(self._slab.mc_ptr())

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

    def describe(self) -> dict:
        return {"bootstrap": "cuMem VMM + POSIX fd over unix socket, rendezvous through %s" % type(self.fabric.store).__name__,
                "multicast": self.has_multicast, "world": self.world, "slab_bytes": self.total,
                "setup_s": round(self.setup_s, 4)}


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def symmetric_memory_supported() -> bool:
    if not torch.cuda.is_available():
        return False
    try:
        from ..ops import native
        return not native().vmm_driver_error()
    except Exception:                                             # noqa: BLE001
        return False
