"""ElasticDataParallel -- the B200-native replacement for Paddle fleet collective DP.

Reference behaviour being replaced: ``fleet.distributed_optimizer(opt, strategy).minimize(loss)``
inserts fused (<=16 MB) NCCL all-reduces after backward (example/distill/resnet/
train_with_fleet.py:332-333,353-364; scripts/train_gpu.sh:67-70) and every membership change
restarts all trainer processes (utils/launcher.py:221-244).

Here:
* gradients live in ONE flat buffer per dtype inside NVSwitch-symmetric memory (``FlatParams``);
  buckets are zero-copy windows of it, planned from (world size, bucket cap) -- re-planned by
  ``rebuild()`` on every elastic stage change without touching model or optimizer state;
* as soon as autograd (or a fused op writing straight into its gradient sink) has produced every
  gradient of a bucket, the bucket's fused all-reduce kernel (csrc/allreduce.cu: P2P two-shot /
  NVLS multimem, fp32 accumulate, 1/world scale, finite check, squared-norm) is enqueued on a side
  stream so it overlaps the rest of backward; the whole thing is CUDA-graph capturable;
* CPU / gloo groups (the fit_a_line plumbing config) fall back to ``dist.all_reduce``.
"""
from __future__ import annotations

import os
import weakref
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import torch
import torch.distributed as dist

from .flat import FlatParams


def _host_id() -> str:
    """Identity of the NVSwitch domain a rank lives in: host name + boot id (containers of one machine share both)."""
    import socket

    boot = ""
    try:
        with open("/proc/sys/kernel/random/boot_id") as f:
            boot = f.read().strip()
    except OSError:
        pass
    return socket.gethostname() + ":" + boot


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


@dataclass
class Bucket:
    dtype: torch.dtype
    start: int                 # element offset inside the group's flat gradient
    numel: int
    entry_ids: List[int] = field(default_factory=list)
    order: int = 0             # launch order (position of the last-ready parameter)
    pending: int = 0
    launched: bool = False
    algo: str = "twoshot"


def plan_buckets(flat: FlatParams, cap_bytes: int) -> List[Bucket]:
    """Cut each dtype group's flat gradient into contiguous buckets of ~cap_bytes.

    Flat layout is reverse registration order, so entry 0 of a group is the *last* layer: buckets are
    produced in the order their gradients become ready."""
    buckets: List[Bucket] = []
    eid = 0
    # global readiness order = interleaving of the dtype groups in reverse-registration order
    order_of: Dict[int, int] = {i: e.order for i, (_, e) in enumerate(flat.entries())}
    for g in flat.groups.values():
        esz = g.grad.element_size()
        cur: Optional[Bucket] = None
        for e in g.entries:
            if cur is None:
                cur = Bucket(dtype=g.dtype, start=e.offset, numel=0)
            cur.entry_ids.append(eid)
            cur.numel = e.offset + e.numel - cur.start
            cur.order = max(cur.order, order_of[eid])
            eid += 1
            if cur.numel * esz >= cap_bytes:
                buckets.append(cur)
                cur = None
        if cur is not None:
            buckets.append(cur)
        # the tail padding of the group rides with the last bucket so the buffer is fully covered
        last = [b for b in buckets if b.dtype == g.dtype][-1]
        last.numel = g.numel - last.start
    buckets.sort(key=lambda b: b.order)
    return buckets


def choose_algo(nbytes: int, world: int, has_multicast: bool, prefer: str = "auto") -> str:
    """Planner rule.  NVSwitch gives every peer full bandwidth, so the crossover is about launch /
    barrier latency, not link count: NVLS (in-switch reduce, half the NVLink traffic) whenever the
    multicast alias exists, P2P two-shot otherwise; tiny out-of-place reductions use one-shot."""
    if prefer in ("twoshot", "multimem", "nccl"):
        return prefer if (prefer != "multimem" or has_multicast) else "twoshot"
    if world <= 1:
        return "none"
    if has_multicast and nbytes >= 64 * 1024:
        return "multimem"
    return "twoshot"


class ElasticDataParallel:
    def __init__(self, module: torch.nn.Module, group: Optional[dist.ProcessGroup] = None,
                 bucket_cap_mb: float = 16.0, overlap: bool = True, comm_blocks: int = 32,
                 algo: str = "auto", timeout_s: float = 60.0, average: bool = True,
                 check_finite: bool = False, track_sqnorm: bool = False, hierarchical: str = "auto"):
        # one engine per module: a second engine on the same parameters would leave the first one's
        # autograd hooks installed (they would fire, and launch reductions, during the new engine's backward)
        old = getattr(module, "_edl_dp_engine", None)
        if old is not None and old() is not None:
            old().detach()
        module._edl_dp_engine = weakref.ref(self)
        self.module = module
        self.bucket_cap = int(bucket_cap_mb * (1 << 20))
        self.overlap = overlap
        self.comm_blocks = comm_blocks
        self.algo_pref = os.environ.get("EDL_ALLREDUCE_ALGO", algo)
        self.timeout_s = timeout_s
        self.average = average
        self.device = next(module.parameters()).device
        self.pool = None
        self.slices = {}
        self.found_inf = None
        self.sqnorm = None
        if check_finite:
            self.found_inf = torch.zeros(1, dtype=torch.int32, device=self.device)
        if track_sqnorm and self.device.type == "cuda":
            self.sqnorm = torch.zeros(1, dtype=torch.float32, device=self.device)
        # "auto": two-level reduction as soon as the ranks span more than one host (NVSwitch domain); "off": never
        # (one flat group; across hosts that means the library path); the reference's use_hierarchical_allreduce knob
        self.hierarchical = os.environ.get("EDL_HIERARCHICAL_ALLREDUCE", hierarchical)
        self.overlap_wgrad = os.environ.get("EDL_OVERLAP_WGRAD", "1") == "1"
        self.enabled = True     # False: gradients stay local (DGC exchanges them itself)
        self._bind_group(group)
        self.flat = FlatParams(module, grad_alloc=self._grad_alloc if self.pool is not None else None)
        self._plan()
        self._install_hooks()
        self.comm_launches = 0

    # ------------------------------------------------------------------ group / memory
    def _bind_group(self, group):
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        self.backend = dist.get_backend(group) if self.world > 1 else "none"
        self._plan_hierarchy(group)
        symm_world = self.local_world if self.hier else self.world
        self.use_symm = (symm_world > 1 and self.device.type == "cuda"
                         and self.algo_pref != "nccl")
        self.pool = None
        if self.use_symm:
            from .symm import SymmetricPool

            need = 0
            for p in self.module.parameters():
                if p.requires_grad:
                    need += (p.numel() + 256) * p.element_size()
            need += 4 << 20
            self.pool = SymmetricPool(need, group=self.local_group if self.hier else group, device=self.device)
        if self.device.type == "cuda" and getattr(self, "comm_stream", None) is None:
            # all-reduce kernels: high priority (few CTAs, on the critical path of the optimizer step);
            # weight-gradient kernels: LOW priority -- they only have to finish before their bucket is
            # reduced, and must not take SMs from the dgrad / BN-backward chain of the main stream
            self.comm_stream = torch.cuda.Stream(device=self.device, priority=-1)
            self.wgrad_stream = torch.cuda.Stream(device=self.device, priority=0)
            self._main_stream = None
            self._comm_used = False

    def _plan_hierarchy(self, group):
        """Peer memory only reaches the GPUs of one NVSwitch domain.  When the ranks span several hosts the
        reduction becomes two-level (the reference's ``use_hierarchical_allreduce``, example/distill/resnet/
        train_with_fleet.py:90-91,360-362): our kernels inside a host, the library collective across hosts on 1/L of
        every bucket per rank.  Hosts must hold the same number of ranks; otherwise one flat library group."""
        self.hier, self.local_group, self.cross_group = False, group, None
        self.local_world, self.local_rank = self.world, self.rank
        if self.world <= 1 or self.hierarchical == "off":
            return
        fake = os.environ.get("EDL_FAKE_HOST")
        if not fake and os.environ.get("LOCAL_WORLD_SIZE") == str(self.world) and group is None:
            return                       # torchrun started every rank of the job on this host: nothing to find out
        me = fake or _host_id()
        hosts = [None] * self.world
        dist.all_gather_object(hosts, me, group=group)
        if len(set(hosts)) <= 1:
            return
        nodes: Dict[str, List[int]] = {}
        for r, h in enumerate(hosts):
            nodes.setdefault(h, []).append(r)
        sizes = {len(v) for v in nodes.values()}
        if len(sizes) != 1:
            self.algo_pref = "nccl"          # uneven hosts: flat library all-reduce
            return
        to_global = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
        L = sizes.pop()
        node_lists = list(nodes.values())
        # every member creates every subgroup in the same order (local synchronisation: ranks outside `group`,
        # e.g. the teachers of a distill job, do not take part)
        for members in node_lists:
            g = dist.new_group(ranks=[to_global(r) for r in members], use_local_synchronization=True) \
                if self.rank in members else None
            if g is not None:
                self.local_group, self.local_rank = g, members.index(self.rank)
        for l in range(L):
            members = [n[l] for n in node_lists]
            if self.rank in members:
                self.cross_group = dist.new_group(ranks=[to_global(r) for r in members], use_local_synchronization=True)
        self.local_world, self.hier = L, True

    def _grad_alloc(self, numel, dtype, device):
        sl = self.pool.alloc(numel, dtype)
        self.slices[dtype] = sl
        return sl.tensor

    def _plan(self):
        self.entries = list(self.flat.entries())
        self.buckets = plan_buckets(self.flat, self.bucket_cap)
        has_mc = self.pool.has_multicast if self.pool is not None else False
        for b in self.buckets:
            esz = 2 if b.dtype in (torch.bfloat16, torch.float16) else 4
            b.algo = choose_algo(b.numel * esz, self.local_world if self.hier else self.world, has_mc, self.algo_pref) \
                if self.use_symm else ("nccl" if self.world > 1 else "none")
        self.bucket_of = {}
        for bi, b in enumerate(self.buckets):
            for eid in b.entry_ids:
                self.bucket_of[eid] = bi
        self._reset_pending()

    def _reset_pending(self):
        for b in self.buckets:
            b.pending = len(b.entry_ids)
            b.launched = False
        self._next = 0
        self._works = []
        self._seen = [False] * len(self.entries)

    def _install_hooks(self):
        self._hook_handles = []
        for eid, (g, e) in enumerate(self.entries):
            cb = self._make_ready(eid)
            e.param._edl_grad_ready = cb
            self._hook_handles.append(
                e.param.register_post_accumulate_grad_hook(lambda p, _cb=cb: _cb()))

    def detach(self):
        """Remove this engine's autograd hooks (another engine is taking over the module)."""
        for h in getattr(self, "_hook_handles", []):
            h.remove()
        self._hook_handles = []

    def _make_ready(self, eid) -> Callable[[], None]:
        def ready():
            # idempotent per step: a parameter may be reported both by a fused op writing into its
            # gradient sink and by autograd's post-accumulate hook
            if self._seen[eid]:
                return
            self._seen[eid] = True
            b = self.buckets[self.bucket_of[eid]]
            b.pending -= 1
            if b.pending == 0 and self.overlap:
                self._launch_ready()
        return ready

    # ------------------------------------------------------------------ launching
    def _launch_ready(self):
        while self._next < len(self.buckets) and self.buckets[self._next].pending <= 0:
            self._launch(self.buckets[self._next])
            self._next += 1

    def _launch(self, b: Bucket):
        if b.launched or self.world <= 1 or not self.enabled:
            b.launched = True
            return
        b.launched = True
        g = self.flat.groups[b.dtype]
        scale = 1.0 / self.world if self.average else 1.0
        if self.hier:
            return self._launch_hier(b, g, scale)
        if b.algo in ("twoshot", "multimem"):
            from ..ops import native, count_launch

            sl = self.slices[b.dtype]
            esz = g.grad.element_size()
            off = b.start * esz
            # the bucket's gradients were written on the main stream (BN, pools, ...) and on the
            # weight-gradient stream: the reduction waits for both
            for st in {torch.cuda.current_stream(self.device), self._main_stream,
                       self.wgrad_stream if self.overlap_wgrad else None}:
                if st is not None and st != self.comm_stream:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    self.comm_stream.wait_event(ev)
            self._comm_used = True
            with torch.cuda.stream(self.comm_stream):
                native().allreduce_twoshot(
                    [p + off for p in sl.data_ptrs], sl.sig_ptrs, (sl.mc_ptr + off) if sl.mc_ptr else 0,
                    self.rank, g.grad, b.numel, scale, self.found_inf, self.sqnorm,
                    b.algo == "multimem", self.comm_blocks, self.timeout_s)
            count_launch()
            self.comm_launches += 1
        else:  # nccl / gloo baseline path
            view = g.grad[b.start:b.start + b.numel]
            if self.device.type == "cuda":
                w = dist.all_reduce(view, group=self.group, async_op=True)
                self._works.append((w, view, scale))
            else:
                dist.all_reduce(view, group=self.group)
                if self.average:
                    view.mul_(scale)

    def _local_allreduce(self, b: Bucket, g, scale: float, final: bool):
        """Sum (x scale) of the bucket over the ranks of this host: our kernel over peer memory, or the library on CPU."""
        view = g.grad[b.start:b.start + b.numel]
        if self.local_world <= 1:
            if scale != 1.0:
                view.mul_(scale)
            return
        if self.use_symm:
            from ..ops import native, count_launch

            sl = self.slices[b.dtype]
            off = b.start * g.grad.element_size()
            native().allreduce_twoshot(
                [p + off for p in sl.data_ptrs], sl.sig_ptrs, (sl.mc_ptr + off) if sl.mc_ptr else 0,
                self.local_rank, g.grad, b.numel, scale, self.found_inf if final else None,
                self.sqnorm if final else None, b.algo == "multimem", self.comm_blocks, self.timeout_s)
            count_launch()
            self.comm_launches += 1
        else:
            dist.all_reduce(view, group=self.local_group)
            if scale != 1.0:
                view.mul_(scale)

    def _launch_hier(self, b: Bucket, g, scale: float):
        """Two-level all-reduce of one bucket: (1) sum inside the host, (2) rank l of every host all-reduces slice l
        across hosts (1/L of the bucket per NIC), (3) the slices are put back together inside the host by a second
        local sum over buffers that are zero outside the owned slice, with the 1/world scale and the finite check."""
        L = self.local_world
        per = -(-b.numel // L)
        per = -(-per // 8) * 8                                  # slices stay 16-byte aligned
        lo = min(b.numel, self.local_rank * per)
        hi = min(b.numel, lo + per)
        view = g.grad[b.start:b.start + b.numel]
        cuda = self.device.type == "cuda"
        if cuda:
            for st in {torch.cuda.current_stream(self.device), self._main_stream,
                       self.wgrad_stream if self.overlap_wgrad else None}:
                if st is not None and st != self.comm_stream:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    self.comm_stream.wait_event(ev)
            self._comm_used = True
        ctx = torch.cuda.stream(self.comm_stream) if cuda else _NullCtx()
        with ctx:
            self._local_allreduce(b, g, 1.0, final=False)
            if hi > lo:
                dist.all_reduce(view[lo:hi], group=self.cross_group)
            if L > 1:
                if lo > 0:
                    view[:lo].zero_()
                if hi < b.numel:
                    view[hi:].zero_()
            self._local_allreduce(b, g, scale, final=True)

    def finish(self):
        """Call after ``backward()``: flush buckets whose parameters received no gradient, then make
        the compute stream wait for the communication stream."""
        for b in self.buckets:
            b.pending = 0
        self._launch_ready()
        if self.device.type == "cuda" and (self.world > 1 or self.overlap_wgrad):
            if self.use_symm or self.overlap_wgrad or self.hier:
                cur = torch.cuda.current_stream(self.device)
                # only streams that took part in this step (a stream without forked work is not part
                # of a CUDA-graph capture and must not be joined into it)
                if self._comm_used:
                    cur.wait_stream(self.comm_stream)
                if self.overlap_wgrad:
                    cur.wait_stream(self.wgrad_stream)
                from ..ops import gemm as _gemm
                _gemm.release_wgrad_keepalive()
            for w, view, scale in self._works:
                w.wait()
                if self.average:
                    view.mul_(scale)
        if self.found_inf is not None and not (self.use_symm and self.world > 1):
            # the fused all-reduce epilogue raises the flag itself; without it (one rank, NCCL / gloo
            # path) the gradients are scanned here.  Summed gradients carry every rank's inf/nan, so
            # all ranks agree without another collective.
            for g in self.flat.groups.values():
                bad = (~torch.isfinite(g.grad).all()).to(torch.int32).view(1)
                self.found_inf.copy_(torch.maximum(self.found_inf, bad))
        self._reset_pending()

    # ------------------------------------------------------------------ user-facing
    def __call__(self, *a, **kw):
        return self.module(*a, **kw)

    def zero_grad(self):
        if self.device.type == "cuda":
            from ..ops import gemm as _gemm
            # the zeroing memsets run on the main stream: the side stream must not start writing
            # gradients before them
            self.flat.zero_grad()
            self._main_stream = torch.cuda.current_stream(self.device)
            self._comm_used = False
            if self.overlap_wgrad:
                self.wgrad_stream.wait_stream(self._main_stream)
                _gemm.set_wgrad_stream(self.wgrad_stream)
            else:
                _gemm.set_wgrad_stream(None)
            if self.found_inf is not None:
                self.found_inf.zero_()
            if self.sqnorm is not None:
                self.sqnorm.zero_()
            return
        self.flat.zero_grad()
        if self.found_inf is not None:
            self.found_inf.zero_()
        if self.sqnorm is not None:
            self.sqnorm.zero_()

    def check_comm_error(self) -> int:
        return self.pool.check_error() if self.pool is not None else 0

    def agree(self, flag: float):
        """(max of ``flag`` over the ranks, comm error word).  The agreement ``ElasticContext.poll`` needs, routed
        through our own scalar all-gather kernel when the ranks share symmetric memory: unlike a library
        all-reduce it has the ``globaltimer`` timeout, so a DEAD peer shows up as a non-zero error word a few
        seconds later instead of a hung process -- the trigger for ``ElasticContext.recover()`` on GPUs."""
        if self.world <= 1:
            return float(flag), 0
        if self.use_symm and self.slices and not self.hier:
            from ..ops import native, count_launch

            sl = max(self.slices.values(), key=lambda x: x.tensor.numel())
            inp = torch.tensor([float(flag), 0.0], device=self.device, dtype=torch.float32)
            out = torch.zeros(self.world * 2, device=self.device, dtype=torch.float32)
            native().comm_allgather_scalars(sl.data_ptrs, sl.sig_ptrs, self.rank, inp, out, min(self.timeout_s, 10.0))
            count_launch()
            err = self.check_comm_error()                     # host sync: the kernel above has finished
            return float(out.view(self.world, 2)[:, 0].max().item()), err
        t = torch.tensor([float(flag)], device=self.device if self.device.type == "cuda" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item()), 0

    def rebuild(self, group: Optional[dist.ProcessGroup]):
        """Elastic stage change: new group => new symmetric slab, new bucket plan.  Parameters,
        master weights and optimizer state stay where they are (no process restart)."""
        for h in self._hook_handles:
            h.remove()
        old_pool = self.pool
        self.slices = {}
        self._bind_group(group)
        if self.pool is not None:
            self.flat.rebind_grads(self._grad_alloc)
        else:
            self.flat.rebind_grads(lambda n, dt, dev: torch.zeros(n, dtype=dt, device=dev))
        del old_pool
        self._plan()
        self._install_hooks()

    @torch.no_grad()
    def broadcast_parameters(self, root: int = 0):
        """Bring every rank's parameters (and fp32 masters) in sync with ``root`` -- used after a
        join so that the newcomer does not need to read the checkpoint from the file system."""
        if self.world <= 1:
            return
        for g in self.flat.groups.values():
            src = g.master if g.master is not None else g.param
            if self.use_symm and not self.hier:
                self._broadcast_tensor(src, root)
                if g.master is not None:
                    g.param.copy_(g.master.to(g.param.dtype))
            else:
                dist.broadcast(src, src=dist.get_global_rank(self.group, root) if self.group else root,
                               group=self.group)
                if g.master is not None:
                    g.param.copy_(g.master.to(g.param.dtype))

    @torch.no_grad()
    def broadcast_tensor(self, t: torch.Tensor, root: int = 0):
        """Broadcast any contiguous device tensor (e.g. optimizer state) from ``root`` of the current group."""
        if self.world <= 1:
            return
        if self.use_symm and not self.hier:
            self._broadcast_tensor(t, root)
        else:
            dist.broadcast(t, src=dist.get_global_rank(self.group, root) if self.group else root, group=self.group)

    def _broadcast_tensor(self, t: torch.Tensor, root: int):
        """Chunked broadcast through the gradient slab (it is idle outside backward)."""
        from ..ops import native, count_launch

        flat = t.view(-1).view(torch.uint8)
        sl = max(self.slices.values(), key=lambda s: s.tensor.numel() * s.tensor.element_size())
        stage = sl.tensor.view(torch.uint8)
        cap = stage.numel() // 16 * 16
        for off in range(0, flat.numel(), cap):
            n = min(cap, flat.numel() - off)
            if self.rank == root:
                stage[:n].copy_(flat[off:off + n])
            native().comm_broadcast(sl.data_ptrs, sl.sig_ptrs, self.rank, root, n, self.comm_blocks,
                                    self.timeout_s, sl.tensor)
            count_launch()
            if self.rank != root:
                flat[off:off + n].copy_(stage[:n])
        sl.tensor.zero_()
