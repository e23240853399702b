"""Top-level trainer-side API (the surface the reference's stale tests document:
``edl.size()``, ``edl.PaddleState(...)``, ``edl.notify_end_one_batch(meta, state)``,
``edl.notify_end_one_epoch(state)`` -- tests/unittests/test_train.py:27-65)."""
from __future__ import annotations

import os
import time
from typing import Optional

from .utils.env import TrainerEnv
from .utils import state as _state

_env: Optional[TrainerEnv] = None
_last_batch_t = None


def trainer_env(refresh: bool = False) -> TrainerEnv:
    global _env
    if _env is None or refresh:
        _env = TrainerEnv()
    return _env


def size() -> int:
    """Number of trainers in the current stage."""
    return trainer_env().size


def rank() -> int:
    return trainer_env().global_rank


def local_rank() -> int:
    return trainer_env().rank_in_pod


def is_leader_trainer() -> bool:
    return rank() == 0


def init_distributed(backend: Optional[str] = None, timeout_s: float = 120.0):
    """``torch.distributed.init_process_group`` from the launcher's environment contract."""
    import datetime

    import torch
    import torch.distributed as dist

    env = trainer_env(refresh=True)
    for k, v in env.torch_distributed_env().items():
        os.environ.setdefault(k, v)
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(env.rank_in_pod % max(1, torch.cuda.device_count()))
    if env.size > 1 and not dist.is_initialized():
        dist.init_process_group(backend, rank=env.global_rank, world_size=env.size,
                                timeout=datetime.timedelta(seconds=timeout_s))
    return env


def notify_end_one_batch(meta, state: _state.State):
    """Record one consumed batch: step counters, avg step time, consumed record ranges (``meta`` is
    the ``{"file_idx", "begin", "end"}`` dict the elastic Reader attaches to every batch)."""
    global _last_batch_t
    now = time.time()
    dt = None if _last_batch_t is None else now - _last_batch_t
    _last_batch_t = now
    state.end_one_batch(world_size=size(), step_time=dt)
    if meta:
        metas = meta if isinstance(meta, (list, tuple)) else [meta]
        for m in metas:
            if m and "file_idx" in m:
                state.data_checkpoint.mark(m["file_idx"], m["begin"], m["end"])


def notify_end_one_epoch(state: _state.State):
    global _last_batch_t
    _last_batch_t = None
    state.end_one_epoch()
