"""Tracing hooks: NVTX ranges for nsys/ncu captures and a step-window torch profiler
(reference: ``--profile`` runs ``paddle.fluid.profiler`` for steps 100..105 on trainer 0 and writes
``./profile_pass_N``, example/distill/resnet/train_with_fleet.py:497-506; distill pipeline stages use
``distill.timeline`` instead)."""
import contextlib
import os

import torch


@contextlib.contextmanager
def nvtx_range(name):
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


class StepProfiler:
    """``with StepProfiler(start=100, stop=105, out='profile_pass_0') as p: ... p.step()`` each
    iteration: kernels of steps [start, stop) are recorded and a table + chrome trace are written."""

    def __init__(self, start=100, stop=105, out="./profile_pass_0", enabled=True, rank=0):
        self.start, self.stop, self.out = start, stop, out
        self.enabled = enabled and rank == 0
        self.n = 0
        self.prof = None

    def __enter__(self):
        return self

    def step(self):
        if not self.enabled:
            return
        if self.n == self.start:
            from torch.profiler import ProfilerActivity, profile

            acts = [ProfilerActivity.CPU] + ([ProfilerActivity.CUDA] if torch.cuda.is_available() else [])
            self.prof = profile(activities=acts)
            self.prof.__enter__()
        elif self.n == self.stop and self.prof is not None:
            self._finish()
        self.n += 1

    def _finish(self):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.prof.__exit__(None, None, None)
        os.makedirs(self.out, exist_ok=True)
        key = "cuda_time_total" if torch.cuda.is_available() else "cpu_time_total"
        with open(os.path.join(self.out, "kernels.txt"), "w") as f:
            f.write(self.prof.key_averages().table(sort_by=key, row_limit=80))
        try:
            self.prof.export_chrome_trace(os.path.join(self.out, "trace.json"))
        except Exception:  # noqa: BLE001
            pass
        self.prof = None

    def __exit__(self, *exc):
        if self.prof is not None:
            self._finish()
