"""Per-stage wall-clock profiler of the distill pipeline: ``DISTILL_READER_PROFILE=1`` prints
``pid op time_ms`` on every ``record()``; otherwise a zero-cost no-op
(reference: python/edl/distill/timeline.py:20-47)."""
import os
import sys
import threading
import time


class _NopTimeLine:
    def record(self, name):
        pass

    def reset(self):
        pass


class _RealTimeLine:
    def __init__(self):
        self.pid = os.getpid()
        self.tid = threading.get_ident() & 0xFFFF
        self.time = time.time()

    def record(self, name):
        now = time.time()
        sys.stderr.write("pid={} tid={} op={} time={:.3f}ms\n".format(self.pid, self.tid, name,
                                                                    (now - self.time) * 1000))
        self.time = now

    def reset(self):
        self.time = time.time()


def TimeLine():
    return _RealTimeLine() if os.environ.get("DISTILL_READER_PROFILE", "0") == "1" else _NopTimeLine()
