"""Training-side observability.

The reference prints ``Pass/trainbatch/loss/acc/lr/time/speed`` every ``fetch_steps`` and writes a
``benchmark_logs/log_<trainer_id>`` dict at the end (example/distill/resnet/train_with_fleet.py:508-523,605-621);
``EpochAttr.avg_step_time`` is meant for the scheduler.  ``StepMeter`` keeps those numbers; ``MetricsExporter``
optionally publishes them (plus world size / stage, so a dashboard sees elastic events) over the Prometheus text
protocol -- ``prometheus_client`` if installed, else a tiny built-in HTTP endpoint."""
from __future__ import annotations

import json
import os
import threading
import time
from http.server import BaseHTTPRequestHandler, HTTPServer
from typing import Dict, Optional


class StepMeter:
    def __init__(self, batch_per_trainer: int, world: int = 1, window: int = 50):
        self.bs, self.world, self.window = batch_per_trainer, world, window
        self.reset()

    def reset(self):
        self.t_last = time.perf_counter()
        self.step_times = []
        self.steps = 0
        self.best_ips = 0.0

    def step(self) -> float:
        now = time.perf_counter()
        dt = now - self.t_last
        self.t_last = now
        self.steps += 1
        self.step_times.append(dt)
        if len(self.step_times) > self.window:
            self.step_times.pop(0)
        return dt

    @property
    def avg_step_time(self) -> float:
        return sum(self.step_times) / max(1, len(self.step_times))

    @property
    def images_per_second(self) -> float:
        ips = self.bs * self.world / max(1e-9, self.avg_step_time)
        if len(self.step_times) >= min(10, self.window):
            self.best_ips = max(self.best_ips, ips)
        return ips

    def summary(self) -> Dict[str, float]:
        return {"steps": self.steps, "avg_step_time_s": self.avg_step_time, "img_per_s": self.images_per_second,
                "best_img_per_s": self.best_ips, "world": self.world}


def write_benchmark_log(trainer_id: int, record: Dict, log_dir: str = "./benchmark_logs") -> str:
    os.makedirs(log_dir, exist_ok=True)
    path = os.path.join(log_dir, "log_%d" % trainer_id)
    with open(path, "w") as f:
        json.dump(record, f, sort_keys=True)
    return path


class MetricsExporter:
    """``exp = MetricsExporter(port).start(); exp.set("edl_img_per_s", 6700, {"job": "rn50"})``"""

    def __init__(self, port: int = 0, host: str = "0.0.0.0"):
        self.host, self.port = host, port
        self._vals: Dict[str, float] = {}
        self._lock = threading.Lock()
        self._srv: Optional[HTTPServer] = None

    @staticmethod
    def _key(name, labels):
        if not labels:
            return name
        return "%s{%s}" % (name, ",".join('%s="%s"' % kv for kv in sorted(labels.items())))

    def set(self, name: str, value: float, labels: Optional[Dict[str, str]] = None):
        with self._lock:
            self._vals[self._key(name, labels)] = float(value)

    def render(self) -> str:
        with self._lock:
            return "".join("%s %.9g\n" % kv for kv in sorted(self._vals.items()))

    def start(self):
        exporter = self

        class H(BaseHTTPRequestHandler):
            def do_GET(self):  # noqa: N802
                body = exporter.render().encode()
                self.send_response(200)
                self.send_header("Content-Type", "text/plain; version=0.0.4")
                self.send_header("Content-Length", str(len(body)))
                self.end_headers()
                self.wfile.write(body)

            def log_message(self, *a):
                pass

        self._srv = HTTPServer((self.host, self.port), H)
        self.port = self._srv.server_address[1]
        threading.Thread(target=self._srv.serve_forever, daemon=True, name="edl-metrics").start()
        return self

    def stop(self):
        if self._srv is not None:
            self._srv.shutdown()
            self._srv.server_close()
            self._srv = None
