"""Handle for the native (C++ / epoll) build of the coordination store, ``store/native/kv_server.cpp``.

Same constructor and surface as ``kv_server.KVServer`` (``start()``, ``stop()``, ``endpoint``, ``port``,
context manager), but the server runs as a child process: one epoll thread, no GIL, snapshots compatible with
the Python server's.  ``build()`` compiles it with g++ (a second or two); ``available()`` tells whether a
compiler or a built binary is around -- callers fall back to the Python server otherwise.

    python -m edl_b200.store.kv_server --native --port 2379
"""
from __future__ import annotations

import os
import shutil
import subprocess
import threading
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.path.join(HERE, "native", "kv_server.cpp")
BINARY = os.path.join(HERE, "edl_kv_server")


def build(force: bool = False) -> str:
    """Compile the server if the binary is missing or older than its source; returns the binary path."""
    if not force and os.path.exists(BINARY) and os.path.getmtime(BINARY) >= os.path.getmtime(SOURCE):
        return BINARY
    cxx = os.environ.get("CXX", "g++")
    tmp = BINARY + ".tmp.%d" % os.getpid()
    r = subprocess.run([cxx, "-O2", "-std=c++17", "-o", tmp, SOURCE], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("building edl_kv_server failed:\n" + r.stdout)
    os.replace(tmp, BINARY)          # atomic: concurrent builders (pytest-xdist, several pods) never see half a file
    return BINARY


def available() -> bool:
    return os.path.exists(BINARY) or shutil.which(os.environ.get("CXX", "g++")) is not None


class NativeKVServer:
    def __init__(self, host: str = "127.0.0.1", port: int = 0, data_dir: Optional[str] = None,
                 snapshot_interval: float = 2.0):
        self.host, self.port = host, port
        self.data_dir, self.snapshot_interval = data_dir, snapshot_interval
        self.proc: Optional[subprocess.Popen] = None

    @property
    def endpoint(self) -> str:
        return "%s:%d" % (self.host, self.port)

    def start(self) -> "NativeKVServer":
        cmd = [build(), "--host", self.host, "--port", str(self.port), "--snapshot_interval", str(self.snapshot_interval)]
        if self.data_dir:
            cmd += ["--data_dir", self.data_dir]
        self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True)
        line = self.proc.stdout.readline()
        if not line.startswith("listening on "):
            self.proc.kill()
            raise RuntimeError("edl_kv_server did not start: %r" % line)
        self.port = int(line.strip().rsplit(":", 1)[1])
        # keep draining stdout so the child can never block on a full pipe
        threading.Thread(target=lambda f=self.proc.stdout: [None for _ in f], daemon=True).start()
        return self

    def stop(self, timeout: float = 5.0):
        if self.proc is None:
            return
        if self.proc.poll() is None:
            self.proc.terminate()              # SIGTERM: final snapshot, then exit
            try:
                self.proc.wait(timeout)
            except subprocess.TimeoutExpired:
                self.proc.kill()
                self.proc.wait()
        self.proc = None

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()
