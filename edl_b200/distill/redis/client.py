"""Student-side client of the redis-flavoured balance server: register, then heartbeat every 2 s on a
blocking socket (reference: python/edl/distill/redis/client.py:24-157)."""
import json
import logging
import random
import socket
import threading

from .balance_server import HEADER, MAGIC, pack_frame

logger = logging.getLogger("edl.distill.redis")


class Client:
    def __init__(self, endpoints, service_name, require_num, token=None, heartbeat_s=2.0):
        if isinstance(endpoints, str):
            endpoints = [e for e in endpoints.split(",") if e]
        self._endpoints = list(endpoints)
        self._service_name, self._require_num = service_name, int(require_num)
        self._heartbeat_s = heartbeat_s
        self._sock = None
        self._servers = []
        self._version = -1
        self._lock = threading.Lock()
        self._stop = threading.Event()
        self._ready = threading.Event()
        self._t = None

    def _recv_exact(self, n):
        buf = bytearray()
        while len(buf) < n:
            b = self._sock.recv(n - len(buf))
            if not b:
                raise ConnectionError("balance server closed the connection")
            buf += b
        return bytes(buf)

    def _call(self, obj):
        self._sock.sendall(pack_frame(obj))
        magic, total = HEADER.unpack(self._recv_exact(HEADER.size))
        if magic != MAGIC:
            raise ConnectionError("bad magic from balance server")
        return json.loads(self._recv_exact(total - HEADER.size).decode("utf-8"))

    def _connect_and_register(self):
        host, port = random.choice(self._endpoints).rsplit(":", 1)
        self._sock = socket.create_connection((host, int(port)), timeout=6)
        self._sock.settimeout(10)
        r = self._call({"type": "register", "service_name": self._service_name, "seq": 0,
                        "num": self._require_num})
        assert r.get("type") == "register", r
        with self._lock:
            self._servers, self._version = list(r.get("servers", [])), int(r.get("version", 0))
        self._ready.set()

    def _loop(self):
        while not self._stop.is_set():
            try:
                if self._sock is None:
                    self._connect_and_register()
                r = self._call({"type": "heartbeat", "version": self._version})
                if r.get("type") == "servers_change":
                    with self._lock:
                        self._servers, self._version = list(r["servers"]), int(r["version"])
                elif r.get("type") == "error":
                    raise ConnectionError(r.get("message"))
            except (OSError, ConnectionError, AssertionError, ValueError) as e:
                logger.warning("balance server link lost (%s); reconnecting", e)
                if self._sock is not None:
                    try:
                        self._sock.close()
                    except OSError:
                        pass
                    self._sock = None
                self._stop.wait(0.5)
                continue
            self._stop.wait(self._heartbeat_s)

    def start(self, wait_s=10.0):
        self._t = threading.Thread(target=self._loop, daemon=True, name="redis-balance-client")
        self._t.start()
        self._ready.wait(wait_s)
        return self

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(3)
        if self._sock is not None:
            try:
                self._sock.close()
            except OSError:
                pass

    def get_servers(self):
        with self._lock:
            return list(self._servers)

    def get_teacher_list(self):
        return self.get_servers()
