"""Teacher registrar sidecar: publish ``ip:port`` of an inference server under a service name with
TTL heart-beats, re-register after an outage, give up when the server stays dead.

CLI parity with the reference (python/edl/discovery/register.py:28-145):
``python -m paddle_edl.discovery.register --db_endpoints h:p --service_name S --server ip:port``.
"""
from __future__ import annotations

import argparse
import json
import logging
import threading
import time

from .etcd_client import EtcdClient
from .server_alive import is_server_alive

logger = logging.getLogger("edl.discovery.register")


def default_load_info() -> str:
    """The reference publishes a placeholder ``"{gpu:20%, net:1}"`` (register.py:35-38); we publish
    real GPU utilisation when NVML is available."""
    info = {"gpu": None, "net": 1}
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        info["gpu"] = pynvml.nvmlDeviceGetUtilizationRates(h).gpu
    except Exception:  # noqa: BLE001 - no GPU / no NVML on this host
        pass
    return json.dumps(info)


class ServerRegister:
    def __init__(self, db_endpoints, service_name, server, ttl=10, heartbeat=1.5, max_dead_probes=45,
                 wait_alive_s=600.0, info_fn=default_load_info, root="service"):
        if isinstance(db_endpoints, str):
            db_endpoints = db_endpoints.split(",")
        self._db = EtcdClient(db_endpoints, root=root)
        self._service_name, self._server = service_name, server
        self._ttl, self._heartbeat, self._max_dead = ttl, heartbeat, max_dead_probes
        self._wait_alive_s = wait_alive_s
        self._info_fn = info_fn
        self._stop = threading.Event()
        self._thread = None

    def _wait_server_up(self) -> bool:
        deadline = time.time() + self._wait_alive_s
        while not self._stop.is_set() and time.time() < deadline:
            alive, _ = is_server_alive(self._server)
            if alive:
                return True
            logger.info("waiting for %s to accept connections", self._server)
            self._stop.wait(self._heartbeat)
        return False

    def register(self, block: bool = True):
        """Wait for the server's port, register it, then heart-beat until stopped / server dead."""
        self._db.init()
        if not self._wait_server_up():
            raise RuntimeError("server %s never came up" % self._server)
        if not self._db.set_server_not_exists(self._service_name, self._server, self._info_fn(),
                                              ttl=self._ttl):
            # somebody (a previous incarnation) still holds the key: take it over
            self._db.refresh(self._service_name, self._server, info=self._info_fn(), ttl=self._ttl)
        logger.info("registered %s under service %s", self._server, self._service_name)
        if block:
            self._beat_loop()
        else:
            self._thread = threading.Thread(target=self._beat_loop, daemon=True, name="teacher-register")
            self._thread.start()
        return self

    def _beat_loop(self):
        dead = 0
        while not self._stop.wait(self._heartbeat):
            alive, _ = is_server_alive(self._server)
            if not alive:
                dead += 1
                logger.warning("%s not reachable (%d/%d)", self._server, dead, self._max_dead)
                if dead >= self._max_dead:
                    logger.error("giving up on %s", self._server)
                    break
                continue
            dead = 0
            try:
                self._db.refresh(self._service_name, self._server, info=self._info_fn(), ttl=self._ttl)
            except Exception as e:  # noqa: BLE001 - store outage: re-register next beat
                logger.warning("refresh failed (%s); will retry", e)
        try:
            self._db.remove_server(self._service_name, self._server)
        except Exception:  # noqa: BLE001
            pass

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(5)


def main(argv=None):
    ap = argparse.ArgumentParser(description="Teacher server registrar")
    ap.add_argument("--db_endpoints", type=str, default="127.0.0.1:2379")
    ap.add_argument("--db_passwd", type=str, default=None)
    ap.add_argument("--db_type", type=str, default="etcd")
    ap.add_argument("--service_name", type=str, required=True)
    ap.add_argument("--server", type=str, required=True, help="ip:port of the teacher server")
    ap.add_argument("--service_token", type=str, default=None)
    ap.add_argument("--ttl", type=int, default=10)
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    ServerRegister(args.db_endpoints, args.service_name, args.server, ttl=args.ttl).register(block=True)


if __name__ == "__main__":
    main()
