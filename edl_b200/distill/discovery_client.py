"""Student-side discovery client: register, heart-beat every 2 s, follow redirects, re-register when
the server forgets us (reference: python/edl/distill/discovery_client.py:49-268)."""
import logging
import os
import random
import threading
import time

import grpc

from ..protos import rpc, schema
from ..protos.schema import Code
from ..utils.network_utils import get_extern_ip

logger = logging.getLogger("edl.distill.discovery")


class DiscoveryClient:
    def __init__(self, endpoints, service_name, require_num, token=None, heartbeat_s=2.0):
        if isinstance(endpoints, str):
            endpoints = [e for e in endpoints.split(",") if e]
        self._discovery_servers = list(endpoints)
        self._service_name, self._require_num, self._token = service_name, int(require_num), token or ""
        self._heartbeat_s = heartbeat_s
        self._client = "{}-{}-{}-{}".format(get_extern_ip(), os.getpid(), id(self) & 0xFFFF, int(time.time() * 1000))
        self._lock = threading.Lock()
        self._servers = []
        self._version = 0
        self._discovery_version = 0
        self._channel = self._stub = None
        self._endpoint = None
        self._stop = threading.Event()
        self._registered = threading.Event()
        self._t = None

    # ------------------------------------------------------------------ connection handling
    def _connect(self, endpoint):
        if self._channel is not None:
            self._channel.close()
        self._endpoint = endpoint
        self._channel = rpc.insecure_channel(endpoint)
        self._stub = rpc.Stub(self._channel, "paddle_edl.distill.DiscoveryService")

    def _pick(self):
        return random.choice(self._discovery_servers)

    def _apply(self, res, registering):
        code = res.status.code
        if res.discovery_servers:
            self._discovery_servers = list(res.discovery_servers)
            self._discovery_version = res.discovery_version
        if code in (Code.OK, Code.ALREADY_REGISTER):
            with self._lock:
                if registering or res.version != self._version:
                    self._servers = list(res.servers)
                    self._version = res.version
            self._registered.set()
            return True
        if code == Code.REDIRECT:
            logger.info("discovery redirect -> %s", res.status.message)
            self._connect(res.status.message)
            self._registered.clear()
            return False
        if code == Code.UNREGISTERED:
            self._registered.clear()
            return False
        if code == Code.NO_READY:
            time.sleep(0.5)
            return False
        logger.warning("discovery error %d: %s", code, res.status.message)
        return False

    def _register_once(self):
        req = schema.distill_discovery.RegisterRequest(client=self._client, service_name=self._service_name,
                                                       require_num=self._require_num, token=self._token)
        return self._apply(self._stub.Register(req, timeout=5), registering=True)

    def _heartbeat_once(self):
        req = schema.distill_discovery.HeartBeatRequest(client=self._client, version=self._version,
                                                        discovery_version=self._discovery_version)
        res = self._stub.HeartBeat(req, timeout=5)
        if res.status.code == Code.OK and res.version != self._version:
            with self._lock:
                self._servers = list(res.servers)
                self._version = res.version
            return True
        return self._apply(res, registering=False)

    def _loop(self):
        failures = 0
        while not self._stop.is_set():
            try:
                if self._stub is None:
                    self._connect(self._pick())
                if not self._registered.is_set():
                    self._register_once()
                    if not self._registered.is_set():
                        self._stop.wait(0.2)
                        continue
                else:
                    self._heartbeat_once()
                failures = 0
            except grpc.RpcError as e:
                failures += 1
                logger.warning("discovery rpc to %s failed (%s)", self._endpoint, e.code())
                if failures >= 3:
                    self._registered.clear()
                    self._connect(self._pick())
                    failures = 0
                self._stop.wait(0.3)
                continue
            self._stop.wait(self._heartbeat_s if self._registered.is_set() else 0.2)

    # ------------------------------------------------------------------ public
    def start(self, daemon=True, wait_s=10.0):
        self._t = threading.Thread(target=self._loop, daemon=daemon, name="discovery-client")
        self._t.start()
        self._registered.wait(wait_s)
        return self

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(3)
        if self._channel is not None:
            self._channel.close()

    def get_servers(self):
        with self._lock:
            return list(self._servers)

    @property
    def version(self):
        return self._version
