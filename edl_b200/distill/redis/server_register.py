"""Teacher registrar for the redis flavour: probe the teacher port, publish it with EXPIRE ttl and
refresh every 1.5 s; give up after 60 failed probes
(reference: python/edl/distill/redis/server_register.py:19-141).

    python -m paddle_edl.distill.redis.server_register --db_endpoints 127.0.0.1:6379 --service_name S --server ip:port
"""
import argparse
import logging
import threading

from ...discovery.register import default_load_info
from ...discovery.server_alive import is_server_alive
from .redis_store import RedisStore

logger = logging.getLogger("edl.distill.redis")


class ServerRegister:
    def __init__(self, db_ip, db_port, service_name, server, ttl=6, heartbeat=1.5, max_dead_probes=60,
                 info_fn=default_load_info):
        self._store = RedisStore(db_ip, db_port, ttl=ttl)
        self._service_name, self._server = service_name, server
        self._heartbeat, self._max_dead = heartbeat, max_dead_probes
        self._info_fn = info_fn
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        dead = 0
        registered = False
        while not self._stop.is_set():
            alive, _ = is_server_alive(self._server)
            if alive:
                dead = 0
                try:
                    self._store.refresh(self._service_name, self._server, info=self._info_fn())
                    if not registered:
                        logger.info("registered %s under %s", self._server, self._service_name)
                        registered = True
                except Exception as e:  # noqa: BLE001
                    logger.warning("redis refresh failed: %s", e)
            else:
                dead += 1
                if dead >= self._max_dead:
                    logger.error("%s is dead; giving up", self._server)
                    break
            self._stop.wait(self._heartbeat)
        try:
            self._store.remove_server(self._service_name, self._server)
        except Exception:  # noqa: BLE001
            pass

    def register(self, block=True):
        if block:
            self._loop()
        else:
            self._t = threading.Thread(target=self._loop, daemon=True, name="redis-teacher-register")
            self._t.start()
        return self

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(5)


def main(argv=None):
    ap = argparse.ArgumentParser(description="Teacher server registrar (redis)")
    ap.add_argument("--db_endpoints", type=str, default="127.0.0.1:6379")
    ap.add_argument("--db_passwd", type=str, default=None)
    ap.add_argument("--db_type", type=str, default="redis")
    ap.add_argument("--service_name", type=str, required=True)
    ap.add_argument("--server", type=str, required=True)
    ap.add_argument("--service_token", type=str, default=None)
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    ip, port = args.db_endpoints.split(",")[0].rsplit(":", 1)
    ServerRegister(ip, port, args.service_name, args.server).register(block=True)


if __name__ == "__main__":
    main()
