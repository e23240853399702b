"""Teacher registry on redis: one hash per ``/service/<name>/nodes/<server>`` with ``EXPIRE ttl``
(reference: python/edl/distill/redis/redis_store.py:18-72, ttl 6 s)."""
from .resp import RespClient


class RedisStore:
    def __init__(self, ip, port, passwd=None, ttl=6):
        self._redis = RespClient(ip, port)
        self._ttl = ttl

    @staticmethod
    def _key(service_name, server):
        return "/service/{}/nodes/{}".format(service_name, server)

    def get_service(self, service_name):
        """-> [{"server":..., "info":...}]"""
        out = []
        for k in self._redis.keys("/service/{}/nodes/*".format(service_name)):
            h = self._redis.hgetall(k)
            if h:
                out.append({"server": h.get("server", k.rsplit("/", 1)[-1]), "info": h.get("info", "")})
        return out

    def set_server(self, service_name, server, info, ttl=None):
        key = self._key(service_name, server)
        self._redis.hset(key, {"server": server, "info": info})
        self._redis.expire(key, ttl or self._ttl)

    def refresh(self, service_name, server, info=None, ttl=None):
        key = self._key(service_name, server)
        if info is not None or not self._redis.exists(key):
            return self.set_server(service_name, server, info or "", ttl)
        self._redis.expire(key, ttl or self._ttl)

    def remove_server(self, service_name, server):
        self._redis.delete(self._key(service_name, server))

    def close(self):
        self._redis.close()
