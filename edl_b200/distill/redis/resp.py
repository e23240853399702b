"""Minimal RESP (REdis Serialization Protocol) client and an in-repo mini server.

The ``redis`` python package and ``redis-server`` are not available offline; this client speaks the
real wire protocol (so it works against a genuine redis-server) and ``MiniRedisServer`` implements
the handful of commands the registry needs (PING, HSET/HMSET, HGETALL, EXPIRE, TTL, DEL, KEYS, SCAN,
EXISTS) with key expiry."""
import fnmatch
import socket
import socketserver
import threading
import time


class RespError(Exception):
    pass


def _encode(args):
    out = [b"*%d\r\n" % len(args)]
    for a in args:
        if not isinstance(a, bytes):
            a = str(a).encode("utf-8")
        out.append(b"$%d\r\n%s\r\n" % (len(a), a))
    return b"".join(out)


class _Reader:
    def __init__(self, sock):
        self.f = sock.makefile("rb")

    def read(self):
        line = self.f.readline()
        if not line:
            raise ConnectionError("connection closed")
        t, rest = line[:1], line[1:-2]
        if t == b"+":
            return rest.decode()
        if t == b"-":
            raise RespError(rest.decode())
        if t == b":":
            return int(rest)
        if t == b"$":
            n = int(rest)
            if n < 0:
                return None
            data = self.f.read(n + 2)
            return data[:-2]
        if t == b"*":
            n = int(rest)
            return None if n < 0 else [self.read() for _ in range(n)]
        raise RespError("bad RESP type %r" % t)


class RespClient:
    def __init__(self, host="127.0.0.1", port=6379, timeout=6.0):
        self.host, self.port, self.timeout = host, int(port), timeout
        self._sock = None
        self._reader = None
        self._lock = threading.Lock()

    def _connect(self):
        self._sock = socket.create_connection((self.host, self.port), timeout=self.timeout)
        self._sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        self._reader = _Reader(self._sock)

    def execute(self, *args):
        with self._lock:
            for attempt in (0, 1):
                try:
                    if self._sock is None:
                        self._connect()
                    self._sock.sendall(_encode(args))
                    return self._reader.read()
                except (OSError, ConnectionError):
                    self.close_nolock()
                    if attempt == 1:
                        raise

    def close_nolock(self):
        if self._sock is not None:
            try:
                self._sock.close()
            except OSError:
                pass
        self._sock = self._reader = None

    def close(self):
        with self._lock:
            self.close_nolock()

    # convenience wrappers (redis-py flavoured names)
    def ping(self): return self.execute("PING") == "PONG"
    def hset(self, key, mapping):
        flat = []
        for k, v in mapping.items():
            flat += [k, v]
        return self.execute("HSET", key, *flat)
    def hgetall(self, key):
        arr = self.execute("HGETALL", key) or []
        return {arr[i].decode(): arr[i + 1].decode() for i in range(0, len(arr), 2)}
    def expire(self, key, ttl): return self.execute("EXPIRE", key, int(ttl))
    def ttl(self, key): return self.execute("TTL", key)
    def delete(self, *keys): return self.execute("DEL", *keys)
    def keys(self, pattern): return [k.decode() for k in (self.execute("KEYS", pattern) or [])]
    def exists(self, key): return self.execute("EXISTS", key) == 1


class _State:
    def __init__(self):
        self.lock = threading.Lock()
        self.data = {}     # key -> dict (hash)
        self.expiry = {}   # key -> monotonic deadline

    def _gc(self, key=None):
        now = time.monotonic()
        keys = [key] if key is not None else list(self.expiry)
        for k in keys:
            d = self.expiry.get(k)
            if d is not None and d <= now:
                self.expiry.pop(k, None)
                self.data.pop(k, None)

    def run(self, cmd, args):
        with self.lock:
            self._gc()
            c = cmd.upper()
            if c == "PING":
                return "+PONG"
            if c in ("HSET", "HMSET"):
                h = self.data.setdefault(args[0], {})
                added = 0
                for i in range(1, len(args) - 1, 2):
                    added += args[i] not in h
                    h[args[i]] = args[i + 1]
                return added if c == "HSET" else "+OK"
            if c == "HGETALL":
                h = self.data.get(args[0], {})
                out = []
                for k, v in h.items():
                    out += [k, v]
                return out
            if c == "EXPIRE":
                if args[0] not in self.data:
                    return 0
                self.expiry[args[0]] = time.monotonic() + float(args[1])
                return 1
            if c == "TTL":
                if args[0] not in self.data:
                    return -2
                d = self.expiry.get(args[0])
                return -1 if d is None else max(0, int(d - time.monotonic()))
            if c == "DEL":
                n = 0
                for k in args:
                    n += self.data.pop(k, None) is not None
                    self.expiry.pop(k, None)
                return n
            if c == "EXISTS":
                return int(args[0] in self.data)
            if c == "KEYS":
                return [k for k in sorted(self.data) if fnmatch.fnmatchcase(k, args[0])]
            if c == "SCAN":
                pat = "*"
                for i, a in enumerate(args):
                    if a.upper() == "MATCH":
                        pat = args[i + 1]
                return ["0", [k for k in sorted(self.data) if fnmatch.fnmatchcase(k, pat)]]
            if c in ("FLUSHALL", "FLUSHDB"):
                self.data.clear()
                self.expiry.clear()
                return "+OK"
            return RespError("ERR unknown command '%s'" % cmd)


def _reply(v):
    if isinstance(v, RespError):
        return b"-%s\r\n" % str(v).encode()
    if isinstance(v, str) and v.startswith("+"):
        return v.encode() + b"\r\n"
    if isinstance(v, int):
        return b":%d\r\n" % v
    if v is None:
        return b"$-1\r\n"
    if isinstance(v, (list, tuple)):
        return b"*%d\r\n" % len(v) + b"".join(_reply(x) for x in v)
    b = v if isinstance(v, bytes) else str(v).encode()
    return b"$%d\r\n%s\r\n" % (len(b), b)


class _Handler(socketserver.StreamRequestHandler):
    def handle(self):
        rd = _Reader(self.connection)
        while True:
            try:
                req = rd.read()
            except (ConnectionError, OSError, RespError):
                return
            if not isinstance(req, list) or not req:
                return
            parts = [p.decode() if isinstance(p, bytes) else str(p) for p in req]
            try:
                self.connection.sendall(_reply(self.server.state.run(parts[0], parts[1:])))
            except OSError:
                return


class MiniRedisServer:
    def __init__(self, host="127.0.0.1", port=0):
        class _S(socketserver.ThreadingTCPServer):
            allow_reuse_address = True
            daemon_threads = True
        self._srv = _S((host, port), _Handler)
        self._srv.state = _State()
        self.host, self.port = self._srv.server_address[:2]

    @property
    def endpoint(self):
        return "%s:%d" % (self.host, self.port)

    def start(self):
        threading.Thread(target=self._srv.serve_forever, kwargs={"poll_interval": 0.1}, daemon=True).start()
        return self

    def stop(self):
        self._srv.shutdown()
        self._srv.server_close()

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()
