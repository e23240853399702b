"""Single-threaded event-loop TCP balance server for the redis flavour.

Wire format (reference: python/edl/distill/redis/balance_server.py:41-44,76-124):
8-byte header ``struct "!4si"`` = magic ``CB EF 00 00`` + int32 TOTAL frame length (header
included), then a UTF-8 JSON body.
  {"type":"register","service_name","seq":0,"num"}  -> {"type":"register","seq":1,"servers":[...],"num"}
  {"type":"heartbeat","version":v}                  -> {"type":"heartbeat"} | {"type":"servers_change","servers","version"}
Bad magic or length mismatch closes the connection.

    python -m paddle_edl.distill.redis.balance_server --server 0.0.0.0:7001 --db_endpoints 127.0.0.1:6379
"""
import argparse
import json
import logging
import selectors
import socket
import struct
import threading

from .service_table import ServiceTable

logger = logging.getLogger("edl.distill.redis")

MAGIC = b"\xCB\xEF\x00\x00"
HEADER = struct.Struct("!4si")
MAX_FRAME = 16 << 20


def pack_frame(obj) -> bytes:
    body = json.dumps(obj).encode("utf-8")
    return HEADER.pack(MAGIC, HEADER.size + len(body)) + body


def set_keepalive(sock, idle=30, interval=10, count=6):
    sock.setsockopt(socket.SOL_SOCKET, socket.SO_KEEPALIVE, 1)
    for opt, val in (("TCP_KEEPIDLE", idle), ("TCP_KEEPINTVL", interval), ("TCP_KEEPCNT", count)):
        if hasattr(socket, opt):
            sock.setsockopt(socket.IPPROTO_TCP, getattr(socket, opt), val)


class _Conn:
    def __init__(self, sock):
        self.sock = sock
        self.inbuf = bytearray()
        self.outbuf = bytearray()


class BalanceServer:
    def __init__(self, host, port, db_host, db_port, passwd=None):
        self._addr = (host, int(port))
        self._table = ServiceTable(db_host, db_port, passwd)
        self._sel = selectors.DefaultSelector()   # epoll on Linux
        self._listen = None
        self._conns = {}
        self._stop = threading.Event()
        self._t = None
        self.port = None

    # ------------------------------------------------------------------ message handling
    def _handle(self, conn, msg):
        fd = conn.sock.fileno()
        t = msg.get("type")
        if t == "register":
            snap = self._table.add_client(fd, msg["service_name"], int(msg.get("num", 1)))
            version, servers = snap if snap else (0, [])
            return {"type": "register", "seq": int(msg.get("seq", 0)) + 1, "servers": servers,
                    "num": int(msg.get("num", 1)), "version": version}
        if t == "heartbeat":
            snap = self._table.get_servers(fd)
            if snap is None:
                return {"type": "error", "message": "not registered"}
            version, servers = snap
            if version != int(msg.get("version", -1)):
                return {"type": "servers_change", "servers": servers, "version": version}
            return {"type": "heartbeat"}
        return {"type": "error", "message": "unknown type %r" % t}

    def _on_readable(self, conn):
        try:
            data = conn.sock.recv(65536)
        except (BlockingIOError, InterruptedError):
            return
        except OSError:
            data = b""
        if not data:
            return self._close(conn)
        conn.inbuf += data
        while len(conn.inbuf) >= HEADER.size:
            magic, total = HEADER.unpack_from(conn.inbuf)
            if magic != MAGIC or total < HEADER.size or total > MAX_FRAME:
                logger.warning("bad frame from %s; closing", conn.sock.getpeername())
                return self._close(conn)
            if len(conn.inbuf) < total:
                break
            body = bytes(conn.inbuf[HEADER.size:total])
            del conn.inbuf[:total]
            try:
                reply = self._handle(conn, json.loads(body.decode("utf-8")))
            except Exception as e:  # noqa: BLE001
                reply = {"type": "error", "message": str(e)}
            conn.outbuf += pack_frame(reply)
        self._flush(conn)

    def _flush(self, conn):
        try:
            while conn.outbuf:
                n = conn.sock.send(conn.outbuf)
                del conn.outbuf[:n]
        except (BlockingIOError, InterruptedError):
            pass
        except OSError:
            return self._close(conn)
        events = selectors.EVENT_READ | (selectors.EVENT_WRITE if conn.outbuf else 0)
        try:
            self._sel.modify(conn.sock, events, conn)
        except (KeyError, ValueError):
            pass

    def _close(self, conn):
        fd = conn.sock.fileno()
        self._table.rm_client(fd)
        try:
            self._sel.unregister(conn.sock)
        except (KeyError, ValueError):
            pass
        conn.sock.close()
        self._conns.pop(fd, None)

    # ------------------------------------------------------------------ loop
    def _loop(self):
        while not self._stop.is_set():
            for key, mask in self._sel.select(timeout=0.2):
                if key.data is None:
                    try:
                        sock, _ = self._listen.accept()
                    except OSError:
                        continue
                    sock.setblocking(False)
                    set_keepalive(sock)
                    conn = _Conn(sock)
                    self._conns[sock.fileno()] = conn
                    self._sel.register(sock, selectors.EVENT_READ, conn)
                else:
                    conn = key.data
                    if mask & selectors.EVENT_READ:
                        self._on_readable(conn)
                    if mask & selectors.EVENT_WRITE and conn.sock.fileno() in self._conns:
                        self._flush(conn)

    def start(self):
        self._listen = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self._listen.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self._listen.bind(self._addr)
        self._listen.listen(512)
        self._listen.setblocking(False)
        self.port = self._listen.getsockname()[1]
        self._sel.register(self._listen, selectors.EVENT_READ, None)
        self._table.start()
        self._t = threading.Thread(target=self._loop, daemon=True, name="balance-server")
        self._t.start()
        logger.info("balance server listening on %s:%d", self._addr[0], self.port)
        return self

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(3)
        for conn in list(self._conns.values()):
            self._close(conn)
        if self._listen is not None:
            self._listen.close()
        self._table.stop()

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()


def main(argv=None):
    ap = argparse.ArgumentParser(description="Discovery server with balance (redis flavour)")
    ap.add_argument("--server", type=str, default="0.0.0.0:7001", help="endpoint of the server, e.g. 0.0.0.0:7001")
    ap.add_argument("--worker_num", type=int, default=1)
    ap.add_argument("--db_endpoints", type=str, default="127.0.0.1:6379")
    ap.add_argument("--db_passwd", type=str, default=None)
    ap.add_argument("--db_type", type=str, default="redis")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    host, port = args.server.rsplit(":", 1)
    db_host, db_port = args.db_endpoints.split(",")[0].rsplit(":", 1)
    srv = BalanceServer(host, port, db_host, db_port, args.db_passwd).start()
    try:
        threading.Event().wait()
    except KeyboardInterrupt:
        srv.stop()


if __name__ == "__main__":
    main()
