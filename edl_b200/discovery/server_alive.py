"""TCP liveness probe for teacher / discovery servers (reference: discovery/server_alive.py:19-34)."""
import socket
from contextlib import closing


def is_server_alive(server: str, timeout: float = 1.5):
    """Returns ``(alive, local_addr)``; ``local_addr`` is the (ip, port) this host used to reach the
    server -- the registrar publishes it so clients know a routable address."""
    host, port = server.rsplit(":", 1)
    with closing(socket.socket(socket.AF_INET, socket.SOCK_STREAM)) as s:
        s.settimeout(timeout)
        try:
            s.connect((host, int(port)))
            addr = s.getsockname()
            try:
                s.shutdown(socket.SHUT_RDWR)
            except OSError:
                pass
            return True, addr
        except OSError:
            return False, None
