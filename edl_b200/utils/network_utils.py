"""Host / port helpers (reference: python/edl/utils/network_utils.py:30-54)."""
import os
import socket
from contextlib import closing


def get_host_name_ip():
    """(hostname, ip).  ``POD_IP`` / ``EDL_POD_IP`` override; falls back to 127.0.0.1 when the
    hostname does not resolve (containers)."""
    ip = os.environ.get("EDL_POD_IP") or os.environ.get("POD_IP")
    name = socket.gethostname()
    if ip:
        return name, ip
    try:
        return name, socket.gethostbyname(name)
    except OSError:
        return name, "127.0.0.1"


def get_extern_ip():
    return get_host_name_ip()[1]


def find_free_ports(num: int):
    """``num`` distinct currently-free TCP ports."""
    ports, socks = [], []
    try:
        while len(ports) < num:
            s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            s.bind(("", 0))
            socks.append(s)
            p = s.getsockname()[1]
            if p not in ports:
                ports.append(p)
    finally:
        for s in socks:
            s.close()
    return ports


def is_port_free(port: int, host: str = "127.0.0.1") -> bool:
    with closing(socket.socket(socket.AF_INET, socket.SOCK_STREAM)) as s:
        return s.connect_ex((host, port)) != 0
