"""Differential and robustness tests of the two coordination-store servers.

* the same random sequence of requests (puts with / without leases, put-if-absent, compare-and-swap transactions,
  single / prefix gets and deletes, lease revokes) sent to the Python server and to the C++ server must produce the
  same responses, revision by revision, and the same watch event stream;
* the C++ server built with AddressSanitizer + UBSan survives the protocol tests and a stream of malformed frames
  (truncated msgpack, wrong types, oversized length prefixes) without a sanitizer report -- the reference has no
  sanitizer or race-detection runs at all."""
import os
import random
import socket
import struct
import subprocess
import time

import msgpack
import pytest

from edl_b200.store import KVClient, KVServer, NativeKVServer, native_server

pytestmark = pytest.mark.skipif(not native_server.available(), reason="no C++ compiler for the native store")


def _strip(resp):
    """Responses minus the fields that legitimately differ (lease ids are time-seeded, ttl floats)."""
    if isinstance(resp, dict):
        return {k: _strip(v) for k, v in resp.items() if k not in ("id", "server")}
    if isinstance(resp, list):
        return [_strip(v) for v in resp]
    return resp


def _random_ops(rng, n):
    keys = ["/d/%s" % c for c in "abcdef"] + ["/d/sub/%d" % i for i in range(4)] + ["/other/x"]
    ops = []
    for _ in range(n):
        k = rng.choice(keys)
        r = rng.random()
        val = bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 12)))
        if r < 0.30:
            ops.append({"method": "put", "key": k, "value": val, "lease": rng.choice([0, 0, 0, 101, 102])})
        elif r < 0.40:
            ops.append({"method": "put", "key": k, "value": val, "lease": 0, "if_not_exists": True})
        elif r < 0.55:
            ops.append({"method": "get", "key": k})
        elif r < 0.65:
            ops.append({"method": "get", "key": rng.choice(["/d/", "/d/sub/", "/", "/nope"]), "prefix": True})
        elif r < 0.75:
            ops.append({"method": "delete", "key": k})
        elif r < 0.78:
            ops.append({"method": "delete", "key": rng.choice(["/d/sub/", "/other"]), "prefix": True})
        elif r < 0.93:
            target = rng.choice(["value", "version", "create", "mod"])
            cmpv = val if target == "value" else rng.randint(0, 6)
            ops.append({"method": "txn",
                        "compare": [{"key": k, "target": target, "op": rng.choice(["==", "!=", ">", "<"]), "value": cmpv}],
                        "success": [{"op": "put", "key": rng.choice(keys), "value": b"s"}, {"op": "get", "key": "/d/", "prefix": True}],
                        "failure": [{"op": "delete", "key": rng.choice(keys)}, {"op": "get", "key": k}]})
        elif r < 0.96:
            ops.append({"method": "lease_revoke", "lease": rng.choice([101, 102])})
        elif r < 0.98:
            ops.append({"method": "lease_grant", "ttl": 60.0, "lease_id": rng.choice([101, 102])})
        else:
            ops.append({"method": "put", "key": k, "value": val, "lease": 999})       # unknown lease -> error
    return ops


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_python_and_native_servers_answer_identically(seed):
    rng = random.Random(seed)
    ops = [{"method": "lease_grant", "ttl": 60.0, "lease_id": 101},
           {"method": "lease_grant", "ttl": 60.0, "lease_id": 102}] + _random_ops(rng, 400)
    transcripts, streams = [], []
    for cls in (KVServer, NativeKVServer):
        with cls() as srv:
            c, w = KVClient(srv.endpoint), KVClient(srv.endpoint)
            events = []
            w.add_watch_prefix_callback("/d/", lambda evs, rev, _e=events: _e.extend(evs))
            out = []
            for op in ops:
                try:
                    out.append(_strip(c.call(dict(op))))
                except Exception as e:  # noqa: BLE001 - error responses must match too (by kind)
                    out.append(("error", type(e).__name__))
            time.sleep(0.2)
            transcripts.append(out)
            streams.append(_strip(events))
            c.close(), w.close()
    for i, (a, b) in enumerate(zip(*transcripts)):
        assert a == b, "request %d %r:\n python %r\n native %r" % (i, ops[i], a, b)
    assert streams[0] == streams[1] and len(streams[0]) > 50


@pytest.fixture(scope="module")
def asan_server_binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("asan") / "edl_kv_server_asan")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                        "-o", out, native_server.SOURCE], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build not available: " + r.stderr[-300:])
    return out


def _start(binary, *extra):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    p = subprocess.Popen([binary, "--host", "127.0.0.1", "--port", "0", *extra], stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, env=env)
    line = p.stdout.readline()
    assert line.startswith("listening on "), line
    return p, int(line.strip().rsplit(":", 1)[1])


def test_sanitized_native_server_survives_protocol_and_garbage(asan_server_binary, tmp_path):
    p, port = _start(asan_server_binary, "--data_dir", str(tmp_path / "d"), "--snapshot_interval", "0.1")
    ep = "127.0.0.1:%d" % port
    try:
        c = KVClient(ep)
        rng = random.Random(7)
        c.call({"method": "lease_grant", "ttl": 0.3, "lease_id": 101})
        c.call({"method": "lease_grant", "ttl": 60.0, "lease_id": 102})
        seen = []
        c.add_watch_prefix_callback("/d/", lambda evs, rev: seen.extend(evs), start_revision=1)
        for op in _random_ops(rng, 300):
            try:
                c.call(dict(op))
            except Exception:  # noqa: BLE001
                pass
        time.sleep(0.6)                                  # lease 101 expires: keys deleted, events pushed
        # malformed input on raw sockets: the server answers with an error or drops the connection, never crashes
        frames = [b"\x00\x00\x00\x05\x81\xa1a",                         # truncated map
                  struct.pack("!I", 3) + b"\xc1\xc1\xc1",              # reserved type byte
                  struct.pack("!I", 1 << 31),                          # absurd length prefix
                  struct.pack("!I", 6) + msgpack.packb([1, 2, 3]),     # not a map
                  struct.pack("!I", 2) + b"\x80\x80",                  # trailing bytes after the object
                  struct.pack("!I", 0)]                                # empty frame
        body = msgpack.packb({"method": "put", "key": 5, "value": {"a": 1}, "id": [1]}, use_bin_type=True)
        frames.append(struct.pack("!I", len(body)) + body)             # wrong field types
        body = msgpack.packb({"method": "txn", "compare": "x", "success": 3, "id": 9}, use_bin_type=True)
        frames.append(struct.pack("!I", len(body)) + body)
        deep = b"\x91" * 200 + b"\xc0"
        frames.append(struct.pack("!I", len(deep)) + deep)             # nesting bomb
        for fr in frames:
            s = socket.create_connection(("127.0.0.1", port), timeout=2)
            s.sendall(fr)
            s.settimeout(0.3)
            try:
                s.recv(4096)
            except (socket.timeout, OSError):
                pass
            s.close()
        for _ in range(20):                               # random byte soup
            s = socket.create_connection(("127.0.0.1", port), timeout=2)
            s.sendall(bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 200))))
            s.close()
        assert c.call({"method": "status"})["ok"]         # still serving
        assert c.put("/alive", b"1")["succeeded"]
        c.close()
    finally:
        p.terminate()
        try:
            _, err = p.communicate(timeout=10)
        except subprocess.TimeoutExpired:
            p.kill()
            _, err = p.communicate()
    assert "ERROR: AddressSanitizer" not in err and "runtime error" not in err and "LeakSanitizer" not in err, err[-3000:]
    assert p.returncode == 0, (p.returncode, err[-2000:])
    assert os.path.exists(tmp_path / "d" / "snapshot.bin")
