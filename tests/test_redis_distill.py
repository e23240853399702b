"""Redis flavour: RESP store with expiry, teacher registrar, balance server framing / protocol, and a
DistillReader driven end-to-end through it (reference: test_redis_distill_reader.sh:19-41 -- there the
"teacher" is just a TCP port the registrar can probe; same trick here)."""
import json
import socket
import time

import numpy as np

from edl_b200.distill import distill_worker
from edl_b200.distill.distill_reader import DistillReader
from edl_b200.distill.redis.balance_server import HEADER, MAGIC, BalanceServer, pack_frame
from edl_b200.distill.redis.client import Client
from edl_b200.distill.redis.redis_store import RedisStore
from edl_b200.distill.redis.resp import MiniRedisServer, RespClient
from edl_b200.distill.redis.server_register import ServerRegister


def test_resp_store_expiry():
    with MiniRedisServer() as rs:
        c = RespClient(rs.host, rs.port)
        assert c.ping()
        st = RedisStore(rs.host, rs.port, ttl=1)
        st.set_server("svc", "1.1.1.1:1", "a")
        st.set_server("svc", "1.1.1.1:2", "b")
        assert sorted(s["server"] for s in st.get_service("svc")) == ["1.1.1.1:1", "1.1.1.1:2"]
        for _ in range(4):
            time.sleep(0.4)
            st.refresh("svc", "1.1.1.1:1")
        assert [s["server"] for s in st.get_service("svc")] == ["1.1.1.1:1"]
        st.remove_server("svc", "1.1.1.1:1")
        assert st.get_service("svc") == []


def test_balance_server_protocol_and_reader():
    with MiniRedisServer() as rs:
        # the registrar only TCP-probes the "teacher": point it at the redis port itself
        teacher = rs.endpoint
        reg = ServerRegister(rs.host, rs.port, "TestService", teacher, ttl=2, heartbeat=0.3).register(block=False)
        time.sleep(0.5)
        with BalanceServer("127.0.0.1", 0, rs.host, rs.port) as bs:
            ep = "127.0.0.1:%d" % bs.port
            # raw protocol: bad magic closes the connection
            s = socket.create_connection(("127.0.0.1", bs.port))
            s.sendall(HEADER.pack(b"\x00\x00\x00\x00", 12) + b"{}  ")
            s.settimeout(2)
            assert s.recv(16) == b""
            s.close()
            # raw protocol: register + heartbeat
            s = socket.create_connection(("127.0.0.1", bs.port))
            s.sendall(pack_frame({"type": "register", "service_name": "TestService", "seq": 0, "num": 1}))
            magic, total = HEADER.unpack(s.recv(HEADER.size))
            body = json.loads(s.recv(total - HEADER.size))
            assert magic == MAGIC and body["type"] == "register" and body["seq"] == 1 and body["servers"] == [teacher]
            s.sendall(pack_frame({"type": "heartbeat", "version": body["version"]}))
            magic, total = HEADER.unpack(s.recv(HEADER.size))
            assert json.loads(s.recv(total - HEADER.size)) == {"type": "heartbeat"}
            s.close()
            # client object
            c = Client([ep], "TestService", 1, heartbeat_s=0.2).start()
            assert c.get_servers() == [teacher]
            c.stop()
            # DistillReader through the redis discover (NOP teacher)
            distill_worker._NOP_PREDICT_TEST = True
            try:
                import edl_b200.distill.distill_reader as drmod
                drmod._service_discover = None
                dr = DistillReader(ins=["image", "label"], predicts=["score"])
                dr.set_teacher_batch_size(4)
                dr.set_dynamic_teacher([ep], "TestService", require_max_teacher=2)
                def gen():
                    for b in range(10):
                        yield [(np.zeros((2, 2), np.float32), np.array([b * 8 + i])) for i in range(8)]
                r = dr.set_sample_list_generator(gen)
                for epoch in range(3):
                    labels = [int(s[1][0]) for batch in r() for s in batch]
                    assert labels == list(range(80))
                dr.stop()
                drmod._service_discover.stop()
                drmod._service_discover = None
            finally:
                distill_worker._NOP_PREDICT_TEST = False
        reg.stop()
