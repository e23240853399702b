"""Teacher discovery + balancing (etcd flavour): balancer properties, register / heartbeat protocol,
teacher up/down propagation, client timeout, redirect between two discovery servers."""
import time


from edl_b200.discovery.etcd_client import EtcdClient
from edl_b200.distill.balance_table import Service
from edl_b200.distill.discovery_client import DiscoveryClient
from edl_b200.distill.discovery_server import DiscoveryServer
from edl_b200.utils.network_utils import find_free_ports


def test_balancer_properties():
    svc = Service("s")
    svc.update_servers(add=["t%d" % i for i in range(8)])
    for i in range(3):
        svc.add_client("c%d" % i, require_num=4)
    assign = {c: svc.snapshot(c)[1] for c in ("c0", "c1", "c2")}
    assert all(len(v) == 2 for v in assign.values())          # floor(8/3) = 2 each
    assert len(set(sum(assign.values(), []))) == 6            # no sharing while servers are plentiful
    v0 = svc.snapshot("c0")[0]
    svc.update_servers(rm=[assign["c0"][0]])                  # one of c0's teachers dies
    ver, servers = svc.snapshot("c0")
    assert ver == v0 + 1 and len(servers) == 2 and assign["c0"][0] not in servers
    others_before = {c: svc.snapshot(c) for c in ("c1", "c2")}
    svc.add_client("c3", 1)                                   # 7 servers / 4 clients -> 1 each
    assert all(len(svc.snapshot(c)[1]) == 1 for c in ("c0", "c1", "c2", "c3"))
    # more clients than servers: servers are shared evenly
    svc2 = Service("x")
    svc2.update_servers(add=["a", "b"])
    for i in range(5):
        svc2.add_client("c%d" % i, 3)
    load = {}
    for i in range(5):
        (s,) = svc2.snapshot("c%d" % i)[1]
        load[s] = load.get(s, 0) + 1
    assert sorted(load.values()) == [2, 3]
    del others_before


def test_register_heartbeat_and_teacher_changes(kv_server):
    reg = EtcdClient([kv_server.endpoint], root="service")
    reg.init()
    for t in ("10.0.0.1:9000", "10.0.0.2:9000"):
        reg.set_server_permanent("Teacher", t, "info")
    port = find_free_ports(1)[0]
    with DiscoveryServer("127.0.0.1:%d" % port, [kv_server.endpoint], idle_seconds=2) as srv:
        c1 = DiscoveryClient([srv.server], "Teacher", require_num=2, heartbeat_s=0.2).start()
        assert sorted(c1.get_servers()) == ["10.0.0.1:9000", "10.0.0.2:9000"]
        c2 = DiscoveryClient([srv.server], "Teacher", require_num=2, heartbeat_s=0.2).start()
        time.sleep(0.8)
        assert len(c1.get_servers()) == 1 and len(c2.get_servers()) == 1
        assert set(c1.get_servers()) | set(c2.get_servers()) == {"10.0.0.1:9000", "10.0.0.2:9000"}
        reg.set_server_permanent("Teacher", "10.0.0.3:9000", "info")   # a teacher joins
        reg.set_server_permanent("Teacher", "10.0.0.4:9000", "info")
        deadline = time.time() + 5
        while time.time() < deadline and not (len(c1.get_servers()) == 2 and len(c2.get_servers()) == 2):
            time.sleep(0.1)
        assert len(c1.get_servers()) == 2 and len(c2.get_servers()) == 2
        assert not set(c1.get_servers()) & set(c2.get_servers())
        gone = c1.get_servers()[0]
        reg.remove_server("Teacher", gone)                              # a teacher leaves
        deadline = time.time() + 5
        while time.time() < deadline and gone in c1.get_servers():
            time.sleep(0.1)
        assert gone not in c1.get_servers()
        c2.stop()                                                       # a student disappears silently
        deadline = time.time() + 8
        while time.time() < deadline and len(c1.get_servers()) < 2:
            time.sleep(0.1)
        assert len(c1.get_servers()) == 2                               # its teachers are re-assigned
        c1.stop()


def test_two_discovery_servers_redirect(kv_server):
    reg = EtcdClient([kv_server.endpoint], root="service")
    reg.init()
    reg.set_server_permanent("SvcA", "t:1", "i")
    p1, p2 = find_free_ports(2)
    with DiscoveryServer("127.0.0.1:%d" % p1, [kv_server.endpoint]) as s1, \
            DiscoveryServer("127.0.0.1:%d" % p2, [kv_server.endpoint]) as s2:
        # enough names that the ring cannot put all of them on one server whatever the (random) ports hash to
        names = ("SvcA",) + tuple("Svc%02d" % i for i in range(40))
        deadline = time.time() + 15            # let both learn about each other (watch-driven, no fixed sleep)
        while time.time() < deadline:
            if all(s1.table._owner(n)[0] == s2.table._owner(n)[0] for n in names) and \
                    {s1.table._owner(n)[0] for n in names} == {s1.server, s2.server}:
                break
            time.sleep(0.1)
        owners = set()
        for name in names:
            owners.add(s1.table._owner(name)[0])
            assert s1.table._owner(name)[0] == s2.table._owner(name)[0]
        assert owners == {s1.server, s2.server}                         # the ring spreads services
        wrong = s2.server if s1.table._owner("SvcA")[0] == s1.server else s1.server
        c = DiscoveryClient([wrong], "SvcA", 1, heartbeat_s=0.2).start()   # must follow the REDIRECT
        assert c.get_servers() == ["t:1"]
        c.stop()
