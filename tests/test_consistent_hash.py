"""Same properties the reference asserts (tests/unittests/test_consistent_hash.py:22-77)."""
from edl_b200.discovery.consistent_hash import ConsistentHash


def _spread(ch, n=10000):
    out = {}
    for i in range(n):
        out.setdefault(ch.get_node("key-%d" % i), []).append(i)
    return out


def test_balance_and_monotonicity():
    nodes = ["127.0.0.1:%d" % p for p in (7001, 7002, 7003)]
    ch = ConsistentHash(nodes)
    before = _spread(ch)
    assert all(len(v) > 2500 for v in before.values()) and len(before) == 3
    owner = {i: n for n, ks in before.items() for i in ks}
    ch.remove_node(nodes[1])
    after = _spread(ch)
    assert nodes[1] not in after
    for n in (nodes[0], nodes[2]):  # keys of surviving nodes never move
        assert set(before[n]) <= set(after[n])
    ch.add_new_node(nodes[1])
    again = _spread(ch)
    assert {i: n for n, ks in again.items() for i in ks} == owner
    ch.add_new_node("127.0.0.1:7004")
    four = _spread(ch)
    assert len(four["127.0.0.1:7004"]) < 3500
    node, all_nodes, version = ch.get_node_nodes("x")
    assert node in all_nodes and version == 4


def test_empty():
    ch = ConsistentHash([])
    assert ch.get_node("a") is None
    assert ch.get_node_nodes("a") == (None, [], 1)
