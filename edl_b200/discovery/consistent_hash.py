"""Consistent hashing of keys (service names) onto nodes (discovery servers).

Behavioural contract of the reference (python/edl/discovery/consistent_hash.py:21-141): MD5 ring,
300 virtual nodes per node, lock-free reads with a single writer, ``get_node_nodes`` returns
``(node, all_nodes, version)``.  Implementation here: every mutation builds a new immutable
``_Ring`` (sorted slot tuple + parallel owner tuple) and swaps one reference, so readers never see
a half-updated ring and lookups are a single ``bisect``.
"""
from __future__ import annotations

import bisect
import hashlib
from typing import Iterable, List, Optional, Tuple


def _slot(key: str) -> int:
    return int.from_bytes(hashlib.md5(key.encode("utf-8")).digest(), "big")


class _Ring:
    __slots__ = ("nodes", "slots", "owners", "version")

    def __init__(self, nodes: Iterable[str], virtual_num: int, version: int):
        self.nodes: Tuple[str, ...] = tuple(nodes)
        table = {}
        for node in self.nodes:
            for i in range(virtual_num):
                s = _slot("%s-v%d" % (node, i))
                prev = table.get(s)
                if prev is None or node < prev:  # deterministic tie-break on (improbable) collisions
                    table[s] = node
        self.slots = tuple(sorted(table))
        self.owners = tuple(table[s] for s in self.slots)
        self.version = version

    def lookup(self, key: str) -> Optional[str]:
        if not self.slots:
            return None
        i = bisect.bisect_left(self.slots, _slot(key))
        if i == len(self.slots):
            i = 0
        return self.owners[i]


class ConsistentHash:
    """One writer thread, any number of reader threads, no locks."""

    def __init__(self, nodes: Iterable[str] = (), virtual_num: int = 300):
        self._virtual_num = virtual_num
        self._ring = _Ring(list(dict.fromkeys(nodes)), virtual_num, 1)

    def add_new_node(self, node: str) -> None:
        ring = self._ring
        if node in ring.nodes:
            return
        self._ring = _Ring(ring.nodes + (node,), self._virtual_num, ring.version + 1)

    def remove_node(self, node: str) -> None:
        ring = self._ring
        if node not in ring.nodes:
            return
        self._ring = _Ring([n for n in ring.nodes if n != node], self._virtual_num, ring.version + 1)

    def get_node(self, key: str) -> Optional[str]:
        return self._ring.lookup(key)

    def get_node_nodes(self, key: str):
        ring = self._ring
        return ring.lookup(key), list(ring.nodes), ring.version

    @property
    def nodes(self) -> List[str]:
        return list(self._ring.nodes)

    @property
    def version(self) -> int:
        return self._ring.version
