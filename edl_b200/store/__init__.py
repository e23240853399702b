"""Self-contained KV store with etcd-v3 semantics (leases, txn, revisions, watches)."""
from .kv_server import KVServer
from .client import KVClient, StoreError, NoValidEndpoint, Lease, StoreLock

__all__ = ["KVServer", "KVClient", "StoreError", "NoValidEndpoint", "Lease", "StoreLock"]
