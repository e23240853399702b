"""Self-contained KV store with etcd-v3 semantics (leases, txn, revisions, watches)."""
from .kv_server import KVServer
from .native_server import NativeKVServer
from .client import KVClient, StoreError, StoreRequestError, NoValidEndpoint, Lease, StoreLock

__all__ = ["KVServer", "NativeKVServer", "KVClient", "StoreError", "StoreRequestError", "NoValidEndpoint", "Lease", "StoreLock"]
