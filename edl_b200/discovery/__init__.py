"""Service registry layer: EtcdClient-compatible KV wrapper, consistent hashing, teacher registrar."""
