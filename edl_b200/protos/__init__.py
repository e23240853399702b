"""Control-plane wire schemas (runtime-built protobuf) and stub-less gRPC helpers."""
from . import schema, rpc  # noqa: F401
