"""Shared gRPC client base (reference: python/edl/utils/client.py:18-26)."""
from ..protos import rpc


class Client:
    def __init__(self, endpoint, service):
        self._endpoint = endpoint
        self._channel = rpc.insecure_channel(endpoint)
        self._stub = rpc.Stub(self._channel, service)

    def close(self):
        self._channel.close()
