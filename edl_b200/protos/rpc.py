"""gRPC plumbing without generated stubs: servers are assembled from
``grpc.method_handlers_generic_handler`` and clients from ``channel.unary_unary`` using the message
classes of :mod:`edl_b200.protos.schema`.  Message size caps follow the reference's 1 GiB
(utils/pod_server.py:130-157)."""
from __future__ import annotations

from concurrent import futures
from typing import Callable, Dict

import grpc

from .schema import SERVICES

MAX_MSG = 1024 * 1024 * 1024
CHANNEL_OPTIONS = [("grpc.max_send_message_length", MAX_MSG), ("grpc.max_receive_message_length", MAX_MSG)]


def make_server(max_workers: int = 20) -> grpc.Server:
    return grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers), options=CHANNEL_OPTIONS)


def add_service(server: grpc.Server, service: str, handlers: Dict[str, Callable]) -> None:
    """handlers: {method_name: fn(request, context) -> response}"""
    spec = SERVICES[service]
    rpc_handlers = {}
    for method, fn in handlers.items():
        req_cls, resp_cls = spec[method]
        rpc_handlers[method] = grpc.unary_unary_rpc_method_handler(
            fn, request_deserializer=req_cls.FromString, response_serializer=resp_cls.SerializeToString)
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(service, rpc_handlers),))


class Stub:
    """Client for one service: ``Stub(channel, 'pod_server.PodServer').Barrier(req, timeout=...)``."""

    def __init__(self, channel: grpc.Channel, service: str):
        for method, (req_cls, resp_cls) in SERVICES[service].items():
            setattr(self, method, channel.unary_unary(
                "/%s/%s" % (service, method), request_serializer=req_cls.SerializeToString,
                response_deserializer=resp_cls.FromString))


def insecure_channel(endpoint: str) -> grpc.Channel:
    return grpc.insecure_channel(endpoint, options=CHANNEL_OPTIONS)
