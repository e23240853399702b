"""The discovery + balance gRPC service (reference: python/edl/distill/discovery_server.py:28-105).

    python -m paddle_edl.distill.discovery_server --server 127.0.0.1:7001 --db_endpoints 127.0.0.1:2379
"""
import argparse
import logging
import time

from ..protos import rpc, schema
from .balance_table import BalanceTable

logger = logging.getLogger("edl.distill.discovery")


class DiscoveryServicer:
    def __init__(self, table: BalanceTable):
        self._table = table

    @staticmethod
    def _response(code, message, version, dversion, servers, dservers):
        r = schema.distill_discovery.Response(version=version, discovery_version=dversion)
        r.status.code = int(code)
        r.status.message = message or ""
        r.servers.extend(servers)
        r.discovery_servers.extend(dservers)
        return r

    def Register(self, request, context):
        return self._response(*self._table.register_client(request.client, request.service_name,
                                                           request.require_num, request.token))

    def HeartBeat(self, request, context):
        return self._response(*self._table.heartbeat(request.client, request.version, request.discovery_version))


class DiscoveryServer:
    def __init__(self, server, db_endpoints, worker_num=4, idle_seconds=7):
        self.server = server
        if isinstance(db_endpoints, str):
            db_endpoints = db_endpoints.split(",")
        self.table = BalanceTable(server, db_endpoints, idle_seconds=idle_seconds)
        self._grpc = None
        self._worker_num = worker_num

    def start(self):
        host, port = self.server.rsplit(":", 1)
        self._grpc = rpc.make_server(max(4, self._worker_num * 4))
        sv = DiscoveryServicer(self.table)
        rpc.add_service(self._grpc, "paddle_edl.distill.DiscoveryService",
                        {"Register": sv.Register, "HeartBeat": sv.HeartBeat})
        bound = self._grpc.add_insecure_port("{}:{}".format("0.0.0.0" if host not in ("127.0.0.1", "localhost") else host, port))
        assert bound > 0, "cannot bind {}".format(self.server)
        if int(port) == 0:
            self.server = "{}:{}".format(host, bound)
            self.table._server = self.server
        self.table.start()
        self._grpc.start()
        logger.info("discovery server %s started", self.server)
        return self

    def stop(self):
        if self._grpc is not None:
            self._grpc.stop(0)
        self.table.stop()

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()


def serve(server, worker_num, db_endpoints):
    srv = DiscoveryServer(server, db_endpoints, worker_num).start()
    try:
        while True:
            time.sleep(3600)
    except KeyboardInterrupt:
        srv.stop()


def main(argv=None):
    ap = argparse.ArgumentParser(description="Discovery server with balance")
    ap.add_argument("--server", type=str, default="127.0.0.1:7001", help="endpoint of this server, ip:port")
    ap.add_argument("--worker_num", type=int, default=1)
    ap.add_argument("--db_endpoints", type=str, default="127.0.0.1:2379", help="registry endpoints, comma separated")
    ap.add_argument("--db_passwd", type=str, default=None)
    ap.add_argument("--db_type", type=str, default="etcd")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    serve(args.server, args.worker_num, args.db_endpoints.split(","))


if __name__ == "__main__":
    main()
