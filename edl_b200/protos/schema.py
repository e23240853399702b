"""Wire schemas of the EDL control-plane RPCs, built at *runtime*.

The reference generates ``*_pb2.py`` with ``grpc_tools.protoc`` (python/edl/protos/run_codegen.py,
generate.sh); ``grpc_tools`` is not available offline, so the same message/service layout (field
names and numbers of python/edl/protos/{common,pod_server,data_server,distill_discovery}.proto --
kept identical for wire compatibility) is declared here with a tiny DSL and turned into real
protobuf message classes through ``descriptor_pb2`` + the default descriptor pool.  The readable
``.proto`` renderings live next to this file for documentation.
"""
from __future__ import annotations

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto
_SCALARS = {
    "string": _F.TYPE_STRING, "bytes": _F.TYPE_BYTES, "int32": _F.TYPE_INT32, "int64": _F.TYPE_INT64,
    "uint64": _F.TYPE_UINT64, "bool": _F.TYPE_BOOL, "float": _F.TYPE_FLOAT, "double": _F.TYPE_DOUBLE,
}


def _build_file(name, package, messages, enums=(), deps=()):
    """messages: {MsgName: [(field_name, number, type, repeated)]}; type is a scalar name or a
    fully-qualified message name starting with '.'."""
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = name
    fd.package = package
    fd.syntax = "proto3"
    for d in deps:
        fd.dependency.append(d)
    for ename, values in enums:
        e = fd.enum_type.add()
        e.name = ename
        for vname, num in values:
            v = e.value.add()
            v.name, v.number = vname, num
    for mname, fields in messages.items():
        m = fd.message_type.add()
        m.name = mname
        for fname, num, ftype, repeated in fields:
            f = m.field.add()
            f.name, f.number = fname, num
            f.label = _F.LABEL_REPEATED if repeated else _F.LABEL_OPTIONAL
            if ftype in _SCALARS:
                f.type = _SCALARS[ftype]
            else:
                f.type = _F.TYPE_MESSAGE
                f.type_name = ftype
    return fd


_pool = descriptor_pool.DescriptorPool()


def _register(fd):
    _pool.Add(fd)
    out = {}
    for m in fd.message_type:
        full = (fd.package + "." if fd.package else "") + m.name
        out[m.name] = message_factory.GetMessageClass(_pool.FindMessageTypeByName(full))
    return out


class _NS:
    def __init__(self, d):
        self.__dict__.update(d)


# ---------------------------------------------------------------------------------- common.proto
common = _NS(_register(_build_file("edl/common.proto", "common", {
    "Status": [("type", 1, "string", False), ("detail", 2, "string", False)],
    "EmptyRet": [("status", 1, ".common.Status", False)],
})))

# ---------------------------------------------------------------------------------- pod_server.proto
pod_server = _NS(_register(_build_file("edl/pod_server.proto", "pod_server", {
    "BarrierRequest": [("job_id", 1, "string", False), ("pod_id", 2, "string", False)],
    "BarrierResponse": [("status", 1, ".common.Status", False), ("cluster_json", 2, "string", False)],
    "ScaleInRequest": [("num", 1, "int32", False)],
    "ScaleOutRequest": [],
}, deps=["edl/common.proto"])))

# ---------------------------------------------------------------------------------- data_server.proto
data_server = _NS(_register(_build_file("edl/data_server.proto", "data_server", {
    "ShutDownRequest": [],
    "EmptyRequest": [],
    "FileListElement": [("idx", 1, "int64", False), ("path", 2, "string", False)],
    "FileListRequest": [("pod_id", 1, "string", False), ("reader_name", 2, "string", False),
                        ("file_list", 3, ".data_server.FileListElement", True)],
    "FileListResponse": [("status", 1, ".common.Status", False),
                         ("file_list", 2, ".data_server.FileListElement", True)],
    "Record": [("record_no", 1, "int64", False), ("field_data", 2, "bytes", True)],
    "BatchData": [("batch_data_id", 1, "string", False), ("records", 2, ".data_server.Record", True)],
    "BatchDataMeta": [("reader_name", 1, "string", False), ("producer_pod_id", 2, "string", False),
                      ("consumer_pod_id", 3, "string", False), ("data_server_endpoint", 4, "string", False),
                      ("batch_data_ids", 5, "string", True)],
    "ReportBatchDataMetaRequest": [("reader_name", 1, "string", False), ("pod_id", 2, "string", False),
                                   ("data_server_endpoint", 3, "string", False),
                                   ("batch_data_ids", 4, "string", True)],
    # ack_seq (extension, field 3): sequence number of the last BatchDataMetaResponse this consumer RECEIVED -- the
    # leader re-sends its previous answer until it is acknowledged, so a lost response loses no batch ids
    "GetBatchDataMetaRequest": [("reader_name", 1, "string", False), ("pod_id", 2, "string", False),
                                ("ack_seq", 3, "uint64", False)],
    "ReachDataEndRequest": [("reader_name", 1, "string", False), ("pod_id", 2, "string", False)],
    "BatchDataMetaResponse": [("status", 1, ".common.Status", False),
                              ("data", 2, ".data_server.BatchDataMeta", True), ("seq", 3, "uint64", False)],
    "BatchDataResponse": [("status", 1, ".common.Status", False), ("data", 2, ".data_server.BatchData", True)],
}, deps=["edl/common.proto"])))

# ---------------------------------------------------------------------------------- distill_discovery.proto
distill_discovery = _NS(_register(_build_file("edl/distill_discovery.proto", "paddle_edl.distill", {
    "Status": [("code", 1, "int32", False), ("message", 2, "string", False)],
    "RegisterRequest": [("client", 1, "string", False), ("service_name", 2, "string", False),
                        ("require_num", 3, "int32", False), ("token", 4, "string", False)],
    "HeartBeatRequest": [("client", 1, "string", False), ("version", 2, "uint64", False),
                         ("discovery_version", 3, "uint64", False)],
    "Response": [("status", 1, ".paddle_edl.distill.Status", False), ("version", 2, "uint64", False),
                 ("discovery_version", 3, "uint64", False), ("servers", 4, "string", True),
                 ("discovery_servers", 5, "string", True)],
}, enums=[("Code", [("OK", 0), ("UNKNOWN", 1), ("NO_READY", 2), ("REDIRECT", 3), ("INVALID_ARGUMENT", 4),
                    ("ALREADY_REGISTER", 5), ("REGISTER_OTHER_SERVICE", 6), ("UNREGISTERED", 7),
                    ("UNAUTHORIZED", 8)])])))


class Code:
    OK, UNKNOWN, NO_READY, REDIRECT, INVALID_ARGUMENT, ALREADY_REGISTER, REGISTER_OTHER_SERVICE, \
        UNREGISTERED, UNAUTHORIZED = range(9)


# ---------------------------------------------------------------------------------- teacher predict RPC
# Not in the reference protos (it delegates to Paddle Serving's brpc); our off-box teacher transport.
predict = _NS(_register(_build_file("edl/predict.proto", "edl.predict", {
    "Tensor": [("name", 1, "string", False), ("dtype", 2, "string", False), ("shape", 3, "int64", True),
               ("data", 4, "bytes", False)],
    "PredictRequest": [("feeds", 1, ".edl.predict.Tensor", True), ("fetch", 2, "string", True)],
    "PredictResponse": [("status", 1, ".common.Status", False), ("outputs", 2, ".edl.predict.Tensor", True)],
    "ConfRequest": [],
    "ConfResponse": [("feed_names", 1, "string", True), ("fetch_names", 2, "string", True),
                     ("feed_shapes_json", 3, "string", False)],
}, deps=["edl/common.proto"])))

# service name -> {method: (request class, response class)}
SERVICES = {
    "pod_server.PodServer": {
        "Barrier": (pod_server.BarrierRequest, pod_server.BarrierResponse),
        "ScaleOut": (pod_server.ScaleOutRequest, common.Status),
        "ScaleIn": (pod_server.ScaleInRequest, common.Status),
    },
    "data_server.DataServer": {
        "ReportBatchDataMeta": (data_server.ReportBatchDataMetaRequest, common.EmptyRet),
        "ReachDataEnd": (data_server.ReachDataEndRequest, common.EmptyRet),
        "GetBatchDataMeta": (data_server.GetBatchDataMetaRequest, data_server.BatchDataMetaResponse),
        "GetFileList": (data_server.FileListRequest, data_server.FileListResponse),
        "GetBatchData": (data_server.BatchDataMeta, data_server.BatchDataResponse),
    },
    "paddle_edl.distill.DiscoveryService": {
        "Register": (distill_discovery.RegisterRequest, distill_discovery.Response),
        "HeartBeat": (distill_discovery.HeartBeatRequest, distill_discovery.Response),
    },
    "edl.predict.PredictService": {
        "Predict": (predict.PredictRequest, predict.PredictResponse),
        "GetConf": (predict.ConfRequest, predict.ConfResponse),
    },
}


# ---------------------------------------------------------------------------------- .proto renderings
_FILES = []   # FileDescriptorProtos in registration order (filled lazily from the pool)
_TYPE_NAMES = {v: k for k, v in _SCALARS.items()}


def render_proto(file_name: str) -> str:
    """Human-readable ``.proto`` text of one runtime-built schema file, services included
    (``python -m edl_b200.protos.schema`` rewrites ``edl_b200/protos/*.proto``)."""
    fdp = descriptor_pb2.FileDescriptorProto()
    _pool.FindFileByName(file_name).CopyToProto(fdp)
    out = ['// Rendered from edl_b200/protos/schema.py -- do not edit.', 'syntax = "proto3";', ""]
    if fdp.package:
        out += ["package %s;" % fdp.package, ""]
    for d in fdp.dependency:
        out.append('import "%s";' % d)
    if fdp.dependency:
        out.append("")
    for e in fdp.enum_type:
        out.append("enum %s {" % e.name)
        out += ["  %s = %d;" % (v.name, v.number) for v in e.value]
        out += ["}", ""]
    for m in fdp.message_type:
        if not m.field:
            out += ["message %s {}" % m.name, ""]
            continue
        out.append("message %s {" % m.name)
        for f in m.field:
            t = f.type_name.lstrip(".") if f.type == _F.TYPE_MESSAGE else _TYPE_NAMES[f.type]
            if fdp.package and t.startswith(fdp.package + "."):
                t = t[len(fdp.package) + 1:]
            out.append("  %s%s %s = %d;" % ("repeated " if f.label == _F.LABEL_REPEATED else "", t, f.name, f.number))
        out += ["}", ""]
    for svc, methods in SERVICES.items():
        pkg, _, sname = svc.rpartition(".")
        if pkg != fdp.package:
            continue
        out.append("service %s {" % sname)
        for mname, (req, resp) in methods.items():
            def short(cls):
                full = cls.DESCRIPTOR.full_name
                return full[len(pkg) + 1:] if full.startswith(pkg + ".") else full
            out.append("  rpc %s(%s) returns (%s) {}" % (mname, short(req), short(resp)))
        out += ["}", ""]
    return "\n".join(out).rstrip() + "\n"


PROTO_FILES = ["edl/common.proto", "edl/pod_server.proto", "edl/data_server.proto", "edl/distill_discovery.proto",
               "edl/predict.proto"]


if __name__ == "__main__":
    import os

    here = os.path.dirname(os.path.abspath(__file__))
    for name in PROTO_FILES:
        path = os.path.join(here, os.path.basename(name))
        with open(path, "w") as f:
            f.write(render_proto(name))
        print("wrote", path)
