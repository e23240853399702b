"""Protobuf <-> python helpers for the data-server messages (reference: python/edl/utils/pb_utils.py)."""
from ..protos import schema


def file_list_to_pb(file_list):
    """[(idx, path)] or [path] -> repeated FileListElement"""
    out = []
    for i, item in enumerate(file_list):
        idx, path = item if isinstance(item, (tuple, list)) else (i, item)
        out.append(schema.data_server.FileListElement(idx=int(idx), path=str(path)))
    return out


def file_list_from_pb(pb_list):
    return [(int(e.idx), e.path) for e in pb_list]


def batch_data_meta_to_dict(meta):
    return {"reader_name": meta.reader_name, "producer_pod_id": meta.producer_pod_id,
            "consumer_pod_id": meta.consumer_pod_id, "data_server_endpoint": meta.data_server_endpoint,
            "batch_data_ids": list(meta.batch_data_ids)}
