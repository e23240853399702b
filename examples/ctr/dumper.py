#!/usr/bin/env python
"""Export the sparse embedding table of a trained CTR-DNN checkpoint as SequenceFile shards for a
key-value serving system (reference: example/ctr/ctr/dumper.py -- key = feature id, value = embedding
row; a ``donefile`` lists the finished shards so that the loader can pick up complete dumps only).

    python examples/ctr/dumper.py --model_path ./ctr_ckpt --output_dir ./cube_dump --shards 4
"""
import argparse
import json
import os
import struct
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")))
from seqfile import SequenceFileWriter  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model_path", required=True, help="checkpoint directory or .pt state dict of models.ctr_dnn.CtrDnn")
    ap.add_argument("--output_dir", required=True)
    ap.add_argument("--shards", type=int, default=1)
    ap.add_argument("--table", type=int, default=-1, help="-1: every slot table, key = slot << 40 | feature id")
    ap.add_argument("--skip_untouched", action="store_true", help="drop rows that still have their init value 0")
    return ap.parse_args()


def load_state(path):
    if os.path.isdir(path):
        from edl_b200.checkpoint import LocalFS, load_check_point

        tensors, _, _ = load_check_point(path, LocalFS(), trainer_id=0, map_location="cpu")
        assert tensors is not None, "no checkpoint under %s" % path
        return tensors["model"] if "model" in tensors else tensors
    sd = torch.load(path, map_location="cpu")
    return sd.get("model", sd)


def dump():
    args = parse_args()
    sd = load_state(args.model_path)
    tables = sorted(k for k in sd if k.startswith("tables.") and k.endswith(".weight"))
    assert tables, "no embedding tables in the checkpoint"
    if args.table >= 0:
        tables = [tables[args.table]]
    os.makedirs(args.output_dir, exist_ok=True)
    stamp = time.strftime("%Y%m%d%H%M%S")
    files = [open(os.path.join(args.output_dir, "part-%05d" % i), "wb") for i in range(args.shards)]
    writers = [SequenceFileWriter(f) for f in files]
    rows = 0
    for slot, name in enumerate(tables):
        w = sd[name].float()
        keep = (w.abs().sum(1) > 0).nonzero().flatten().tolist() if args.skip_untouched else range(w.shape[0])
        for fid in keep:
            key = (slot << 40) | int(fid)
            writers[key % args.shards].write(struct.pack(">Q", key), w[fid].numpy().astype("<f4").tobytes())
            rows += 1
    for f in files:
        f.close()
    done = {"id": stamp, "key": stamp, "input": os.path.abspath(args.output_dir), "shards": args.shards, "rows": rows,
            "dim": int(sd[tables[0]].shape[1])}
    with open(os.path.join(args.output_dir, "donefile"), "a") as f:
        f.write(json.dumps(done) + "\n")
    print("dumped %d rows into %d shard(s) under %s" % (rows, args.shards, args.output_dir))


if __name__ == "__main__":
    dump()
