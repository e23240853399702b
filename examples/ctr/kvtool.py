#!/usr/bin/env python
"""Key-value file tool of the CTR example (the role of the reference's example/ctr/ctr/kvtool.py:31-351: read and
write the SequenceFiles its sparse-table dumper produces for the key-value serving system).  Built on
``seqfile.py`` (format implemented from the Hadoop SequenceFile specification) plus a plain length-prefixed
"kv" stream for pipes:

    python kvtool.py info  table.seq                      # header, record count, key / value size statistics
    python kvtool.py cat   table.seq [--limit 5] [--keys] # records as  <u64 key>\\t<fp32 values...>
    python kvtool.py write out.seq < pairs.tsv            # lines "key<TAB>v0,v1,..." -> SequenceFile
    python kvtool.py merge out.seq a.seq b.seq ...        # later files win on duplicate keys
    python kvtool.py tokv  table.seq > table.kv           # [u32 klen][key][u32 vlen][value] stream (writekv)
"""
import argparse
import os
import struct
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from seqfile import SequenceFileReader, SequenceFileWriter  # noqa: E402


def writekv(key: bytes, val: bytes, f):
    """One record of the pipe format: 4-byte little-endian lengths in front of key and value."""
    f.write(struct.pack("<I", len(key)) + key + struct.pack("<I", len(val)) + val)


class KvFileReader:
    """Iterates (key, value) of a ``writekv`` stream."""

    def __init__(self, f):
        self.f = f

    def __iter__(self):
        while True:
            head = self.f.read(4)
            if len(head) < 4:
                return
            k = self.f.read(struct.unpack("<I", head)[0])
            v = self.f.read(struct.unpack("<I", self.f.read(4))[0])
            yield k, v


def get_reader(f, type="seqfile"):          # noqa: A002 - the reference's parameter name
    return SequenceFileReader(f) if type == "seqfile" else KvFileReader(f)


def _fmt(key: bytes, val: bytes, keys_only=False):
    k = str(struct.unpack("<Q", key)[0]) if len(key) == 8 else key.hex()
    if keys_only:
        return k
    if len(val) % 4 == 0 and val:
        return k + "\t" + ",".join("%.6g" % x for x in struct.unpack("<%df" % (len(val) // 4), val))
    return k + "\t" + val.hex()


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    for name in ("info", "cat", "tokv"):
        p = sub.add_parser(name)
        p.add_argument("path")
        if name == "cat":
            p.add_argument("--limit", type=int, default=0)
            p.add_argument("--keys", action="store_true")
    p = sub.add_parser("write")
    p.add_argument("out")
    p = sub.add_parser("merge")
    p.add_argument("out")
    p.add_argument("inputs", nargs="+")
    a = ap.parse_args()
    if a.cmd == "info":
        n = kb = vb = 0
        with open(a.path, "rb") as f:
            r = SequenceFileReader(f)
            for k, v in r:
                n, kb, vb = n + 1, kb + len(k), vb + len(v)
            print("key class %s, value class %s" % (r.key_class.decode(), r.value_class.decode()))
        print("%d records, %.1f key bytes and %.1f value bytes on average" % (n, kb / max(n, 1), vb / max(n, 1)))
    elif a.cmd == "cat":
        with open(a.path, "rb") as f:
            for i, (k, v) in enumerate(SequenceFileReader(f)):
                if a.limit and i >= a.limit:
                    break
                print(_fmt(k, v, a.keys))
    elif a.cmd == "tokv":
        with open(a.path, "rb") as f:
            for k, v in SequenceFileReader(f):
                writekv(k, v, sys.stdout.buffer)
    elif a.cmd == "write":
        with open(a.out, "wb") as f:
            w = SequenceFileWriter(f)
            for ln in sys.stdin:
                key, _, vals = ln.rstrip("\n").partition("\t")
                v = [float(x) for x in vals.split(",") if x]
                w.write(struct.pack("<Q", int(key)), struct.pack("<%df" % len(v), *v))
    elif a.cmd == "merge":
        table = {}
        for path in a.inputs:
            with open(path, "rb") as f:
                table.update(SequenceFileReader(f))
        with open(a.out, "wb") as f:
            w = SequenceFileWriter(f)
            for k in sorted(table):
                w.write(k, table[k])
        print("%d records -> %s" % (len(table), a.out))


if __name__ == "__main__":
    main()
