"""Minimal Hadoop SequenceFile (version 6, uncompressed, BytesWritable key/value) writer and reader.

The CTR example of the reference exports its sparse embedding table as SequenceFiles that a key-value
serving system ("cube") bulk-loads (example/ctr/ctr/{dumper.py,kvtool.py}).  This module implements the
public file format from its specification: header ``SEQ\\x06``, the two Writable class names, two
compression flags, an empty metadata map, a 16-byte sync marker; then records ``[record_len][key_len]
[key][value]`` (big-endian int32 lengths, BytesWritable = 4-byte length + payload) with ``-1`` + sync
marker escapes every ~2000 bytes.
"""
import hashlib
import struct

_BYTES_WRITABLE = b"org.apache.hadoop.io.BytesWritable"
SYNC_INTERVAL = 2000


def _vint(n: int) -> bytes:
    """Hadoop WritableUtils.writeVInt for small non-negative values (class-name lengths)."""
    if -112 <= n <= 127:
        return struct.pack("b", n)
    raw = n.to_bytes((n.bit_length() + 7) // 8, "big")
    return struct.pack("b", -112 - len(raw)) + raw


def _read_vint(f) -> int:
    first = struct.unpack("b", f.read(1))[0]
    if first >= -112:
        return first
    size = -(first + 112) if first >= -120 else -(first + 120)
    val = int.from_bytes(f.read(size), "big")
    return val if first >= -120 else ~val


class SequenceFileWriter:
    def __init__(self, f, sync_seed: bytes = b"edl-b200"):
        self.f = f
        self.sync = hashlib.md5(sync_seed).digest()
        self._since_sync = 0
        f.write(b"SEQ\x06")
        for cls in (_BYTES_WRITABLE, _BYTES_WRITABLE):
            f.write(_vint(len(cls)) + cls)
        f.write(b"\x00\x00")                  # no value compression, no block compression
        f.write(struct.pack(">i", 0))         # empty metadata
        f.write(self.sync)

    def write(self, key: bytes, value: bytes):
        if self._since_sync >= SYNC_INTERVAL:
            self.f.write(struct.pack(">i", -1) + self.sync)
            self._since_sync = 0
        k = struct.pack(">i", len(key)) + key
        v = struct.pack(">i", len(value)) + value
        rec = struct.pack(">ii", len(k) + len(v), len(k)) + k + v
        self.f.write(rec)
        self._since_sync += len(rec)


class SequenceFileReader:
    def __init__(self, f):
        self.f = f
        assert f.read(4) == b"SEQ\x06", "not a version-6 SequenceFile"
        self.key_class = f.read(_read_vint(f))
        self.value_class = f.read(_read_vint(f))
        flags = f.read(2)
        assert flags == b"\x00\x00", "compressed SequenceFiles are not supported"
        for _ in range(struct.unpack(">i", f.read(4))[0]):
            for _ in range(2):                     # Text key / value: vint length + bytes
                f.read(_read_vint(f))
        self.sync = f.read(16)

    def __iter__(self):
        while True:
            head = self.f.read(4)
            if len(head) < 4:
                return
            rec_len = struct.unpack(">i", head)[0]
            if rec_len == -1:
                assert self.f.read(16) == self.sync, "corrupt sync marker"
                continue
            key_len = struct.unpack(">i", self.f.read(4))[0]
            k = self.f.read(key_len)
            v = self.f.read(rec_len - key_len)
            yield k[4:], v[4:]
