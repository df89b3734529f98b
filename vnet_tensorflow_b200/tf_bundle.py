"""TensorFlow checkpoint ("tensor bundle") reader and writer in NumPy, without TensorFlow.

The reference saves and restores its variables with `tf.train.Saver` (model.py:689-699,758-764,803-809;
evaluation restores at model.py:1138-1139), which writes two files per checkpoint prefix:

  `<prefix>.data-00000-of-00001`   the raw little-endian bytes of every tensor, back to back;
  `<prefix>.index`                 an immutable sorted string table (the LevelDB table format TensorFlow
                                   vendors under core/lib/io) from variable name to a BundleEntryProto
                                   {dtype, shape, shard_id, offset, size, crc32c}; the key "" holds the
                                   BundleHeaderProto {num_shards, endianness, version}.

TensorFlow is a third-party dependency of the reference and absent here; this file restates the
published on-disk format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/{table,block,format},
tensorflow/core/lib/hash/crc32c, TF 1.15) so that checkpoints of old runs can be resumed by this engine and
checkpoints of this engine can be restored by the reference's Saver.  What is pinned here without TensorFlow:
the CRC (against its published check values and TensorBoard's TF-compatible record CRC) and the nested protos
(against the TF .proto classes TensorBoard ships) - tests/test_tf_bundle.py.  Whole files written by a real
TensorFlow are not available in this environment; DESIGN.md §2 says so.
"""
from __future__ import annotations

import os
import struct
import sys
from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

# ---------------------------------------------------------------------------------------------------------------
# CRC-32C (Castagnoli), as tensorflow/core/lib/hash/crc32c: reflected polynomial 0x82f63b78, initial value and
# final xor 0xffffffff; stored "masked" so that a CRC of bytes that embed CRCs stays well distributed.
# ---------------------------------------------------------------------------------------------------------------
_POLY = 0x82F63B78
_MASK_DELTA = 0xA282EAD8


def _make_table() -> np.ndarray:
    t = np.arange(256, dtype=np.uint32)
    for _ in range(8):
        t = np.where(t & 1, (t >> 1) ^ np.uint32(_POLY), t >> 1).astype(np.uint32)
    return t


_TABLE = _make_table()
_TABLE_LIST = [int(v) for v in _TABLE]
_LOCKSTEP_MIN = 1 << 16  # below this a byte loop is quicker than setting up the lanes


def _raw_update_loop(r: int, data: bytes) -> int:
    tab = _TABLE_LIST
    for b in data:
        r = tab[(r ^ b) & 0xFF] ^ (r >> 8)
    return r


def _raw_update_lockstep(r: int, data: np.ndarray) -> int:
    """The register after `data` (a whole number of equal chunks), computed with one lane per chunk.

    The register update is linear over GF(2): reg(r, A||B) = Z(reg(r, A)) ^ reg(0, B) where Z feeds len(B) zero
    bytes.  All chunks run in lock step from a zero register; 32 extra lanes fed with zeros from the unit
    vectors give Z for the chunk length in the same loop; the chunks are then chained in order."""
    n_lanes, chunk = data.shape
    cols = np.zeros((chunk, n_lanes + 32), np.uint8)
    cols[:, :n_lanes] = data.T
    s = np.zeros(n_lanes + 32, np.uint32)
    s[n_lanes:] = np.uint32(1) << np.arange(32, dtype=np.uint32)
    for k in range(chunk):
        s = _TABLE[(s ^ cols[k]) & np.uint32(0xFF)] ^ (s >> np.uint32(8))
    basis = [int(v) for v in s[n_lanes:]]
    ztab = []
    for byte in range(4):  # Z as four 256-entry tables, one per register byte
        tab = [0] * 256
        for v in range(256):
            acc = 0
            for bit in range(8):
                if v >> bit & 1:
                    acc ^= basis[8 * byte + bit]
            tab[v] = acc
        ztab.append(tab)
    z0, z1, z2, z3 = ztab
    for lane in s[:n_lanes].tolist():
        r = z0[r & 0xFF] ^ z1[(r >> 8) & 0xFF] ^ z2[(r >> 16) & 0xFF] ^ z3[r >> 24] ^ lane
    return r


def crc32c(data, crc: int = 0) -> int:
    """crc32c::Extend(crc, data): the CRC-32C of the concatenation whose CRC so far is `crc`."""
    buf = memoryview(data).cast("B")
    n = len(buf)
    r = (crc ^ 0xFFFFFFFF) & 0xFFFFFFFF
    if n >= _LOCKSTEP_MIN:
        chunk = 1 << max(8, (n.bit_length() + 1) // 2)  # about sqrt(n): balances the two loops
        lanes = n // chunk
        body = np.frombuffer(buf, np.uint8, lanes * chunk).reshape(lanes, chunk)
        r = _raw_update_lockstep(r, body)
        buf = buf[lanes * chunk:]
    r = _raw_update_loop(r, bytes(buf))
    return r ^ 0xFFFFFFFF


def crc_mask(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF


def crc_unmask(masked: int) -> int:
    rot = (masked - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------------------------
# Protocol-buffer wire helpers for the three small messages of tensor_bundle.proto.
# ---------------------------------------------------------------------------------------------------------------
def _varint(n: int) -> bytes:
    if n < 0:
        n += 1 << 64  # two's complement, ten bytes, as protobuf encodes negative int64
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def _read_varint(buf, pos: int) -> Tuple[int, int]:
    shift = 0
    val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if b < 0x80:
            return val, pos
        shift += 7
        if shift > 63:
            raise ValueError("malformed varint")


def _fields(buf) -> Iterable[Tuple[int, int, object]]:
    """(field number, wire type, value) for every field of a serialized message."""
    pos = 0
    n = len(buf)
    while pos < n:
        tag, pos = _read_varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _read_varint(buf, pos)
        elif wt == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            val = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            val = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield num, wt, val


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= 1 << 63 else v


# tensorflow/core/framework/types.proto
_DTYPE_TO_NP = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
                10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_NP_TO_DTYPE = {np.dtype(v): k for k, v in _DTYPE_TO_NP.items()}


def encode_shape(dims: Iterable[int]) -> bytes:
    """TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}."""
    out = b""
    for d in dims:
        dim = (b"\x08" + _varint(int(d))) if d else b""
        out += b"\x12" + _varint(len(dim)) + dim
    return out


def decode_shape(buf: bytes) -> Tuple[int, ...]:
    dims = []
    for num, _, val in _fields(buf):
        if num == 2:
            size = 0
            for n2, _, v2 in _fields(val):
                if n2 == 1:
                    size = _signed64(v2)
            dims.append(size)
        elif num == 3 and val:
            raise ValueError("tensor of unknown rank in a checkpoint")
    return tuple(dims)


class BundleEntry:
    """BundleEntryProto: dtype = 1, shape = 2, shard_id = 3, offset = 4, size = 5, fixed32 crc32c = 6, slices = 7."""
    __slots__ = ("dtype", "shape", "shard_id", "offset", "size", "crc32c", "sliced")

    def __init__(self, dtype=0, shape=(), shard_id=0, offset=0, size=0, crc=0, sliced=False):
        self.dtype, self.shape, self.shard_id, self.offset, self.size, self.crc32c = dtype, tuple(shape), shard_id, offset, size, crc
        self.sliced = sliced

    def encode(self) -> bytes:
        out = b""
        if self.dtype:
            out += b"\x08" + _varint(self.dtype)
        shp = encode_shape(self.shape)
        out += b"\x12" + _varint(len(shp)) + shp  # the writer always touches mutable_shape(), scalars included
        if self.shard_id:
            out += b"\x18" + _varint(self.shard_id)
        if self.offset:
            out += b"\x20" + _varint(self.offset)
        if self.size:
            out += b"\x28" + _varint(self.size)
        if self.crc32c:
            out += b"\x35" + struct.pack("<I", self.crc32c)
        return out

    @classmethod
    def decode(cls, buf: bytes) -> "BundleEntry":
        e = cls()
        for num, _, val in _fields(buf):
            if num == 1:
                e.dtype = val
            elif num == 2:
                e.shape = decode_shape(val)
            elif num == 3:
                e.shard_id = val
            elif num == 4:
                e.offset = _signed64(val)
            elif num == 5:
                e.size = _signed64(val)
            elif num == 6:
                e.crc32c = val
            elif num == 7:
                e.sliced = True
        return e


def encode_header(num_shards: int = 1, producer: int = 1) -> bytes:
    """BundleHeaderProto: num_shards = 1, endianness = 2 (LITTLE = 0, omitted), VersionDef version = 3 {producer = 1}."""
    ver = b"\x08" + _varint(producer)
    return b"\x08" + _varint(num_shards) + b"\x1a" + _varint(len(ver)) + ver


def decode_header(buf: bytes) -> Dict[str, int]:
    h = {"num_shards": 0, "endianness": 0, "producer": 0}
    for num, _, val in _fields(buf):
        if num == 1:
            h["num_shards"] = val
        elif num == 2:
            h["endianness"] = val
        elif num == 3:
            for n2, _, v2 in _fields(val):
                if n2 == 1:
                    h["producer"] = v2
    return h


# ---------------------------------------------------------------------------------------------------------------
# The table file (tensorflow/core/lib/io/table_builder.cc, block_builder.cc, format.cc).
#   file   := data_block* metaindex_block index_block footer
#   block  := entry* restart_offset:u32* n_restarts:u32 | type:u8 (0 = uncompressed) | masked crc32c(block+type):u32
#   entry  := shared:varint32 non_shared:varint32 value_len:varint32 key[shared:] value
#   footer := metaindex BlockHandle, index BlockHandle (varint64 offset, size), zero padding to 40 bytes, magic:u64
# ---------------------------------------------------------------------------------------------------------------
_MAGIC = 0xDB4775248B80FB57
_BLOCK_SIZE = 262144        # table::Options defaults, which BundleWriter keeps
_RESTART_INTERVAL = 16
_FOOTER = 48


class _BlockBuilder:
    def __init__(self, restart_interval: int):
        self.interval = restart_interval
        self.buf = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.last_key = b""

    def empty(self) -> bool:
        return not self.buf

    def size_estimate(self) -> int:
        return len(self.buf) + 4 * len(self.restarts) + 4

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.counter < self.interval:
            lim = min(len(key), len(self.last_key))
            while shared < lim and key[shared] == self.last_key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        self.last_key = key
        self.counter += 1

    def finish(self) -> bytes:
        return bytes(self.buf) + struct.pack("<%dI" % len(self.restarts), *self.restarts) + struct.pack("<I", len(self.restarts))


def _shortest_separator(start: bytes, limit: bytes) -> bytes:
    """BytewiseComparator::FindShortestSeparator: a short key in [start, limit)."""
    n = min(len(start), len(limit))
    i = 0
    while i < n and start[i] == limit[i]:
        i += 1
    if i < n and start[i] < 0xFF and start[i] + 1 < limit[i]:
        return start[:i] + bytes([start[i] + 1])
    return start


def _short_successor(key: bytes) -> bytes:
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def write_table(path: str, items: List[Tuple[bytes, bytes]], block_size: int = _BLOCK_SIZE):
    """Write sorted (key, value) pairs as an uncompressed table, as table::TableBuilder does."""
    out = bytearray()

    def write_block(raw: bytes) -> bytes:
        handle = _varint(len(out)) + _varint(len(raw))
        out.extend(raw)
        out.extend(b"\x00" + struct.pack("<I", crc_mask(crc32c(b"\x00", crc32c(raw)))))
        return handle

    data = _BlockBuilder(_RESTART_INTERVAL)
    index = _BlockBuilder(1)
    pending: Optional[bytes] = None  # handle of a flushed data block whose index key waits for the next key
    last_key = b""
    for n, (key, value) in enumerate(items):
        if n and key <= last_key:
            raise ValueError("table keys must be strictly increasing")
        if pending is not None:
            index.add(_shortest_separator(last_key, key), pending)
            pending = None
        data.add(key, value)
        last_key = key
        if data.size_estimate() >= block_size:
            pending = write_block(data.finish())
            data = _BlockBuilder(_RESTART_INTERVAL)
    if not data.empty():
        pending = write_block(data.finish())
    meta_handle = write_block(_BlockBuilder(_RESTART_INTERVAL).finish())
    if pending is not None:
        index.add(_short_successor(last_key), pending)
    index_handle = write_block(index.finish())
    footer = meta_handle + index_handle
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC))
    with open(path, "wb") as f:
        f.write(out)


def _read_block(buf: bytes, offset: int, size: int, verify: bool) -> bytes:
    raw = buf[offset:offset + size]
    if len(raw) != size or offset + size + 5 > len(buf):
        raise ValueError("truncated table block")
    kind = buf[offset + size]
    if verify:
        stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
        if crc_unmask(stored) != crc32c(buf[offset:offset + size + 1]):
            raise ValueError("table block checksum mismatch")
    if kind != 0:
        raise ValueError("compressed table block (type %d); tensor bundles are written uncompressed" % kind)
    return raw


def _block_entries(raw: bytes) -> Iterable[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack_from("<I", raw, len(raw) - 4)[0]
    end = len(raw) - 4 - 4 * n_restarts
    pos = 0
    key = b""
    while pos < end:
        shared, pos = _read_varint(raw, pos)
        non_shared, pos = _read_varint(raw, pos)
        vlen, pos = _read_varint(raw, pos)
        key = key[:shared] + raw[pos:pos + non_shared]
        pos += non_shared
        yield key, raw[pos:pos + vlen]
        pos += vlen


def read_table(path: str, verify: bool = True) -> "OrderedDict[bytes, bytes]":
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < _FOOTER or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != _MAGIC:
        raise ValueError("%s is not a TensorFlow table file (bad magic number)" % path)
    foot = buf[len(buf) - _FOOTER:]
    pos = 0
    _, pos = _read_varint(foot, pos)
    _, pos = _read_varint(foot, pos)
    i_off, pos = _read_varint(foot, pos)
    i_size, pos = _read_varint(foot, pos)
    out: "OrderedDict[bytes, bytes]" = OrderedDict()
    for _, handle in _block_entries(_read_block(buf, i_off, i_size, verify)):
        off, p = _read_varint(handle, 0)
        size, _ = _read_varint(handle, p)
        for key, value in _block_entries(_read_block(buf, off, size, verify)):
            out[key] = value
    return out


# ---------------------------------------------------------------------------------------------------------------
# The bundle (tensorflow/core/util/tensor_bundle/tensor_bundle.cc).
# ---------------------------------------------------------------------------------------------------------------
def data_path(prefix: str, shard: int = 0, num_shards: int = 1) -> str:
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


def index_path(prefix: str) -> str:
    return prefix + ".index"


def is_bundle(prefix: str) -> bool:
    return os.path.exists(index_path(prefix))


class BundleReader:
    """Variables of a checkpoint prefix by name (BundleReader::Lookup)."""

    def __init__(self, prefix: str, verify: bool = True):
        self.prefix = prefix
        self.verify = verify
        table = read_table(index_path(prefix), verify)
        if b"" not in table:
            raise ValueError("%s has no bundle header" % index_path(prefix))
        self.header = decode_header(table.pop(b""))
        if self.header["endianness"] != 0:
            raise ValueError("big-endian tensor bundle")
        self.entries: "OrderedDict[str, BundleEntry]" = OrderedDict(
            (k.decode(), BundleEntry.decode(v)) for k, v in table.items())
        self._shards: Dict[int, np.memmap] = {}

    def keys(self) -> List[str]:
        return list(self.entries)

    def __contains__(self, name: str) -> bool:
        return name in self.entries

    def shape_and_dtype(self, name: str):
        e = self.entries[name]
        return e.shape, np.dtype(_DTYPE_TO_NP[e.dtype]) if e.dtype in _DTYPE_TO_NP else None

    def _shard(self, shard_id: int):
        if shard_id not in self._shards:
            path = data_path(self.prefix, shard_id, self.header["num_shards"])
            self._shards[shard_id] = np.memmap(path, np.uint8, "r") if os.path.getsize(path) else np.zeros(0, np.uint8)
        return self._shards[shard_id]

    def get_tensor(self, name: str) -> np.ndarray:
        if name not in self.entries:
            raise KeyError("Key %s not found in checkpoint %s" % (name, self.prefix))
        e = self.entries[name]
        if e.sliced:
            raise ValueError("%s is a partitioned variable; the reference never saves one" % name)
        if e.dtype not in _DTYPE_TO_NP:
            raise ValueError("%s has TensorFlow dtype %d, which this reader does not decode" % (name, e.dtype))
        dt = np.dtype(_DTYPE_TO_NP[e.dtype])
        count = int(np.prod(e.shape, dtype=np.int64))
        if count * dt.itemsize != e.size:
            raise ValueError("%s: %d bytes stored for shape %s of %s" % (name, e.size, e.shape, dt))
        raw = self._shard(e.shard_id)[e.offset:e.offset + e.size]
        if raw.size != e.size:
            raise ValueError("%s: data file is shorter than the index says" % name)
        raw = np.ascontiguousarray(raw)
        if self.verify and crc_unmask(e.crc32c) != crc32c(raw):
            raise ValueError("%s: tensor checksum mismatch" % name)
        return raw.view(dt).reshape(e.shape).copy()


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray]):
    """One-shard bundle of `tensors` (BundleWriter::Add in key order, then Finish)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items: List[Tuple[bytes, bytes]] = [(b"", encode_header(1))]
    offset = 0
    tmp = data_path(prefix) + ".tempstate"
    with open(tmp, "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode()):
            if not name:
                raise ValueError("the empty key is reserved for the bundle header")
            a = np.asarray(tensors[name], order="C")  # (ascontiguousarray would turn a scalar into shape [1])
            if a.dtype not in _NP_TO_DTYPE:
                raise ValueError("%s: dtype %s has no TensorFlow counterpart here" % (name, a.dtype))
            if a.dtype.byteorder == ">":
                a = a.astype(a.dtype.newbyteorder("<"))
            raw = a.tobytes()
            f.write(raw)
            entry = BundleEntry(_NP_TO_DTYPE[a.dtype], a.shape, 0, offset, len(raw), crc_mask(crc32c(raw)))
            items.append((name.encode(), entry.encode()))
            offset += len(raw)
    os.replace(tmp, data_path(prefix))
    write_table(index_path(prefix) + ".tempstate", items)
    os.replace(index_path(prefix) + ".tempstate", index_path(prefix))


def _main(argv: List[str]) -> int:
    if len(argv) == 2 and argv[0] == "list":
        r = BundleReader(argv[1])
        for k, e in r.entries.items():
            print("%-90s %-8s %s" % (k, np.dtype(_DTYPE_TO_NP.get(e.dtype, np.void)).name, list(e.shape)))
        return 0
    if len(argv) == 3 and argv[0] == "to-npz":
        r = BundleReader(argv[1])
        np.savez(argv[2], **{k: r.get_tensor(k) for k in r.keys()})
        return 0
    if len(argv) == 3 and argv[0] == "from-npz":
        with np.load(argv[1]) as z:
            write_bundle(argv[2], {k: z[k] for k in z.files})
        return 0
    print("usage: python -m vnet_tensorflow_b200.tf_bundle list PREFIX | to-npz PREFIX OUT.npz | from-npz IN.npz PREFIX",
          file=sys.stderr)
    return 2


if __name__ == "__main__":
    sys.exit(_main(sys.argv[1:]))
