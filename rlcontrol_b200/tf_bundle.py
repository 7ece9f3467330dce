"""Reader / writer of TensorFlow V2 checkpoint bundles (``<prefix>.index`` + ``<prefix>.data-00000-of-00001``),
without TensorFlow.

The reference restores its "true Q" critics with ``tf.train.Saver.restore`` of the ``main/qf`` variables
(agents/SoftActorCritic.py:37-49; the five ``Bimodal1DEnv_trueQ_ckpt/*`` bundles); this module lets the
drop-in critics load those files and write checkpoints in the same container.

Format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table*): the index is a LevelDB-style table
-- data blocks of prefix-compressed (key, value) entries + restart array, each block followed by a 1-byte
compression tag and a masked crc32c, an index block mapping last keys to block handles, and a 48-byte footer
(metaindex handle, index handle, padding, magic 0xdb4775248b80fb57).  Key "" holds a BundleHeaderProto, every
other key a BundleEntryProto {1: dtype, 2: shape{2: dim{1: size}}, 3: shard_id, 4: offset, 5: size,
6: crc32c (fixed32, masked)}.  Tensors are raw little-endian bytes in the data shard.
Only uncompressed blocks, one shard and dense (unsliced) tensors are supported -- what ``Saver`` writes."""
from __future__ import annotations

import os
import struct
from typing import Dict

import numpy as np

MAGIC = 0xDB4775248B80FB57
DT = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8")}     # DT_FLOAT, DOUBLE, INT32, INT64
DT_INV = {v: k for k, v in DT.items()}

# ---- crc32c (Castagnoli), masked as LevelDB / TF store it ------------------------------------
_TABLE = None


def _table():
    global _TABLE
    if _TABLE is None:
        t = np.zeros(256, np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t[i] = c
        _TABLE = t
    return _TABLE


def crc32c(data: bytes) -> int:
    t = _table()
    c = 0xFFFFFFFF
    for b in data:
        c = int(t[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c: int) -> int:
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints / minimal protobuf -----------------------------------------------------------------
def _varint(buf: bytes, pos: int):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(x: int) -> bytes:
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        if x:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _proto_fields(buf: bytes):
    """Yield (field number, wire type, value) of one protobuf message (varint, fixed32/64, length-delimited)."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        f, w = tag >> 3, tag & 7
        if w == 0:
            v, pos = _varint(buf, pos)
        elif w == 1:
            v, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif w == 2:
            n, pos = _varint(buf, pos)
            v, pos = buf[pos:pos + n], pos + n
        elif w == 5:
            v, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % w)
        yield f, w, v


def _parse_entry(buf: bytes):
    e = dict(dtype=0, shape=[], shard=0, offset=0, size=0, crc=0, sliced=False)
    for f, _, v in _proto_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            for f2, _, v2 in _proto_fields(v):
                if f2 == 2:
                    e["shape"].append(next((v3 for f3, _, v3 in _proto_fields(v2) if f3 == 1), 0))
        elif f == 3:
            e["shard"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc"] = v
        elif f == 7:
            e["sliced"] = True
    return e


# ---- table blocks ---------------------------------------------------------------------------------
def _block_entries(block: bytes):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _read_block(raw: bytes, offset: int, size: int, verify: bool) -> bytes:
    block, tag = raw[offset:offset + size], raw[offset + size]
    if tag != 0:
        raise ValueError("compressed table blocks are not supported")
    if verify:
        stored = struct.unpack_from("<I", raw, offset + size + 1)[0]
        if mask_crc(crc32c(raw[offset:offset + size + 1])) != stored:
            raise ValueError("table block checksum mismatch")
    return block


def read_index(prefix: str, verify: bool = True):
    """Entries of ``<prefix>.index``: {tensor name: dict(dtype, shape, shard, offset, size, crc)}."""
    raw = open(prefix + ".index", "rb").read()
    if len(raw) < 48 or struct.unpack_from("<Q", raw, len(raw) - 8)[0] != MAGIC:
        raise ValueError("not a TensorFlow bundle index (bad magic)")
    foot = raw[-48:]
    pos = 0
    _, pos = _varint(foot, pos)                # metaindex handle
    _, pos = _varint(foot, pos)
    ioff, pos = _varint(foot, pos)
    isize, pos = _varint(foot, pos)
    entries = {}
    for _, handle in _block_entries(_read_block(raw, ioff, isize, verify)):
        boff, p = _varint(handle, 0)
        bsize, _ = _varint(handle, p)
        for key, val in _block_entries(_read_block(raw, boff, bsize, verify)):
            if key:                            # key "" is the BundleHeaderProto
                entries[key.decode()] = _parse_entry(val)
    return entries


def read_bundle(prefix: str, verify: bool = True) -> Dict[str, np.ndarray]:
    """All dense tensors of a checkpoint as numpy arrays (``verify``: check the per-tensor crc32c)."""
    entries = read_index(prefix, verify)
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    out = {}
    for name, e in entries.items():
        if e["sliced"] or e["shard"] != 0 or e["dtype"] not in DT:
            continue
        buf = data[e["offset"]:e["offset"] + e["size"]]
        if verify and mask_crc(crc32c(buf)) != e["crc"]:
            raise ValueError("tensor %r: crc32c mismatch" % name)
        out[name] = np.frombuffer(buf, dtype=DT[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def read_critic(prefix: str, scope: str = "main/qf"):
    """The six tensors of a T-mid / T-in critic saved under ``scope`` by ``tf.contrib.layers.fully_connected``
    (``fully_connected{,_1,_2}/{weights,biases}``), TF layout [in,out]: (W1, b1, W2, b2, W3, b3)."""
    t = read_bundle(prefix)
    names = ["fully_connected", "fully_connected_1", "fully_connected_2"]
    return tuple(t["%s/%s/%s" % (scope, n, k)] for n in names for k in ("weights", "biases"))


# ---- writer -----------------------------------------------------------------------------------------
def _entry_proto(dtype: int, shape, offset: int, size: int, crc: int) -> bytes:
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    out = b"\x08" + _put_varint(dtype) + b"\x12" + _put_varint(len(dims)) + dims
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
    return out


def _build_block(items) -> bytes:
    """One table block, restart interval 16 as LevelDB writes it."""
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % 16 == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    out += b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))
    return bytes(out)


def _short_successor(key: bytes) -> bytes:
    """LevelDB's BytewiseComparator::FindShortSuccessor: the index block stores a short key >= the block's last key."""
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    """Write ``tensors`` as a one-shard V2 bundle readable by :func:`read_bundle` (and laid out like the files
    ``tf.train.Saver`` writes: sorted keys, header entry under the empty key, masked crc32c everywhere)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items, data = [], bytearray()
    for name in sorted(tensors):
        a = np.asarray(tensors[name])
        a = a.astype(a.dtype.newbyteorder("<"), order="C", copy=False)     # keeps 0-d scalars 0-d
        if a.dtype not in DT_INV:
            raise ValueError("unsupported dtype %s" % a.dtype)
        buf = a.tobytes()
        items.append((name.encode(), _entry_proto(DT_INV[a.dtype], a.shape, len(data), len(buf), mask_crc(crc32c(buf)))))
        data += buf
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"          # num_shards = 1, version { producer: 1 }
    blocks = bytearray()

    def emit(block: bytes):
        off = len(blocks)
        blocks.extend(block + b"\x00")
        blocks.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))
    all_items = [(b"", header)] + items
    dh = emit(_build_block(all_items))
    mh = emit(_build_block([]))
    ih = emit(_build_block([(_short_successor(all_items[-1][0]), dh)]))
    foot = mh + ih
    foot += b"\x00" * (40 - len(foot)) + struct.pack("<Q", MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(blocks) + foot)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
