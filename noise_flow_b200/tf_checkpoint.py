"""TensorFlow-free reader/writer for TF "checkpoint V2" bundles.

The reference persists the model with ``tf.train.Saver`` (reference
``train_noise_flow.py:322,411-417``) and ``NoiseFlowWrapper`` restores
``<path>/ckpt/model.ckpt.best`` (reference ``borealisflows/NoiseFlowWrapper.py:43,77``).
TensorFlow cannot run in this environment, so the bundle format is parsed directly:

* ``<prefix>.index``  -- an uncompressed LevelDB-format SSTable.  Key ``""`` holds a
  ``BundleHeaderProto``; every other key is a variable name whose value is a
  ``BundleEntryProto`` (dtype, shape, shard_id, offset, size, crc32c).
* ``<prefix>.data-00000-of-00001`` -- raw little-endian tensor bytes addressed by
  ``(offset, size)``.

Only what the Noise Flow artefacts need is implemented (float32/float64/int32/int64
dense tensors, one shard, no compression).
"""
from __future__ import annotations

import os
import struct
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

_TABLE_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8")}
_DTYPE_ENUM = {v: k for k, v in _DTYPES.items()}


# ----------------------------------------------------------------------------- varints / protobuf
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _proto_fields(buf: bytes):
    """Yield (field_number, wire_type, value) for a serialized protobuf message."""
    pos = 0
    n = len(buf)
    while pos < n:
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, val


def _parse_shape(buf: bytes):
    dims = []
    for field, _, val in _proto_fields(buf):
        if field == 2:  # repeated Dim
            size = 0
            for f2, _, v2 in _proto_fields(val):
                if f2 == 1:
                    size = v2
            dims.append(size)
    return tuple(dims)


def _parse_entry(buf: bytes):
    ent = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None}
    for field, wt, val in _proto_fields(buf):
        if field == 1:
            ent["dtype"] = val
        elif field == 2:
            ent["shape"] = _parse_shape(val)
        elif field == 3:
            ent["shard_id"] = val
        elif field == 4:
            ent["offset"] = val
        elif field == 5:
            ent["size"] = val
        elif field == 6:
            ent["crc32c"] = struct.unpack("<I", val)[0]
    return ent


# ----------------------------------------------------------------------------- SSTable
def _read_block(data: bytes, offset: int, size: int):
    """Return the list of (key, value) entries of one table block."""
    if data[offset + size] != 0:
        raise ValueError("compressed SSTable blocks are not supported")
    block = data[offset:offset + size]
    n_restarts = struct.unpack("<I", block[-4:])[0]
    limit = len(block) - 4 - 4 * n_restarts
    pos = 0
    key = b""
    out = []
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_index(index_path: str) -> "OrderedDict[str, dict]":
    """Parse ``<prefix>.index`` into ``{variable_name: entry}`` (insertion = key order)."""
    with open(index_path, "rb") as f:
        data = f.read()
    footer = data[-48:]
    if struct.unpack("<Q", footer[-8:])[0] != _TABLE_MAGIC:
        raise ValueError("%s is not an SSTable (bad magic)" % index_path)
    pos = 0
    _, pos = _varint(footer, pos)   # metaindex offset
    _, pos = _varint(footer, pos)   # metaindex size
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    entries: "OrderedDict[str, dict]" = OrderedDict()
    for _, handle in _read_block(data, idx_off, idx_size):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for key, val in _read_block(data, boff, bsize):
            if key == b"":
                continue  # BundleHeaderProto
            entries[key.decode("utf-8")] = _parse_entry(val)
    return entries


def load_checkpoint(prefix: str) -> "OrderedDict[str, np.ndarray]":
    """Load every tensor of the bundle ``prefix`` (e.g. ``.../ckpt/model.ckpt.best``)."""
    entries = read_index(prefix + ".index")
    shards: Dict[int, bytes] = {}
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, ent in entries.items():
        sid = ent["shard_id"]
        if sid not in shards:
            cands = [p for p in os.listdir(os.path.dirname(prefix) or ".")
                     if p.startswith(os.path.basename(prefix) + ".data-%05d-of-" % sid)]
            if not cands:
                raise FileNotFoundError("data shard %d of %s" % (sid, prefix))
            with open(os.path.join(os.path.dirname(prefix) or ".", cands[0]), "rb") as f:
                shards[sid] = f.read()
        if ent["dtype"] not in _DTYPES:
            raise ValueError("unsupported dtype enum %d for %s" % (ent["dtype"], name))
        dt = _DTYPES[ent["dtype"]]
        raw = shards[sid][ent["offset"]:ent["offset"] + ent["size"]]
        arr = np.frombuffer(raw, dtype=dt).reshape(ent["shape"]).copy()
        out[name] = arr
    return out


# ----------------------------------------------------------------------------- writer
def _crc32c_table():
    poly = 0x82F63B78
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        tbl.append(c)
    return tbl


_CRC_TBL = _crc32c_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    crc ^= 0xFFFFFFFF
    for b in data:
        crc = _CRC_TBL[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _mask_crc(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _block_bytes(items) -> bytes:
    """One restart point per entry (no prefix sharing) -- valid, if not minimal."""
    body = bytearray()
    restarts = []
    for key, val in items:
        restarts.append(len(body))
        body += _put_varint(0) + _put_varint(len(key)) + _put_varint(len(val)) + key + val
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    return bytes(body)


def save_checkpoint(prefix: str, tensors: "Dict[str, np.ndarray]") -> None:
    """Write a single-shard V2 bundle that :func:`load_checkpoint` (and TF) can read."""
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    names = sorted(tensors.keys())
    data = bytearray()
    items = []
    # BundleHeaderProto: num_shards=1 (field 1), endianness LITTLE=0 (omitted), version{producer=1}
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"
    items.append((b"", header))
    for name in names:
        arr = np.asarray(tensors[name], order="C")  # (ascontiguousarray would promote 0-d to 1-d)
        dt = arr.dtype.newbyteorder("<") if arr.dtype.byteorder == ">" else arr.dtype
        if np.dtype(dt) not in _DTYPE_ENUM:
            raise ValueError("unsupported dtype %s for %s" % (arr.dtype, name))
        raw = arr.astype(dt, copy=False).tobytes()
        shape_msg = b"".join(b"\x12" + _put_varint(len(d)) + d
                             for d in (b"\x08" + _put_varint(int(s)) for s in arr.shape))
        ent = b"\x08" + _put_varint(_DTYPE_ENUM[np.dtype(dt)])
        ent += b"\x12" + _put_varint(len(shape_msg)) + shape_msg
        if len(data):
            ent += b"\x20" + _put_varint(len(data))
        ent += b"\x28" + _put_varint(len(raw))
        ent += b"\x35" + struct.pack("<I", _mask_crc(crc32c(raw)))
        items.append((name.encode("utf-8"), ent))
        data += raw
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))

    out = bytearray()

    def emit(block: bytes) -> Tuple[int, int]:
        off = len(out)
        out.extend(block)
        trailer = b"\x00"
        out.extend(trailer + struct.pack("<I", _mask_crc(crc32c(block + trailer))))
        return off, len(block)

    d_off, d_size = emit(_block_bytes(items))
    m_off, m_size = emit(_block_bytes([]))
    last_key = items[-1][0] + b"\x00"
    i_off, i_size = emit(_block_bytes([(last_key, _put_varint(d_off) + _put_varint(d_size))]))
    footer = _put_varint(m_off) + _put_varint(m_size) + _put_varint(i_off) + _put_varint(i_size)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _TABLE_MAGIC)
    out.extend(footer)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
