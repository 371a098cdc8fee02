"""Reader for TensorFlow V2 checkpoints (``ckpt-N.index`` + ``ckpt-N.data-*``) with no TensorFlow.

The reference restores its generator with ``tf.train.Checkpoint(generator=...)``
(/root/reference/train_test_GSC.py:143-148, 362-365).  The ``.index`` file is a leveldb *table*
(uncompressed blocks, prefix-compressed keys, 48-byte footer with magic 0xdb4775248b80fb57) whose
values are ``BundleEntryProto`` messages {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6};
tensor bytes are raw little-endian at ``offset`` in the data shard.  Only what the weight converter
needs is implemented: listing (name, dtype, shape, shard, offset, size) and reading fp32 tensors.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


@dataclass(frozen=True)
class BundleEntry:
    name: str
    dtype: int            # tensorflow DataType enum; 1 = DT_FLOAT
    shape: Tuple[int, ...]
    shard_id: int
    offset: int
    size: int
    crc32c: int


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


def _read_block(buf: bytes, offset: int, size: int) -> List[Tuple[bytes, bytes]]:
    """Decode one uncompressed leveldb block into (key, value) pairs."""
    if buf[offset + size] != 0:
        raise ValueError("compressed leveldb blocks are not supported (type %d)" % buf[offset + size])
    blk = buf[offset:offset + size]
    n_restarts = struct.unpack_from("<I", blk, len(blk) - 4)[0]
    end = len(blk) - 4 - 4 * n_restarts
    pos = 0
    key = b""
    out = []
    while pos < end:
        shared, pos = _varint(blk, pos)
        non_shared, pos = _varint(blk, pos)
        vlen, pos = _varint(blk, pos)
        key = key[:shared] + blk[pos:pos + non_shared]
        pos += non_shared
        out.append((key, blk[pos:pos + vlen]))
        pos += vlen
    return out


def _parse_shape(msg: bytes) -> Tuple[int, ...]:
    # TensorShapeProto: repeated Dim dim = 2 { int64 size = 1; }
    dims = []
    pos = 0
    while pos < len(msg):
        tag, pos = _varint(msg, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 2:
            ln, pos = _varint(msg, pos)
            sub = msg[pos:pos + ln]
            pos += ln
            if field == 2:
                sp = 0
                size = 0
                while sp < len(sub):
                    t2, sp = _varint(sub, sp)
                    if t2 & 7 == 0:
                        v, sp = _varint(sub, sp)
                        if t2 >> 3 == 1:
                            size = v
                    elif t2 & 7 == 2:
                        l2, sp = _varint(sub, sp)
                        sp += l2
                    else:
                        raise ValueError("unexpected wire type in Dim")
                dims.append(size)
        elif wt == 0:
            _, pos = _varint(msg, pos)
        else:
            raise ValueError("unexpected wire type in TensorShapeProto")
    return tuple(dims)


def _parse_entry(name: str, msg: bytes) -> BundleEntry:
    f = {1: 0, 3: 0, 4: 0, 5: 0, 6: 0}
    shape: Tuple[int, ...] = ()
    pos = 0
    while pos < len(msg):
        tag, pos = _varint(msg, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(msg, pos)
            f[field] = v
        elif wt == 5:
            f[field] = struct.unpack_from("<I", msg, pos)[0]
            pos += 4
        elif wt == 1:
            pos += 8
        elif wt == 2:
            ln, pos = _varint(msg, pos)
            if field == 2:
                shape = _parse_shape(msg[pos:pos + ln])
            pos += ln
        else:
            raise ValueError("unexpected wire type %d" % wt)
    return BundleEntry(name, f[1], shape, f[3], f[4], f[5], f[6])


def read_index(index_path: str) -> Dict[str, BundleEntry]:
    """All tensor entries of a ``.index`` file keyed by full object-graph path."""
    with open(index_path, "rb") as fh:
        buf = fh.read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != _MAGIC:
        raise ValueError("%s is not a leveldb table (bad footer magic)" % index_path)
    footer = buf[-48:]
    pos = 0
    _, pos = _varint(footer, pos)          # metaindex offset
    _, pos = _varint(footer, pos)          # metaindex size
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    entries: Dict[str, BundleEntry] = {}
    for _, handle in _read_block(buf, idx_off, idx_size):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for key, val in _read_block(buf, boff, bsize):
            if key == b"":
                continue                     # BundleHeaderProto
            name = key.decode("utf-8")
            entries[name] = _parse_entry(name, val)
    return entries


def generator_variables(index_path: str) -> Dict[str, Tuple[int, ...]]:
    """{short name: shape} of the generator's own variables (Adam slots and discriminators dropped).

    Short name = object-graph path with the ``generator/`` prefix and the
    ``/.ATTRIBUTES/VARIABLE_VALUE`` suffix removed, e.g. ``res_stack/0/non_local/theta/kernel``.
    """
    out = {}
    for name, e in read_index(index_path).items():
        if not name.startswith("generator/") or not name.endswith(_SUFFIX):
            continue
        if "/.OPTIMIZER_SLOT/" in name:
            continue
        out[name[len("generator/"):-len(_SUFFIX)]] = e.shape
    return out


def read_generator_weights(index_path: str) -> Dict[str, np.ndarray]:
    """Read the generator's fp32 tensors from the data shard(s) next to ``index_path``.

    Raises FileNotFoundError if the shard is missing (the reference repo ships only the index,
    /root/reference/.MISSING_LARGE_BLOBS).
    """
    prefix = index_path[:-len(".index")]
    entries = read_index(index_path)
    n_shards = 1 + max(e.shard_id for e in entries.values())
    out = {}
    shards = {}
    for name, e in entries.items():
        if not name.startswith("generator/") or not name.endswith(_SUFFIX) or "/.OPTIMIZER_SLOT/" in name:
            continue
        if e.dtype != 1:
            raise ValueError("%s: dtype %d is not DT_FLOAT" % (name, e.dtype))
        path = "%s.data-%05d-of-%05d" % (prefix, e.shard_id, n_shards)
        if path not in shards:
            if not os.path.exists(path):
                raise FileNotFoundError(path)
            shards[path] = np.memmap(path, dtype=np.uint8, mode="r")
        raw = shards[path][e.offset:e.offset + e.size]
        out[name[len("generator/"):-len(_SUFFIX)]] = np.frombuffer(raw.tobytes(), dtype="<f4").reshape(e.shape).copy()
    return out
