"""Reader for TensorFlow V2 checkpoints (``ckpt-N.index`` + ``ckpt-N.data-*``) with no TensorFlow.

The reference restores its generator with ``tf.train.Checkpoint(generator=...)``
(/root/reference/train_test_GSC.py:143-148, 362-365).  The ``.index`` file is a leveldb *table*
(uncompressed blocks, prefix-compressed keys, 48-byte footer with magic 0xdb4775248b80fb57) whose
values are ``BundleEntryProto`` messages {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6};
tensor bytes are raw little-endian at ``offset`` in the data shard.  Only what the weight converter
needs is implemented: listing (name, dtype, shape, shard, offset, size) and reading fp32 tensors from one or
several data shards, with the two checksums TensorFlow writes verified: the masked crc32c in every block trailer of
the index and the masked crc32c of every tensor's bytes.
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


_CRC_TABLE = None
_CRC_NATIVE = None
_MASK_DELTA = 0xA282EAD8


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
            tab.append(c)
        _CRC_TABLE = tab
    return _CRC_TABLE


def crc32c(data, crc: int = 0) -> int:
    """CRC-32C (Castagnoli) as TensorFlow's ``crc32c::Value``.  Uses ``bsr_crc32c`` of libbsr.so when the library is
    built (the same routine, ~1 GB/s); the pure-Python table loop otherwise (fine for the index, slow for tensors)."""
    global _CRC_NATIVE
    data = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    if _CRC_NATIVE is None:
        try:
            import ctypes
            lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libbsr.so"))
            lib.bsr_crc32c.restype = ctypes.c_uint32
            lib.bsr_crc32c.argtypes = [ctypes.c_uint32, ctypes.c_char_p, ctypes.c_size_t]
            _CRC_NATIVE = lib.bsr_crc32c
        except (OSError, AttributeError):
            _CRC_NATIVE = False
    if _CRC_NATIVE:
        return int(_CRC_NATIVE(crc, bytes(data), len(data)))
    tab = _crc_table()
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = (c >> 8) ^ tab[(c ^ b) & 0xFF]
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    """``crc32c::Mask``: what is stored in block trailers and BundleEntryProto.crc32c."""
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(masked: int) -> int:
    rot = (masked - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


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


def _read_block(buf: bytes, offset: int, size: int, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    """Decode one uncompressed leveldb block into (key, value) pairs.  The 5-byte trailer is
    [type][masked crc32c(contents + type)]; a mismatch means a damaged index file."""
    if offset + size + 5 > len(buf):
        raise ValueError("leveldb block at %d (+%d) runs past the end of the file" % (offset, size))
    if buf[offset + size] != 0:
        raise ValueError("compressed leveldb blocks are not supported (type %d)" % buf[offset + size])
    blk = buf[offset:offset + size]
    if verify:
        stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
        if unmask_crc(stored) != crc32c(buf[offset:offset + size + 1]):
            raise ValueError("leveldb block at offset %d fails its crc32c check" % offset)
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


def read_index(index_path: str, verify_crc: bool = True) -> Dict[str, BundleEntry]:
    """All tensor entries of a ``.index`` file keyed by full object-graph path (block checksums verified)."""
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
    for _, handle in _read_block(buf, idx_off, idx_size, verify_crc):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for key, val in _read_block(buf, boff, bsize, verify_crc):
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


def read_generator_weights(index_path: str, verify_crc: bool = True) -> Dict[str, np.ndarray]:
    """Read the generator's fp32 tensors from the data shard(s) ``<prefix>.data-XXXXX-of-YYYYY`` next to
    ``index_path``; every tensor's bytes are checked against its BundleEntryProto.crc32c (as
    ``checkpoint.restore`` does) and against the shard's length.

    Raises FileNotFoundError if a shard is missing (the reference repo ships only the index,
    /root/reference/.MISSING_LARGE_BLOBS), ValueError on a checksum / size mismatch.
    """
    prefix = index_path[:-len(".index")]
    entries = read_index(index_path, verify_crc)
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
        if e.offset + e.size > shards[path].shape[0]:
            raise ValueError("%s: bytes [%d, %d) lie outside %s" % (name, e.offset, e.offset + e.size, path))
        if e.size != 4 * int(np.prod(e.shape, dtype=np.int64)):
            raise ValueError("%s: %d bytes do not match shape %r" % (name, e.size, e.shape))
        raw = shards[path][e.offset:e.offset + e.size].tobytes()
        if verify_crc and e.crc32c and unmask_crc(e.crc32c) != crc32c(raw):
            raise ValueError("%s: tensor bytes fail their crc32c check" % name)
        out[name[len("generator/"):-len(_SUFFIX)]] = np.frombuffer(raw, dtype="<f4").reshape(e.shape).copy()
    return out
