"""TF-checkpoint reader: crc32c, block trailers, sharded data files.  CPU only.  The bundle used here is
written by this test in the on-disk format TensorFlow uses (leveldb table + raw little-endian shards); the reference's
own .index files (no data shards are shipped) are checked when /root/reference is present."""
import glob
import os
import struct

import numpy as np
import pytest

from blindshadowremoval_b200 import tf_checkpoint as T

SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def block(entries):
    """leveldb block, restart point at every entry, + 5-byte trailer."""
    body, restarts = bytearray(), []
    for k, v in entries:
        restarts.append(len(body))
        body += varint(0) + varint(len(k)) + varint(len(v)) + k + v
    for r in restarts or [0]:
        body += struct.pack("<I", r)
    body += struct.pack("<I", max(1, len(restarts)))
    trailer = b"\x00" + struct.pack("<I", T.mask_crc(T.crc32c(bytes(body) + b"\x00")))
    return bytes(body), trailer


def entry_proto(shape, shard, offset, size, crc):
    dims = b"".join(b"\x12" + varint(len(d)) + d for d in (b"\x08" + varint(s) for s in shape))
    return (b"\x08" + varint(1) + b"\x12" + varint(len(dims)) + dims + b"\x18" + varint(shard) + b"\x20" + varint(offset) +
            b"\x28" + varint(size) + b"\x35" + struct.pack("<I", crc))


def write_bundle(prefix, tensors, n_shards=2):
    shards = [bytearray() for _ in range(n_shards)]
    entries = [(b"", b"\x08\x02")]                                   # BundleHeaderProto{num_shards}
    for i, (name, arr) in enumerate(sorted(tensors.items())):
        raw = np.ascontiguousarray(arr, "<f4").tobytes()
        sh = i % n_shards
        entries.append((name.encode(), entry_proto(arr.shape, sh, len(shards[sh]), len(raw), T.mask_crc(T.crc32c(raw)))))
        shards[sh] += raw
    data, dtrail = block(entries)
    meta, mtrail = block([])
    out = bytearray(data + dtrail)
    meta_off = len(out)
    out += meta + mtrail
    idx, itrail = block([(entries[-1][0] + b"\xff", varint(0) + varint(len(data)))])
    idx_off = len(out)
    out += idx + itrail
    footer = varint(meta_off) + varint(len(meta)) + varint(idx_off) + varint(len(idx))
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    open(prefix + ".index", "wb").write(out)
    for s, b in enumerate(shards):
        open("%s.data-%05d-of-%05d" % (prefix, s, n_shards), "wb").write(b)
    return len(data)


def test_crc32c_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283 and T.crc32c(b"") == 0            # the standard check value
    assert T.crc32c(b"\x00" * 32) == 0x8A9136AA and T.crc32c(b"\xff" * 32) == 0x62A8AB43   # RFC 3720 B.4
    assert T.crc32c(b"6789", T.crc32c(b"12345")) == 0xE3069283                    # incremental
    native, T._CRC_NATIVE = T._CRC_NATIVE, False                                  # pure-Python path gives the same
    try:
        assert T.crc32c(b"123456789") == 0xE3069283 and T.crc32c(bytes(range(256)) * 3) == T.crc32c(bytes(range(256)) * 3)
        py = T.crc32c(bytes(range(251)) * 7)
    finally:
        T._CRC_NATIVE = None
    assert T.crc32c(bytes(range(251)) * 7) == py
    assert T.unmask_crc(T.mask_crc(0xE3069283)) == 0xE3069283 and T.mask_crc(0xE3069283) != 0xE3069283
    del native


def test_sharded_bundle_roundtrip_and_corruption(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {"generator/conv1/conv/kernel" + SUFFIX: rng.standard_normal((7, 7, 3, 32)).astype(np.float32),
               "generator/conv1/conv/bias" + SUFFIX: rng.standard_normal((32,)).astype(np.float32),
               "generator/conv1/conv/kernel/.OPTIMIZER_SLOT/gen_opt/m" + SUFFIX: np.zeros((7, 7, 3, 32), np.float32),
               "disc1/conv/kernel" + SUFFIX: np.ones((3, 3), np.float32)}
    prefix = str(tmp_path / "ckpt-1")
    data_len = write_bundle(prefix, tensors, n_shards=2)
    ents = T.read_index(prefix + ".index")
    assert len(ents) == 4 and {e.shard_id for e in ents.values()} == {0, 1}
    got = T.read_generator_weights(prefix + ".index")
    assert sorted(got) == ["conv1/conv/bias", "conv1/conv/kernel"]                # Adam slots and discriminators dropped
    assert np.array_equal(got["conv1/conv/kernel"], tensors["generator/conv1/conv/kernel" + SUFFIX])
    # a flipped bit in a tensor is caught by its crc32c, one in the index by the block trailer
    shard = "%s.data-%05d-of-%05d" % (prefix, ents["generator/conv1/conv/kernel" + SUFFIX].shard_id, 2)
    raw = bytearray(open(shard, "rb").read())
    raw[ents["generator/conv1/conv/kernel" + SUFFIX].offset + 5] ^= 0x10
    open(shard, "wb").write(raw)
    with pytest.raises(ValueError, match="crc32c"):
        T.read_generator_weights(prefix + ".index")
    assert T.read_generator_weights(prefix + ".index", verify_crc=False)["conv1/conv/kernel"].shape == (7, 7, 3, 32)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[data_len // 2] ^= 0x01
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(ValueError, match="crc32c"):
        T.read_index(prefix + ".index")
    os.remove(shard)
    open(prefix + ".index", "wb").write(bytes(idx[:data_len // 2]) + bytes([idx[data_len // 2] ^ 0x01]) + bytes(idx[data_len // 2 + 1:]))
    with pytest.raises(FileNotFoundError):
        T.read_generator_weights(prefix + ".index")
    with pytest.raises(ValueError, match="magic"):
        open(prefix + ".bad.index", "wb").write(b"\x00" * 64)
        T.read_index(prefix + ".bad.index")


def test_reference_index_files_pass_their_block_checksums():
    hits = sorted(glob.glob("/root/reference/log/*/ckpt-*.index"))
    if not hits:
        pytest.skip("reference checkout not present")
    for path in hits[:4]:
        ents = T.read_index(path, verify_crc=True)                                # raises on any damaged block
        assert len(ents) > 500 and all(e.crc32c for e in ents.values())
