"""Host-side logic that needs no GPU: converter folding, blob layout, C-ABI surface, sharding."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest
import torch

from blindshadowremoval_b200 import convert, generator, sharding
from blindshadowremoval_b200.metrics import psnr, sfw_auc
from blindshadowremoval_b200.weights import check_weights, random_weights, variable_shapes
from oracle import generator_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bsr.h")).read()
    declared = set(re.findall(r"\b(bsr_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("bsr_handle")
    assert len(declared) >= 14
    lib = ctypes.CDLL(generator.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert set(generator._SYMBOLS) == declared      # the ctypes binding covers the whole header
    lib2 = generator.load_library()
    assert b"sm_100a" in lib2.bsr_version()


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(generator.BsrError, match="no CPU path"):
        generator.Generator("gsc", seed=1)


def test_bad_arguments_rejected_before_any_device_work():
    with pytest.raises(ValueError):
        generator.Generator("rgb")
    with pytest.raises(ValueError):
        generator.Generator("gsc", precision="fp8")


@pytest.mark.parametrize("variant", ["gsc", "tsm"])
def test_blob_layout(variant):
    w = random_weights(variant, 5)
    check_weights(variant, w)
    blob = convert.build_blob(variant, w)
    assert blob[:8] == b"BSRW0001"
    v, n = struct.unpack_from("<ii", blob, 8)
    assert v == ("gsc", "tsm").index(variant) and n == 4 + 6 * 5 + 3 + 1 + 3 + 3
    ent = struct.Struct("<32s6i2Q")
    names = []
    for i in range(n):
        name, kh, kw, cin, cout, tr, _, w_off, b_off = ent.unpack_from(blob, 16 + i * ent.size)
        names.append(name.rstrip(b"\0").decode())
        assert b_off == w_off + kh * kw * cin * cout * 4
        assert b_off + cout * 4 <= len(blob)
    assert names[:5] == ["conv1", "down1", "down2", "down3", "res0.conv1"] and names[-1] == "clr_conv3"
    bad = dict(w)
    bad.pop("conv1/conv/bias")
    with pytest.raises(ValueError):
        convert.build_blob(variant, bad)


def test_bn_folding_equals_conv_then_bn():
    w = random_weights("gsc", 9)
    layers = {l[0]: l for l in convert.canonical_layers("gsc", w)}
    wt = {k: torch.from_numpy(v).double() for k, v in w.items()}
    x = torch.randn(1, 8, 8, 64, dtype=torch.float64)
    # Conv + BN: down3
    ref = R.batchnorm(R.conv2d_same(x, wt["down3/conv/kernel"], wt["down3/conv/bias"], 2), wt, "down3/bnorm")
    _, kh, kw, tr, kk, bb = layers["down3"]
    got = R.conv2d_same(x, torch.from_numpy(kk).double().reshape(kh, kw, 64, 96), torch.from_numpy(bb).double(), 2)
    assert (ref - got).abs().max() < 1e-5
    # ConvT + BN: up3 (kernel transposed to [tap][cin][cout])
    x = torch.randn(1, 4, 4, 128, dtype=torch.float64)
    ref = R.batchnorm(R.conv2d_transpose_same(x, wt["up3/conv/kernel"], wt["up3/conv/bias"]), wt, "up3/bnorm")
    _, kh, kw, tr, kk, bb = layers["up3"]
    assert tr == 1 and kk.shape == (9, 128, 64)
    k_tf = torch.from_numpy(kk).double().reshape(3, 3, 128, 64).permute(0, 1, 3, 2)
    got = R.conv2d_transpose_same(x, k_tf, torch.from_numpy(bb).double())
    assert (ref - got).abs().max() < 1e-5
    # fused layers
    assert layers["res0.qkv"][4].shape == (1, 257, 384) and layers["heads"][4].shape == (49, 64, 2)
    # clr_conv1 channel reorder: TF channel 0 (gs) becomes canonical channel 64
    k1 = w["clr_conv1/conv/kernel"]
    s = w["clr_conv1/bnorm/gamma"] / np.sqrt(w["clr_conv1/bnorm/moving_variance"] + 1e-3)
    assert np.allclose(layers["clr_conv1"][4][4, 64], k1[1, 1, 0] * s, atol=1e-6)
    assert np.allclose(layers["clr_conv1"][4][4, 0], k1[1, 1, 1] * s, atol=1e-6)


def test_shard_units_partition():
    for n in (0, 1, 7, 8, 100, 256):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_units(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.shard_images(40, 10, 1, 3) == (20, 30)       # whole chunks only
    with pytest.raises(ValueError):
        sharding.shard_images(41, 10, 0, 2)
    with pytest.raises(ValueError):
        sharding.shard_units(4, 2, 2)


def test_auc_sentinels_and_psnr():
    # one-class label: sklearn alone would raise; the reference's [1,0] sentinels make it defined
    lab = np.zeros((4, 4))
    assert 0.0 <= sfw_auc(lab, np.linspace(0, 0.5, 16)) <= 1.0
    lab[0, 0] = 1
    pred = np.zeros(16)
    pred[0] = 0.9
    assert sfw_auc(lab, pred) == 1.0
    assert psnr(np.zeros(4), np.zeros(4)) == float("inf")
    assert abs(psnr(np.zeros(4), np.full(4, 0.1)) - 20.0) < 1e-9


def test_variable_shapes_wide_layers():
    g, t = variable_shapes("gsc"), variable_shapes("tsm")
    assert g["up1/conv/kernel"] == (3, 3, 96, 257) and t["up1/conv/kernel"] == (3, 3, 96, 291)
    assert g["clr_up1/conv/kernel"] == (3, 3, 128, 261) and t["clr_up1/conv/kernel"] == (3, 3, 128, 877)


def test_compact_input_helpers_match_reference_arithmetic():
    """what the compact host path assumes about the fp32 chunk it replaces."""
    # dataset.py:119 `cv2.imread(...) / 255.` is a float64 division later cast to fp32 by tf; the device kernel does the
    # fp32 division float(u8)/255.f -- identical for every byte value
    u = np.arange(256)
    assert np.array_equal((u / 255.).astype(np.float32), u.astype(np.float32) / np.float32(255.))
    # downsample8 == tf.image.resize(x,[32,32]) as restated by the oracle (model.py:237, warp.py:137)
    x = np.random.default_rng(0).random((2, 256, 256, 6), dtype=np.float32)
    ref = R.resize_bilinear(torch.from_numpy(x), 32, 32).numpy()
    got = generator.downsample8(x)
    assert got.shape == (2, 32, 32, 6) and np.abs(got - ref).max() < 1e-6
    with pytest.raises(ValueError):
        generator.downsample8(x[:, :128])


def test_unpack_chunk_matches_reference_split_sizes():
    """tf.split(img, [3,3,3,6,1] / [3,3,1,3,6,1] / [3,3,6,1], 3) after the reshape to [F,256,256,-1]."""
    rng = np.random.default_rng(1)
    for c, names in ((16, ["img", "gt", "uv", "reg", "face"]), (17, ["img", "cmap", "mask", "uv", "reg", "face"]),
                     (13, ["img", "uv", "reg", "face"])):
        chunk = rng.random((2, 256, 256, c), dtype=np.float32)
        parts = generator.unpack_chunk(chunk)
        assert list(parts) == names
        assert np.array_equal(np.concatenate([parts[k] for k in names], axis=3), chunk)
        assert parts["img"].shape[3] == 3 and parts["uv"].shape[3] == 3 and parts["reg"].shape[3] == 6 and parts["face"].shape[3] == 1
        flat = generator.unpack_chunk(chunk.reshape(-1), frames=2)                # the reference reshapes a flat record
        assert np.array_equal(flat["uv"], parts["uv"])
        t = generator.unpack_chunk(torch.from_numpy(chunk))                       # torch tensors slice the same way
        assert torch.equal(t["reg"], torch.from_numpy(parts["reg"]))
    with pytest.raises(ValueError):
        generator.unpack_chunk(np.zeros((1, 256, 256, 12), np.float32))
    with pytest.raises(ValueError):
        generator.unpack_chunk(np.zeros(5, np.float32))


def test_ssim_matches_definition_on_simple_cases():
    from blindshadowremoval_b200.metrics import ssim
    rng = np.random.default_rng(2)
    a = rng.random((64, 64, 1))
    assert abs(ssim(a, a) - 1.0) < 1e-12                                          # identical images
    assert ssim(a, 1.0 - a) < 0.0                                                 # anti-correlated structure
    c = np.full((32, 32, 1), 0.25)
    d = np.full((32, 32, 1), 0.75)                                                # constants: luminance term only
    lum = (2 * 0.25 * 0.75 + 1e-4) / (0.25 ** 2 + 0.75 ** 2 + 1e-4)
    assert abs(ssim(c, d) - lum) < 1e-9
    assert abs(ssim(a[..., 0], a[..., 0] * 0.5 + 0.1) - ssim(a, a * 0.5 + 0.1)) < 1e-12   # [H,W] accepted


def test_bench_layer_work_matches_survey_mac_table():
    """The roofline numerators bench.py reports are SURVEY.md Appendix B / section 8d: per-layer MACs and the network
    totals 9052.1 (GSC) / 10085.1 (TSM) MMAC per image."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for variant, total in (("gsc", 9052.1), ("tsm", 10085.1)):
        L = bench.layer_work(variant)
        mmac = {k: v[0] / 2e6 for k, v in L.items()}
        convs = sum(v for k, v in mmac.items() if k not in ("compose", "clr_tail"))
        tail = mmac["clr_tail"]                                   # clr_conv2 + clr_conv3 = 16.8 + 3.1
        assert abs(tail - 19.9) < 0.1
        assert abs(convs + tail - total) < 0.6, (variant, convs + tail)
    g = {k: v[0] / 2e6 for k, v in bench.layer_work("gsc").items()}
    for name, want in (("conv1", 308.3), ("down1", 302.0), ("down2", 151.0), ("down3", 56.6), ("res0.conv1", 13.0),
                       ("res1.conv1", 33.7), ("res3.conv1", 34.2), ("res2.conv2", 151.0), ("res4.conv3", 33.7),
                       ("res0.qkv", 3 * 33.7), ("res5.attention", 2 * 134.2), ("res1.w", 33.7), ("up1", 227.4),
                       ("up2", 377.5), ("up3", 1208.0), ("heads", 2 * 205.5), ("clr_up1", 307.9), ("clr_up2", 453.0),
                       ("clr_up3", 906.0), ("clr_conv1", 613.4)):
        assert abs(g[name] - want) < 0.06 * max(1.0, want / 100), (name, g[name], want)
    t = {k: v[0] / 2e6 for k, v in bench.layer_work("tsm").items()}
    assert abs(t["res0.conv1"] - 38.1) < 0.1 and abs(t["res4.conv1"] - 115.0) < 0.1 and abs(t["clr_up1"] - 1034.6) < 0.1
    # bytes: every fused unit moves at least its unique input + output once
    assert all(v[1] > 0 for v in bench.layer_work("gsc").values())
