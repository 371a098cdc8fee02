"""The oracle pinned: against an independent fp64 restatement, hand-derived TF-semantics cases, SciPy's
map_coordinates (the reference's own sp_batch_map_offsets), the checkpoint index, and golden fingerprints."""
import json
import os

import numpy as np
import pytest
import torch
from scipy.ndimage import map_coordinates

from blindshadowremoval_b200.synthetic import make_inputs
from blindshadowremoval_b200.weights import random_weights, variable_shapes
from oracle import generator_ref as R
from oracle import np64_ref as N


def test_same_padding_stride2_known_answer():
    # TF SAME, k=3, s=2, even input: pad 0 before / 1 after  -> out[0] sees in[0..2], out[1] in[2..4(pad)]
    x = torch.arange(1.0, 5.0).reshape(1, 1, 4, 1).expand(1, 4, 4, 1).contiguous()
    k = torch.ones(3, 3, 1, 1)
    y = R.conv2d_same(x, k, None, stride=2)[0, :, :, 0]
    assert y.shape == (2, 2)
    # columns: (1+2+3)=6 over 3 rows (rows 0..2) = 18 ; second col (3+4+0)=7*3=21; second row has 2 valid rows
    assert y.tolist() == [[18.0, 21.0], [12.0, 14.0]]


def test_same_padding_7x7_symmetric():
    x = torch.ones(1, 8, 8, 1)
    y = R.conv2d_same(x, torch.ones(7, 7, 1, 1), None)[0, :, :, 0]
    assert y[0, 0] == 16 and y[4, 4] == 49 and y[7, 7] == 16      # 3 before / 3 after


def test_conv_transpose_known_answer():
    # out[2i+kh, 2j+kw] += x[i,j] * W[kh,kw]; cropped to 2H
    x = torch.tensor([[1.0, 2.0], [3.0, 4.0]]).reshape(1, 2, 2, 1)
    k = torch.arange(1.0, 10.0).reshape(3, 3, 1, 1)
    y = R.conv2d_transpose_same(x, k, None)[0, :, :, 0].numpy()
    exp = np.zeros((5, 5))
    for i in range(2):
        for j in range(2):
            exp[2 * i:2 * i + 3, 2 * j:2 * j + 3] += x[0, i, j, 0].item() * np.arange(1.0, 10.0).reshape(3, 3)
    assert np.array_equal(y, exp[:4, :4])


def test_resize_256_to_32_is_mean_of_centre_2x2():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 256, 256, 2, generator=g, dtype=torch.float64)
    y = R.resize_bilinear(x, 32, 32)
    i, j = 5, 17
    exp = x[0, 8 * i + 3:8 * i + 5, 8 * j + 3:8 * j + 5].mean(dim=(0, 1))
    assert torch.allclose(y[0, i, j], exp, atol=1e-14)


def test_resize_upsample_edges_clamped():
    x = torch.tensor([0.0, 1.0]).reshape(1, 1, 2, 1)
    y = R.resize_bilinear(x, 1, 4)[0, 0, :, 0]
    assert torch.allclose(y, torch.tensor([0.0, 0.25, 0.75, 1.0]))


def test_warp_matches_scipy_reference():
    # /root/reference/warp.py:118-131 (sp_batch_map_offsets) on already-resized offsets
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 32, 32, 5))
    reg = np.zeros((3, 256, 256, 3))
    reg[..., :2] = rng.uniform(-0.15, 0.15, (3, 1, 1, 2)) + 0.02 * rng.standard_normal((3, 256, 256, 2))
    got = R.batch_map_offsets(torch.from_numpy(x), torch.from_numpy(reg)).numpy()
    off = (R.resize_bilinear(torch.from_numpy(reg), 32, 32).numpy() * 32)[..., :2].reshape(3, -1, 2)
    grid = np.stack(np.mgrid[:32, :32], -1).reshape(-1, 2)
    coords = (off + grid).clip(0, 31)
    for b in range(3):
        for c in range(5):
            exp = map_coordinates(x[b, :, :, c], coords[b].T, mode="nearest", order=1).reshape(32, 32)
            assert np.abs(got[b, :, :, c] - exp).max() < 1e-12


@pytest.mark.parametrize("variant,frame", [("gsc", 1), ("tsm", 2)])
def test_two_restatements_agree(variant, frame):
    w = random_weights(variant, 11)
    d = make_inputs(frame, 3, with_reg=True)
    a = R.generator_forward(w, d["img"], d["uv"], d["reg"], variant=variant, frame=frame, dtype=torch.float64)
    b = N.forward(w, d["img"], d["uv"], d["reg"], variant=variant, frame=frame)
    c = R.generator_forward(w, d["img"], d["uv"], d["reg"], variant=variant, frame=frame, dtype=torch.float32)
    for k in ("gs", "con_rgb", "mask22", "dif", "dif_small"):
        assert np.abs(a[k] - b[k]).max() < 1e-10, k            # fp64 vs fp64, different formulations
        assert np.abs(c[k] - b[k]).max() < 1e-5, k             # fp32 torch vs fp64 numpy
    assert np.array_equal(a["bmask"], b["bmask"])


def test_share_false_is_concat():
    x = torch.rand(4, 32, 32, 6)
    reg = torch.zeros(4, 256, 256, 6)
    assert torch.equal(R.share_layer(x, reg, 2, False), torch.cat([x, x], -1))
    # zero offsets: warp is the identity, so sharing = [max, mean] over the chunk broadcast to its frames
    out = R.share_layer(x, reg, 2, True).reshape(2, 2, 32, 32, 12)
    g = x.reshape(2, 2, 32, 32, 6)
    assert torch.allclose(out[:, 0, ..., :6], g.max(1).values) and torch.allclose(out[:, 1, ..., 6:], g.mean(1))


@pytest.mark.parametrize("variant", ["gsc", "tsm"])
def test_variable_inventory_matches_checkpoint_index(variant, golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "ckpt_variables.json")))[variant]
    spec = {k: list(v) for k, v in variable_shapes(variant).items()}
    assert spec == gold
    assert len(spec) == 258


def test_index_parser_on_reference_if_present(golden_dir):
    import glob
    from blindshadowremoval_b200.tf_checkpoint import generator_variables, read_index
    hits = glob.glob("/root/reference/log/*-gradients/ckpt-94.index")
    if not hits:
        pytest.skip("reference checkout not present (GPU box)")
    assert len(read_index(hits[0])) == 828
    gold = json.load(open(os.path.join(golden_dir, "ckpt_variables.json")))["gsc"]
    assert {k: list(v) for k, v in generator_variables(hits[0]).items()} == gold


@pytest.mark.parametrize("variant,frame", [("gsc", 1), ("tsm", 2)])
def test_oracle_golden_fingerprint(variant, frame, golden_dir):
    g = np.load(os.path.join(golden_dir, "oracle_%s.npz" % variant))
    w = random_weights(variant, 1234)
    w["conv3/conv/bias"] = g["conv3_bias"]
    d = make_inputs(2, 0, with_reg=True)
    o = R.generator_forward(w, d["img"], d["uv"], d["reg"], variant=variant, frame=frame)
    assert np.abs(o["dif_small"] - g["dif_small"]).max() < 1e-4
    assert (o["bmask"] != g["bmask"]).mean() < 0.005
    for k in ("con_rgb", "dif", "gs"):
        assert np.abs(o[k][:, ::8, ::8] - g[k]).max() < 2e-4, k
    assert 0.3 < g["bmask"].mean() < 0.7       # both sides of the hole threshold are exercised
