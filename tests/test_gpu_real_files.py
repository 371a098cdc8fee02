"""-m gpu: the reference's own golden vectors and real files through the CUDA path.

* CUDA ShareLayer (bsr_share_layer) against the outputs of the reference's sp_batch_map_offsets (warp.py:118-131,
  executed by tests/golden/make_reference_golden.py) and against the oracle's share_layer at the generator's sizes;
* BASELINE config 1 (sample_imgs/02165, batch 1) and config 2 (UCB inputs, batch 32 + ragged tail of 4) built by the
  matplotlib-free feed, through the chunk entry point, against the oracle with the same (random-init) weights."""
import glob
import os

import numpy as np
import pytest
import torch

from blindshadowremoval_b200 import feed
from blindshadowremoval_b200.metrics import psnr
from blindshadowremoval_b200.weights import random_weights

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "fixtures")
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def G():
    from blindshadowremoval_b200 import generator
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a B200; there is no CPU fallback")
    os.environ["BSR_DEBUG_KEEP"] = "1"
    return generator


def reg_from_pixel_offsets(off_in, off_out=None):
    """[B,256,256,6] registration field whose resize-to-32 x 32 (warp.py:137) is exactly the given pixel offsets."""
    b = off_in.shape[0]
    reg = np.zeros((b, 256, 256, 6), np.float32)
    reg[..., 0:2] = np.repeat(np.repeat(off_in / np.float32(32), 8, axis=1), 8, axis=2)
    if off_out is not None:
        reg[..., 3:5] = np.repeat(np.repeat(off_out / np.float32(32), 8, axis=1), 8, axis=2)
    return reg


@pytest.mark.parametrize("name", ["small", "border", "integer"])
def test_cuda_share_layer_equals_reference_warp_golden(G, name):
    """frame = 1 and a zero un-warp field make ShareLayer = concat(warp(x), warp(x)): the CUDA warp must reproduce what
    the reference's SciPy implementation produced (fp32 check handle: 1e-5; 16-bit handle: storage rounding)."""
    g = np.load(os.path.join(GOLD, "warp_reference.npz"))
    x, off, want = g[name + "_x"], g[name + "_off"], g[name + "_out"]
    b = x.shape[0]
    reg = torch.from_numpy(reg_from_pixel_offsets(off)).cuda()
    xt = torch.from_numpy(np.ascontiguousarray(x[..., None])).cuda()
    for precision, tol in (("fp32check", 2e-5), ("tc16", 4e-3)):
        gen = G.Generator("tsm", precision, device=0, micro_batch=b, seed=1)
        out = gen.share_layer(xt, reg, frame=1, share=True).cpu().numpy()
        gen.check()
        assert np.abs(out[..., 0] - want).max() < tol, (precision, np.abs(out[..., 0] - want).max())
        assert np.abs(out[..., 1] - want).max() < tol
        gen.close()


@pytest.mark.parametrize("C,frame,n", [(96, 2, 8), (291, 2, 4), (291, 10, 10), (5, 2, 2)])
def test_cuda_share_layer_matches_oracle(G, C, frame, n):
    """The generator's two ShareLayer sizes (96 and 291 channels; frame 2 and 10) plus an odd width: warp-in, max | mean
    over the frames of a chunk, warp-out (model_with_TSM.py:204-229) and the share=False branch, against the oracle."""
    from oracle.generator_ref import share_layer
    from blindshadowremoval_b200.synthetic import make_inputs
    rng = np.random.default_rng(C + frame)
    x = rng.standard_normal((n, 32, 32, C)).astype(np.float32)
    reg = make_inputs(n, seed=7, with_reg=True)["reg"] * np.float32(2.0)
    want = share_layer(torch.from_numpy(x), torch.from_numpy(reg), frame, True).numpy()
    xt, rt = torch.from_numpy(x).cuda(), torch.from_numpy(reg).cuda()
    for precision, tol in (("fp32check", 2e-5), ("tc16", 6e-3)):
        gen = G.Generator("tsm", precision, device=0, micro_batch=n, seed=1)
        out = gen.share_layer(xt, rt, frame=frame, share=True).cpu().numpy()
        assert out.shape == want.shape and np.abs(out - want).max() < tol, (precision, np.abs(out - want).max())
        dup = gen.share_layer(xt, rt, frame=frame, share=False).cpu().numpy()
        ref_dup = np.concatenate([x, x], axis=-1)
        assert np.abs(dup - ref_dup).max() < (1e-7 if precision == "fp32check" else 4e-3)
        gen.check()
        gen.close()


def _oracle_chunk(w, frames, bm):
    from oracle.generator_ref import caller_glue, generator_forward
    img = np.stack([f["img"] for f in frames])
    uv = np.stack([f["uv"] for f in frames])
    face = np.stack([f["face"] for f in frames])
    ref = generator_forward(w, img, uv, variant="gsc", bmask_override=bm)
    ref0 = generator_forward(w, img, uv, variant="gsc")
    rgb_c, mp = caller_glue(ref["con_rgb"], ref["dif"], face)
    return rgb_c, mp, ref, float((ref0["bmask"] != bm).mean())


def test_config1_sample_image_batch1_through_chunk_entry(G):
    """BASELINE config 1: sample_imgs/02165/{png,npy} -> feed.load_frame -> [1,256,256,16] chunk -> bsr_forward_chunk,
    against the oracle (same random-init weights: no trained checkpoint ships with the reference)."""
    f = feed.load_frame(os.path.join(FIX, "sample_imgs", "02165", "02165.png"))
    chunk = feed.build_chunk([f])
    w = random_weights("gsc", 1234)
    gen = G.Generator("gsc", "tc16", device=0, micro_batch=1, weights=w)
    rgb, mp, gs, m22 = gen.forward_chunk(chunk, want_raw=True)
    gen.check()
    bm = gen.debug_read("bmask").reshape(1, 32, 32, 1)
    want_rgb, want_mp, ref, flips = _oracle_chunk(w, [f], bm)
    rep = {"rgb": float(np.abs(rgb - want_rgb).max()), "mask_pred": float(np.abs(mp - want_mp).max()),
           "gs": float(np.abs(gs - ref["gs"]).max()), "flips": flips}
    print("config 1 (02165):", rep, "psnr rgb %.1f dB" % psnr(rgb, want_rgb))
    assert rep["rgb"] <= 1e-2 and rep["mask_pred"] <= 1e-2 and rep["gs"] <= 1e-2 and flips < 0.01
    assert psnr(rgb, want_rgb) >= 40.0
    # the compact entry point fed with the uint8 file content + 32 x 32 maps gives the bits of the fp32 call
    lm = f["lm"]
    png = (f["img"] * 255.0)
    if np.abs(png - np.rint(png)).max() < 1e-4:          # an un-resized 256 x 256 crop keeps exact byte values
        u8 = np.rint(png).astype(np.uint8)[None]
        c = feed.frame_maps_compact(lm, with_face=False)
        out = gen.forward_compact(u8, c["uv32"][None], want=("con_rgb",))
        _, rgb_f, _, _ = gen(u8.astype(np.float32) / np.float32(255.0), f["uv"][None], None, want=("con_rgb",))
        assert np.array_equal(out["con_rgb"], rgb_f)
    gen.close()


def test_config2_ucb_batch32_plus_ragged_tail(G):
    """BASELINE config 2: the UCB inputs in batches of 32 plus a tail of 4 (100 = 3 x 32 + 4 in the reference run;
    here the 8 committed pairs tiled to 36) through the chunk entry, every distinct image against the oracle."""
    files = sorted(glob.glob(os.path.join(FIX, "UCB", "input", "*", "*.png")))
    assert len(files) == 8
    frames = [feed.load_frame(p, gt_path=p.replace(os.sep + "input" + os.sep, os.sep + "gt" + os.sep)) for p in files]
    order = [i % 8 for i in range(36)]
    chunk = feed.build_chunk([frames[i] for i in order])
    w = random_weights("gsc", 1234)
    gen = G.Generator("gsc", "tc16", device=0, micro_batch=32, weights=w)
    t = torch.from_numpy(chunk).cuda()
    rgb_a, mp_a = (o.cpu().numpy() for o in gen.forward_chunk(t[:32]))
    bm_a = gen.debug_read("bmask").reshape(32, 32, 32, 1)
    rgb_b, mp_b = (o.cpu().numpy() for o in gen.forward_chunk(t[32:]))
    bm_b = gen.debug_read("bmask").reshape(4, 32, 32, 1)
    rgb_all, mp_all = (o.cpu().numpy() for o in gen.forward_chunk(t))                 # 32 + ragged 4 in one call
    gen.check()
    assert np.array_equal(rgb_all, np.concatenate([rgb_a, rgb_b])) and np.array_equal(mp_all, np.concatenate([mp_a, mp_b]))
    assert np.array_equal(rgb_a[:8], rgb_a[8:16]) and np.array_equal(rgb_b, rgb_a[:4])   # repeats are bit-identical
    want_rgb, want_mp, _, flips = _oracle_chunk(w, frames, bm_a[:8])
    worst = (float(np.abs(rgb_a[:8] - want_rgb).max()), float(np.abs(mp_a[:8] - want_mp).max()))
    print("config 2 (8 UCB pairs): max-abs rgb %.2e mask_pred %.2e flips %.4f psnr %.1f dB" % (worst + (flips, psnr(rgb_a[:8], want_rgb))))
    assert worst[0] <= 1e-2 and worst[1] <= 1e-2 and flips < 0.01 and psnr(rgb_a[:8], want_rgb) >= 40.0
    assert np.array_equal(bm_b, bm_a[:4])
    gen.close()
