"""Golden vectors produced by the reference's OWN functions (ast-extracted and executed in the build container by
tests/golden/make_reference_golden.py) against the oracle and the host-side feed.  The -m gpu counterparts (CUDA
ShareLayer vs the same vectors, real files through the generator) are in tests/test_gpu_real_files.py."""
import glob
import os

import numpy as np
import pytest
import torch

from blindshadowremoval_b200 import feed
from oracle import generator_ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "fixtures")
GOLD = os.path.join(HERE, "golden")


def reg_from_pixel_offsets(off):
    """A [B,256,256,3] registration field whose tf.image.resize(...,[32,32]) * 32 (warp.py:137) is EXACTLY the pixel
    offsets `off` [B,32,32,2]: constant 8x8 cells of off / 32 (means of equal values and x32 are exact in fp32)."""
    b = off.shape[0]
    reg = np.zeros((b, 256, 256, 3), np.float32)
    reg[..., :2] = np.repeat(np.repeat(off / np.float32(32), 8, axis=1), 8, axis=2)
    return reg


@pytest.mark.parametrize("name", ["small", "border", "integer"])
def test_oracle_warp_equals_reference_sp_batch_map_offsets(name):
    """/root/reference/warp.py:118-131 executed on seeded data -> golden; the oracle's batch_map_offsets (the
    restatement of tf_batch_map_offsets, warp.py:134-165) must reproduce it."""
    g = np.load(os.path.join(GOLD, "warp_reference.npz"))
    x, off, want = g[name + "_x"], g[name + "_off"], g[name + "_out"]
    reg = reg_from_pixel_offsets(off)
    got = R.batch_map_offsets(torch.from_numpy(x[..., None]).double(), torch.from_numpy(reg).double()).numpy()[..., 0]
    assert np.abs(got - want).max() < 1e-12
    got32 = R.batch_map_offsets(torch.from_numpy(x[..., None]), torch.from_numpy(reg)).numpy()[..., 0]
    assert np.abs(got32 - want).max() < 2e-5
    if name == "border":
        grid = np.stack(np.mgrid[:32, :32], -1)
        assert ((off + grid < 0) | (off + grid > 31)).mean() > 0.05        # the clip branch is really exercised


def _fixture_files():
    return sorted(glob.glob(os.path.join(FIX, "sample_imgs", "*", "*.npy"))) + \
        sorted(glob.glob(os.path.join(FIX, "UCB", "input", "*", "*.npy")))


def test_crop_and_resize_equals_reference_face_crop_and_resize():
    """/root/reference/utils.py:356-433 executed on the reference's real files -> golden; feed.crop_and_resize must
    give the same landmarks, box and pixels."""
    import cv2
    g = np.load(os.path.join(GOLD, "crop_reference.npz"))
    files = _fixture_files()
    assert len(files) == 9
    for f in files:
        key = os.path.basename(f)[:-4].replace("-", "_")
        img = cv2.cvtColor(cv2.imread(f[:-4] + ".png"), cv2.COLOR_BGR2RGB) / 255.
        crop, lm, lm_mirror, box = feed.crop_and_resize(img, np.load(f), 256)
        assert list(box) == list(g[key + "_box"]), key
        assert np.abs(lm - g[key + "_lm"]).max() < 1e-6 and np.abs(lm_mirror - g[key + "_lm_mirror"]).max() < 1e-6
        assert np.array_equal(crop[::8, ::8], g[key + "_crop8"]), key
        assert np.allclose(crop.mean(axis=(0, 1)), g[key + "_crop_mean"], atol=1e-12)


def test_gaussian5_equals_cv2_gaussian_blur():
    """generate_face_region blurs with cv2.GaussianBlur(x, (5,5), 0) (utils.py:274); feed._gaussian5 restates it."""
    import cv2
    rng = np.random.default_rng(0)
    for a in ((rng.random((256, 256)) > 0.5).astype(np.float32), rng.random((64, 96)).astype(np.float32)):
        want = cv2.GaussianBlur(a, (5, 5), 0)
        assert np.abs(feed._gaussian5(a) - want).max() < 1e-6


def test_real_files_build_reference_shaped_chunks():
    """Config 1 / config 2 inputs: sample_imgs/02165 and the UCB pairs through the matplotlib-free feed give the
    [F,256,256,16] float32 chunk of dataset.py:296-302 with sane maps (uv zero outside the hull, face in [0,1],
    registration offsets small)."""
    f = feed.load_frame(os.path.join(FIX, "sample_imgs", "02165", "02165.png"))
    chunk = feed.build_chunk([f])
    assert chunk.shape == (1, 256, 256, 16) and chunk.dtype == np.float32 and np.isfinite(chunk).all()
    assert 0.0 <= chunk[..., 0:6].min() and chunk[..., 0:6].max() <= 1.0
    face = chunk[0, ..., 15]
    assert 0.2 < face.mean() < 0.8 and face.min() >= 0.0 and face.max() <= 1.0 + 1e-6
    uv = chunk[0, ..., 6:9]
    assert np.abs(uv[face == 0]).max() == 0.0 and uv.max() > 0.5
    assert np.abs(chunk[0, ..., 9:15]).max() < 0.5
    u = sorted(glob.glob(os.path.join(FIX, "UCB", "input", "*", "*.png")))[0]
    g = feed.load_frame(u, gt_path=u.replace(os.sep + "input" + os.sep, os.sep + "gt" + os.sep))
    assert g["img"].shape == (256, 256, 3) and np.abs(g["img"] - g["gt"]).mean() > 1e-3   # shadowed input differs from gt
