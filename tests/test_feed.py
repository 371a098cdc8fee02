"""Host feed: landmark-driven uv / registration / face maps without matplotlib, and their 32x32
compact forms.  CPU only.  The reference's own generators cannot run here (matplotlib missing), so the checks are an
independent restatement (oracle/feed_ref.py) plus the properties a piecewise-linear Delaunay interpolant must have."""
import numpy as np
import pytest

from blindshadowremoval_b200 import feed
from blindshadowremoval_b200.generator import downsample8
from oracle import feed_ref


def landmarks(seed):
    """A plausible 68-point face: the reference template jittered and shifted, in [0,1] image coordinates."""
    _, lm_ref = feed.face_template()
    rng = np.random.default_rng(seed)
    return (lm_ref + rng.normal(0, 0.01, lm_ref.shape) + rng.uniform(-0.03, 0.03, (1, 2))).astype(np.float32)


def test_template_shapes_and_ranges():
    uv, lm_ref = feed.face_template()
    assert uv.shape == (68, 3) and lm_ref.shape == (68, 2)
    assert 0 < uv.min() and uv.max() < 1 and 0 < lm_ref.min() and lm_ref.max() < 1


@pytest.mark.parametrize("seed", [0, 1])
def test_maps_match_independent_restatement(seed):
    lm = landmarks(seed)
    uv, lm_ref = feed.face_template()
    assert np.allclose(feed.generate_uv_map(lm, uv, 256), feed_ref.generate_uv_map(lm, uv, 256), atol=1e-9)
    a, b = feed.generate_offset_map(lm, lm_ref, 256), feed_ref.generate_offset_map(lm, lm_ref, 256)
    assert not np.isnan(a).any() and np.allclose(a, np.nan_to_num(b), atol=1e-9)
    assert np.abs(feed.generate_face_region(lm, 256) - feed_ref.generate_face_region(lm, 256)).max() < 1e-6


def test_interpolant_properties():
    lm = landmarks(3)
    uv, lm_ref = feed.face_template()
    it = feed.TriInterpolator(lm)
    it.locate(lm[:, 0].astype(np.float64), lm[:, 1].astype(np.float64))
    assert np.allclose(it(uv), uv, atol=1e-6)                                   # exact at the landmarks
    xi, yi = np.meshgrid(np.linspace(0, 1, 64), np.linspace(0, 1, 64))
    it.locate(xi, yi)
    affine = 0.3 * lm[:, 0] - 1.7 * lm[:, 1] + 0.25                             # affine data is reproduced exactly
    got = it(affine)
    inside = ~np.isnan(got)
    assert inside.sum() > 500 and not inside[0, 0] and not inside[-1, -1]       # corners of the image are outside the hull
    assert np.allclose(got[inside], (0.3 * xi - 1.7 * yi + 0.25)[inside], atol=1e-6)
    m = feed.generate_uv_map(lm, uv, 256)
    assert m.shape == (256, 256, 3) and m[0, 0].tolist() == [0.0, 0.0, 0.0] and 0.1 < m[128, 128, 0] < 0.9
    assert np.abs(feed.generate_offset_map(lm, lm, 256)).max() == 0.0           # identical landmark sets: no offset
    off = feed.generate_offset_map(lm, lm_ref, 256)
    assert off.shape == (256, 256, 3) and np.abs(off[..., 2]).max() == 0.0 and np.abs(off[0, 0]).max() < 1e-12   # anchors pin the border
    face = feed.generate_face_region(lm, 256)
    assert face.dtype == np.float32 and face[128, 128, 0] == 1.0 and face[0, 0, 0] == 0.0 and 0.2 < face.mean() < 0.8


def test_compact_maps_equal_downsampled_full_maps_bit_for_bit():
    """The generator reads uv / reg only through tf.image.resize(., [32,32]) (model.py:237, warp.py:137): evaluating the
    interpolants at the 4 centre samples of each 8x8 cell gives exactly downsample8(full maps)."""
    lm = landmarks(7)
    full = feed.frame_maps(lm)
    comp = feed.frame_maps_compact(lm)
    assert full["uv"].dtype == np.float32 and full["reg"].shape == (256, 256, 6) and full["face"].shape == (256, 256, 1)
    assert np.array_equal(comp["uv32"], downsample8(full["uv"][None])[0])
    assert np.array_equal(comp["reg32"], downsample8(full["reg"][None])[0])
    assert np.array_equal(comp["face"], full["face"])
    frame = feed.build_frame(np.zeros((256, 256, 6), np.float32), lm)
    assert frame.shape == (256, 256, 16) and np.array_equal(frame[..., 6:9], full["uv"]) and np.array_equal(frame[..., 15:], full["face"])
