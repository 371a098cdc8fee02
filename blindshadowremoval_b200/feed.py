"""Host-side batch feeding: the landmark-driven maps the reference's dataset code attaches to every face crop
(SURVEY.md 8f rows 1-2), restated without matplotlib.

The reference builds, per frame, ``[img(3) | gt(3) | uv map(3) | reg_in(3) | reg_out(3) | face(1)]``
(/root/reference/dataset.py:616-640, 111-130) with three generators that interpolate values given at the 68 facial
landmarks over a Delaunay triangulation (``matplotlib.tri.Triangulation`` + ``LinearTriInterpolator``):

  generate_uv_map(source, uv, S)         warp.py:215-232   canonical-UV map, 0 outside the landmark hull
  generate_offset_map(source, target, S) warp.py:194-213   registration offsets source - target, on `target` + 16 anchors
  generate_face_region(source, S)        utils.py:255-276  {0,1} hull mask (jaw mirrored upwards), 5x5 Gaussian blur

Here the triangulation is ``scipy.spatial.Delaunay`` with matplotlib's Qhull options and the interpolation is the
barycentric form of the same piecewise-linear function.  Because the generator consumes uv / reg only after an 8x
down-sample (model.py:237, warp.py:137), ``frame_maps_compact`` evaluates the interpolants only at the four centre
samples of every 8x8 cell (4096 instead of 65536 points per map) and returns exactly ``downsample8`` of the full maps -
the inputs of ``Generator.forward_compact``.

Parity: matplotlib is not installable in the build container, so these functions are pinned by properties (exact at
the landmarks, exact for affine data, zero outside the hull, compact == downsample8(full)) and by an independent
restatement in oracle/feed_ref.py, not by reference outputs ("parity unpinned" for this row).
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np
from scipy.spatial import Delaunay

IMG = 256
FEAT = 32
_HERE = os.path.dirname(os.path.abspath(__file__))
_TEMPLATE = None

# the 16 fixed border points of generate_offset_map (warp.py:195-199), in pixel units of a 256 grid
_ANCHORS = np.asarray([[0, 0], [0, 255], [255, 0], [255, 255], [0, 127], [127, 0], [255, 127], [127, 255],
                       [0, 63], [0, 191], [255, 63], [255, 191], [63, 0], [191, 0], [63, 255], [191, 255]], np.float64) / 255


def face_template() -> Tuple[np.ndarray, np.ndarray]:
    """(uv[68,3], lm_ref[68,2]) of dataset.py:10-16, extracted by tools/extract_face_template.py."""
    global _TEMPLATE
    if _TEMPLATE is None:
        z = np.load(os.path.join(_HERE, "data", "face_template.npz"))
        _TEMPLATE = (z["uv"].astype(np.float32), z["lm_ref"].astype(np.float32))
    return _TEMPLATE


class TriInterpolator:
    """Piecewise-linear interpolation over the Delaunay triangulation of ``points`` (= ``mtri.Triangulation(x, y)`` +
    ``mtri.LinearTriInterpolator``); NaN outside the convex hull, like matplotlib's masked result after ``np.stack``."""

    def __init__(self, points: np.ndarray):
        self.points = np.asarray(points, np.float64)
        self.tri = Delaunay(self.points, qhull_options="Qt Qbb Qc Qz")     # matplotlib's Qhull options
        self._simplex = None
        self._bary = None

    def locate(self, xq: np.ndarray, yq: np.ndarray) -> None:
        """Find the triangle and barycentric weights of every query point once; ``__call__`` reuses them."""
        q = np.stack([np.asarray(xq, np.float64).ravel(), np.asarray(yq, np.float64).ravel()], axis=1)
        s = self.tri.find_simplex(q)
        t = self.tri.transform[np.maximum(s, 0)]                            # [Q,3,2]: inverse edge matrix and origin
        b2 = np.einsum("qij,qj->qi", t[:, :2, :], q - t[:, 2, :])
        self._bary = np.concatenate([b2, 1.0 - b2.sum(axis=1, keepdims=True)], axis=1)
        self._simplex = s
        self._shape = np.shape(xq)

    def __call__(self, values: np.ndarray) -> np.ndarray:
        """values[P] or [P,K] at the points -> interpolated [*query shape(,K)] float64."""
        v = np.asarray(values, np.float64)
        vk = v[:, None] if v.ndim == 1 else v
        corner = vk[self.tri.simplices[np.maximum(self._simplex, 0)]]      # [Q,3,K]
        out = np.einsum("qc,qck->qk", self._bary, corner)
        out[self._simplex < 0] = np.nan
        out = out.reshape(self._shape + (vk.shape[1],))
        return out[..., 0] if v.ndim == 1 else out


def _grid(img_size: int):
    return np.meshgrid(np.linspace(0, 1, img_size), np.linspace(0, 1, img_size))      # xi (columns), yi (rows)


def _cell_centre_grid():
    """The 4 x 1024 pixel positions tf.image.resize(x, [32,32]) reads from a 256 map: rows/cols 8i+3, 8i+4."""
    lin = np.linspace(0, 1, IMG)
    idx = np.stack([8 * np.arange(FEAT) + 3, 8 * np.arange(FEAT) + 4], axis=1).ravel()   # 3,4,11,12,...
    return np.meshgrid(lin[idx], lin[idx])                                                 # [64,64] each


def _pool_centres(v64: np.ndarray) -> np.ndarray:
    """[64,64,K] samples at the cell-centre grid -> [32,32,K], same arithmetic as generator.downsample8."""
    v = v64.astype(np.float32)
    h = np.float32(0.5)
    top = v[0::2, 0::2] * h + v[0::2, 1::2] * h
    bot = v[1::2, 0::2] * h + v[1::2, 1::2] * h
    return np.ascontiguousarray(top * h + bot * h)


_LOCATED: Dict[tuple, "TriInterpolator"] = {}


def _located(points: np.ndarray, grid: str) -> "TriInterpolator":
    """Triangulation of `points` with every pixel of the 256 x 256 grid ('full') or of the cell-centre grid ('centres')
    already located.  The un-warp field of every frame interpolates over the SAME point set - the template landmarks plus
    the 16 border anchors (dataset.py:635) - so its Delaunay triangulation, point location and barycentric weights are
    computed once per process instead of once per frame; other point sets (the frame's own landmarks) are not cached."""
    key = (grid, points.shape, points.dtype.str, points.tobytes())
    it = _LOCATED.get(key)
    if it is None:
        it = TriInterpolator(points)
        it.locate(*(_grid(IMG) if grid == "full" else _cell_centre_grid()))
        if len(_LOCATED) >= 8:
            _LOCATED.clear()
        _LOCATED[key] = it
    return it


def _uv_values(uv: np.ndarray) -> np.ndarray:
    return np.stack([uv[:, 1], uv[:, 0], uv[:, 2]], axis=1)       # stack([_offsetmapy, _offsetmapx, _offsetmapz]), warp.py:230


def generate_uv_map(source: np.ndarray, uv: np.ndarray, img_size: int = IMG) -> np.ndarray:
    """warp.py:215-232.  source[68,2] landmarks in [0,1] (x, y); uv[68,3] -> [S,S,3] float64, 0 outside the hull."""
    it = TriInterpolator(source)
    it.locate(*_grid(img_size))
    return np.nan_to_num(it(_uv_values(np.asarray(uv))))


def generate_offset_map(source: np.ndarray, target: np.ndarray, img_size: int = IMG) -> np.ndarray:
    """warp.py:194-213.  Offsets source - target interpolated over `target` + 16 border anchors -> [S,S,3] float64
    = (dy, dx, 0).  No nan_to_num in the reference; with the anchors the hull is the whole image."""
    src = np.concatenate([np.asarray(source, np.float64), _ANCHORS], axis=0).astype(np.float32)
    tgt = np.concatenate([np.asarray(target, np.float64), _ANCHORS], axis=0).astype(np.float32)
    off = (src - tgt).astype(np.float64)
    if img_size == IMG and _is_template(target):
        it = _located(tgt, "full")
    else:
        it = TriInterpolator(tgt)
        it.locate(*_grid(img_size))
    m = it(np.stack([off[:, 1], off[:, 0]], axis=1))
    return np.concatenate([m, m[..., 1:2] * 0], axis=2)


def _is_template(points: np.ndarray) -> bool:
    """True for the (constant) template landmarks: only those are worth caching."""
    ref = face_template()[1]
    p = np.asarray(points)
    return p.shape == ref.shape and np.array_equal(p.astype(np.float32), ref)


def _hull_points(source: np.ndarray) -> np.ndarray:
    morelm = np.copy(source[0:17, :])                                       # utils.py:256-258: jaw line mirrored upwards
    morelm[:, 1] = morelm[0, 1] - (morelm[:, 1] - morelm[0, 1]) * 0.8
    return np.concatenate([source, morelm], axis=0)


def _gaussian5(mask: np.ndarray) -> np.ndarray:
    """cv2.GaussianBlur(x, (5,5), 0) for float32: separable [1,4,6,4,1]/16, BORDER_REFLECT_101."""
    k = np.asarray([1, 4, 6, 4, 1], np.float32) / np.float32(16)
    p = np.pad(mask.astype(np.float32), 2, mode="reflect")
    rows = sum(k[i] * p[:, i:i + mask.shape[1]] for i in range(5))
    return sum(k[i] * rows[i:i + mask.shape[0], :] for i in range(5)).astype(np.float32)


def generate_face_region(source: np.ndarray, img_size: int = IMG) -> np.ndarray:
    """utils.py:255-276 -> [S,S,1] float32: 1 inside the hull of landmarks + mirrored jaw (where the interpolated
    x coordinate is > 0), blurred 5x5."""
    pts = _hull_points(np.asarray(source))
    it = TriInterpolator(pts)
    it.locate(*_grid(img_size))
    inside = np.asarray(np.nan_to_num(it(pts[:, 0].astype(np.float64))) > 0, np.float32)
    return _gaussian5(inside).reshape(img_size, img_size, 1)


def frame_maps(lm: np.ndarray, uv: Optional[np.ndarray] = None, lm_ref: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """The per-frame maps of dataset.py:633-637 at 256x256: uv[256,256,3], reg[256,256,6] = reg_in | reg_out,
    face[256,256,1] (float32, as the chunk is cast at dataset.py:142)."""
    tuv, tref = face_template()
    uv = tuv if uv is None else uv
    lm_ref = tref if lm_ref is None else lm_ref
    reg = np.concatenate([generate_offset_map(lm, lm_ref), generate_offset_map(lm_ref, lm)], axis=2)
    return {"uv": generate_uv_map(lm, uv).astype(np.float32), "reg": reg.astype(np.float32),
            "face": generate_face_region(lm)}


def frame_maps_compact(lm: np.ndarray, uv: Optional[np.ndarray] = None, lm_ref: Optional[np.ndarray] = None,
                       with_face: bool = True) -> Dict[str, np.ndarray]:
    """uv32[32,32,3], reg32[32,32,6] (and face[256,256,1]) = ``downsample8`` of ``frame_maps`` bit for bit, from 16x
    fewer interpolation points: the inputs of ``Generator.forward_compact`` / ``bsr_forward_*_host_compact``."""
    tuv, tref = face_template()
    uv = tuv if uv is None else uv
    lm_ref = tref if lm_ref is None else lm_ref
    xq, yq = _cell_centre_grid()
    it = TriInterpolator(lm)
    it.locate(xq, yq)
    uv64 = np.nan_to_num(it(_uv_values(np.asarray(uv))))
    regs = []
    for source, target in ((lm, lm_ref), (lm_ref, lm)):
        src = np.concatenate([np.asarray(source, np.float64), _ANCHORS], axis=0).astype(np.float32)
        tgt = np.concatenate([np.asarray(target, np.float64), _ANCHORS], axis=0).astype(np.float32)
        off = (src - tgt).astype(np.float64)
        if _is_template(target):
            io = _located(tgt, "centres")
        else:
            io = TriInterpolator(tgt)
            io.locate(xq, yq)
        m = io(np.stack([off[:, 1], off[:, 0]], axis=1))
        regs.append(np.concatenate([m, m[..., 1:2] * 0], axis=2))
    out = {"uv32": _pool_centres(uv64), "reg32": _pool_centres(np.concatenate(regs, axis=2))}
    if with_face:
        out["face"] = generate_face_region(lm)
    return out


# landmark order after a horizontal flip (utils.py:360-364, 1-based there)
_MIRROR_ORDER = np.asarray([17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 27, 26, 25, 24, 23, 22, 21, 20, 19, 18,
                            28, 29, 30, 31, 36, 35, 34, 33, 32, 46, 45, 44, 43, 48, 47, 40, 39, 38, 37, 42, 41,
                            55, 54, 53, 52, 51, 50, 49, 60, 59, 58, 57, 56, 65, 64, 63, 62, 61, 68, 67, 66], np.int64) - 1


def crop_and_resize(img: np.ndarray, lm: np.ndarray, fsize: int = IMG):
    """``face_crop_and_resize(img, lm, fsize)`` of utils.py:356-433 with aug=False (the test-time call of
    dataset.py:165, 632): square box of half-size 1.4 x the larger landmark half-extent around the landmark centre,
    shifted up by 20 % of it; the image is zero-padded when the box leaves it; ``cv2.resize`` (bilinear) to fsize.
    Returns (crop[fsize,fsize,C] float64, lm / (2 length), mirrored lm / (2 length), box[4]) like the reference.
    Pinned against the reference's own function on its real files (tests/golden/crop_reference.npz)."""
    import cv2
    img = np.asarray(img)
    lm = np.array(lm, copy=True)
    h, w = img.shape[0], img.shape[1]
    lm_mirror = np.array(lm, copy=True)
    lm_mirror[:, 0] = w - lm_mirror[:, 0]
    lm_mirror = lm_mirror[_MIRROR_ORDER, :]
    x0, x1, y0, y1 = np.min(lm[:, 0]), np.max(lm[:, 0]), np.min(lm[:, 1]), np.max(lm[:, 1])
    cx, cy = (x0 + x1) / 2, (y0 + y1) / 2
    length = np.max([(x1 - x0) / 2, (y1 - y0) / 2]) * 1.4
    il, iu = int(length), int(length * 1.2)
    box = [int(cx) - il, int(cy) - iu, int(cx) + il, int(cy) + il + il - iu]
    box0 = list(box)
    box_m = [w - box[2], box[1], w - box[0], box[3]]
    lm[:, 0] -= box[0]
    lm[:, 1] -= box[1]
    lm_mirror[:, 0] -= box_m[0]
    lm_mirror[:, 1] -= box_m[1]
    pad_x = max(-box[0], box[2] - w) if (box[0] < 0 or box[2] > w) else 0
    pad_y = max(-box[1], box[3] - h) if (box[1] < 0 or box[3] > h) else 0
    if pad_x > 0 or pad_y > 0:
        big = np.zeros((h + 2 * pad_y + 2, w + 2 * pad_x + 2, img.shape[2]))
        big[pad_y:pad_y + h, pad_x:pad_x + w, :] = img
        img = big
        box = [box[0] + pad_x, box[1] + pad_y, box[2] + pad_x, box[3] + pad_y]
    crop = img[box[1]:box[3], box[0]:box[2], :]
    if crop.shape[0] == crop.shape[1] and crop.shape[0] > 0:
        crop = cv2.resize(crop, (fsize, fsize))
    else:
        crop = np.zeros((fsize, fsize, img.shape[2]))
    return crop, lm / (length * 2), lm_mirror / (length * 2), box0


def load_frame(png_path: str, npy_path: Optional[str] = None, gt_path: Optional[str] = None) -> Dict[str, np.ndarray]:
    """One real sample as dataset.py:157-170 / 625-637 builds it: PNG -> RGB / 255, ground truth (the image itself for
    in-the-wild data, dataset.py:623), crop, landmark maps.  Returns img[256,256,3], gt[256,256,3], uv, reg, face
    (float32), lm[68,2] and box."""
    import cv2
    npy_path = npy_path or png_path[:-4] + ".npy"
    img = cv2.cvtColor(cv2.imread(png_path), cv2.COLOR_BGR2RGB) / 255.
    gt = cv2.cvtColor(cv2.imread(gt_path or png_path), cv2.COLOR_BGR2RGB) / 255.
    crop, lm, _, box = crop_and_resize(np.concatenate([img, gt], axis=2), np.load(npy_path), IMG)
    m = frame_maps(lm)
    return {"img": crop[..., 0:3].astype(np.float32), "gt": crop[..., 3:6].astype(np.float32), "uv": m["uv"], "reg": m["reg"],
            "face": m["face"], "lm": lm, "box": np.asarray(box)}


def build_chunk(frames) -> np.ndarray:
    """[F,256,256,16] = img3|gt3|uv3|reg6|face1 per frame (dataset.py:170, 296-302), float32."""
    return np.stack([np.concatenate([f["img"], f["gt"], f["uv"], f["reg"], f["face"]], axis=2) for f in frames]).astype(np.float32)


def build_frame(img6: np.ndarray, lm: np.ndarray) -> np.ndarray:
    """One frame of the 16-channel chunk, dataset.py:637: concat([img|gt (6), uvm, reg_in, reg_out, face])."""
    m = frame_maps(lm)
    return np.concatenate([np.asarray(img6, np.float32), m["uv"], m["reg"], m["face"]], axis=2).astype(np.float32)
