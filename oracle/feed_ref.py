"""TEST INFRASTRUCTURE ONLY - independent CPU restatement of the reference's landmark map generators, used by tests/ to
check blindshadowremoval_b200/feed.py.  PARITY UNPINNED: matplotlib (which the reference uses, warp.py:5) is not
installable here, so neither implementation has been compared with reference outputs; they are two separate codings
(SciPy's compiled LinearNDInterpolator here, explicit barycentric algebra in feed.py) of

  generate_uv_map      /root/reference/warp.py:215-232
  generate_offset_map  /root/reference/warp.py:194-213
  generate_face_region /root/reference/utils.py:255-276

``mtri.Triangulation(x, y)`` = Delaunay triangulation (Qhull, options "Qt Qbb Qc Qz"); ``LinearTriInterpolator`` = the
plane through the three corner values of the containing triangle, masked (NaN after np.stack) outside the hull.
"""
import cv2
import numpy as np
from scipy.interpolate import LinearNDInterpolator
from scipy.spatial import Delaunay

ANCHORS = np.asarray([[0, 0], [0, 255], [255, 0], [255, 255], [0, 127], [127, 0], [255, 127], [127, 255],
                      [0, 63], [0, 191], [255, 63], [255, 191], [63, 0], [191, 0], [63, 255], [191, 255]]) / 255


def _interp(points, values, img_size):
    xi, yi = np.meshgrid(np.linspace(0, 1, img_size), np.linspace(0, 1, img_size))
    f = LinearNDInterpolator(Delaunay(np.asarray(points, np.float64), qhull_options="Qt Qbb Qc Qz"), np.asarray(values, np.float64))
    return f(xi, yi)


def generate_uv_map(source, uv, img_size):
    x, y, z = (_interp(source, uv[:, k], img_size) for k in range(3))
    return np.nan_to_num(np.stack([y, x, z], axis=2))


def generate_offset_map(source, target, img_size):
    s = np.concatenate([source, ANCHORS], axis=0).astype(np.float32)
    t = np.concatenate([target, ANCHORS], axis=0).astype(np.float32)
    off = s - t
    x, y = _interp(t, off[:, 0], img_size), _interp(t, off[:, 1], img_size)
    return np.stack([y, x, x * 0], axis=2)


def generate_face_region(source, img_size):
    morelm = np.copy(source[0:17, :])
    morelm[:, 1] = morelm[0, 1] - (morelm[:, 1] - morelm[0, 1]) * 0.8
    pts = np.concatenate([source, morelm], axis=0)
    m = np.nan_to_num(np.stack([_interp(pts, pts[:, 0], img_size)], axis=2))
    m = np.asarray(m > 0, np.float32)
    return cv2.GaussianBlur(m, (5, 5), 0).reshape([img_size, img_size, 1])
