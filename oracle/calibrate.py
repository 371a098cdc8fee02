"""ORACLE helper (test infrastructure): centre ``dif`` on the hole threshold for random weights.

``dif = gray*tanh(conv2(y)) + conv3(y)`` (model.py:246-251) shifts linearly with the conv3 bias, so
one oracle forward is enough to move the median of the 32x32 ``dif_small`` onto 0.1 (model.py:256):
roughly half the cells then take each ``bmask`` value, which exercises both sides of the
discontinuity instead of whatever constant a random seed happens to give.
"""
import numpy as np

from .generator_ref import generator_forward, HOLE_THRESHOLD


def centre_hole_threshold(weights, img, uv, reg=None, **kw):
    out = generator_forward(weights, img, uv, reg, **kw)
    shift = HOLE_THRESHOLD - float(np.median(out["dif_small"]))
    w = dict(weights)
    w["conv3/conv/bias"] = (weights["conv3/conv/bias"] + np.float32(shift)).astype(np.float32)
    return w
