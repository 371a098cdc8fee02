"""ORACLE (test infrastructure, not product code): CPU restatement of the reference generator forward.

Parity pinning: the reference (TensorFlow 2.3 / Keras 2.4) cannot be imported in the build image and
ships no tests, golden vectors or trained weights for this path, so **parity against the reference's
own outputs is unpinned**.  This restatement is pinned instead by (a) ``oracle/np64_ref.py``, an
independent NumPy float64 restatement written from the op definitions (scatter-form transposed
conv, explicit window sums, explicit half-pixel resize) that must agree to 1e-5, (b) the variable
names/shapes of ``/root/reference/log/*/ckpt-*.index`` and (c) SciPy ``map_coordinates`` for the warp
(the reference's own ``sp_batch_map_offsets``, warp.py:118-131).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg
may import this module.  Weights are a dict of NumPy arrays in TF layout (see
``blindshadowremoval_b200/weights.py``); activations are NHWC like the reference.

Reference lines followed: /root/reference/model.py:6-61 (NonLocalBlock), 81-113 (ResBottleneck),
115-147 (Conv), 149-177 (ConvT), 228-290 (Generator.call); model_with_TSM.py:199-229 (ShareLayer),
261-325 (Generator.call); warp.py:71-115, 134-165; caller glue train_test_GSC.py:808-809, 711-718.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3        # keras.layers.BatchNormalization default epsilon
LEAKY_ALPHA = 0.3    # keras.layers.LeakyReLU default alpha
GRAY = (0.2989, 0.5870, 0.1140)   # tf.image.rgb_to_grayscale
HOLE_THRESHOLD = 0.1  # model.py:256


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a)).to(dtype)


def _same_pads(size: int, k: int, s: int):
    """TF 'SAME': out = ceil(in/s); total = max((out-1)*s + k - in, 0); before = total//2."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv2d_same(x, kernel, bias, stride=1):
    """keras Conv2D(padding='same') on NHWC ``x``; kernel [kh,kw,in,out] (model.py:119,140)."""
    kh, kw = kernel.shape[0], kernel.shape[1]
    pt, pb = _same_pads(x.shape[1], kh, stride)
    pl, pr = _same_pads(x.shape[2], kw, stride)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xn, kernel.permute(3, 2, 0, 1), bias, stride=stride)
    return y.permute(0, 2, 3, 1)


def conv2d_transpose_same(x, kernel, bias):
    """keras Conv2DTranspose(3x3, strides 2, 'same'); kernel [kh,kw,out,in] (model.py:153,170).

    TF defines it as the gradient of the SAME stride-2 conv: out[2i+kh, 2j+kw, o] += x[i,j,c] *
    W[kh,kw,o,c], output cropped to [2H, 2W] (the SAME conv pads 0 before / 1 after on even sizes).
    """
    n, h, w, _ = x.shape
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2), kernel.permute(3, 2, 0, 1), bias, stride=2, padding=0)
    return y[:, :, :2 * h, :2 * w].permute(0, 2, 3, 1)


def batchnorm(x, w, prefix):
    g, b = w[prefix + "/gamma"], w[prefix + "/beta"]
    m, v = w[prefix + "/moving_mean"], w[prefix + "/moving_variance"]
    return (x - m) * (g / torch.sqrt(v + BN_EPS)) + b


def leaky(x):
    return torch.where(x >= 0, x, x * LEAKY_ALPHA)


def resize_bilinear(x, out_h: int, out_w: int):
    """tf.image.resize default: bilinear, half-pixel centres, antialias=False (model.py:237,256)."""
    def axis(n_in, n_out):
        src = (torch.arange(n_out, dtype=x.dtype) + 0.5) * (n_in / n_out) - 0.5
        lo = torch.floor(src)
        frac = src - lo
        lo_i = lo.long().clamp(0, n_in - 1)
        hi_i = (lo.long() + 1).clamp(0, n_in - 1)
        return lo_i, hi_i, frac
    r0, r1, fr = axis(x.shape[1], out_h)
    c0, c1, fc = axis(x.shape[2], out_w)
    rows = x[:, r0] * (1 - fr)[None, :, None, None] + x[:, r1] * fr[None, :, None, None]
    return rows[:, :, c0] * (1 - fc)[None, None, :, None] + rows[:, :, c1] * fc[None, None, :, None]


def rgb_to_gray(x):
    return x[..., 0:1] * GRAY[0] + x[..., 1:2] * GRAY[1] + x[..., 2:3] * GRAY[2]


def conv_block(x, w, name, stride=1, norm=True, act=True):
    """model.py:115-147 (``Conv``)."""
    y = conv2d_same(x, w[name + "/conv/kernel"], w[name + "/conv/bias"], stride)
    if norm:
        y = batchnorm(y, w, name + "/bnorm")
    return leaky(y) if act else y


def convt_block(x, w, name):
    """model.py:149-177 (``ConvT``)."""
    y = conv2d_transpose_same(x, w[name + "/conv/kernel"], w[name + "/conv/bias"])
    return leaky(batchnorm(y, w, name + "/bnorm"))


def non_local(x, w, p):
    """model.py:23-61 with pool=False; softmax logits are *not* scaled."""
    n, h, wd, _ = x.shape
    g = conv2d_same(x, w[p + "/g/kernel"], w[p + "/g/bias"]).reshape(n, h * wd, -1)
    phi = conv2d_same(x, w[p + "/phi/kernel"], w[p + "/phi/bias"]).reshape(n, h * wd, -1)
    theta = conv2d_same(x, w[p + "/theta/kernel"], w[p + "/theta/bias"]).reshape(n, h * wd, -1)
    f = torch.matmul(theta, phi.transpose(1, 2))
    y = torch.matmul(torch.softmax(f, dim=-1), g).reshape(n, h, wd, -1)
    w_y = batchnorm(conv2d_same(y, w[p + "/w/kernel"], w[p + "/w/bias"]), w, p + "/bnorm")
    return x + w_y


def res_bottleneck(x, w, p):
    """model.py:98-113 (stride 1): narrower of (x, y) is zero-extended in channels."""
    y = leaky(batchnorm(conv2d_same(x, w[p + "/conv1/kernel"], w[p + "/conv1/bias"]), w, p + "/bnorm1"))
    y = leaky(batchnorm(conv2d_same(y, w[p + "/conv2/kernel"], w[p + "/conv2/bias"]), w, p + "/bnorm2"))
    y = batchnorm(conv2d_same(y, w[p + "/conv3/kernel"], w[p + "/conv3/bias"]), w, p + "/bnorm3")
    y = non_local(y, w, p + "/non_local")
    cx, cy = x.shape[-1], y.shape[-1]
    if cx < cy:
        x = F.pad(x, (0, cy - cx))
    elif cy < cx:
        y = F.pad(y, (0, cx - cy))
    return leaky(x + y)


def batch_map_offsets(x, offsets):
    """warp.py:134-165 + 71-115: bilinear warp of ``x`` [B,s,s,C] by ``offsets`` [B,256,256,3]."""
    b, s = x.shape[0], x.shape[1]
    off = resize_bilinear(offsets, s, s) * s           # warp.py:137
    off = off[..., 0:2]                                # warp.py:139 (Δrow, Δcol)
    gi, gj = torch.meshgrid(torch.arange(s, dtype=x.dtype), torch.arange(s, dtype=x.dtype), indexing="ij")
    c0 = (off[..., 0] + gi).clamp(0, s - 1)            # warp.py:85
    c1 = (off[..., 1] + gj).clamp(0, s - 1)
    f0, f1 = torch.floor(c0), torch.floor(c1)
    lt0, lt1 = f0.long(), f1.long()
    rb0, rb1 = torch.ceil(c0).long(), torch.ceil(c1).long()
    bi = torch.arange(b)[:, None, None].expand(b, s, s)
    v_lt = x[bi, lt0, lt1]
    v_rb = x[bi, rb0, rb1]
    v_lb = x[bi, lt0, rb1]                             # warp.py:88: (lt[0], rb[1])
    v_rt = x[bi, rb0, lt1]                             # warp.py:89: (rb[0], lt[1])
    o0 = (c0 - f0)[..., None]
    o1 = (c1 - f1)[..., None]
    v_t = v_lt + (v_rt - v_lt) * o0
    v_b = v_lb + (v_rb - v_lb) * o0
    return v_t + (v_b - v_t) * o1


def share_layer(x, reg, frame: int, share: bool):
    """model_with_TSM.py:204-229.  ``x`` is [n_chunks*frame, s, s, C]; the reference handles one chunk
    per call (reshape [1, frame, ...]); chunks here are independent groups of ``frame``."""
    if not share:
        return torch.cat([x, x], dim=-1)
    reg_in, reg_out = reg[..., 0:3], reg[..., 3:6]
    x_reg = batch_map_offsets(x, reg_in)
    n, s, _, c = x_reg.shape
    g = x_reg.reshape(n // frame, frame, s, s, c)
    sh = torch.cat([g.max(dim=1).values, g.mean(dim=1)], dim=-1)
    sh = sh[:, None].expand(n // frame, frame, s, s, 2 * c).reshape(n, s, s, 2 * c)
    return batch_map_offsets(sh, reg_out)


def generator_forward(weights: Dict[str, np.ndarray], img, uv, reg=None, *, variant: str = "gsc",
                      frame: int = 1, share: bool = True, dtype=torch.float32,
                      bmask_override: Optional[np.ndarray] = None, keep: bool = False):
    """Forward of ``Generator.call`` (model.py:228-290 / model_with_TSM.py:261-325), inference mode.

    Returns a dict with the four reference outputs ``gs, con_rgb, mask22, dif`` plus ``bmask`` and
    ``dif_small`` (the 32x32 resize that is thresholded); with ``keep`` also every intermediate.
    ``bmask_override`` [N,32,32,1] replaces the thresholded hole mask (used to compare a bf16 run
    whose near-threshold cells flipped).
    """
    w = {k: _t(v, dtype) for k, v in weights.items()}
    img, uv = _t(img, dtype), _t(uv, dtype)
    tsm = variant == "tsm"
    if tsm:
        reg = _t(reg, dtype)
        if img.shape[0] % frame:
            raise ValueError("batch %d is not a multiple of frame %d" % (img.shape[0], frame))
    t = {}
    x1 = conv_block(img, w, "conv1")
    x2 = conv_block(x1, w, "down1", stride=2)
    x3 = conv_block(x2, w, "down2", stride=2)
    x = conv_block(x3, w, "down3", stride=2)
    s = x.shape[1]
    uv_s = resize_bilinear(uv, s, s)
    if tsm:
        x = torch.cat([x, share_layer(x, reg, frame, share), uv_s], dim=-1)
    else:
        x = torch.cat([x, uv_s], dim=-1)
    t.update(x1=x1, x2=x2, x3=x3, x_in0=x)
    for i in range(3):
        x = res_bottleneck(x, w, "res_stack/%d" % i)
        t["res%d" % i] = x
    y = convt_block(x, w, "up1")
    t["up1"] = y
    y = convt_block(torch.cat([y, x3], dim=-1), w, "up2")
    t["up2"] = y
    y = convt_block(torch.cat([y, x2], dim=-1), w, "up3")
    t["up3"] = y
    mask = torch.tanh(conv_block(y, w, "conv2", norm=False, act=False))
    con = conv_block(y, w, "conv3", norm=False, act=False)
    gray = rgb_to_gray(img)
    gs = gray * (1 + mask) + con
    dif_gs = gs - gray
    mask22 = torch.cat([torch.relu(mask), mask * 0, torch.relu(-mask)], dim=-1)
    dif_small = resize_bilinear(dif_gs, s, s)
    bmask = (dif_small > HOLE_THRESHOLD).to(dtype)
    if bmask_override is not None:
        bmask = _t(bmask_override, dtype)
    x_hole = x * (1 - bmask)
    if tsm:
        x = torch.cat([x_hole, bmask, share_layer(x_hole, reg, frame, share), uv_s], dim=-1)
    else:
        x = torch.cat([x_hole, bmask, uv_s], dim=-1)
    t["x_in3"] = x
    for i in range(3, 6):
        x = res_bottleneck(x, w, "res_stack/%d" % i)
        t["res%d" % i] = x
    f = convt_block(x, w, "clr_up1")
    t["clr_up1"] = f
    f = convt_block(f, w, "clr_up2")
    t["clr_up2"] = f
    f = convt_block(f, w, "clr_up3")
    t["clr_up3"] = f
    c = conv_block(torch.cat([gs, f], dim=-1), w, "clr_conv1")
    c = conv_block(c, w, "clr_conv2")
    con_rgb = conv_block(c, w, "clr_conv3", norm=False, act=False)
    dif = rgb_to_gray(con_rgb) - gray
    out = dict(gs=gs, con_rgb=con_rgb, mask22=mask22, dif=dif, bmask=bmask, dif_small=dif_small)
    if keep:
        out.update(t)
    return {k: v.detach().cpu().numpy() for k, v in out.items()}


def caller_glue(con_rgb, dif, face):
    """train_test_GSC.py:808-809 / 872-873 / 902-903: ``mask_pred = dif*face``; ``clip(con_rgb,0,1)``."""
    return np.clip(con_rgb, 0.0, 1.0), dif * face


def composite(pred, inp, m):
    """train_test_GSC.py:711,718: ``clip(pred*m + inp*(1-m), 0, 1)``."""
    return np.clip(pred * m + inp * (1.0 - m), 0.0, 1.0)
