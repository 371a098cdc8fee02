"""ORACLE PIN (test infrastructure): independent NumPy float64 restatement of the generator forward.

Written from the op *definitions*, deliberately not sharing code or formulations with
``oracle/generator_ref.py`` (which leans on torch's conv / conv_transpose):
  * Conv2D SAME      = explicit zero-padded window, one einsum per filter tap (model.py:119,140);
  * Conv2DTranspose  = scatter ``out[2i+kh, 2j+kw] += x[i,j] @ W[kh,kw].T`` then crop (model.py:153);
  * resize           = explicit half-pixel source coordinates per output pixel;
  * warp             = SciPy ``map_coordinates(order=1, mode='nearest')`` exactly as the reference's own
                       ``sp_batch_map_offsets`` / ``sp_batch_map_coordinates`` (warp.py:61-68, 118-131),
                       preceded by the resize-and-scale step of ``tf_batch_map_offsets`` (warp.py:137-139).
``tests/test_oracle.py`` requires both restatements to agree to 1e-5; parity with TensorFlow itself
stays unpinned (TF is not installable here).
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import map_coordinates

EPS = 1e-3
ALPHA = 0.3


def _pads(n, k, s):
    out = (n + s - 1) // s
    tot = max((out - 1) * s + k - n, 0)
    return out, tot // 2, tot - tot // 2


def conv_same(x, k, b, s=1):
    n, h, w, ci = x.shape
    kh, kw, _, co = k.shape
    oh, pt, pb = _pads(h, kh, s)
    ow, pl, pr = _pads(w, kw, s)
    xp = np.zeros((n, h + pt + pb, w + pl + pr, ci))
    xp[:, pt:pt + h, pl:pl + w] = x
    y = np.zeros((n, oh, ow, co))
    for i in range(kh):
        for j in range(kw):
            win = xp[:, i:i + (oh - 1) * s + 1:s, j:j + (ow - 1) * s + 1:s]
            y += np.einsum("nhwc,co->nhwo", win, k[i, j])
    return y + b


def convt_same(x, k, b):
    n, h, w, ci = x.shape
    kh, kw, co, _ = k.shape
    full = np.zeros((n, 2 * h + kh - 1, 2 * w + kw - 1, co))
    for i in range(kh):
        for j in range(kw):
            full[:, i:i + 2 * h:2, j:j + 2 * w:2] += np.einsum("nhwc,oc->nhwo", x, k[i, j])
    return full[:, :2 * h, :2 * w] + b


def bn(x, w, p):
    return w[p + "/gamma"] * (x - w[p + "/moving_mean"]) / np.sqrt(w[p + "/moving_variance"] + EPS) + w[p + "/beta"]


def lrelu(x):
    return np.maximum(x, 0) + ALPHA * np.minimum(x, 0)


def resize(x, oh, ow):
    n, h, w, c = x.shape
    out = np.zeros((n, oh, ow, c))
    for i in range(oh):
        sy = (i + 0.5) * h / oh - 0.5
        y0 = int(np.floor(sy)); fy = sy - y0
        ya, yb = min(max(y0, 0), h - 1), min(max(y0 + 1, 0), h - 1)
        for j in range(ow):
            sx = (j + 0.5) * w / ow - 0.5
            x0 = int(np.floor(sx)); fx = sx - x0
            xa, xb = min(max(x0, 0), w - 1), min(max(x0 + 1, 0), w - 1)
            top = x[:, ya, xa] * (1 - fx) + x[:, ya, xb] * fx
            bot = x[:, yb, xa] * (1 - fx) + x[:, yb, xb] * fx
            out[:, i, j] = top * (1 - fy) + bot * fy
    return out


def gray(x):
    return (x[..., :3] @ np.array([0.2989, 0.5870, 0.1140]))[..., None]


def warp(x, offsets):
    b, s = x.shape[0], x.shape[1]
    off = (resize(offsets, s, s) * s)[..., 0:2].reshape(b, -1, 2)
    grid = np.stack(np.mgrid[:s, :s], -1).reshape(-1, 2)
    coords = (off + grid[None]).clip(0, s - 1)
    out = np.zeros_like(x)
    for n in range(b):
        for c in range(x.shape[3]):
            out[n, :, :, c] = map_coordinates(x[n, :, :, c], coords[n].T, mode="nearest", order=1).reshape(s, s)
    return out


def share(x, reg, frame, do_share):
    if not do_share:
        return np.concatenate([x, x], -1)
    xr = warp(x, reg[..., 0:3])
    n, s, _, c = xr.shape
    g = xr.reshape(n // frame, frame, s, s, c)
    sh = np.concatenate([g.max(1), g.mean(1)], -1)
    sh = np.repeat(sh[:, None], frame, axis=1).reshape(n, s, s, 2 * c)
    return warp(sh, reg[..., 3:6])


def nonlocal_(x, w, p):
    n, h, wd, _ = x.shape
    g = conv_same(x, w[p + "/g/kernel"], w[p + "/g/bias"]).reshape(n, h * wd, -1)
    ph = conv_same(x, w[p + "/phi/kernel"], w[p + "/phi/bias"]).reshape(n, h * wd, -1)
    th = conv_same(x, w[p + "/theta/kernel"], w[p + "/theta/bias"]).reshape(n, h * wd, -1)
    f = np.einsum("nqd,nkd->nqk", th, ph)
    f = np.exp(f - f.max(-1, keepdims=True))
    f = f / f.sum(-1, keepdims=True)
    y = np.einsum("nqk,nkd->nqd", f, g).reshape(n, h, wd, -1)
    return x + bn(conv_same(y, w[p + "/w/kernel"], w[p + "/w/bias"]), w, p + "/bnorm")


def resblock(x, w, p):
    y = lrelu(bn(conv_same(x, w[p + "/conv1/kernel"], w[p + "/conv1/bias"]), w, p + "/bnorm1"))
    y = lrelu(bn(conv_same(y, w[p + "/conv2/kernel"], w[p + "/conv2/bias"]), w, p + "/bnorm2"))
    y = bn(conv_same(y, w[p + "/conv3/kernel"], w[p + "/conv3/bias"]), w, p + "/bnorm3")
    y = nonlocal_(y, w, p + "/non_local")
    c = max(x.shape[-1], y.shape[-1])
    xe = np.zeros(x.shape[:3] + (c,)); xe[..., :x.shape[-1]] = x
    ye = np.zeros(y.shape[:3] + (c,)); ye[..., :y.shape[-1]] = y
    return lrelu(xe + ye)


def forward(weights, img, uv, reg=None, variant="gsc", frame=1, do_share=True):
    w = {k: np.asarray(v, np.float64) for k, v in weights.items()}
    img = np.asarray(img, np.float64); uv = np.asarray(uv, np.float64)
    tsm = variant == "tsm"
    if tsm:
        reg = np.asarray(reg, np.float64)

    def C(x, name, s=1, norm=True, act=True):
        y = conv_same(x, w[name + "/conv/kernel"], w[name + "/conv/bias"], s)
        if norm:
            y = bn(y, w, name + "/bnorm")
        return lrelu(y) if act else y

    def T(x, name):
        return lrelu(bn(convt_same(x, w[name + "/conv/kernel"], w[name + "/conv/bias"]), w, name + "/bnorm"))

    x1 = C(img, "conv1"); x2 = C(x1, "down1", 2); x3 = C(x2, "down2", 2); x = C(x3, "down3", 2)
    s = x.shape[1]
    uvs = resize(uv, s, s)
    x = np.concatenate([x, share(x, reg, frame, do_share), uvs], -1) if tsm else np.concatenate([x, uvs], -1)
    for i in range(3):
        x = resblock(x, w, "res_stack/%d" % i)
    y = T(x, "up1"); y = T(np.concatenate([y, x3], -1), "up2"); y = T(np.concatenate([y, x2], -1), "up3")
    mask = np.tanh(C(y, "conv2", norm=False, act=False))
    con = C(y, "conv3", norm=False, act=False)
    g0 = gray(img)
    gs = g0 * (1 + mask) + con
    dif_gs = gs - g0
    mask22 = np.concatenate([np.maximum(mask, 0), mask * 0, np.maximum(-mask, 0)], -1)
    dif_small = resize(dif_gs, s, s)
    bmask = (dif_small > 0.1).astype(np.float64)
    xh = x * (1 - bmask)
    if tsm:
        x = np.concatenate([xh, bmask, share(xh, reg, frame, do_share), uvs], -1)
    else:
        x = np.concatenate([xh, bmask, uvs], -1)
    for i in range(3, 6):
        x = resblock(x, w, "res_stack/%d" % i)
    f = T(T(T(x, "clr_up1"), "clr_up2"), "clr_up3")
    c = C(C(C(np.concatenate([gs, f], -1), "clr_conv1"), "clr_conv2"), "clr_conv3", norm=False, act=False)
    return dict(gs=gs, con_rgb=c, mask22=mask22, dif=gray(c) - g0, bmask=bmask, dif_small=dif_small)
