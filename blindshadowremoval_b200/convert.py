"""One-time weight converter: Keras/TF generator variables -> the library's canonical blob.

Input: {name: ndarray} in TF layout (``weights.variable_shapes``) from ``tf_checkpoint.read_generator_weights``
or ``weights.random_weights``.  Output: bytes for ``bsr_load_weights``: per layer a BN-folded fp32
kernel ``[tap][cin][cout]`` + bias.  Device-specific packing (bf16, K-major, channel padding, TMA
tensor maps) happens inside libbsr at load time.

Folding (inference BatchNormalization, eps 1e-3; /root/reference/model.py:99-101,142,172,58):
    s = gamma / sqrt(moving_variance + eps);  W' = W * s[out];  b' = (b - moving_mean) * s + beta
Fusions: theta|phi|g -> one 1x1 ``qkv`` (257->384); conv2|conv3 -> one 7x7 ``heads`` (64->2);
clr_conv1's input channels are reordered from TF's [gs, f0..f63] (model.py:267) to [f0..f63, gs] so
that ``gs`` is the last channel of the buffer clr_up3 writes into.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

from .weights import BN_EPS, N_RES, VARIANTS, check_weights

MAGIC = b"BSRW0001"


def _fold(kernel_tap_cin_cout: np.ndarray, bias: np.ndarray, w: Dict[str, np.ndarray], bn_prefix):
    k = kernel_tap_cin_cout.astype(np.float64)
    b = bias.astype(np.float64)
    if bn_prefix is not None:
        s = w[bn_prefix + "/gamma"].astype(np.float64) / np.sqrt(w[bn_prefix + "/moving_variance"].astype(np.float64) + BN_EPS)
        k = k * s[None, None, :]
        b = (b - w[bn_prefix + "/moving_mean"]) * s + w[bn_prefix + "/beta"]
    return np.ascontiguousarray(k, np.float32), np.ascontiguousarray(b, np.float32)


def canonical_layers(variant: str, w: Dict[str, np.ndarray]) -> List[Tuple[str, int, int, int, np.ndarray, np.ndarray]]:
    """[(name, kh, kw, transposed, kernel[tap,cin,cout], bias[cout])] in execution order."""
    check_weights(variant, w)
    out = []

    def conv(dst, src, bn=True):
        k = w[src + "/conv/kernel"]
        kh, kw, ci, co = k.shape
        kk, bb = _fold(k.reshape(kh * kw, ci, co), w[src + "/conv/bias"], w, src + "/bnorm" if bn else None)
        out.append((dst, kh, kw, 0, kk, bb))

    def convt(dst, src):
        k = w[src + "/conv/kernel"]                       # [kh, kw, out, in]
        kh, kw, co, ci = k.shape
        kk, bb = _fold(k.transpose(0, 1, 3, 2).reshape(kh * kw, ci, co), w[src + "/conv/bias"], w, src + "/bnorm")
        out.append((dst, kh, kw, 1, kk, bb))

    conv("conv1", "conv1"); conv("down1", "down1"); conv("down2", "down2"); conv("down3", "down3")
    for i in range(N_RES):
        p = "res_stack/%d" % i
        for j in (1, 2, 3):
            k = w["%s/conv%d/kernel" % (p, j)]
            kh, kw, ci, co = k.shape
            kk, bb = _fold(k.reshape(kh * kw, ci, co), w["%s/conv%d/bias" % (p, j)], w, "%s/bnorm%d" % (p, j))
            out.append(("res%d.conv%d" % (i, j), kh, kw, 0, kk, bb))
        nl = p + "/non_local"
        qkv = np.concatenate([w[nl + "/theta/kernel"], w[nl + "/phi/kernel"], w[nl + "/g/kernel"]], axis=3)
        qb = np.concatenate([w[nl + "/theta/bias"], w[nl + "/phi/bias"], w[nl + "/g/bias"]])
        kk, bb = _fold(qkv.reshape(1, qkv.shape[2], qkv.shape[3]), qb, w, None)
        out.append(("res%d.qkv" % i, 1, 1, 0, kk, bb))
        k = w[nl + "/w/kernel"]
        kk, bb = _fold(k.reshape(1, k.shape[2], k.shape[3]), w[nl + "/w/bias"], w, nl + "/bnorm")
        out.append(("res%d.w" % i, 1, 1, 0, kk, bb))
    convt("up1", "up1"); convt("up2", "up2"); convt("up3", "up3")
    hk = np.concatenate([w["conv2/conv/kernel"], w["conv3/conv/kernel"]], axis=3)       # [7,7,64,2]
    hb = np.concatenate([w["conv2/conv/bias"], w["conv3/conv/bias"]])
    kk, bb = _fold(hk.reshape(49, 64, 2), hb, w, None)
    out.append(("heads", 7, 7, 0, kk, bb))
    convt("clr_up1", "clr_up1"); convt("clr_up2", "clr_up2"); convt("clr_up3", "clr_up3")
    k = w["clr_conv1/conv/kernel"]                                                       # [3,3,65,16], in = [gs, f]
    k = np.concatenate([k[:, :, 1:, :], k[:, :, 0:1, :]], axis=2)                        # -> [f, gs]
    kk, bb = _fold(k.reshape(9, 65, 16), w["clr_conv1/conv/bias"], w, "clr_conv1/bnorm")
    out.append(("clr_conv1", 3, 3, 0, kk, bb))
    conv("clr_conv2", "clr_conv2"); conv("clr_conv3", "clr_conv3", bn=False)
    return out


def build_blob(variant: str, w: Dict[str, np.ndarray]) -> bytes:
    layers = canonical_layers(variant, w)
    entry = struct.Struct("<32s6i2Q")
    header = 16 + entry.size * len(layers)
    payload = bytearray()
    table = bytearray()
    for name, kh, kw, tr, kk, bb in layers:
        w_off = header + len(payload)
        payload += kk.tobytes()
        b_off = header + len(payload)
        payload += bb.tobytes()
        table += entry.pack(name.encode(), kh, kw, kk.shape[1], kk.shape[2], tr, 0, w_off, b_off)
    return MAGIC + struct.pack("<ii", VARIANTS.index(variant), len(layers)) + bytes(table) + bytes(payload)


def convert_checkpoint(index_path: str, variant: str) -> bytes:
    """TF checkpoint (index + data shard) -> blob, no TensorFlow needed."""
    from .tf_checkpoint import read_generator_weights
    return build_blob(variant, read_generator_weights(index_path))
