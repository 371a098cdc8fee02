"""Variable inventory and seeded random initialisation of the shadow-removal generator.

Names and shapes are the object-graph paths that ``tf.train.Checkpoint(generator=...)`` writes
(/root/reference/train_test_GSC.py:143-148), i.e. what ``log/*/ckpt-*.index`` holds: Conv2D kernels
are ``[kh, kw, in, out]``, Conv2DTranspose kernels ``[kh, kw, out, in]``
(/root/reference/model.py:198-226 builds the layers; model_with_TSM.py:231-259 the TSM variant).

No trained weights ship with the reference (/root/reference/.MISSING_LARGE_BLOBS), so parity and
benchmarks use ``random_weights``: a seeded, *non-trivial* init (BN statistics away from 0/1 so a
folding bug cannot hide, attention projections scaled so the unscaled softmax logits stay in a range
where the comparison is well conditioned).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

VARIANTS = ("gsc", "tsm")
N_RES = 6
RES_CH = 257            # n_ch[5] + 1, model.py:226
BN_EPS = 1e-3           # Keras BatchNormalization default
LEAKY_ALPHA = 0.3       # Keras LeakyReLU default


def res_in_channels(variant: str):
    """Input channel count of each ResBottleneck (model.py:238,259; model_with_TSM.py:272,294)."""
    if variant == "gsc":
        return [99, 257, 257, 261, 261, 261]
    if variant == "tsm":
        return [291, 291, 291, 877, 877, 877]
    raise ValueError("variant must be one of %r" % (VARIANTS,))


def variable_shapes(variant: str) -> "OrderedDict[str, Tuple[int, ...]]":
    """Ordered {name: shape} of every generator variable for ``variant``."""
    rin = res_in_channels(variant)
    wide1 = max(rin[0], RES_CH)      # channels leaving res block 2 (feeds up1)
    wide2 = max(rin[3], RES_CH)      # channels leaving res block 5 (feeds clr_up1)
    out: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def bn(prefix, c):
        for n in ("gamma", "beta", "moving_mean", "moving_variance"):
            out["%s/%s" % (prefix, n)] = (c,)

    def conv(name, k, cin, cout, norm=True):
        out["%s/conv/kernel" % name] = (k, k, cin, cout)
        out["%s/conv/bias" % name] = (cout,)
        if norm:
            bn("%s/bnorm" % name, cout)

    def convt(name, cin, cout):
        out["%s/conv/kernel" % name] = (3, 3, cout, cin)
        out["%s/conv/bias" % name] = (cout,)
        bn("%s/bnorm" % name, cout)

    conv("conv1", 7, 3, 32)
    conv("down1", 3, 32, 64)
    conv("down2", 3, 64, 64)
    conv("down3", 3, 64, 96)
    for i in range(N_RES):
        p = "res_stack/%d" % i
        out[p + "/conv1/kernel"] = (1, 1, rin[i], 128)
        out[p + "/conv1/bias"] = (128,)
        bn(p + "/bnorm1", 128)
        out[p + "/conv2/kernel"] = (3, 3, 128, 128)
        out[p + "/conv2/bias"] = (128,)
        bn(p + "/bnorm2", 128)
        out[p + "/conv3/kernel"] = (1, 1, 128, RES_CH)
        out[p + "/conv3/bias"] = (RES_CH,)
        bn(p + "/bnorm3", RES_CH)
        for n in ("g", "phi", "theta"):
            out["%s/non_local/%s/kernel" % (p, n)] = (1, 1, RES_CH, 128)
            out["%s/non_local/%s/bias" % (p, n)] = (128,)
        out[p + "/non_local/w/kernel"] = (1, 1, 128, RES_CH)
        out[p + "/non_local/w/bias"] = (RES_CH,)
        bn(p + "/non_local/bnorm", RES_CH)
    convt("up1", wide1, 96)
    convt("up2", 160, 64)
    convt("up3", 128, 64)
    conv("conv2", 7, 64, 1, norm=False)
    conv("conv3", 7, 64, 1, norm=False)
    convt("clr_up1", wide2, 128)
    convt("clr_up2", 128, 96)
    convt("clr_up3", 96, 64)
    conv("clr_conv1", 3, 65, 16)
    conv("clr_conv2", 1, 16, 16)
    conv("clr_conv3", 1, 16, 3, norm=False)
    return out


def random_weights(variant: str, seed: int = 1234) -> Dict[str, np.ndarray]:
    """Seeded fp32 weights of the identical architecture.

    Kernels: He-style normal with fan-in gain (keeps activations O(1) through LeakyReLU stacks);
    biases uniform +-0.1; BN gamma in [0.5,1.5], beta/mean in [-0.2,0.2], variance in [0.5,1.5].
    theta/phi are scaled down so logits theta.phi (no 1/sqrt(d) in the reference, model.py:51-52) have
    a standard deviation of a few units; conv2/conv3 (mask / con heads) are scaled so that ``dif``
    straddles the 0.1 hole threshold (model.py:256) on a useful fraction of cells.
    """
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape in variable_shapes(variant).items():
        leaf = name.rsplit("/", 1)[1]
        if leaf == "kernel":
            kh, kw, a, b = shape
            is_t = name.split("/")[0] in ("up1", "up2", "up3", "clr_up1", "clr_up2", "clr_up3")
            cin = b if is_t else a
            # a stride-2 transposed conv touches on average 9/4 taps per output pixel
            fan_in = cin * (kh * kw / 4.0 if is_t else kh * kw)
            std = np.sqrt(1.4 / fan_in)
            if "/non_local/theta/" in name or "/non_local/phi/" in name:
                std *= 1.1
            if "/non_local/w/" in name:
                std *= 0.35
            if name.startswith("res_stack/") and "/conv3/" in name:
                std *= 0.5
            if name.startswith("conv2/") or name.startswith("conv3/"):
                std *= 0.6
            if name.startswith("clr_conv3/"):
                std *= 0.25
            w = rng.standard_normal(shape) * std
            if name.startswith("conv2/") or name.startswith("conv3/"):
                # zero response to a per-channel constant: keeps ``dif`` centred on the conv3 bias
                w = w - w.mean(axis=(0, 1), keepdims=True)
        elif leaf == "bias":
            w = rng.uniform(-0.1, 0.1, shape)
            if name.startswith("clr_conv3/"):
                w = w + 0.45
            if name.startswith("conv3/"):
                w = w * 0 + 0.1          # E[dif] ~ hole threshold (model.py:256): both bmask values occur
        elif leaf == "gamma":
            w = rng.uniform(0.5, 1.5, shape)
        elif leaf in ("beta", "moving_mean"):
            w = rng.uniform(-0.2, 0.2, shape)
        elif leaf == "moving_variance":
            w = rng.uniform(0.5, 1.5, shape)
        else:
            raise AssertionError(name)
        out[name] = np.ascontiguousarray(w, dtype=np.float32)
    return out


def check_weights(variant: str, weights: Dict[str, np.ndarray]) -> None:
    """Raise ValueError unless ``weights`` has exactly the variables of ``variant``."""
    spec = variable_shapes(variant)
    missing = [k for k in spec if k not in weights]
    extra = [k for k in weights if k not in spec]
    if missing or extra:
        raise ValueError("weight set mismatch: missing %s, unexpected %s" % (missing[:4], extra[:4]))
    for k, shp in spec.items():
        if tuple(weights[k].shape) != tuple(shp):
            raise ValueError("%s: shape %s, expected %s" % (k, tuple(weights[k].shape), shp))
