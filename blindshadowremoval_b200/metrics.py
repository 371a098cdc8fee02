"""Evaluation helpers used next to the forward by the reference's test steps."""
from __future__ import annotations

import numpy as np


def sfw_auc(label: np.ndarray, mask_pred: np.ndarray) -> float:
    """Shadow-segmentation AUC exactly as train_test_GSC.py:820-832 / train_with_TSM.py:689-701:
    both vectors get the sentinels [1, 0] prepended so the score is defined for one-class labels."""
    from sklearn import metrics
    y = np.concatenate([np.array([1.0, 0.0]), np.asarray(label, np.float64).reshape(-1)])
    s = np.concatenate([np.array([1.0, 0.0]), np.asarray(mask_pred, np.float64).reshape(-1)])
    return float(metrics.roc_auc_score(y, s))


def psnr(a: np.ndarray, b: np.ndarray, max_val: float = 1.0) -> float:
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return float("inf") if mse == 0 else 10.0 * np.log10(max_val * max_val / mse)


def ssim(a: np.ndarray, b: np.ndarray, max_val: float = 1.0) -> float:
    """``tf.image.ssim(a, b, max_val)`` for one [H,W,C] pair (defaults: 11x11 Gaussian window, sigma 1.5, k1 = 0.01,
    k2 = 0.03, VALID convolution; mean over positions, then over channels), as used at train_with_TSM.py:686 and
    train_test_GSC.py:817."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.ndim == 2:
        a, b = a[..., None], b[..., None]
    g = np.exp(-((np.arange(11) - 5.0) ** 2) / (2.0 * 1.5 ** 2))
    g /= g.sum()

    def blur(x):                                                  # separable VALID filter over H and W
        x = sum(g[i] * x[i:x.shape[0] - 10 + i] for i in range(11))
        return sum(g[i] * x[:, i:x.shape[1] - 10 + i] for i in range(11))

    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    mu_a, mu_b = blur(a), blur(b)
    var_a, var_b, cov = blur(a * a) - mu_a * mu_a, blur(b * b) - mu_b * mu_b, blur(a * b) - mu_a * mu_b
    lum = (2.0 * mu_a * mu_b + c1) / (mu_a * mu_a + mu_b * mu_b + c1)
    cs = (2.0 * cov + c2) / (var_a + var_b + c2)
    return float(np.mean(np.mean(lum * cs, axis=(0, 1))))
