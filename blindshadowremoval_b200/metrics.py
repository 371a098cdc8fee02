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
