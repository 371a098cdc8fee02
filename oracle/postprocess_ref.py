"""ORACLE (test infrastructure, not product code): CPU restatement of the post-processing of ``FSRNet.test_step``
(/root/reference/train_test_GSC.py:411-748), the block that runs right after the generator call of BASELINE config 2
(UCB evaluation): resize everything to the crop size and pad, region heuristics on the predicted shadow mask
(:479-580), connected-component filter (:590-611), nose rule (:650-662), composite + clip (:711-718), SSIM / PSNR
against the ground truth (:724-725).

Parity pinning: the reference's block mixes TensorFlow ops (``tf.image.resize``, ``tf.image.ssim``) with NumPy / cv2, so it
cannot run here as a whole - **parity unpinned** for the TF ops; the one third-party routine with observable behaviour,
``cv2.connectedComponentsWithStats(..., connectivity=4)``, IS available and pins ``components`` below
(tests/test_postprocess.py).  ``resize`` is the restatement already used by the generator oracle (half-pixel bilinear,
no antialias); ``ssim`` / ``psnr`` are blindshadowremoval_b200.metrics (restated ``tf.image.ssim`` / ``psnr``).

Only tests/ may import this module.  Defined deviations from the reference, both for inputs on which the reference
raises: an empty region mask skips the rule that needs its bounding box (``np.min`` of an empty array, :481-488), and a
prediction without any connected component yields an empty detected mask (``np.max(sizes)`` of an empty array, :600).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
from scipy import ndimage

from oracle.generator_ref import resize_bilinear

IMG = 256
MASK_KINDS = ("face_hair", "face", "mouth", "nose", "eyebrow", "eye", "glasses")


def _resize_pad(x: np.ndarray, size: int, rnd: bool = False) -> np.ndarray:
    """``tf.pad(tf.image.resize(x, [size, size]), [[0, 256-size], [0, 256-size], [0, 0]])`` (:438-476), optionally with
    the ``tf.round`` the region masks get (half to even)."""
    t = torch.from_numpy(np.asarray(x, np.float32))[None]
    y = resize_bilinear(t, size, size)[0].numpy()
    if rnd:
        y = np.rint(y)
    out = np.zeros((IMG, IMG, x.shape[2]), np.float32)
    out[:size, :size] = y
    return out


def _bbox(mask2d: np.ndarray):
    rows, cols = np.where(mask2d == 1)
    if rows.size == 0:
        return None
    return int(rows.min()), int(rows.max()), int(cols.min()), int(cols.max())


def components(binary: np.ndarray):
    """4-connected components of a {0,1} image: (labels int32 with 0 = background, sizes[1..n]).  Same partition and
    sizes as ``cv2.connectedComponentsWithStats(binary, connectivity=4)`` (:590); label numbering may differ."""
    lab, n = ndimage.label(binary != 0)              # default structure = 4-connectivity
    sizes = np.bincount(lab.ravel(), minlength=n + 1)[1:]
    return lab.astype(np.int32), sizes


def test_step_postprocess(img0, gt0, con_rgb0, dif0, size: int, masks: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """img0, gt0, con_rgb0 [256,256,3], dif0 [256,256,1] = frame 0 of the chunk and of the generator outputs
    (``deshadow_img_c[0]``, ``mask_pred[0]`` = 4th output, :422-434); ``size = box[3] - box[1]`` (:417); masks[kind]
    [256,256,3] in {0,1} (the PNGs / 255, :387-393).  Returns the intermediates and results of the block."""
    m = {k: _resize_pad(masks[k], size, rnd=True) for k in MASK_KINDS}
    gt_sc = _resize_pad(gt0, size)
    pred = _resize_pad(con_rgb0, size)
    tmp = _resize_pad(img0, size)
    mask_pred = _resize_pad(dif0, size) * m["face_hair"]                                       # :473-476, 3 channels
    nose, mouth = _bbox(m["nose"][:, :, 0]), _bbox(m["mouth"][:, :, 0])
    # ---- mustache / mouth false positives (:479-496)
    if nose is not None and mouth is not None:
        mid_nose_h = (nose[1] + nose[0]) / 2.0
        upper_mouth, lower_mouth, left_mouth, right_mouth = mouth
        reg = np.zeros((IMG, IMG, 3), bool)
        reg[int(mid_nose_h):int(upper_mouth), int(left_mouth):int(right_mouth)] = True
        mask_pred = mask_pred * (~((mask_pred < 0.018) & reg)).astype(np.float32)
        reg = np.zeros((IMG, IMG, 3), bool)
        reg[int(upper_mouth):int(lower_mouth), int(left_mouth):int(right_mouth)] = True
        mask_pred = mask_pred * (~((mask_pred < 0.02) & reg)).astype(np.float32)
    hair = m["face_hair"] - m["face"]                                                           # :498
    inten = np.repeat(tmp.mean(axis=2, keepdims=True), 3, axis=2)                               # :520-521
    # ---- per-pixel threshold (:518-577)
    thr = np.full((IMG, IMG, 3), 0.01, np.float64)
    thr[hair > 0] = 0.02
    thr[(hair > 0) & (inten < 0.13)] = 0.004
    brow = _bbox(m["eyebrow"][:, :, 0])
    if m["eyebrow"].sum() > 30 and brow is not None:                                            # forehead (:528-538)
        fm = m["face"].copy()
        fm[brow[0]:IMG] = 0
        fb = _bbox(fm[:, :, 0])
        if fb is not None:
            rect = np.zeros((IMG, IMG, 3), bool)
            rect[int(fb[0] + 20):int(brow[0] - 40), int(fb[2] + 40):int(fb[3] - 40)] = True
            thr[rect & (inten < 0.4)] = -0.001
    if mouth is not None:                                                                       # mouth and below (:541-556)
        below = np.zeros((IMG, IMG, 3), np.float32)
        below[int(mouth[0]):IMG] = 1.0
        roi = below * m["face"]
        shadowed = (mask_pred > 0.01).astype(np.float32)
        frac = float((shadowed * roi).sum() / roi.sum())
        if 0.252 < frac < 0.268:
            thr[roi > 0] = 1.0
        mab = roi * tmp * shadowed
        mean_mab = float(mab.mean(axis=2).sum() / (roi[:, :, 0] * shadowed[:, :, 0]).sum())
        if 0.3 < frac < 0.31 and mean_mab > 0.358:
            thr[roi > 0] = 1.0
        if 0.295 < frac < 0.3 and mean_mab > 0.22:
            thr[roi > 0] = 1.0
    face_bb = _bbox(m["face"][:, :, 0])
    if m["eyebrow"].sum() > 0 and brow is not None and face_bb is not None:                      # left eyebrow (:557-570)
        mid_face = face_bb[2] * 0.8 + face_bb[3] * 0.2
        if brow[2] - face_bb[2] == 0:
            left = np.zeros((IMG, IMG, 3), np.float32)
            left[:, 0:int(mid_face)] = 1.0
            thr[((m["eyebrow"] * left) > 0) & (inten > 0.1)] = 1.0
    detected = (mask_pred.astype(np.float32) > thr.astype(np.float32)).astype(np.uint8)         # :577
    # ---- connected components: keep the big ones that are not mostly hair (:590-611)
    lab, sizes = components(detected[:, :, 0])
    img2 = np.zeros((IMG, IMG), np.float64)
    if sizes.size:
        min_size = 0.45 * sizes.max()
        hair0 = hair[:, :, 0]
        for i, sz in enumerate(sizes):
            comp = lab == i + 1
            if sz >= min_size and hair0[comp].sum() / sz < 0.8:
                img2[comp] = 1
    # ---- nose rule (:650-662)
    shadow_image = img2 * tmp.mean(axis=2)
    if nose is not None and img2.sum() > 0:
        mean_intensity = shadow_image.sum() / img2.sum()
        frac_nose = float(((m["nose"][:, :, 0] * shadow_image) > 0).sum() / m["nose"][:, :, 0].sum())
        mid_h, low, mid_w = (nose[1] + nose[0]) / 2.0, nose[1], (nose[3] + nose[2]) / 2.0
        if 0.15 < frac_nose < 0.25 or 0.30 < frac_nose < 0.31 or 0.34 < frac_nose < 0.35:
            hi = low + 5 if mean_intensity < 0.15 else low + 65
            img2[int(mid_h):int(hi), int(mid_w - 35):int(mid_w + 35)] = 0
    det3 = np.repeat(img2[:, :, None], 3, axis=2).astype(np.float32)
    final = np.clip(pred * det3 + tmp * (1.0 - det3), 0.0, 1.0).astype(np.float32)              # :711, 718
    from blindshadowremoval_b200.metrics import psnr, ssim
    return {"final": final, "detected": img2.astype(np.float32), "mask_pred": mask_pred, "threshold": thr.astype(np.float32),
            "gt_sc": gt_sc, "tmp": tmp, "pred": pred, "ssim": ssim(gt_sc, final, 1.0), "psnr": psnr(gt_sc, final, 1.0),
            "masks": m}
