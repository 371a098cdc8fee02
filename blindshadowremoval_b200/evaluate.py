"""Sharded evaluation loops around the forward, with the units dealt to ranks and ONE all-reduce at the end:
``evaluate_sfw`` = the reference's ``testsfw`` / ``test_step_sfw`` (/root/reference/train_with_TSM.py:619-639, 668-701),
``evaluate_ucb`` = ``test`` / ``test_step`` (/root/reference/train_test_GSC.py:360-408, 411-748; BASELINE config 2).

Per chunk ``[frame,256,256,17]`` (img | cmap | mask | uv | reg | face): ``rgb, mask_pred = gen.forward_chunk(chunk)``,
then on frame 0 (``masksc = mask[0]``, ``mask_predsc = mask_pred[0]``, :683-684): SSIM and PSNR of the predicted mask
against the label map (:686-687) and the AUC of ``mask == 2`` with the ``[1, 0]`` sentinels (:689-701).  ``Logging``
keeps running means over the steps (utils.py:136-171); here every rank sums its own chunks and
``sharding.reduce_metrics`` produces the same global means on all ranks (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np

from . import sharding
from .metrics import psnr, sfw_auc, ssim


def evaluate_sfw(gen, get_chunk: Callable[[int], np.ndarray], n_chunks: int, frame: int = 2, rank: int = 0, world: int = 1,
                 device=None, on_result: Optional[Callable] = None, reduce: bool = True) -> Dict[str, float]:
    """Run chunks ``[b, e)`` of this rank through ``gen.forward_chunk`` and return the global metric means.

    ``get_chunk(i)`` returns chunk ``i`` as a NumPy array or CUDA tensor ``[frame,256,256,17]``; ``on_result(i, rgb,
    mask_pred)`` (optional) receives every result, e.g. to write the figures ``Logging.save_img`` writes."""
    b, e = sharding.shard_units(n_chunks, rank, world)
    sums = {"ssim": 0.0, "psnr": 0.0, "auc": 0.0}
    for i in range(b, e):
        chunk = get_chunk(i)
        rgb, mask_pred = gen.forward_chunk(chunk, frame=frame, share=True)
        to_np = lambda t: t if isinstance(t, np.ndarray) else t.detach().cpu().numpy()
        c0 = to_np(chunk[0:1])[0]
        mask0 = c0[..., 6:7]                                      # tf.split(img, [3,3,1,3,6,1], 3)[2][0]
        pred0 = to_np(mask_pred[0:1])[0]
        sums["ssim"] += ssim(mask0, pred0, 1.0)
        sums["psnr"] += psnr(mask0, pred0, 1.0)
        sums["auc"] += sfw_auc((mask0 == 2).astype(np.float32), pred0)
        if on_result is not None:
            on_result(i, rgb, mask_pred)
    if not reduce:          # this process's own means, no collective (used to cross-check the sharded result)
        return {k: v / max(e - b, 1) for k, v in sums.items()} | {"count": float(e - b)}
    return sharding.reduce_metrics(sums, e - b, device=device)


def evaluate_ucb(gen, get_sample: Callable[[int], dict], n_samples: int, batch: int = 32, rank: int = 0, world: int = 1,
                 device=None, on_result: Optional[Callable] = None) -> Dict[str, float]:
    """``FSRNet.test`` (train_test_GSC.py:360-408) on this rank's share of ``n_samples`` UCB samples, ``batch`` samples per
    generator call (BASELINE config 2: batch 32 + ragged tail).  ``get_sample(i)`` returns a dict with img, gt, uv
    [256,256,3] (frame 0 of the reference's 10-frame chunk: the other nine frames never reach its outputs, model.py has no
    cross-sample op), ``size`` = box[3] - box[1] and ``masks`` uint8 [7,256,256] in ``Generator.MASK_KINDS`` order.
    Per sample: generator -> ``deshadow_img_c[0]``, ``mask_pred[0]`` (4th output) -> device post-processing
    (``bsr_postprocess_ucb``) -> SSIM / PSNR against the ground truth (:724-725); returns the global means
    (``Logging``'s running means, utils.py:136-171).  ``on_result(i, final, detected)`` receives every result."""
    import torch
    b, e = sharding.shard_units(n_samples, rank, world)
    sums = {"ssim": 0.0, "psnr": 0.0}
    dev = torch.device("cuda", gen.device)
    for i0 in range(b, e, batch):
        idx = list(range(i0, min(i0 + batch, e)))
        samples = [get_sample(i) for i in idx]
        t = lambda k, dt=torch.float32: torch.from_numpy(np.stack([np.asarray(s[k]) for s in samples])).to(dt).to(dev)
        img, gt, uv = t("img"), t("gt"), t("uv")
        _, rgb, _, dif = gen(img, uv, None, chuck=4, training=False, want=("con_rgb", "dif"))
        sizes = torch.tensor([int(s["size"]) for s in samples], dtype=torch.int32, device=dev)
        final, det, met = gen.postprocess_ucb(img, gt, rgb, dif, sizes, t("masks", torch.uint8))
        m = met.double().cpu().numpy()
        sums["ssim"] += float(m[:, 0].sum())
        sums["psnr"] += float(m[:, 1].sum())
        if on_result is not None:
            for k, i in enumerate(idx):
                on_result(i, final[k], det[k])
    return sharding.reduce_metrics(sums, e - b, device=device)
