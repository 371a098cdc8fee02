"""Sharded evaluation loop around the forward: the reference's ``testsfw`` / ``test_step_sfw``
(/root/reference/train_with_TSM.py:619-639, 668-701) with the chunks dealt to ranks and ONE all-reduce at the end.

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
                 device=None, on_result: Optional[Callable] = None) -> Dict[str, float]:
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
    return sharding.reduce_metrics(sums, e - b, device=device)
