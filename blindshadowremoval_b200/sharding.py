"""Multi-GPU plumbing: the path shards with NO data-path collective.

GSC images are independent (model.py has no cross-sample op at inference); the TSM unit is a chunk
of ``frame`` images (model_with_TSM.py:218-224).  Units are dealt to ranks in contiguous, balanced
blocks; weights are replicated.  The only collective is one all-reduce (sum) of the evaluation
accumulators, reproducing ``Logging.update``'s running means (/root/reference/utils.py:136-171).
"""
from __future__ import annotations

from typing import Dict, Tuple


def shard_units(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the units owned by ``rank``; sizes differ by at most one, empty if n < world."""
    if world <= 0 or not 0 <= rank < world or n_units < 0:
        raise ValueError("bad shard request n=%d rank=%d world=%d" % (n_units, rank, world))
    base, rem = divmod(n_units, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_images(n_images: int, frame: int, rank: int, world: int) -> Tuple[int, int]:
    """Image range for ``rank`` when whole chunks of ``frame`` images must stay together."""
    if frame <= 0 or n_images % frame:
        raise ValueError("n_images %d is not a multiple of frame %d" % (n_images, frame))
    b, e = shard_units(n_images // frame, rank, world)
    return b * frame, e * frame


def reduce_metrics(sums: Dict[str, float], count: int, device=None) -> Dict[str, float]:
    """All-reduce {metric: sum} and count over the default process group; returns the global means.

    Works without an initialised process group (single process).  NCCL when the tensors are CUDA
    (NVLink/NVSwitch on the B200 box), gloo on CPU.
    """
    import torch
    import torch.distributed as dist
    keys = sorted(sums)
    t = torch.tensor([float(sums[k]) for k in keys] + [float(count)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    total = t[-1].item()
    return {k: (t[i].item() / total if total > 0 else float("nan")) for i, k in enumerate(keys)} | {"count": total}
