"""Seeded synthetic inputs of the reference's crop shape (256x256 face crops).

Shapes/ranges follow what ``dataset.py:parse_fn_test`` / ``parse_fn_test_FFHQ`` hand to the model
(/root/reference/dataset.py:616-770): ``img`` = RGB/255 in [0,1]; ``uv`` = canonical-UV map that is 0
outside the landmark hull (warp.py:215-232); ``reg`` = [reg_in(3) | reg_out(3)] landmark-registration
offset fields in normalised units with a zero third channel (warp.py:194-213); ``face`` = {0,1} mask.
SFW/UCB data is not needed: these are smooth random fields with the same statistics class.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

IMG = 256


def _lowpass(gen: torch.Generator, n: int, c: int, coarse: int) -> torch.Tensor:
    z = torch.randn(n, c, coarse, coarse, generator=gen)
    return F.interpolate(z, size=(IMG, IMG), mode="bicubic", align_corners=False)


def make_inputs(n: int, seed: int = 0, with_reg: bool = False):
    """Return dict of NHWC float32 arrays: img[n,256,256,3], uv[...,3], face[...,1] (+ reg[...,6])."""
    gen = torch.Generator().manual_seed(seed)
    img = (0.5 + 0.25 * _lowpass(gen, n, 3, 16) + 0.03 * torch.randn(n, 3, IMG, IMG, generator=gen)).clamp(0, 1)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, IMG), torch.linspace(-1, 1, IMG), indexing="ij")
    cx = 0.1 * (torch.rand(n, 1, 1, generator=gen) - 0.5)
    cy = 0.1 * (torch.rand(n, 1, 1, generator=gen) - 0.5)
    ax = 0.55 + 0.1 * torch.rand(n, 1, 1, generator=gen)
    ay = 0.70 + 0.1 * torch.rand(n, 1, 1, generator=gen)
    face = ((((xx - cx) / ax) ** 2 + ((yy - cy) / ay) ** 2) <= 1.0).float()[:, None]
    ramp = torch.stack([(yy + 1) / 2, (xx + 1) / 2, 0.5 + 0.5 * torch.cos(3.0 * xx) * torch.cos(2.0 * yy)])[None]
    uv = (ramp + 0.05 * _lowpass(gen, n, 3, 8)) * face
    out = {
        "img": img.permute(0, 2, 3, 1).contiguous().numpy().astype(np.float32),
        "uv": uv.permute(0, 2, 3, 1).contiguous().numpy().astype(np.float32),
        "face": face.permute(0, 2, 3, 1).contiguous().numpy().astype(np.float32),
    }
    if with_reg:
        r = 0.035 * _lowpass(gen, n, 4, 6).clamp(-2.8, 2.8)        # |offset| <= ~0.1 (normalised units)
        z = torch.zeros(n, 1, IMG, IMG)
        reg = torch.cat([r[:, 0:2], z, r[:, 2:4], z], dim=1)
        out["reg"] = reg.permute(0, 2, 3, 1).contiguous().numpy().astype(np.float32)
    return out
