"""BASELINE config 5: 2048-image synthetic stress run, 16-bit tensor-core path vs fp32 check mode.

2048 distinct seeded crops in batches of 128 through both precisions of the SAME library (the fp32 check mode is
itself pinned to the oracle at 1e-4 by tests/test_gpu_parity.py).  Reports, over all 2048 images: max-abs / mean-abs /
PSNR of the [0,1]-clipped colour output and of dif, hole-mask cells that differ between the two modes, share of pixels
within 1e-2; for the last batch: error growth through the named intermediates; and the per-layer device times of the
bf16 path.  Usage (GPU box):  python tools/stress_2048.py [n_images] > profiles/r1c_stress_2048.txt
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BSR_DEBUG_KEEP"] = "1"
os.environ["BSR_PROFILE"] = "1"
from blindshadowremoval_b200.generator import Generator  # noqa: E402
from blindshadowremoval_b200.synthetic import make_inputs  # noqa: E402
from blindshadowremoval_b200.weights import random_weights  # noqa: E402

NAMES = ["x1", "x2", "x3", "x_in0", "res0", "res1", "res2", "up1", "up2", "up3", "x_in3", "res3", "res4", "res5",
         "clr_up1", "clr_up2", "clr_up3"]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
MB = 128
w = random_weights("gsc", 1234)
bf = Generator("gsc", "tc16", device=0, micro_batch=MB, weights=w)
fp = Generator("gsc", "fp32check", device=0, micro_batch=MB, weights=w)
acc = dict(n=0, max_rgb=0.0, sum_rgb=0.0, se_rgb=0.0, max_dif=0.0, sum_dif=0.0, se_dif=0.0, within=0, px=0, flips=0,
           cells=0, max_rgb_same_mask=0.0)
t_bf = t_fp = 0.0
for b in range(N // MB):
    d = make_inputs(MB, seed=1000 + b)
    img, uv = torch.from_numpy(d["img"]).cuda(), torch.from_numpy(d["uv"]).cuda()
    torch.cuda.synchronize()
    t0 = time.time()
    _, rgb_b, _, dif_b = bf(img, uv, None, want=("con_rgb", "dif"))
    torch.cuda.synchronize()
    t1 = time.time()
    _, rgb_f, _, dif_f = fp(img, uv, None, want=("con_rgb", "dif"))
    torch.cuda.synchronize()
    t2 = time.time()
    t_bf += t1 - t0
    t_fp += t2 - t1
    bm_b = torch.from_numpy(bf.debug_read("bmask")).reshape(MB, 32, 32)
    bm_f = torch.from_numpy(fp.debug_read("bmask")).reshape(MB, 32, 32)
    e_rgb = (rgb_b.clamp(0, 1) - rgb_f.clamp(0, 1)).abs()
    e_dif = (dif_b - dif_f).abs()
    acc["n"] += MB
    acc["max_rgb"] = max(acc["max_rgb"], float(e_rgb.max()))
    acc["sum_rgb"] += float(e_rgb.sum())
    acc["se_rgb"] += float((e_rgb ** 2).sum())
    acc["max_dif"] = max(acc["max_dif"], float(e_dif.max()))
    acc["sum_dif"] += float(e_dif.sum())
    acc["se_dif"] += float((e_dif ** 2).sum())
    acc["within"] += int((e_rgb <= 1e-2).sum())
    acc["px"] += e_rgb.numel()
    flipped = bm_b != bm_f
    acc["flips"] += int(flipped.sum())
    acc["cells"] += bm_b.numel()
    same = ~flipped.reshape(MB, -1).any(dim=1)                 # images whose hole mask agrees in every cell
    if same.any():
        acc["max_rgb_same_mask"] = max(acc["max_rgb_same_mask"], float(e_rgb[same.cuda()].max()))
    if b == N // MB - 1:
        print("error growth through the network (last batch of %d images; bf16 vs fp32 check mode)" % MB)
        for name in NAMES:
            a, r = bf.debug_read(name), fp.debug_read(name)
            err = np.abs(a - r)
            print("  %-8s max|err| %.3e  mean|err| %.3e  rms(ref) %.3f  rel %.2e" %
                  (name, err.max(), err.mean(), np.sqrt((r ** 2).mean()), err.mean() / max(np.abs(r).mean(), 1e-12)))
        print("per-layer device time of the bf16 path, %d images per launch (CUDA events, BSR_PROFILE=1)" % MB)
        for name, ms in bf.layer_times():
            print("  %-14s %.4f ms" % (name, ms))
px = acc["px"]
print("=" * 100)
print("images %d (seeds 1000..%d, %d per batch), weights seed 1234, errflags bf16 %d fp32 %d" %
      (acc["n"], 1000 + N // MB - 1, MB, bf.debug_read("errflag")[0], fp.debug_read("errflag")[0]))
print("clip(con_rgb): max|bf16 - fp32| %.3e   mean %.3e   PSNR %.2f dB   pixels within 1e-2: %.4f %%" %
      (acc["max_rgb"], acc["sum_rgb"] / px, 10 * np.log10(1.0 / max(acc["se_rgb"] / px, 1e-20)), 100.0 * acc["within"] / px))
print("clip(con_rgb), images whose hole mask agrees in all cells: max|err| %.3e" % acc["max_rgb_same_mask"])
print("dif:           max|bf16 - fp32| %.3e   mean %.3e   PSNR %.2f dB" %
      (acc["max_dif"], acc["sum_dif"] / (px / 3), 10 * np.log10(1.0 / max(acc["se_dif"] / (px / 3), 1e-20))))
print("hole-mask cells that differ between the modes: %d of %d (%.4f %%)" %
      (acc["flips"], acc["cells"], 100.0 * acc["flips"] / acc["cells"]))
print("wall time incl. first-call setup: bf16 %.2f s (%.0f img/s), fp32 check %.2f s (%.0f img/s)" %
      (t_bf, acc["n"] / t_bf, t_fp, acc["n"] / t_fp))
