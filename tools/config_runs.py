"""BASELINE configs 1 and 2 on the reference's real files (tests/fixtures): timings + errors for BASELINE.md section 4.
Run on the GPU box:  python tools/config_runs.py > profiles/r2_config_runs.txt"""
import glob
import os
import sys
import time

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BSR_DEBUG_KEEP"] = "1"
from blindshadowremoval_b200 import feed  # noqa: E402
from blindshadowremoval_b200.evaluate import evaluate_ucb  # noqa: E402
from blindshadowremoval_b200.generator import Generator, act_dtype  # noqa: E402
from blindshadowremoval_b200.metrics import psnr  # noqa: E402
from blindshadowremoval_b200.weights import random_weights  # noqa: E402
from oracle.generator_ref import caller_glue, generator_forward  # noqa: E402

FIX = os.path.join(ROOT, "tests", "fixtures")
w = random_weights("gsc", 1234)
print("storage type of the 16-bit path:", act_dtype(), "| weights: random init seed 1234 (no trained checkpoint ships)")

# ---- config 1: sample_imgs/02165, batch 1
t0 = time.time()
f = feed.load_frame(os.path.join(FIX, "sample_imgs", "02165", "02165.png"))
t_feed = time.time() - t0
chunk = feed.build_chunk([f])
gen = Generator("gsc", "tc16", device=0, micro_batch=1, weights=w)
rgb, mp, gs, m22 = gen.forward_chunk(chunk, want_raw=True)
bm = gen.debug_read("bmask").reshape(1, 32, 32, 1)
ref = generator_forward(w, f["img"][None], f["uv"][None], variant="gsc", bmask_override=bm)
r_rgb, r_mp = caller_glue(ref["con_rgb"], ref["dif"], f["face"][None])
torch.set_num_threads(os.cpu_count() or 1)
t0 = time.time()
for _ in range(3):
    generator_forward(w, f["img"][None], f["uv"][None], variant="gsc")
t_cpu = (time.time() - t0) / 3
gen.close()
os.environ.pop("BSR_DEBUG_KEEP")
gen = Generator("gsc", "tc16", device=0, micro_batch=1, weights=w)
tc = torch.from_numpy(chunk).cuda()
for _ in range(10):
    gen.forward_chunk(tc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    gen.forward_chunk(tc)
e1.record()
torch.cuda.synchronize()
ms_dev = e0.elapsed_time(e1) / 100
t0 = time.perf_counter()
for _ in range(50):
    gen.forward_chunk(chunk)                      # NumPy in -> NumPy out (upload, forward, download)
ms_host = (time.perf_counter() - t0) / 50 * 1e3
print("config 1 (sample_imgs/02165, batch 1): device-resident %.3f ms/image, host NumPy in/out %.3f ms/image, oracle on %d CPU threads "
      "%.1f ms/image, feed (crop + landmark maps, 1 core) %.1f ms" % (ms_dev, ms_host, torch.get_num_threads(), t_cpu * 1e3, t_feed * 1e3))
print("  max-abs vs oracle: clip(con_rgb) %.2e  mask_pred %.2e  gs %.2e  | PSNR rgb %.1f dB" % (
    np.abs(rgb - r_rgb).max(), np.abs(mp - r_mp).max(), np.abs(gs - ref["gs"]).max(), psnr(rgb, r_rgb)))
gen.close()

# ---- config 2: UCB via fsr.test, batch 32 (+ ragged tail of 4): the 8 committed pairs tiled to 100 samples
from oracle import postprocess_ref as PP  # noqa: E402
files = sorted(glob.glob(os.path.join(FIX, "UCB", "input", "*", "*.png")))
loaded = []
for p in files:
    stem = os.path.basename(p)[:-4]
    fr = feed.load_frame(p, gt_path=p.replace(os.sep + "input" + os.sep, os.sep + "gt" + os.sep))
    masks = np.stack([np.rint(cv2.imread(os.path.join(FIX, "UCB_masks", k, stem + ".png"))[..., 0] / 255.0).astype(np.uint8)
                      for k in PP.MASK_KINDS])
    loaded.append({"img": fr["img"], "gt": fr["gt"], "uv": fr["uv"], "size": int(fr["box"][3] - fr["box"][1]), "masks": masks})
gen = Generator("gsc", "tc16", device=0, micro_batch=32, weights=w)
evaluate_ucb(gen, lambda i: loaded[i % 8], 100, batch=32)
torch.cuda.synchronize()
t0 = time.perf_counter()
out = evaluate_ucb(gen, lambda i: loaded[i % 8], 100, batch=32)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("config 2 (UCB, 100 samples = the 8 committed pairs tiled, batches of 32 + 4): %.1f images/s end to end (host stacking + H2D + "
      "generator + device post-processing + metrics readback), mean SSIM %.4f PSNR %.3f dB (random weights)" % (100 / dt, out["ssim"], out["psnr"]))
# device-only: generator + post-processing of one batch of 32
dev = torch.device("cuda", 0)
t = lambda k, dt_=torch.float32: torch.from_numpy(np.stack([loaded[i % 8][k] for i in range(32)])).to(dt_).to(dev)
img, gt, uv, mk = t("img"), t("gt"), t("uv"), t("masks", torch.uint8)
sizes = torch.tensor([loaded[i % 8]["size"] for i in range(32)], dtype=torch.int32, device=dev)
def step():
    _, rgb_, _, dif_ = gen(img, uv, None, want=("con_rgb", "dif"))
    return gen.postprocess_ucb(img, gt, rgb_, dif_, sizes, mk)
for _ in range(5):
    step()
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
_, rgb_, _, dif_ = gen(img, uv, None, want=("con_rgb", "dif"))
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    gen.postprocess_ucb(img, gt, rgb_, dif_, sizes, mk)
e1.record()
torch.cuda.synchronize()
ms_pp = e0.elapsed_time(e1) / 20
print("  device-resident batch of 32: generator + post-processing %.3f ms (%.0f images/s); post-processing alone %.3f ms "
      "(%.1f us/image, 11 launches)" % (ms, 32 / ms * 1e3, ms_pp, ms_pp / 32 * 1e3))
t0 = time.time()
for i in range(4):
    o = generator_forward(w, loaded[i]["img"][None], loaded[i]["uv"][None], variant="gsc")
    PP.test_step_postprocess(loaded[i]["img"], loaded[i]["gt"], o["con_rgb"][0], o["dif"][0], loaded[i]["size"],
                             {k: np.repeat(loaded[i]["masks"][j][..., None], 3, axis=2).astype(np.float64) for j, k in enumerate(PP.MASK_KINDS)})
print("  oracle (generator + post-processing restated on CPU, %d threads): %.1f images/s" % (torch.get_num_threads(), 4 / (time.time() - t0)))
gen.close()
