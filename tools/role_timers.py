"""Per-launch role timers (BSR_ABLATE=8): where does a persistent conv CTA spend its cycles?"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BSR_ABLATE"] = str(8 | int(os.environ.get("ABL", "0")))
os.environ["BSR_DEBUG_KEEP"] = "1"
os.environ["BSR_PROFILE"] = "1"          # per-launch CUDA-event times next to the cycle counters -> effective SM clock
from blindshadowremoval_b200.generator import Generator  # noqa: E402
from blindshadowremoval_b200.synthetic import make_inputs  # noqa: E402

mb = int(os.environ.get("MB", "128"))
LAUNCHES = open(os.path.join(ROOT, "profiles", "launch_names_gsc.txt")).read().strip().split(",")
gen = Generator("gsc", "bf16", device=0, micro_batch=mb, seed=1234)
d = make_inputs(mb, 0)
img, uv = torch.from_numpy(d["img"]).cuda(), torch.from_numpy(d["uv"]).cuda()
for _ in range(int(os.environ.get("REPS", "3"))):
    gen(img, uv, None, want=("con_rgb", "dif"))
torch.cuda.synchronize()
lt = gen.layer_times()
t = gen.debug_read("timers").reshape(64, 16)
names = dict(enumerate(LAUNCHES))           # launch index inside one forward (h->launches at launch time)
print("launch  name       | producer: total wait_empty dep_wait steps | mma: total wait_full wait_tempty wait_res tiles | epi: total wait_tfull tiles")
for i in range(64):
    if t[i, 0] == 0 and t[i, 4] == 0:
        continue
    if t[i, 9] == -1:       # single-pass attention kernel (attention_fa.cuh): row 0 of tile A in CTA 0
        it = max(t[i, 8], 1)
        print("%3d attn+w (fa) | softmax warp: total %d = %d items; per item: main loop %d (waiting for S %d) | O->smem %d | wait D2 %d | tail %d (waiting for residuals %d, staging barrier %d)" % (
            i, t[i, 0], t[i, 8], t[i, 1] / it, t[i, 2] / it, t[i, 3] / it, t[i, 4] / it, t[i, 5] / it, t[i, 6] / it, t[i, 7] / it))
        continue
    print("%3d %-10s | %9d %9d %8d %5d | %9d %9d %9d %8d %4d | %9d %9d %4d" % (
        i, names.get(i, ""), t[i, 0], t[i, 1], t[i, 2], t[i, 3], t[i, 4], t[i, 5], t[i, 6], t[i, 7], t[i, 8], t[i, 9],
        t[i, 10], t[i, 11]))
    ms = lt[i][1] if i < len(lt) else 0.0
    print("      mma: fence %d issue %d commit %d | producer tma-issue %d | %.4f ms (%s) -> %.0f MHz if CTA 0 spans the launch" % (
        t[i, 12], t[i, 13], t[i, 14], t[i, 15], ms, lt[i][0] if i < len(lt) else "", max(t[i, 0], t[i, 4], t[i, 9]) / max(ms, 1e-9) / 1e3))
