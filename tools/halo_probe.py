"""Bring-up probe for the halo-tile transposed convs: up3 / clr_up3 against the oracle at a batch where the weights are
resident (the halo program needs that), for both descriptor variants."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BSR_DEBUG_KEEP"] = "1"
from blindshadowremoval_b200.generator import Generator
from blindshadowremoval_b200.synthetic import make_inputs
from blindshadowremoval_b200.weights import random_weights
from oracle.generator_ref import generator_forward
n = 12
w = random_weights("gsc", 1234)
d = make_inputs(n, 3)
gen = Generator("gsc", "tc16", device=0, micro_batch=n, weights=w)
t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
out = gen(t["img"], t["uv"], None)
torch.cuda.synchronize()
bm = gen.debug_read("bmask").reshape(n, 32, 32, 1)
ref = generator_forward(w, d["img"], d["uv"], variant="gsc", keep=True, bmask_override=bm)
print("env", {k: v for k, v in os.environ.items() if k.startswith("BSR_")}, "plan", gen.plan_counters(), "errflag", gen.debug_read("errflag")[0])
for name in ("up2", "up3", "clr_up2", "clr_up3"):
    a, b = gen.debug_read(name), ref[name].reshape(-1)
    print("  %-8s rel rms err %.3e  max %.3e" % (name, np.sqrt(((a - b) ** 2).mean()) / np.sqrt((b ** 2).mean()), np.abs(a - b).max()))
print("  con_rgb max-abs %.3e" % np.abs(out[1].cpu().numpy() - ref["con_rgb"]).max())
