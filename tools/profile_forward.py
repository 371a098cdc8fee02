"""One profiled forward for ncu:  ncu --profile-from-start off ... python tools/profile_forward.py [mb] [variant]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blindshadowremoval_b200.generator import Generator  # noqa: E402
from blindshadowremoval_b200.synthetic import make_inputs  # noqa: E402

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
variant = sys.argv[2] if len(sys.argv) > 2 else "gsc"
frame = 2
gen = Generator(variant, "bf16", device=0, micro_batch=mb, seed=1234)
d = make_inputs(mb, 0, with_reg=True)
img, uv, reg = (torch.from_numpy(d[k]).cuda() for k in ("img", "uv", "reg"))
for _ in range(3):
    gen(img, uv, reg, frame=frame, want=("con_rgb", "dif"))
torch.cuda.synchronize()
torch.cuda.profiler.start()
gen(img, uv, reg, frame=frame, want=("con_rgb", "dif"))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches per forward:", gen.launch_count())
