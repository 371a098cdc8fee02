"""One profiled pass over the glue / post-processing kernels that are not part of the generator's 46 launches:
chunk unpack + caller glue (bsr_forward_chunk), composite, the UCB post-processing (bsr_postprocess_ucb), the compact
converters.  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum ...
python tools/profile_glue.py [n]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blindshadowremoval_b200.generator import Generator  # noqa: E402
from blindshadowremoval_b200.synthetic import make_inputs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
gen = Generator("gsc", "tc16", device=0, micro_batch=n, seed=1234)
d = make_inputs(n, 0, with_reg=True)
rng = np.random.default_rng(0)
face = (rng.random((n, 256, 256, 1)) > 0.3).astype(np.float32)
chunk = torch.from_numpy(np.concatenate([d["img"], d["uv"], d["reg"], face], axis=3)).cuda()       # 13-channel layout
img = torch.from_numpy(d["img"]).cuda()
gt = torch.from_numpy(np.clip(d["img"] * 1.1, 0, 1).astype(np.float32)).cuda()
sizes = torch.full((n,), 200, dtype=torch.int32).cuda()
masks = torch.from_numpy((rng.random((n, 7, 256, 256)) > 0.7).astype(np.uint8) * 255).cuda()
img_u8 = (d["img"] * 255).astype(np.uint8)
uv32 = np.ascontiguousarray(d["uv"][:, 3::8, 3::8, :])                       # any 32x32 field will do for timing


def once():
    rgb, mp = gen.forward_chunk(chunk)
    out = gen.composite(rgb, img, mp)
    final, det, met = gen.postprocess_ucb(img, gt, rgb, mp, sizes, masks)
    gen.forward_compact(img_u8, uv32)
    return out, final, met


for _ in range(2):
    once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("glue pass done, n =", n)
