"""Conditioning probe for the peaked-softmax cases (run on the GPU box): for theta gains g, error of the 16-bit path and
of the fp32 check mode against the fp32 oracle, and of the fp32 oracle against the fp64 oracle (the network's own
sensitivity to 1e-7 perturbations).  Prints one line per gain."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BSR_DEBUG_KEEP"] = "1"
from blindshadowremoval_b200.generator import Generator  # noqa: E402
from blindshadowremoval_b200.synthetic import make_inputs  # noqa: E402
from blindshadowremoval_b200.weights import random_weights  # noqa: E402
from oracle.calibrate import centre_hole_threshold  # noqa: E402
from oracle.generator_ref import generator_forward  # noqa: E402


def peaked(gain, undamp):
    w = dict(random_weights("gsc", 1234))
    for i in range(6):
        w["res_stack/%d/non_local/theta/kernel" % i] = w["res_stack/%d/non_local/theta/kernel" % i] * np.float32(gain)
        if undamp:
            w["res_stack/%d/non_local/w/kernel" % i] = w["res_stack/%d/non_local/w/kernel" % i] * np.float32(1 / 0.35)
    return w


def main():
    d = make_inputs(2, seed=6, with_reg=True)
    for undamp in (False, True):
        for gain in (1.0, 2.0, 3.0, 4.0, 8.0):
            w = centre_hole_threshold(peaked(gain, undamp), d["img"], d["uv"], None, variant="gsc", frame=1)
            row = {}
            for prec in ("tc16", "fp32check"):
                gen = Generator("gsc", prec, device=0, micro_batch=2, weights=w)
                t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
                out = [o.cpu().numpy() for o in gen(t["img"], t["uv"], None)]
                bm = gen.debug_read("bmask").reshape(2, 32, 32, 1)
                ref = generator_forward(w, d["img"], d["uv"], variant="gsc", bmask_override=bm)
                row[prec] = {k: float(np.abs(o - ref[k]).max()) for k, o in zip(("gs", "con_rgb", "mask22", "dif"), out)}
                if prec == "tc16":
                    qk = gen.debug_read("qk5").reshape(2, 1024, 256).astype(np.float64)
                    s = qk[..., :128] @ qk[..., 128:].transpose(0, 2, 1)
                    row["spread5"] = float(np.median(s.max(-1) - s.min(-1)))
                    row["rowmax5"] = float(np.median(s.max(-1)))
                    ref64 = generator_forward(w, d["img"], d["uv"], variant="gsc", bmask_override=bm, dtype=torch.float64)
                    row["oracle32_vs_64"] = float(np.abs(ref64["con_rgb"] - ref["con_rgb"]).max())
                gen.close()
            print("undamped_w=%d gain=%.0f block5 logits: row-max median %.1f spread %.1f | con_rgb max-abs: tc16 %.2e fp32check %.2e "
                  "oracle fp32-vs-fp64 %.2e | gs tc16 %.2e" % (undamp, gain, row["rowmax5"], row["spread5"], row["tc16"]["con_rgb"],
                                                              row["fp32check"]["con_rgb"], row["oracle32_vs_64"], row["tc16"]["gs"]),
                  flush=True)


if __name__ == "__main__":
    main()
