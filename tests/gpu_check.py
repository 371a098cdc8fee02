"""Bring-up report (run on the GPU box): per-intermediate error of every precision mode vs the oracle."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blindshadowremoval_b200.generator import Generator  # noqa: E402
from blindshadowremoval_b200.synthetic import make_inputs  # noqa: E402
from blindshadowremoval_b200.weights import random_weights  # noqa: E402
from oracle.calibrate import centre_hole_threshold  # noqa: E402
from oracle.generator_ref import generator_forward  # noqa: E402

NAMES = ["x1", "x2", "x3", "x_in0", "res0", "res1", "res2", "up1", "up2", "up3", "dif_small", "bmask", "x_in3",
         "res3", "res4", "res5", "clr_up1", "clr_up2", "clr_up3"]


def run(variant, precision, n, frame, env, report):
    for k in ("BSR_FORCE_DIRECT", "BSR_TC_DISABLE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    os.environ["BSR_DEBUG_KEEP"] = "1"
    tag = "%s/%s/%s" % (variant, precision, ",".join("%s=%s" % kv for kv in env.items()) or "-")
    print("=" * 100)
    print(tag)
    w = random_weights(variant, 1234)
    d = make_inputs(n, 0, with_reg=True)
    w = centre_hole_threshold(w, d["img"], d["uv"], d["reg"], variant=variant, frame=frame)
    gen = Generator(variant, precision, device=0, micro_batch=n, weights=w)
    img, uv, reg = (torch.from_numpy(d[k]).cuda() for k in ("img", "uv", "reg"))
    t0 = time.time()
    gs, rgb, m22, dif = gen(img, uv, reg, frame=frame, share=True, training=False)
    torch.cuda.synchronize()
    dt = time.time() - t0
    flag = gen.debug_read("errflag")[0]
    bm = gen.debug_read("bmask").reshape(n, 32, 32, 1)
    ref0 = generator_forward(w, d["img"], d["uv"], d["reg"], variant=variant, frame=frame, keep=True)
    flips = int((bm != ref0["bmask"]).sum())
    ref = ref0 if flips == 0 else generator_forward(w, d["img"], d["uv"], d["reg"], variant=variant, frame=frame,
                                                    keep=True, bmask_override=bm)
    rows = {}
    print("first call %.3fs  launches %d  errflag %d  bmask flips %d / %d (mean %.2f)" %
          (dt, gen.launch_count(), flag, flips, bm.size, bm.mean()))
    for name in NAMES:
        a = gen.debug_read(name)
        b = ref[name].reshape(-1)
        if a.size != b.size:
            print("  %-10s SIZE MISMATCH %d vs %d" % (name, a.size, b.size))
            continue
        err = np.abs(a - b)
        rows[name] = float(err.max())
        print("  %-10s max|err| %.3e  mean|err| %.3e  ref rms %.3f  nan %d" %
              (name, err.max(), err.mean(), np.sqrt((b ** 2).mean()), int(np.isnan(a).sum())))
    for name, t in (("gs", gs), ("con_rgb", rgb), ("mask22", m22), ("dif", dif)):
        a = t.cpu().numpy()
        err = np.abs(a - ref[name])
        mse = float(((a - ref[name]) ** 2).mean())
        rows[name] = float(err.max())
        print("  %-10s max|err| %.3e  mean|err| %.3e  psnr %.1f dB" %
              (name, err.max(), err.mean(), 10 * np.log10(1.0 / max(mse, 1e-20))))
    report[tag] = dict(errflag=float(flag), flips=flips, launches=gen.launch_count(), err=rows)
    gen.close()


def main():
    report = {}
    which = sys.argv[1:] or ["fp32", "direct", "tc"]
    try:
        if "fp32" in which:
            run("gsc", "fp32check", 2, 1, {}, report)
            run("tsm", "fp32check", 4, 2, {}, report)
        if "direct" in which:
            run("gsc", "bf16", 2, 1, {"BSR_FORCE_DIRECT": "1"}, report)
        if "tc" in which:
            run("gsc", "bf16", 2, 1, {}, report)
            run("tsm", "bf16", 4, 2, {}, report)
    finally:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as fh:
            json.dump(report, fh, indent=1)


if __name__ == "__main__":
    main()
