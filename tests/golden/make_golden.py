"""Regenerates tests/golden/* (run in the build container, where /root/reference exists).

  ckpt_variables.json   generator variable names/shapes parsed from the reference's own
                        log/*/ckpt-*.index files (the only structural ground truth the reference ships)
  oracle_gsc.npz / oracle_tsm.npz
                        small fingerprints of the oracle's outputs on seeded inputs/weights (32x32
                        dif_small + bmask in full, 8x-subsampled con_rgb/dif/gs), so a later edit of the
                        oracle cannot drift silently.  They pin the oracle against ITSELF, not against
                        TensorFlow (parity with TF is unpinned, see oracle/generator_ref.py).
"""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from blindshadowremoval_b200.synthetic import make_inputs  # noqa: E402
from blindshadowremoval_b200.tf_checkpoint import generator_variables  # noqa: E402
from blindshadowremoval_b200.weights import random_weights  # noqa: E402
from oracle.calibrate import centre_hole_threshold  # noqa: E402
from oracle.generator_ref import generator_forward  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(GOLD, exist_ok=True)
    out = {}
    for tag, pat in (("gsc", "/root/reference/log/*-gradients/ckpt-94.index"),
                     ("tsm", "/root/reference/log/*-with-TSM/ckpt-110.index")):
        path = glob.glob(pat)[0]
        out[tag] = {k: list(v) for k, v in sorted(generator_variables(path).items())}
    with open(os.path.join(GOLD, "ckpt_variables.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)
    for variant, frame, n in (("gsc", 1, 2), ("tsm", 2, 2)):
        w = random_weights(variant, 1234)
        d = make_inputs(n, 0, with_reg=True)
        w = centre_hole_threshold(w, d["img"], d["uv"], d["reg"], variant=variant, frame=frame)
        o = generator_forward(w, d["img"], d["uv"], d["reg"], variant=variant, frame=frame)
        np.savez_compressed(os.path.join(GOLD, "oracle_%s.npz" % variant),
                            conv3_bias=w["conv3/conv/bias"], dif_small=o["dif_small"], bmask=o["bmask"],
                            con_rgb=o["con_rgb"][:, ::8, ::8], dif=o["dif"][:, ::8, ::8], gs=o["gs"][:, ::8, ::8])
    print("wrote", os.listdir(GOLD))


if __name__ == "__main__":
    main()
