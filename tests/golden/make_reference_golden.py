"""Golden vectors produced BY THE REFERENCE'S OWN CODE, executed in the build container.

TensorFlow / matplotlib cannot be imported here, so the reference modules cannot be imported as a whole.  A few of
their functions are pure NumPy / SciPy / cv2 though; this script pulls exactly those function definitions out of the
reference sources with ``ast`` (no source is copied into the repository), executes them unmodified and stores their
outputs:

  warp_reference.npz   /root/reference/warp.py:61-68  sp_batch_map_coordinates   (docstring: "Reference implementation
                       /root/reference/warp.py:118-131 sp_batch_map_offsets       for tf_batch_map_offsets")
                       on seeded feature maps + offset fields -> pins the warp of the oracle's ShareLayer and, through the
                       -m gpu test, the CUDA ShareLayer kernels (model_with_TSM.py:204-229).
  crop_reference.npz   /root/reference/utils.py:356-433 face_crop_and_resize (aug=False) on the reference's real files
                       (sample_imgs/02165 and the 8 UCB pairs copied to tests/fixtures/) -> pins feed.crop_and_resize.

Run:  python tests/golden/make_reference_golden.py      (needs /root/reference; the tests only read the .npz files)
"""
import ast
import glob
import os
import random
import sys

import cv2
import numpy as np
from scipy import ndimage
from scipy.ndimage import map_coordinates

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
FIX = os.path.join(ROOT, "tests", "fixtures")


def extract(path, names, namespace):
    """exec the FunctionDef nodes `names` of the reference file `path` inside `namespace`."""
    tree = ast.parse(open(path).read(), filename=path)
    found = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), namespace)
            found[node.name] = (node.lineno, node.end_lineno)
    missing = set(names) - set(found)
    if missing:
        raise RuntimeError("not found in %s: %s" % (path, sorted(missing)))
    return found


def warp_golden():
    # warp.py:4 imports map_coordinates from the removed scipy.ndimage.interpolation path; same function
    ns = {"np": np, "sp_map_coordinates": map_coordinates}
    lines = extract(os.path.join(REF, "warp.py"), ["sp_batch_map_coordinates", "sp_batch_map_offsets"], ns)
    rng = np.random.default_rng(2022)
    cases = {}
    # (name, batch, size, offset magnitude in pixels): inside the map, crossing the border (clip), integer offsets
    for name, b, s, mag in (("small", 6, 32, 1.5), ("border", 4, 32, 9.0), ("integer", 3, 32, 0.0)):
        x = rng.standard_normal((b, s, s)).astype(np.float32)
        off = (rng.uniform(-mag, mag, (b, 1, 1, 2)) + 0.3 * mag * rng.standard_normal((b, s, s, 2))).astype(np.float32)
        if name == "integer":
            off = rng.integers(-3, 4, (b, s, s, 2)).astype(np.float32)
        out = ns["sp_batch_map_offsets"](x.astype(np.float64), off.astype(np.float64))       # [b, s*s]
        cases[name + "_x"] = x
        cases[name + "_off"] = off
        cases[name + "_out"] = out.reshape(b, s, s)
    np.savez_compressed(os.path.join(GOLD, "warp_reference.npz"), **cases)
    return lines


def crop_golden():
    ns = {"np": np, "cv2": cv2, "random": random, "ndimage": ndimage}
    lines = extract(os.path.join(REF, "utils.py"), ["face_crop_and_resize"], ns)
    files = sorted(glob.glob(os.path.join(FIX, "sample_imgs", "*", "*.npy"))) + \
        sorted(glob.glob(os.path.join(FIX, "UCB", "input", "*", "*.npy")))
    out = {}
    for f in files:
        key = os.path.basename(f)[:-4].replace("-", "_")
        img = cv2.cvtColor(cv2.imread(f[:-4] + ".png"), cv2.COLOR_BGR2RGB) / 255.          # dataset.py:159, 627
        crop, lm, lm_mirror, box = ns["face_crop_and_resize"](img, np.load(f), 256)
        out[key + "_lm"] = lm.astype(np.float64)
        out[key + "_lm_mirror"] = lm_mirror.astype(np.float64)
        out[key + "_box"] = np.asarray(box, np.int64)
        out[key + "_crop8"] = crop[::8, ::8].astype(np.float64)                              # fingerprint of the resized crop
        out[key + "_crop_mean"] = np.asarray(crop.mean(axis=(0, 1)), np.float64)
    np.savez_compressed(os.path.join(GOLD, "crop_reference.npz"), **out)
    return lines, [os.path.relpath(f, ROOT) for f in files]


def main():
    if not os.path.isdir(REF):
        sys.exit("needs the reference checkout at /root/reference (build container only)")
    print("warp.py functions at lines", warp_golden())
    print("utils.py functions at lines", crop_golden())


if __name__ == "__main__":
    main()
