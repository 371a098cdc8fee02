"""Extract the two numeric templates the reference's dataset code feeds to its map generators and store them as
blindshadowremoval_b200/data/face_template.npz (run once in the build container, where /root/reference exists):

  uv      [68,3]  canonical UV coordinates of the 68 landmarks    (dataset.py:10-13: ``uv = np.transpose(...)``)
  lm_ref  [68,2]  reference landmark positions, normalised by 256 (dataset.py:14-16: ``lm_ref = np.transpose(...)/256.``)

dataset.py imports tensorflow and cannot be imported here, so the two literals are read with ``ast``.
"""
import ast
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/dataset.py").read()
tree = ast.parse(src)
lits = {}
for node in tree.body:
    if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
        name = node.targets[0].id
        if name in ("uv", "lm_ref") and isinstance(node.value, ast.List) and name not in lits:
            lits[name] = np.asarray(ast.literal_eval(node.value), dtype=np.float32)
uv = np.transpose(lits["uv"]).astype(np.float32)                       # dataset.py:13
lm_ref = (np.transpose(lits["lm_ref"]) / 256.0).astype(np.float32)     # dataset.py:16
assert uv.shape == (68, 3) and lm_ref.shape == (68, 2), (uv.shape, lm_ref.shape)
out = os.path.join(ROOT, "blindshadowremoval_b200", "data", "face_template.npz")
np.savez_compressed(out, uv=uv, lm_ref=lm_ref)
print("wrote", out, uv.shape, lm_ref.shape, float(uv.min()), float(uv.max()), float(lm_ref.min()), float(lm_ref.max()))
