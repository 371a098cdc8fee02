"""Read one bench.py JSON line on stdin and print `tag n_gpus value e2e e2e_compact ms_per_step` (sweep helper)."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
g = lambda k: (d.get(k) or {}).get("value")
print(sys.argv[1] if len(sys.argv) > 1 else "-", d["n_gpus"], d["value"], g("e2e"), g("e2e_compact"), d["ms_per_step"])
