"""Compare per-layer times of two `bench.py --layers` JSON lines (files given as argv[1], argv[2])."""
import json
import sys


def load(path):
    d = json.loads(open(path).read().strip().splitlines()[-1])
    return d["value"], {r["layer"]: r for r in d.get("layers", [])}


va, a = load(sys.argv[1])
vb, b = load(sys.argv[2])
print("throughput: %.0f -> %.0f img/s" % (va, vb))
for k in a:
    if k in b:
        ta, tb = a[k].get("ms_per_forward", 0.0), b[k].get("ms_per_forward", 0.0)
        print("%-16s %8.3f -> %8.3f ms  (%+.1f%%)" % (k, ta, tb, 100.0 * (tb - ta) / ta if ta else 0.0))
