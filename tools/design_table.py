"""Markdown per-layer roofline table for DESIGN.md from a `bench.py --layers` JSON line (+ DRAM bytes of the committed ncu
capture):  python tools/design_table.py gpurun_out/<run>/bench.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
tj = json.load(open("profiles/ncu_traffic.json"))
imgs = d["roofline"]["images_per_launch"]
alias = {"attention+w": "r1.attn+w"}
rows = {}
for r in d["layers"]:
    name = r["layer"]
    key = name.split(".")[1] if name.startswith("res") and "." in name else name
    if name.startswith("res"):
        key = "res " + key + " x6"
    a = rows.setdefault(key, dict(ms=0.0, calls=0, tf=[], gb=[], frac=[], bound=r["bound"], src=name))
    a["ms"] += r["ms_per_forward"]
    a["calls"] += r["calls_per_forward"]
    a["tf"].append(r["tflops"]); a["gb"].append(r["gbs"]); a["frac"].append(r["frac"])
tot = sum(a["ms"] for a in rows.values())
print("| unit (launches per forward) | ms per forward (%d images) | share | bound | achieved | frac of measured peak | ncu DRAM MB / image |" % imgs)
print("|---|---|---|---|---|---|---|")
for key, a in sorted(rows.items(), key=lambda kv: -kv[1]["ms"]):
    src = a["src"]
    lk = alias.get(src, src.replace("res", "r") if src.startswith("res") else src)
    ent = tj["layers"].get(lk) or tj["layers"].get(lk.replace("r0", "r1"))
    if src == "clr_up1" and "clr_up1" not in tj["layers"]:      # captures from before the fused halo kernel: 4 phase launches
        b = sum(tj["layers"].get("clr_up1.p%d" % i, {"dram_bytes": 0})["dram_bytes"] for i in range(4))
        ent = {"dram_bytes": b}
    mb = "%.2f" % (ent["dram_bytes"] / tj["images_per_launch"] / 1e6) if ent else "-"
    mean = lambda v: sum(v) / len(v)
    ach = "%.0f TFLOP/s" % mean(a["tf"]) if a["bound"] == "tensor" else ("%.0f GB/s" % mean(a["gb"]) if mean(a["gb"]) else "-")
    print("| %s (%d) | %.3f | %.1f %% | %s | %s | %s | %s |" % (key, a["calls"], a["ms"], 100 * a["ms"] / tot, a["bound"] if mean(a["frac"]) else "glue", ach,
                                                   ("%.2f" % mean(a["frac"])) if mean(a["frac"]) else "-", mb))
print("| **sum of per-layer events** | **%.2f** | | | | | |" % tot)
