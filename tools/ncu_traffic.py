"""profiles/ncu_traffic.json from `ncu -i X.ncu-rep --page raw --csv` of one profiled forward (tools/profile_forward.py):
dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by the launch names of profiles/launch_names_gsc.txt;
"attention+w" = the mean over the six fused attention launches (bench.py's roofline.traffic reads it)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
images = int(sys.argv[2])
source = sys.argv[3]
hdr, units, data = rows[0], rows[1], rows[2:]
ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
names = open(os.path.join(ROOT, "profiles", "launch_names_gsc.txt")).read().strip().split(",")
assert len(data) == len(names), (len(data), len(names))
layers = {}
for n, r in zip(names, data):
    b = float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
    layers[n] = {"dram_bytes": int(round(b))}
att = [v["dram_bytes"] for k, v in layers.items() if k.endswith("attn+w")]
layers["attention+w"] = {"dram_bytes": int(round(sum(att) / len(att)))}
json.dump({"source": source, "images_per_launch": images, "layers": layers}, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=0)
print("attention+w", layers["attention+w"], "total", sum(v["dram_bytes"] for k, v in layers.items() if k != "attention+w"))
