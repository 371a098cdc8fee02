"""Per-launch table of the TSM forward from the ncu metrics csv (gpu__time_duration + dram bytes): tools/gpu/r2x.sh."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ii, ki, mi, ui, vi = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
d = {}
for r in rows[1:]:
    e = d.setdefault(int(r[ii]), {"k": r[ki].split("(")[0][:44]})
    e[r[mi]] = float(r[vi].replace(",", "")) * scale[r[ui]]
print("# TSM forward (frame 2, 128 frames per launch): ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none")
print("# id kernel time_us dram_MB achieved_GB/s (cold, serialised)")
tot = 0.0
for k in sorted(d):
    e = d[k]
    t = e["gpu__time_duration.sum"]
    b = e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"]
    tot += t
    print("%3d %-44s %10.1f %9.1f %9.0f" % (k, e["k"], t, b / 1e6, b / t / 1e3))
print("# total %.1f us over %d launches" % (tot, len(d)))
