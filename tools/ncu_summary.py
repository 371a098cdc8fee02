"""Compact per-launch table from `ncu -i X.ncu-rep --page raw --csv` (one forward, profiler-range capture)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [("gpu__time_duration.sum", "t_us"), ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
idx = [(hdr.index(c), n) for c, n in cols if c in hdr]
names = sys.argv[2].split(",") if len(sys.argv) > 2 else []
kn = hdr.index("Kernel Name")
print("units: time %s, dram %s" % (units[hdr.index("gpu__time_duration.sum")], units[hdr.index("dram__bytes_read.sum")]))
print("%3s %-12s %-28s " % ("#", "layer", "kernel") + " ".join("%9s" % n for _, n in idx))
tot = 0.0
for i, r in enumerate(data):
    tot += float(r[idx[0][0]].replace(",", ""))
    print("%3d %-12s %-28s " % (i, names[i] if i < len(names) else "", r[kn].split("(")[0][-28:]) +
          " ".join("%9s" % r[j][:9] for j, _ in idx))
print("total time", round(tot, 1))
