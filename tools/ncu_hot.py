"""Summarise an `ncu --page source --csv` export: hottest SASS instructions by stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    try:
        n = int(r[isamp])
    except Exception:
        continue
    data.append((n, r))
tot = sum(n for n, _ in data)
print("total samples", tot)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for n, r in sorted(data, key=lambda t: -t[0])[:top]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stalls), reverse=True)[:2]
    print("%6d %5.1f%%  %-70s %s" % (n, 100.0 * n / tot, r[isrc][:70], " ".join("%s=%d" % (h[6:], v) for v, h in st if v)))
