"""Per-kernel histogram of the Blackwell-specific SASS opcodes in libbsr.so (evidence that the hot path really is
tcgen05 / TMA code):  python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt
UTCHMMA = tcgen05.mma kind::f16, UTMALDG / UTMASTG = TMA tensor loads / stores, UTMAPF = TMA prefetch, LDTM / STTM =
tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, UTCATOMSWS = TMEM allocation."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "blindshadowremoval_b200", "libbsr.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
keys = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "MUFU.EX2", "F2FP"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in keys:
        if re.search(r"\b" + re.escape(k), line):
            per[cur][k] += 1
    if re.search(r"^\s+/\*[0-9a-f]{4,}\*/", line):
        per[cur]["instructions"] += 1
print("library:", os.path.relpath(lib, ROOT), "\n")
print("%-62s %7s " % ("kernel", "instrs") + " ".join("%8s" % k[:8] for k in keys))
tot = collections.Counter()
for name, c in per.items():
    if not any(c[k] for k in keys[:7]):
        continue
    print("%-62s %7d " % (name[-62:], c["instructions"]) + " ".join("%8d" % c[k] for k in keys))
    tot.update(c)
print("%-62s %7d " % ("TOTAL (kernels with tcgen05 / TMA opcodes)", tot["instructions"]) + " ".join("%8d" % tot[k] for k in keys))
print("\nkernels without tcgen05 / TMA opcodes (glue, post-processing, fp32 check mode):")
print(", ".join(sorted({n.split("<")[0].split("::")[-1] for n, c in per.items() if not any(c[k] for k in keys[:7])})))
