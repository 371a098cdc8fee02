"""L1 data-pipe occupancy per launch from an ncu metrics csv (tools/gpu/r2v.sh): LSU wavefronts (global / shared) and the
tensor-core operand wavefronts share one 128-byte-per-cycle pipe per SM."""
import csv
import os
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ii, mi, vi = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value")
d = {}
for r in rows[1:]:
    d.setdefault(int(r[ii]), {})[r[mi]] = float(r[vi].replace(",", ""))
names = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "launch_names_gsc.txt")).read().strip().split(",")
print("%-11s %9s %10s %10s %10s %10s %9s %8s %6s" % ("launch", "ms", "cycles", "lsu_glob/SM", "lsu_smem/SM", "tc_smem/SM", "requests", "wf/req", "util"))
for k in sorted(d):
    m = d[k]
    lg, sh, tc = (m["l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum"] / 148, m["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"] / 148,
                  m["l1tex__data_pipe_tc_wavefronts_mem_shared.sum"] / 148)
    el = m["sm__cycles_elapsed.avg"]
    rq = m["l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"] + m["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]
    print("%-11s %9.4f %10.0f %10.0f %10.0f %10.0f %9.0f %8.1f %6.2f" % (
        names[k] if k < len(names) else str(k), m["gpu__time_duration.sum"] / 1e6, el, lg, sh, tc, rq, lg * 148 / max(rq, 1), (lg + sh + tc) / el))
