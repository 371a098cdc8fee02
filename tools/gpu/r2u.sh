#!/bin/bash
# round 2, GPU call U: store-pattern micro-benchmark (L1 wavefronts per warp-wide global store)
mkdir -p gpurun_out/r2u; O=gpurun_out/r2u
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/store_pattern tools/micro/store_pattern.cu
/tmp/store_pattern | tee $O/store_pattern_time.txt
ncu --metrics l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__m_l1tex2xbar_write_sectors_mem_lg_op_st.sum,gpu__time_duration.sum --csv --log-file $O/store_pattern_ncu.csv /tmp/store_pattern > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2u/store_pattern_ncu.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki][:40]), {})[r[mi]] = r[vi]
for k in sorted(d)[::3]:
    print(k, {m[-45:]: v for m, v in d[k].items()})
PY
