#!/bin/bash
# round 2, GPU call F: natural-V attention + L2-prefetched tail; full tests; bench with extras; ncu captures
mkdir -p gpurun_out/r2f; O=gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|Error" $O/pytest_all.log | tail -8
timeout 600 python bench.py --steps 8 --warmup 3 --layers > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
head -12 $O/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f/bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','e2e_compact','strong_scaling','tsm','config3_sfw_eval','cpu_baseline','clocks','gpu_launches'):
    print(k, d.get(k))
print('roofline', {k:v for k,v in d['roofline'].items() if k!='traffic'})
PY
# ncu: launch list + full metrics of one 256-image forward, source-level capture of the two heaviest kernels
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ncu_launch_list_gsc_mb256.csv python tools/profile_forward.py 256 > /dev/null 2>&1
timeout 1200 ncu --profile-from-start off --set full --clock-control none -o /tmp/prof_all -f python tools/profile_forward.py 256 > $O/prof_full.log 2>&1
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > $O/prof_all_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/prof_all_raw.csv > $O/ncu_full_per_launch_mb256.txt 2>&1; tail -55 $O/ncu_full_per_launch_mb256.txt
for L in 11 48 24; do
  timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -s $L -c 1 -o /tmp/prof_l$L -f python tools/profile_forward.py 256 > /dev/null 2>&1
  ncu -i /tmp/prof_l$L.ncu-rep --page source --csv > /tmp/src_l$L.csv 2>/dev/null
  echo "=== launch $L" >> $O/ncu_hot_instructions.txt; python tools/ncu_hot.py /tmp/src_l$L.csv 30 >> $O/ncu_hot_instructions.txt 2>&1
done
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/ncu_launch_list_tsm2_mb128.csv python tools/profile_forward.py 128 tsm > /dev/null 2>&1
cat $O/summary.txt
