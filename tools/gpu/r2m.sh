#!/bin/bash
# round 2, GPU call M: final single-GPU measurement suite of the round
mkdir -p gpurun_out/r2m; O=gpurun_out/r2m
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" >> $O/summary.txt; tail -3 $O/pytest_all.log
timeout 900 python bench.py --steps 10 --warmup 3 --layers > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?" >> $O/summary.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>/dev/null; cat $O/bench_reference.json | cut -c1-400
timeout 600 python tools/config_runs.py > $O/config_runs.txt 2>&1; cat $O/config_runs.txt
timeout 900 python tools/stress_2048.py > $O/stress_2048_f16.txt 2>&1; tail -8 $O/stress_2048_f16.txt
BSR_LIB=$PWD/blindshadowremoval_b200/libbsr_bf16.so timeout 900 python tools/stress_2048.py > $O/stress_2048_bf16.txt 2>&1; tail -8 $O/stress_2048_bf16.txt
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ncu_launch_list_gsc_mb256.csv python tools/profile_forward.py 256 > /dev/null 2>&1
timeout 1200 ncu --profile-from-start off --set full --clock-control none -o /tmp/prof_all -f python tools/profile_forward.py 256 > $O/prof_full.log 2>&1
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > $O/prof_all_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/prof_all_raw.csv "$(cat profiles/launch_names_gsc.txt)" > $O/ncu_full_per_launch_mb256.txt 2>&1; tail -52 $O/ncu_full_per_launch_mb256.txt | cut -c1-160
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/ncu_launch_list_tsm2_mb128.csv python tools/profile_forward.py 128 tsm > /dev/null 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m/bench_final.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','e2e_compact','strong_scaling','tsm','config3_sfw_eval','cpu_baseline','clocks','gpu_launches'):
    print(k, d.get(k))
print('roofline', {k:v for k,v in d['roofline'].items()})
PY
cat $O/summary.txt
