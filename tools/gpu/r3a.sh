#!/bin/bash
# round 2 (second session), GPU call A: verification of the HEAD tree: smoke, all GPU tests, per-layer bench, ncu launch list
mkdir -p gpurun_out/r3a; O=gpurun_out/r3a; rm -f $O/summary.txt
S=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s)-S ))" >> $O/summary.txt; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_all.log 2>&1; echo "pytest all rc=$? t=$(( $(date +%s)-S ))" >> $O/summary.txt; tail -14 $O/pytest_all.log
timeout 600 python bench.py --steps 10 --warmup 3 --layers > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$? t=$(( $(date +%s)-S ))" >> $O/summary.txt
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ncu_launch_list_gsc_mb256.csv python tools/profile_forward.py 256 > /dev/null 2>&1
echo "ncu list t=$(( $(date +%s)-S ))" >> $O/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3a/bench_final.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','e2e_compact','strong_scaling','tsm','cpu_baseline','clocks','gpu_launches'):
    print(k, d.get(k))
print('roofline', {k:v for k,v in d['roofline'].items()})
PY
grep -E "^ *[a-z_0-9.]+ +[0-9]" $O/bench_final.err | head -60
cat $O/summary.txt
