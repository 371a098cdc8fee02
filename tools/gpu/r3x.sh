#!/bin/bash
# round 2, final single-GPU measurement suite of the HEAD tree (evidence for profiles/r2_*)
mkdir -p gpurun_out/r3x; O=gpurun_out/r3x; rm -f $O/summary.txt
S=$(date +%s); T() { echo "$1 rc=$2 t=$(( $(date +%s)-S ))" >> $O/summary.txt; }
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; T smoke $?; tail -1 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; T pytest_all $?; tail -2 $O/pytest_all.log
timeout 900 python bench.py --steps 10 --warmup 3 --layers > $O/bench_final.json 2> $O/bench_final.err; T bench $?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>/dev/null; T bench_ref $?; cut -c1-200 $O/bench_reference.json
timeout 600 python tools/config_runs.py > $O/config_runs.txt 2>&1; T config_runs $?; tail -6 $O/config_runs.txt
timeout 900 python tools/stress_2048.py > $O/stress_2048_f16.txt 2>&1; T stress $?; tail -4 $O/stress_2048_f16.txt
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ncu_launch_list_gsc_mb256.csv python tools/profile_forward.py 256 > /dev/null 2>&1; T ncu_list $?
timeout 1200 ncu --profile-from-start off --set full --clock-control none -o /tmp/prof_all -f python tools/profile_forward.py 256 > $O/prof_full.log 2>&1; T ncu_full $?
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > $O/prof_all_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/prof_all_raw.csv "$(cat profiles/launch_names_gsc.txt)" > $O/ncu_full_per_launch_mb256.txt 2>&1; tail -2 $O/ncu_full_per_launch_mb256.txt | cut -c1-160
python tools/datapipe_table.py $O/prof_all_raw.csv "$(cat profiles/launch_names_gsc.txt)" > $O/ncu_datapipe_mb256.txt 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/ncu_launch_list_tsm2_mb128.csv python tools/profile_forward.py 128 tsm > /dev/null 2>&1; T ncu_tsm $?
python tools/ncu_tsm_table.py $O/ncu_launch_list_tsm2_mb128.csv > $O/ncu_tsm2_per_launch_mb128.txt 2>&1
BSR_LIB=$PWD/blindshadowremoval_b200/libbsr_timers.so MB=256 timeout 300 python tools/role_timers.py > $O/role_timers_mb256.txt 2>&1; T timers $?
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/profile_forward.py 2 gsc > $O/racecheck_gsc.log 2>&1; grep -E "RACECHECK SUMMARY|ERROR SUMMARY" $O/racecheck_gsc.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/profile_forward.py 2 tsm > $O/racecheck_tsm.log 2>&1; grep -E "RACECHECK SUMMARY|ERROR SUMMARY" $O/racecheck_tsm.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/profile_forward.py 2 gsc > $O/memcheck_gsc.log 2>&1; grep -E "ERROR SUMMARY" $O/memcheck_gsc.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 5 python tools/profile_forward.py 2 gsc > $O/synccheck_gsc.log 2>&1; grep -E "ERROR SUMMARY" $O/synccheck_gsc.log
T sanitizers 0
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3x/bench_final.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','e2e_compact','strong_scaling','tsm','config3_sfw_eval','cpu_baseline','clocks','gpu_launches'):
    print(k, str(d.get(k))[:300])
print('roofline', {k:v for k,v in d['roofline'].items()})
PY
cat $O/summary.txt
