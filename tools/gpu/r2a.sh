#!/bin/bash
# round 2, GPU call A: f16 storage vs bf16 A/B, new parity tests, sanitizer triage of the r1 attention kernel
mkdir -p gpurun_out/r2a; O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q -s > $O/pytest_f16.log 2>&1; echo "pytest f16 rc=$?" >> $O/summary.txt
tail -5 $O/pytest_f16.log
timeout 300 python tests/gpu_check.py tc > $O/gpu_check_f16.log 2>&1
BSR_LIB=$PWD/blindshadowremoval_b200/libbsr_bf16.so timeout 300 python tests/gpu_check.py tc > $O/gpu_check_bf16.log 2>&1
grep -E "^(gsc|tsm)|con_rgb|gs  |dif  |mask22|flips" $O/gpu_check_f16.log | head -30
echo ---- bf16; grep -E "^(gsc|tsm)|con_rgb|gs  |dif  |mask22|flips" $O/gpu_check_bf16.log | head -30
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --layers > $O/bench_f16.json 2> $O/bench_f16.err; tail -c 1500 $O/bench_f16.json
# synccheck triage of the round-1 attention kernel (ADVICE medium #1)
for v in default NO_FUSE_W NO_TMA_STORE; do
  case $v in default) E="";; NO_FUSE_W) E="BSR_NO_FUSE_W=1";; NO_TMA_STORE) E="BSR_NO_TMA_STORE=1";; esac
  env $E timeout 600 compute-sanitizer --tool synccheck --print-limit 5 python tools/profile_forward.py 2 > $O/synccheck_$v.log 2>&1
  echo "synccheck $v: $(grep -c 'Barrier error' $O/synccheck_$v.log) barrier errors; $(grep 'ERROR SUMMARY' $O/synccheck_$v.log | head -1)" >> $O/summary.txt
done
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/profile_forward.py 1 > $O/racecheck_gsc.log 2>&1
echo "racecheck gsc: $(grep 'RACECHECK SUMMARY' $O/racecheck_gsc.log)" >> $O/summary.txt
cat $O/summary.txt
