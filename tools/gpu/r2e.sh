#!/bin/bash
# round 2, GPU call E: attention kernel with the fused output conv + block tail: correctness, timing, sanitizers
mkdir -p gpurun_out/r2e; O=gpurun_out/r2e
timeout 300 python tests/gpu_check.py tc > $O/gpu_check.log 2>&1
grep -E "^(gsc|tsm)|errflag|res0 |res5 |con_rgb|gs  " $O/gpu_check.log | head -20
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|Error" $O/pytest_all.log | tail -8
for mb in 128 256; do
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --layers --micro-batch $mb > $O/bench_mb$mb.json 2> $O/bench_mb$mb.err
  grep -E "attention" $O/bench_mb$mb.err | head -3; python tools/bench_pick.py mb$mb < $O/bench_mb$mb.json
done
timeout 600 compute-sanitizer --tool synccheck --print-limit 5 python tools/profile_forward.py 2 > $O/synccheck.log 2>&1
echo "synccheck: $(grep -c 'Barrier error' $O/synccheck.log) barrier errors; $(grep 'ERROR SUMMARY' $O/synccheck.log | head -1)" >> $O/summary.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/profile_forward.py 2 > $O/racecheck.log 2>&1
echo "racecheck gsc: $(grep 'RACECHECK SUMMARY' $O/racecheck.log)" >> $O/summary.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python tools/profile_forward.py 2 tsm > $O/memcheck_tsm.log 2>&1
echo "memcheck tsm: $(grep 'ERROR SUMMARY' $O/memcheck_tsm.log)" >> $O/summary.txt
cat $O/summary.txt
