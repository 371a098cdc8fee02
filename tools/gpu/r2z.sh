#!/bin/bash
# round 2, GPU call Z: skip the all-padding K steps of the last K block (cin 257 -> 4.25 blocks instead of 5): parity, A/B
mkdir -p gpurun_out/r2z; O=gpurun_out/r2z
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_real_files.py -m gpu -q -x > $O/pytest_parity.log 2>&1; echo "pytest parity rc=$?" > $O/summary.txt
grep -E "passed|failed|FAILED|Error" $O/pytest_parity.log | tail -8
for v in prev new prev new; do
  if [ $v = prev ]; then export BSR_LIB=$PWD/blindshadowremoval_b200/libbsr_prev.so; else unset BSR_LIB; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --layers > $O/bench_$v.json 2> $O/bench_$v.err
  echo "== $v"; grep -E "clr_conv1|clr_up3|^conv1|res1.qkv|res1.conv1|^up1|clr_up1|heads" $O/bench_$v.err | cut -c1-100; python tools/bench_pick.py $v < $O/bench_$v.json
done
unset BSR_LIB
BSR_LIB=$PWD/blindshadowremoval_b200/libbsr_timers.so MB=256 timeout 300 python tools/role_timers.py > $O/role_timers_mb256.txt 2>&1
grep -A1 -E "^ +(7|10|22|46) " $O/role_timers_mb256.txt | cut -c1-200
cat $O/summary.txt
