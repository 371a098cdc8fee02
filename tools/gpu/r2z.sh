#!/bin/bash
# round 2, GPU call Z: qkv in two 192-column tiles: parity, A/B against the previous build
mkdir -p gpurun_out/r2z; O=gpurun_out/r2z
timeout 300 python tests/gpu_check.py tc > $O/gpu_check.log 2>&1; grep -E "^(gsc|tsm)|res0|res5|con_rgb|gs  |flips|errflag" $O/gpu_check.log | head -14
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_real_files.py -m gpu -q -x > $O/pytest_parity.log 2>&1; echo "pytest parity rc=$?" > $O/summary.txt
grep -E "passed|failed|FAILED|Error" $O/pytest_parity.log | tail -8
for v in prev new prev new; do
  if [ $v = prev ]; then export BSR_LIB=$PWD/blindshadowremoval_b200/libbsr_prev.so; else unset BSR_LIB; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --layers > $O/bench_$v.json 2> $O/bench_$v.err
  echo "== $v"; grep -E "res[15].conv3|res1.qkv|res0.conv1" $O/bench_$v.err | cut -c1-100; python tools/bench_pick.py $v < $O/bench_$v.json
done
cat $O/summary.txt
