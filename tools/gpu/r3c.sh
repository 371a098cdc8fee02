#!/bin/bash
# in-place hole mask (GSC) + packed fp32 math in the attention tail: parity, sanitizer tests, A/B against the previous build
mkdir -p gpurun_out/r3c; O=gpurun_out/r3c
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_real_files.py tests/test_gpu_sanitizer.py -m gpu -q -x > $O/pytest_parity.log 2>&1; echo "pytest parity rc=$?" > $O/summary.txt
grep -E "passed|failed|FAILED|Error" $O/pytest_parity.log | tail -8
for v in prev new prev new; do
  if [ $v = prev ]; then export BSR_LIB=$PWD/blindshadowremoval_b200/libbsr_prev.so; else unset BSR_LIB; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --layers > $O/bench_$v.json 2> $O/bench_$v.err
  echo "== $v"; grep -E "^(hole|attention|assemble)" $O/bench_$v.err | sort -u | cut -c1-100; python tools/bench_pick.py $v < $O/bench_$v.json
done
cat $O/summary.txt
