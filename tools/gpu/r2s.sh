#!/bin/bash
# round 2, GPU call S: fp32 host path that uploads only the uv / reg rows the model reads
mkdir -p gpurun_out/r2s; O=gpurun_out/r2s
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|Error" $O/pytest_all.log | tail -8
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err; python tools/bench_pick.py rows < $O/bench.json
BSR_HOST_FULL_UV=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_full.json 2> $O/bench_full.err; python tools/bench_pick.py fulluv < $O/bench_full.json
timeout 300 python bench.py --variant tsm --frame 2 --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_tsm2.json 2> $O/bench_tsm2.err; python tools/bench_pick.py tsm2 < $O/bench_tsm2.json
cat $O/summary.txt
