#!/bin/bash
# round 2, GPU call R: zero-initialised accumulators (merged first-tap MMAs) for conv1 / heads
mkdir -p gpurun_out/r2r; O=gpurun_out/r2r
timeout 300 python tools/halo_probe.py > $O/probe.log 2>&1; tail -6 $O/probe.log
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|Error" $O/pytest_all.log | tail -8
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --layers > $O/bench.json 2> $O/bench.err; head -12 $O/bench.err; python tools/bench_pick.py r2r < $O/bench.json
BSR_LIB=$PWD/blindshadowremoval_b200/libbsr_timers.so MB=256 timeout 300 python tools/role_timers.py > $O/role_timers_mb256.txt 2>&1
grep -A1 -E "^ +(1|2|24|25|47|48) " $O/role_timers_mb256.txt
cat $O/summary.txt
