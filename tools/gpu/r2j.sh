#!/bin/bash
# round 2, GPU call J: halo-tile transposed convs (up3 / clr_up3): descriptor probe, tests, bench
mkdir -p gpurun_out/r2j; O=gpurun_out/r2j
timeout 300 python tools/halo_probe.py > $O/probe_bo1.log 2>&1; cat $O/probe_bo1.log | tail -7

BSR_NO_HALO=1 timeout 300 python tools/halo_probe.py > $O/probe_nohalo.log 2>&1; cat $O/probe_nohalo.log | tail -7
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --layers > $O/bench.json 2> $O/bench.err; head -8 $O/bench.err; python tools/bench_pick.py halo < $O/bench.json
BSR_NO_HALO=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --layers > $O/bench_nohalo.json 2> $O/bench_nohalo.err; head -5 $O/bench_nohalo.err; python tools/bench_pick.py nohalo < $O/bench_nohalo.json
