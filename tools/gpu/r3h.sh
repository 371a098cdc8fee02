#!/bin/bash
# SM-aligned host chunks as the default: host-path tests + bench (e2e, e2e_compact), TSM host path
mkdir -p gpurun_out/r3h; O=gpurun_out/r3h
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sanitizer.py -m gpu -q -x -k "host or compact or eager or bench or racecheck or chunk" > $O/pytest_host.log 2>&1; echo "pytest host rc=$?"
tail -3 $O/pytest_host.log
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_$i.json 2> $O/bench_$i.err; python tools/bench_pick.py run$i < $O/bench_$i.json
done
timeout 300 python bench.py --variant tsm --frame 2 --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_tsm.json 2> $O/bench_tsm.err; python tools/bench_pick.py tsm2 < $O/bench_tsm.json
