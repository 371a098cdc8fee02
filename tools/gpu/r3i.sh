#!/bin/bash
# thread-per-cell in-place hole kernel: parity + per-layer time
mkdir -p gpurun_out/r3i; O=gpurun_out/r3i
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_real_files.py -m gpu -q -x > $O/pytest_parity.log 2>&1; echo "pytest parity rc=$?"
tail -2 $O/pytest_parity.log
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --layers > $O/bench_$i.json 2> $O/bench_$i.err
grep -E "^(hole|pack_img|assemble)" $O/bench_$i.err | sort -u | cut -c1-100; python tools/bench_pick.py run$i < $O/bench_$i.json
done
