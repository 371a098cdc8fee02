#!/bin/bash
# round 2, GPU call D: post-processing on device, large-logit tests, TSM per-layer timing (vectorised ShareLayer)
mkdir -p gpurun_out/r2d; O=gpurun_out/r2d
timeout 900 python -m pytest tests/test_postprocess.py tests/test_gpu_real_files.py tests/test_gpu_parity.py -m gpu -q -s -k "postprocess or real or share or config or large_logit" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|Error|assert|large-logit|config|post-processing" $O/pytest.log | tail -40
timeout 300 python bench.py --variant tsm --frame 2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --layers > $O/bench_tsm2.json 2> $O/bench_tsm2.err
grep -E "share_layer|attention|res_tail|hole|assemble" $O/bench_tsm2.err | head; python tools/bench_pick.py tsm2 < $O/bench_tsm2.json
timeout 300 python bench.py --variant tsm --frame 10 --batch 250 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --layers > $O/bench_tsm10.json 2> $O/bench_tsm10.err
grep -E "share_layer" $O/bench_tsm10.err | head -3; python tools/bench_pick.py tsm10 < $O/bench_tsm10.json
cat $O/summary.txt
