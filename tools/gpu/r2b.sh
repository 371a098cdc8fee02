#!/bin/bash
# round 2, GPU call B: f16 storage (idesc fixed): full GPU test-suite, error table, bench with per-layer table
mkdir -p gpurun_out/r2b; O=gpurun_out/r2b
timeout 1200 python -m pytest tests -m gpu -q -s > $O/pytest_f16.log 2>&1; echo "pytest f16 rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|worst|gain|large-logit|flips" $O/pytest_f16.log | tail -40
timeout 300 python tests/gpu_check.py tc > $O/gpu_check_f16.log 2>&1
grep -E "^(gsc|tsm)|con_rgb|gs  |dif  |mask22|flips" $O/gpu_check_f16.log | head -30
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --layers > $O/bench_f16.json 2> $O/bench_f16.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b/bench_f16.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','roofline') if k in d})
PY
cat $O/summary.txt
