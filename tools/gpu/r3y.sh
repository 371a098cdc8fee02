#!/bin/bash
# final-tree check: smoke, all GPU tests (incl. racecheck), bench line with --layers, reference arm
mkdir -p gpurun_out/r3y; O=gpurun_out/r3y; rm -f $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt; tail -1 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt; tail -2 $O/pytest_all.log
timeout 900 python bench.py --steps 10 --warmup 3 --layers > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?" >> $O/summary.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3y/bench_final.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','e2e_compact','strong_scaling','tsm','clocks','gpu_launches'):
    print(k, str(d.get(k))[:200])
print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('traffic',)})
PY
cat $O/summary.txt
