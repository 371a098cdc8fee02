#!/bin/bash
# last bench lines of the round on the final tree (1 GPU, with the per-layer table) + reference arm + GSC launch list
mkdir -p gpurun_out/r3z; O=gpurun_out/r3z
timeout 900 python bench.py --steps 10 --warmup 3 --layers > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>/dev/null
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ncu_launch_list_gsc_mb256.csv python tools/profile_forward.py 256 > /dev/null 2>&1
python tools/bench_pick.py final < $O/bench_final.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3z/bench_final.json').read().strip().splitlines()[-1])
print(d['roofline']['network'], d['roofline']['frac'], d['tsm'], d['clocks'])
PY
