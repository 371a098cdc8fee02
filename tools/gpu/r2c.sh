#!/bin/bash
# round 2, GPU call C: single-pass attention kernel (attention_fa.cuh): correctness, timing vs the r1 kernel, sanitizers
mkdir -p gpurun_out/r2c; O=gpurun_out/r2c
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "attention_kernel_alone" > $O/pytest_attn.log 2>&1; echo "pytest attn rc=$?" >> $O/summary.txt
grep -E "passed|failed|gain" $O/pytest_attn.log | tail
timeout 1200 python -m pytest tests -m gpu -q -s > $O/pytest_all.log 2>&1; echo "pytest all rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|worst" $O/pytest_all.log | tail -12
for v in fa v1; do for mb in 128 256; do
  E=""; [ $v = v1 ] && E="BSR_ATTN_V1=1"
  env $E timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --layers --micro-batch $mb > $O/bench_${v}_mb$mb.json 2> $O/bench_${v}_mb$mb.err
  python - $O/bench_${v}_mb$mb.json $v $mb <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
L={r['layer']:r for r in d.get('layers',[])}
pick=[k for k in L if 'attention' in k or k.endswith('.w')][:3]
print(sys.argv[2], 'mb', sys.argv[3], 'img/s', d['value'], 'ms/step', d['ms_per_step'], {k:(L[k]['ms_per_launch'], L[k].get('tflops')) for k in pick})
PY
done; done
timeout 900 python tools/large_logit_probe.py > $O/large_logit_probe.log 2>&1; cat $O/large_logit_probe.log | tail -12
timeout 600 compute-sanitizer --tool synccheck --print-limit 5 python tools/profile_forward.py 2 > $O/synccheck_fa.log 2>&1
echo "synccheck fa: $(grep -c 'Barrier error' $O/synccheck_fa.log) barrier errors; $(grep 'ERROR SUMMARY' $O/synccheck_fa.log | head -1)" >> $O/summary.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/profile_forward.py 2 > $O/racecheck_fa.log 2>&1
echo "racecheck gsc fa: $(grep 'RACECHECK SUMMARY' $O/racecheck_fa.log)" >> $O/summary.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python tools/profile_forward.py 2 > $O/memcheck_fa.log 2>&1
echo "memcheck gsc fa: $(grep 'ERROR SUMMARY' $O/memcheck_fa.log)" >> $O/summary.txt
cat $O/summary.txt
