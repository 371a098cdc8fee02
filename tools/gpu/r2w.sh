#!/bin/bash
# round 2, GPU call W: A/B on one box: rotated step order of the streamed-weight convs (BSR_NO_ROT=1 = old), + parity
mkdir -p gpurun_out/r2w; O=gpurun_out/r2w
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/pytest_parity.log 2>&1; echo "pytest parity rc=$?" > $O/summary.txt
grep -E "passed|failed|FAILED|Error" $O/pytest_parity.log | tail -8
for v in norot default norot default; do
  if [ $v = norot ]; then export BSR_NO_ROT=1; else unset BSR_NO_ROT; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --layers > $O/bench_$v.json 2> $O/bench_$v.err
  echo "== $v"; grep -E "conv1 |conv2|up1|up2|down2|down3|qkv|conv3" $O/bench_$v.err | sort | awk '{k=$1; sub(/res[0-9]\./,"res.",k); n[k]++; t[k]+=$3} END {for (k in n) printf "%s %.4f  ", k, t[k]/n[k]; print ""}'; python tools/bench_pick.py $v < $O/bench_$v.json
done
cat $O/summary.txt
