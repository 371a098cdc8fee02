#!/bin/bash
# round 2, GPU call T: role timers + per-launch event times (effective clock), short and sustained
mkdir -p gpurun_out/r2t; O=gpurun_out/r2t
for reps in 3 60; do
  BSR_LIB=$PWD/blindshadowremoval_b200/libbsr_timers.so MB=256 REPS=$reps timeout 300 python tools/role_timers.py > $O/role_timers_reps$reps.txt 2>&1
  echo "== reps $reps"; grep -A1 -E "^ +(1|2|8|10|24|25|47|48) " $O/role_timers_reps$reps.txt | cut -c1-230
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_throttle_reasons.active --format=csv
