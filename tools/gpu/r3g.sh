#!/bin/bash
# host-path chunk size sweep: multiples of num_sms / 4 = 37 images fill the 148 SMs of the 32x32-resolution kernels exactly
mkdir -p gpurun_out/r3g; O=gpurun_out/r3g
for c in 0 37 64 74 111 128 0 74; do
  if [ $c = 0 ]; then unset BSR_HOST_CHUNK; else export BSR_HOST_CHUNK=$c; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_c$c.json 2> $O/bench_c$c.err
  python tools/bench_pick.py chunk$c < $O/bench_c$c.json
done
