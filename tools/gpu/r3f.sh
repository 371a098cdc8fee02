#!/bin/bash
# role timers incl. the single-pass attention kernel
mkdir -p gpurun_out/r3f; O=gpurun_out/r3f
BSR_LIB=$PWD/blindshadowremoval_b200/libbsr_timers.so MB=256 timeout 300 python tools/role_timers.py > $O/role_timers_mb256.txt 2>&1
grep -E "attn" $O/role_timers_mb256.txt | cut -c1-400
tail -3 $O/role_timers_mb256.txt | cut -c1-200
