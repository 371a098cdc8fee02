#!/bin/bash
# ncu --set full of the ShareLayer kernels only (old warp-per-cell kernels and the new thread-per-vector ones)
mkdir -p gpurun_out/r3e; O=gpurun_out/r3e
for v in v1 new; do
  if [ $v = v1 ]; then export BSR_SHARE_V1=1; else unset BSR_SHARE_V1; fi
  timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:share -o /tmp/share_$v -f python tools/profile_forward.py 128 tsm > $O/ncu_$v.log 2>&1
  ncu -i /tmp/share_$v.ncu-rep --page details --csv > $O/share_${v}_details.csv 2>/dev/null
  ncu -i /tmp/share_$v.ncu-rep --page source --csv --print-source sass > $O/share_${v}_source.csv 2>/dev/null
done
ls -la $O
