#!/bin/bash
# ncu --set full of the final ShareLayer kernels
mkdir -p gpurun_out/r3e; O=gpurun_out/r3e
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:share -o /tmp/share_new3 -f python tools/profile_forward.py 128 tsm > $O/ncu_new3.log 2>&1
ncu -i /tmp/share_new3.ncu-rep --page details --csv > $O/share_new3_details.csv 2>/dev/null
ncu -i /tmp/share_new3.ncu-rep --page source --csv --print-source sass --kernel-name regex:share_reduce > $O/share_new3_reduce_source.csv 2>/dev/null
ls -la $O | tail -3
