#!/bin/bash
# ncu --set full of the new ShareLayer kernels
mkdir -p gpurun_out/r3e; O=gpurun_out/r3e
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:share -o /tmp/share_new2 -f python tools/profile_forward.py 128 tsm > $O/ncu_new2.log 2>&1
ncu -i /tmp/share_new2.ncu-rep --page details --csv > $O/share_new2_details.csv 2>/dev/null
ncu -i /tmp/share_new2.ncu-rep --page raw --csv > $O/share_new2_raw.csv 2>/dev/null
ls -la $O | tail -4
