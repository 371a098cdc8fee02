#!/bin/bash
# per-kernel time and DRAM bytes of the glue / post-processing kernels outside the generator's 46 launches
mkdir -p gpurun_out/r3l; O=gpurun_out/r3l
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/ncu_glue.csv python tools/profile_glue.py 32 > $O/glue.log 2>&1
tail -3 $O/glue.log
python tools/ncu_tsm_table.py $O/ncu_glue.csv 2>&1 | grep -v -E "conv_tc|attention_fa|conv3x3|convt_halo" | cut -c1-120 | head -60
