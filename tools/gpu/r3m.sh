#!/bin/bash
# tiled chunk split: bit-exact chunk tests (3 layouts) + real-file configs + kernel time
mkdir -p gpurun_out/r3m; O=gpurun_out/r3m
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_real_files.py -m gpu -q -x -k "chunk or config or evaluate" > $O/pytest_chunk.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_chunk.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/ncu_glue.csv python tools/profile_glue.py 32 > $O/glue.log 2>&1
python tools/ncu_tsm_table.py $O/ncu_glue.csv 2>&1 | grep -v -E "conv_tc|attention_fa|conv3x3|convt_halo" > $O/ncu_glue_per_launch.txt; grep -E "unpack|caller|composite" $O/ncu_glue_per_launch.txt | cut -c1-120
