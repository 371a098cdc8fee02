#!/bin/bash
# 16-bit ShareLayer kernels (tap records, 8 lanes per cell, packed blends): tests, TSM parity, A/B (BSR_SHARE_V1=1 = previous kernels)
mkdir -p gpurun_out/r3d; O=gpurun_out/r3d
timeout 900 python -m pytest tests/test_gpu_real_files.py -m gpu -q -x -k share > $O/pytest_share.log 2>&1; echo "pytest share rc=$?" > $O/summary.txt
tail -5 $O/pytest_share.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tsm or oracle or chunks or launch" > $O/pytest_tsm.log 2>&1; echo "pytest tsm rc=$?" >> $O/summary.txt
tail -5 $O/pytest_tsm.log
for v in v1 new v1 new; do
  if [ $v = v1 ]; then export BSR_SHARE_V1=1; else unset BSR_SHARE_V1; fi
  timeout 300 python bench.py --variant tsm --frame 2 --steps 8 --warmup 3 --no-cpu-baseline --no-extras --layers > $O/bench_$v.json 2> $O/bench_$v.err
  echo "== $v frame 2"; grep -E "^(share|hole)" $O/bench_$v.err | sort -u | cut -c1-100; python tools/bench_pick.py $v < $O/bench_$v.json
  timeout 300 python bench.py --variant tsm --frame 10 --steps 8 --warmup 3 --no-cpu-baseline --no-extras --layers > $O/bench10_$v.json 2> $O/bench10_$v.err
  echo "== $v frame 10"; grep -E "^(share)" $O/bench10_$v.err | sort -u | cut -c1-100; python tools/bench_pick.py $v < $O/bench10_$v.json
done
unset BSR_SHARE_V1
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/ncu_launch_list_tsm2_mb128.csv python tools/profile_forward.py 128 tsm > /dev/null 2>&1
grep -E "share" $O/ncu_launch_list_tsm2_mb128.csv | cut -c1-260 | head -12
cat $O/summary.txt
