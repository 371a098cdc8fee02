#!/bin/bash
# round 2, GPU call V: rotated chunk stores (up3 / clr_up3 / qkv): parity, bench, L1 data-pipe metrics per launch
mkdir -p gpurun_out/r2v; O=gpurun_out/r2v
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/pytest_parity.log 2>&1; echo "pytest parity rc=$?" > $O/summary.txt
grep -E "passed|failed|FAILED|Error" $O/pytest_parity.log | tail -8
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --layers > $O/bench.json 2> $O/bench.err; head -20 $O/bench.err; python tools/bench_pick.py r2v < $O/bench.json
ncu --profile-from-start off --metrics l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_tc_wavefronts_mem_shared.sum,sm__cycles_elapsed.avg,gpu__time_duration.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum --clock-control none --csv --log-file $O/ncu_datapipe_mb256.csv python tools/profile_forward.py 256 > /dev/null 2>&1
python tools/datapipe_table.py $O/ncu_datapipe_mb256.csv | tee $O/datapipe_mb256.txt
cat $O/summary.txt
