#!/bin/bash
# round 2, GPU call G: role timers of the conv layers (where do the persistent CTAs wait?), graph replay test, config runs
mkdir -p gpurun_out/r2g; O=gpurun_out/r2g
BSR_LIB=$PWD/blindshadowremoval_b200/libbsr_timers.so MB=256 timeout 300 python tools/role_timers.py > $O/role_timers_mb256.txt 2>&1
grep -E "conv1|down1|up3|heads|clr_up3|clr_conv1|clr_up2|up2|qkv|launch" $O/role_timers_mb256.txt | grep -v "r[1-5]\." | head -40
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "device_error_flag or benchmark_configuration_gsc or layer_times" > $O/pytest_graph.log 2>&1; echo "pytest graph rc=$?" >> $O/summary.txt
tail -5 $O/pytest_graph.log
timeout 600 python tools/config_runs.py > $O/config_runs.txt 2>&1; cat $O/config_runs.txt
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_graph.json 2> $O/bench_graph.err; python tools/bench_pick.py graphs < $O/bench_graph.json
BSR_NO_GRAPH=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_nograph.json 2> $O/bench_nograph.err; python tools/bench_pick.py nograph < $O/bench_nograph.json
cat $O/summary.txt
