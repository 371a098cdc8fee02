timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --layers > gpurun_out/bench_r1c_layers.json 2> gpurun_out/bench_r1c.err; tail -c 600 gpurun_out/bench_r1c_layers.json | head -c 100; echo
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --micro-batch 256 2>/dev/null | tail -1 | python tools/bench_pick.py mb256
timeout 200 python bench.py --variant tsm --frame 2 --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python tools/bench_pick.py tsm2
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1c_ncu_launch_list_mb128.csv python tools/profile_forward.py 128 > /dev/null 2>&1
timeout 1000 ncu --profile-from-start off --set full --clock-control none -o /tmp/prof_all -f python tools/profile_forward.py 128 > gpurun_out/prof_full_c.log 2>&1
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw_c.csv 2>/dev/null
for L in 11 24; do timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -s $L -c 1 -o gpurun_out/prof_c_l$L -f python tools/profile_forward.py 128 > /dev/null 2>&1; done
ls -la gpurun_out/ | grep -E "prof_c|r1c|raw_c"
