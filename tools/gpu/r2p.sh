#!/bin/bash
# round 2, multi-GPU call: bench.py under torchrun on all GPUs of the box (N = $NGPU), as the driver launches it
N=${NGPU:-2}
mkdir -p gpurun_out/r2p; O=gpurun_out/r2p
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > $O/gpus_$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "rc=$?"; tail -3 $O/bench_n$N.err | cut -c1-300
python - $O/bench_n$N.json <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ('n_gpus','value','ms_per_step','e2e','e2e_compact','strong_scaling','tsm','config3_sfw_eval','clocks'):
    print(k, d.get(k))
PY
