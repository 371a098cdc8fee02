#!/bin/bash
# usage: tools/gpu/run_n.sh <name> <timeout_s> <ngpus>
name=$1; to=$2; n=$3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $to -- "NGPU=$n bash tools/gpu/$name.sh" > gpurun_out/${name}_n${n}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
