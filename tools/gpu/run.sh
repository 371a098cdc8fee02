#!/bin/bash
# usage: tools/gpu/run.sh <name> <timeout_s> [--gpus N]   -> runs tools/gpu/<name>.sh on a GPU box, retrying while the pod is busy
name=$1; to=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" -- "bash tools/gpu/$name.sh" > gpurun_out/${name}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
