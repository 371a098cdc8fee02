#!/bin/bash
# memcheck + racecheck over the glue pass (chunk split through shared memory, caller glue, composite, post-processing, compact path)
mkdir -p gpurun_out/r3n; O=gpurun_out/r3n
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python tools/profile_glue.py 2 > $O/memcheck_glue.log 2>&1; grep -E "ERROR SUMMARY|glue pass" $O/memcheck_glue.log
timeout 200 compute-sanitizer --tool racecheck --print-limit 5 python tools/profile_glue.py 2 > $O/racecheck_glue.log 2>&1; grep -E "RACECHECK SUMMARY|glue pass" $O/racecheck_glue.log
