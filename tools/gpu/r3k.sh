#!/bin/bash
# new parity test at the bench's default configuration (n = 256, micro-batch 256)
mkdir -p gpurun_out/r3k; O=gpurun_out/r3k
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "mb256" > $O/pytest_mb256.log 2>&1; echo "pytest rc=$?"
grep -E "worst|passed|failed|Error|assert" $O/pytest_mb256.log | head -10
