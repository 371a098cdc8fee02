#!/bin/bash
# final tree: the whole GPU test step once more (sanitizers cover the thread-per-cell hole kernel) + smoke
mkdir -p gpurun_out/r3j; O=gpurun_out/r3j
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_all.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/profile_forward.py 2 gsc > $O/memcheck_gsc.log 2>&1; grep -E "ERROR SUMMARY" $O/memcheck_gsc.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/profile_forward.py 2 gsc > $O/racecheck_gsc.log 2>&1; grep -E "RACECHECK SUMMARY" $O/racecheck_gsc.log
