#!/bin/bash
# synccheck triage of the halo kernels: old layout (barriers behind the TMA tiles), old + A loads ablated, new layout (barriers first)
mkdir -p gpurun_out/r3b; O=gpurun_out/r3b
run() { name=$1; shift; env "$@" timeout 600 compute-sanitizer --tool synccheck --print-limit 1 python tools/profile_forward.py 2 gsc > $O/sc_$name.log 2>&1; echo "== $name: $(grep -E 'ERROR SUMMARY' $O/sc_$name.log | head -1) $(grep -E 'launches per' $O/sc_$name.log) $(grep -m1 -E ' in [a-z0-9_]+\.cuh:[0-9]+' $O/sc_$name.log | grep -v mbar_try)"; }
run old BSR_LIB=$PWD/blindshadowremoval_b200/libbsr_prev.so
run old_noA BSR_LIB=$PWD/blindshadowremoval_b200/libbsr_prev.so BSR_ABLATE=4
run new X=1
run new_tsm X=1
grep -B2 -A8 "Barrier error" $O/sc_new.log | head -40 | cut -c1-200
