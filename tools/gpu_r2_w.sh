#!/usr/bin/env bash
# round 2, call W: compute-sanitizer on the final kernels (memcheck, synccheck, racecheck) and racecheck on the build with the
# explicit thread-to-thread handoff barrier (-DCUMF_TC2_EXPLICIT_HANDOFF), phase lines of doALS
set -x
OUT=gpurun_out/r2w
mkdir -p $OUT
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_small.py 100 10 200 > $OUT/sanitizer_$tool.log 2>&1
  grep -E "sanitize_small|SUMMARY|returned" $OUT/sanitizer_$tool.log
done
CUMF_ALS_LIB=$PWD/cumf_als_b200/libcumf_als_b200_handoff.so timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize_small.py 100 10 200 > $OUT/sanitizer_racecheck_explicit_handoff.log 2>&1
grep -E "sanitize_small|SUMMARY|returned" $OUT/sanitizer_racecheck_explicit_handoff.log
CUMF_DEBUG=1 timeout 300 python tools/e2e_phases.py > $OUT/e2e_phases.log 2>&1
grep -E "setup|release|wall|download" $OUT/e2e_phases.log
