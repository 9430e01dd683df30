#!/usr/bin/env bash
# round 2, call K: same-device groups with the event barrier (debug pass included), release-phase breakdown of doALS, and the
# theta-side bound probe (which of gather / MMA / solver sets the time of the short-row launch)
set -x
OUT=gpurun_out/r2k
mkdir -p $OUT
timeout 600 python tools/multi_gpu_check.py 2 same > $OUT/multi_check_same.log 2>&1
grep -E "^\[|returned|error" $OUT/multi_check_same.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_hugewiki_replica.py -q -m gpu > $OUT/pytest_multi.log 2>&1; tail -n 4 $OUT/pytest_multi.log
CUMF_DEBUG=1 timeout 300 python tools/e2e_phases.py > $OUT/e2e_phases.log 2>&1
grep -E "release|wall" $OUT/e2e_phases.log
timeout 300 python tools/theta_probe.py prepare
PROBE_TAG=shipped timeout 200 python tools/theta_probe.py | tee $OUT/theta_probe.log
PROBE_TAG=shipped_cg0 PROBE_CG=0 timeout 200 python tools/theta_probe.py | tee -a $OUT/theta_probe.log
PROBE_TAG=shipped_impl2_both CUMF_TC_IMPL=2 timeout 200 python tools/theta_probe.py | tee -a $OUT/theta_probe.log
PROBE_TAG=shipped_hi_only CUMF_TT_FP16=1 timeout 200 python tools/theta_probe.py | tee -a $OUT/theta_probe.log
for v in NOMMA NOGATHER GATHER1 NOSOLVE NOMMANOSOLVE NOGATHERNOSOLVE NOGATHERNOMMA NOGATHERNOMMANOSOLVE; do
  CUMF_ALS_LIB=$PWD/cumf_als_b200/libcumf_als_b200_exp_$v.so PROBE_TAG=exp_$v timeout 200 python tools/theta_probe.py 2>&1 | tail -n 2 | tee -a $OUT/theta_probe.log
done
