#!/usr/bin/env bash
# round 2, call L: ratings as a second MMA operand (no patching of landed rows) -- parity for every f, A/B against the patching
# build on one box, per-role cycle counters of both, release-phase breakdown of doALS
set -x
OUT=gpurun_out/r2l
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_generic_f.py -q -m gpu -x > $OUT/pytest_generic.log 2>&1; tail -n 5 $OUT/pytest_generic.log
timeout 300 python tools/theta_probe.py prepare
L=$PWD/cumf_als_b200/libcumf_als_b200
PROBE_TAG=shipped timeout 200 python tools/theta_probe.py | tee $OUT/theta_probe.log
CUMF_ALS_LIB=${L}_old.so PROBE_TAG=old_patching timeout 200 python tools/theta_probe.py | tail -n 1 | tee -a $OUT/theta_probe.log
CUMF_ALS_LIB=${L}_spin.so PROBE_TAG=issuer_spin timeout 200 python tools/theta_probe.py | tail -n 1 | tee -a $OUT/theta_probe.log
PROBE_TAG=shipped_again timeout 200 python tools/theta_probe.py | tee -a $OUT/theta_probe.log
PROBE_TAG=shipped_cg0 PROBE_CG=0 timeout 200 python tools/theta_probe.py | tee -a $OUT/theta_probe.log
CUMF_TC2_PROF=1 CUMF_ALS_LIB=${L}_prof.so PROBE_TAG=prof timeout 200 python tools/theta_probe.py > $OUT/prof_new.log 2>&1; tail -n 12 $OUT/prof_new.log
CUMF_TC2_PROF=1 CUMF_ALS_LIB=${L}_oldprof.so PROBE_TAG=oldprof timeout 200 python tools/theta_probe.py > $OUT/prof_old.log 2>&1; tail -n 12 $OUT/prof_old.log
CUMF_DEBUG=1 timeout 300 python tools/e2e_phases.py > $OUT/e2e_phases.log 2>&1
grep -E "release|wall" $OUT/e2e_phases.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/bench_ours.json 2> $OUT/bench_ours.err; cut -c1-300 $OUT/bench_ours.json
