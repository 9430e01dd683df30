#!/usr/bin/env bash
# round 2, call M: solver chain changes (8 partial sums, batched drain, chunk head prefetch, descriptors two refills ahead) on
# the patching kernel -- parity, A/B on one box, per-step cycle counters of the in-kernel CG
set -x
OUT=gpurun_out/r2m
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_generic_f.py tests/test_gpu_parity.py -q -m gpu -x > $OUT/pytest.log 2>&1; tail -n 5 $OUT/pytest.log
timeout 300 python tools/theta_probe.py prepare
L=$PWD/cumf_als_b200/libcumf_als_b200
PROBE_TAG=shipped timeout 200 python tools/theta_probe.py | tee $OUT/theta_probe.log
CUMF_ALS_LIB=${L}_spmv4.so PROBE_TAG=spmv4 timeout 200 python tools/theta_probe.py | tail -n 1 | tee -a $OUT/theta_probe.log
CUMF_ALS_LIB=${L}_drainserial.so PROBE_TAG=drain_serial timeout 200 python tools/theta_probe.py | tail -n 1 | tee -a $OUT/theta_probe.log
PROBE_TAG=shipped_again timeout 200 python tools/theta_probe.py | tee -a $OUT/theta_probe.log
PROBE_TAG=shipped_cg0 PROBE_CG=0 timeout 200 python tools/theta_probe.py | tee -a $OUT/theta_probe.log
CUMF_TC2_PROF=1 CUMF_ALS_LIB=${L}_prof.so PROBE_TAG=prof timeout 200 python tools/theta_probe.py > $OUT/prof.log 2>&1; tail -n 13 $OUT/prof.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/bench_ours.json 2> $OUT/bench_ours.err; cut -c1-300 $OUT/bench_ours.json
