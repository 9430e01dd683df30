#!/usr/bin/env bash
# round 2, call N: split-table rows padded so that every gathered 64-half box is one aligned 128-byte line (512-byte rows at
# f = 100 instead of 448) -- parity, A/B against the packed layout on one box, other ranks, Yahoo shape
set -x
OUT=gpurun_out/r2n
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_generic_f.py tests/test_gpu_parity.py tests/test_gpu_multi.py -q -m gpu -x > $OUT/pytest.log 2>&1; tail -n 5 $OUT/pytest.log
timeout 300 python tools/theta_probe.py prepare
L=$PWD/cumf_als_b200/libcumf_als_b200
PROBE_TAG=aligned timeout 200 python tools/theta_probe.py | tee $OUT/theta_probe.log
CUMF_ALS_LIB=${L}_packed.so PROBE_TAG=packed timeout 200 python tools/theta_probe.py | tail -n 1 | tee -a $OUT/theta_probe.log
PROBE_TAG=aligned_again timeout 200 python tools/theta_probe.py | tee -a $OUT/theta_probe.log
PROBE_TAG=aligned_impl2_both CUMF_TC_IMPL=2 timeout 200 python tools/theta_probe.py | tee -a $OUT/theta_probe.log
CUMF_ALS_LIB=${L}_packed.so PROBE_TAG=packed_impl2_both CUMF_TC_IMPL=2 timeout 200 python tools/theta_probe.py | tail -n 1 | tee -a $OUT/theta_probe.log
CUMF_TC2_PROF=1 CUMF_ALS_LIB=${L}_prof.so PROBE_TAG=prof timeout 200 python tools/theta_probe.py > $OUT/prof.log 2>&1; tail -n 10 $OUT/prof.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/bench_ours.json 2> $OUT/bench_ours.err; cut -c1-300 $OUT/bench_ours.json
timeout 300 python bench.py --workload netflix_f200 --steps 5 --warmup 2 --no-e2e --no-cpu > $OUT/bench_f200.json 2> $OUT/bench_f200.err; cut -c1-300 $OUT/bench_f200.json
CUMF_ALS_LIB=${L}_packed.so timeout 300 python bench.py --workload netflix_f200 --steps 5 --warmup 2 --no-e2e --no-cpu > $OUT/bench_f200_packed.json 2> $OUT/bench_f200_packed.err; cut -c1-300 $OUT/bench_f200_packed.json
timeout 300 python bench.py --workload yahoo --steps 5 --warmup 2 --no-e2e --no-cpu > $OUT/bench_yahoo.json 2> $OUT/bench_yahoo.err; cut -c1-300 $OUT/bench_yahoo.json
CUMF_ALS_LIB=${L}_packed.so timeout 300 python bench.py --workload yahoo --steps 5 --warmup 2 --no-e2e --no-cpu > $OUT/bench_yahoo_packed.json 2> $OUT/bench_yahoo_packed.err; cut -c1-300 $OUT/bench_yahoo_packed.json
timeout 300 python bench.py --workload ml10m --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_ml10m.json 2> $OUT/bench_ml10m.err; cut -c1-300 $OUT/bench_ml10m.json
