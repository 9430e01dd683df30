#!/usr/bin/env bash
# round 2, call P: group shards with replicas + flags inside the arena (same-device tests), full GPU suite, two last A/B builds
# of the short-row launch (160-register solvers, MMA order), bench line with the out-of-process clock sampler
set -x
OUT=gpurun_out/r2p
mkdir -p $OUT
timeout 1800 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -n 5 $OUT/pytest_gpu.log
timeout 300 python tools/theta_probe.py prepare
L=$PWD/cumf_als_b200/libcumf_als_b200
PROBE_TAG=shipped timeout 200 python tools/theta_probe.py | tee $OUT/theta_probe.log
CUMF_ALS_LIB=${L}_regs160.so PROBE_TAG=regs_32_160 timeout 200 python tools/theta_probe.py | tail -n 1 | tee -a $OUT/theta_probe.log
CUMF_ALS_LIB=${L}_orderb.so PROBE_TAG=mma_order_b timeout 200 python tools/theta_probe.py | tail -n 1 | tee -a $OUT/theta_probe.log
PROBE_TAG=shipped_again timeout 200 python tools/theta_probe.py | tee -a $OUT/theta_probe.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/bench_ours.json 2> $OUT/bench_ours.err; cut -c1-300 $OUT/bench_ours.json; grep -o '"clocks": {[^}]*}' $OUT/bench_ours.json
timeout 300 python bench.py --workload ml10m --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_ml10m.json 2> $OUT/bench_ml10m.err; cut -c1-200 $OUT/bench_ml10m.json; grep -o '"clocks": {[^}]*}' $OUT/bench_ml10m.json
CUMF_DEBUG=1 CUMF_GPUS=2 CUMF_GROUP_SAME_DEVICE=1 timeout 300 python tools/e2e_phases.py 0.25 > $OUT/e2e_phases_same2.log 2>&1; grep -E "shard|release|wall" $OUT/e2e_phases_same2.log | tail -n 12
