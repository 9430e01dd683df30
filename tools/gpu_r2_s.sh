#!/usr/bin/env bash
# round 2, call S: theta half-step keeps the X-side gather table current (fused split rows) -- tests on one GPU (single solver and
# same-device groups), bench line, doALS phases
set -x
OUT=gpurun_out/r2s
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_hugewiki_replica.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_contract_sizes.py -q -m gpu -s > $OUT/pytest.log 2>&1; tail -n 6 $OUT/pytest.log; grep -E "launches:|reference twice" $OUT/pytest.log
timeout 600 python tools/multi_gpu_check.py 2 same > $OUT/multi_check_same.log 2>&1
grep -E "^\[|returned|error" $OUT/multi_check_same.log | cut -c1-250
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/bench_ours.json 2> $OUT/bench_ours.err; cut -c1-300 $OUT/bench_ours.json
CUMF_FUSED_SPLIT=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > $OUT/bench_nofused.json 2> $OUT/bench_nofused.err; cut -c1-300 $OUT/bench_nofused.json
