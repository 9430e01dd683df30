#!/usr/bin/env bash
# round 2, call F: generic kernel with the issuer-side rating drop -- parity + A/B numbers; replica / multi tests after fixes
set -x
OUT=gpurun_out/r2f
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_generic_f.py tests/test_gpu_multi.py tests/test_gpu_hugewiki_replica.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu > $OUT/pytest_gpu.log 2>&1
tail -n 12 $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_default.json 2> $OUT/bench_default.err
CUMF_TC_IMPL=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_v2.json 2> $OUT/bench_v2.err
CUMF_TC_IMPL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_v1.json 2> $OUT/bench_v1.err
timeout 300 python bench.py --workload netflix_f200 --steps 5 --warmup 2 --no-e2e --no-cpu > $OUT/bench_f200.json 2> $OUT/bench_f200.err
timeout 300 python bench.py --workload ml10m --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_ml10m.json 2> $OUT/bench_ml10m.err
timeout 300 python bench.py --workload yahoo --steps 5 --warmup 2 --no-e2e --no-cpu > $OUT/bench_yahoo.json 2> $OUT/bench_yahoo.err
CUMF_TC_IMPL=2 timeout 300 python tools/cg_share.py > $OUT/cg_share_v2.log 2>&1
CUMF_TT_FP16=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_tt_fp16.json 2> $OUT/bench_tt_fp16.err
cat $OUT/*.json $OUT/cg_share_v2.log
tail -n 4 $OUT/*.err
