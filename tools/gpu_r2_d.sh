#!/usr/bin/env bash
# round 2, call D: multi-GPU logic on one device (group, doALS under CUMF_GPUS), Hugewiki replica parity, 32- vs 64-rating
# stages of the generic kernel, default bench line
set -x
OUT=gpurun_out/r2d
mkdir -p $OUT
timeout 600 python tools/multi_gpu_check.py 2 same > $OUT/multi_check.log 2>&1
tail -n 12 $OUT/multi_check.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_hugewiki_replica.py tests/test_gpu_generic_f.py tests/test_gpu_parity.py -q -m gpu -s > $OUT/pytest_gpu.log 2>&1
tail -n 12 $OUT/pytest_gpu.log
CUMF_TC_IMPL=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_v2_k64.json 2> $OUT/bench_v2_k64.err
CUMF_ALS_LIB=$PWD/cumf_als_b200/libcumf_als_b200_k32.so CUMF_TC_IMPL=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_v2_k32.json 2> $OUT/bench_v2_k32.err
CUMF_ALS_LIB=$PWD/cumf_als_b200/libcumf_als_b200_k32.so timeout 300 python bench.py --workload ml10m --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_ml10m_k32.json 2> $OUT/bench_ml10m_k32.err
timeout 300 python bench.py --workload ml10m --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_ml10m_k64.json 2> $OUT/bench_ml10m_k64.err
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_ours.json 2> $OUT/bench_ours.err
timeout 600 python tools/hugewiki_bench.py 1 0.0625 3 > $OUT/hugewiki_1gpu_16th.json 2> $OUT/hugewiki_1gpu_16th.err
CUMF_GROUP_SAME_DEVICE=1 timeout 600 python tools/hugewiki_bench.py 2 0.03125 2 > $OUT/hugewiki_2shards_32nd.json 2> $OUT/hugewiki_2shards_32nd.err
cat $OUT/*.json
tail -n 5 $OUT/*.err
