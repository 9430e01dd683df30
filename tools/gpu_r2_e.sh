#!/usr/bin/env bash
# round 2, call E (2 GPUs): the library's multi-GPU paths on real peers -- group, doALS under CUMF_GPUS, torchrun bench through
# CUDA IPC, partial-Gram (E2) with and without overlap
set -x
OUT=gpurun_out/r2e
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 600 python tools/multi_gpu_check.py 2 > $OUT/multi_check.log 2>&1
grep -E "^\[|returned|error" $OUT/multi_check.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/bench_2gpu_rows.json 2> $OUT/bench_2gpu_rows.err
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --sharding partial-gram --no-e2e > $OUT/bench_2gpu_e2.json 2> $OUT/bench_2gpu_e2.err
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --sharding partial-gram --e2-overlap --e2-batch-mb 128 --free-sms 16 --no-e2e > $OUT/bench_2gpu_e2_overlap.json 2> $OUT/bench_2gpu_e2_overlap.err
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --sharding partial-gram --e2-batch-mb 128 --no-e2e > $OUT/bench_2gpu_e2_batched.json 2> $OUT/bench_2gpu_e2_batched.err
timeout 600 $TR bench.py --gpus 2 --workload yahoo --steps 5 --warmup 2 --no-e2e > $OUT/bench_2gpu_yahoo_rows.json 2> $OUT/bench_2gpu_yahoo_rows.err
timeout 900 $TR bench.py --gpus 2 --workload yahoo --steps 3 --warmup 1 --sharding partial-gram --e2-overlap --free-sms 16 --no-e2e > $OUT/bench_2gpu_yahoo_e2.json 2> $OUT/bench_2gpu_yahoo_e2.err
timeout 600 python tools/hugewiki_bench.py 2 0.125 3 > $OUT/hugewiki_2gpu_8th.json 2> $OUT/hugewiki_2gpu_8th.err
cat $OUT/*.json
tail -n 6 $OUT/*.err
