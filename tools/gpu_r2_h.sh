#!/usr/bin/env bash
# round 2, call H (8 GPUs): scaling of the library's multi-GPU path, Yahoo shape E1 / E2, Hugewiki-scale synthetic
set -x
OUT=gpurun_out/r2h
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.err
cut -c1-600 $OUT/bench_8gpu.json
timeout 900 python tools/hugewiki_bench.py 8 1.0 3 > $OUT/hugewiki_8gpu.json 2> $OUT/hugewiki_8gpu.err
cat $OUT/hugewiki_8gpu.json
timeout 600 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e > $OUT/bench_4gpu.json 2> $OUT/bench_4gpu.err
timeout 600 python tools/multi_gpu_check.py 8 > $OUT/multi_check_8.log 2>&1
grep -E "^\[|returned|error" $OUT/multi_check_8.log
timeout 900 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --workload yahoo --steps 10 --warmup 3 --no-e2e > $OUT/bench_8gpu_yahoo_rows.json 2> $OUT/bench_8gpu_yahoo_rows.err
timeout 1200 $TR --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --workload yahoo --steps 2 --warmup 1 --sharding partial-gram --e2-overlap --free-sms 16 --no-e2e > $OUT/bench_8gpu_yahoo_e2.json 2> $OUT/bench_8gpu_yahoo_e2.err
cat $OUT/*.json | cut -c1-700
tail -n 5 $OUT/*.err
