#!/usr/bin/env bash
# round 2, call O (8 GPUs): bench line at 8 GPUs with the aligned-table kernel (resident + e2e through doALS, CUMF_GPUS=8), the
# phase lines of that doALS call, and the bit-for-bit check of group / doALS against one GPU
set -x
OUT=gpurun_out/r2o
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.err
cut -c1-700 $OUT/bench_8gpu.json
CUMF_DEBUG=1 CUMF_GPUS=8 timeout 300 python tools/e2e_phases.py > $OUT/e2e_phases_8gpu.log 2>&1
grep -E "shard|setup|release|wall|download" $OUT/e2e_phases_8gpu.log | tail -n 40
timeout 600 python tools/multi_gpu_check.py 8 > $OUT/multi_check_8.log 2>&1
grep -E "^\[|returned|error" $OUT/multi_check_8.log
timeout 600 $TR --nproc-per-node 8 --master-port 29533 bench.py --gpus 8 --workload yahoo --steps 10 --warmup 3 --no-e2e > $OUT/bench_8gpu_yahoo_rows.json 2> $OUT/bench_8gpu_yahoo_rows.err
cut -c1-400 $OUT/bench_8gpu_yahoo_rows.json
tail -n 3 $OUT/*.err
