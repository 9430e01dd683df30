#!/usr/bin/env bash
# round 2, call T (2 GPUs): one process per GPU (CUDA IPC) with the fused split rows -- bench line at 2 GPUs with and without,
# RMSE against the 1-GPU line, the bit-for-bit checks of group / doALS
set -x
OUT=gpurun_out/r2t
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
cut -c1-400 $OUT/bench_2gpu.json
CUMF_FUSED_SPLIT=0 timeout 600 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > $OUT/bench_2gpu_nofused.json 2> $OUT/bench_2gpu_nofused.err
cut -c1-400 $OUT/bench_2gpu_nofused.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err
timeout 600 python tools/multi_gpu_check.py 2 > $OUT/multi_check_2.log 2>&1
grep -E "^\[|returned|error" $OUT/multi_check_2.log | cut -c1-250
python - <<'PY'
import json
for n in ("bench_2gpu","bench_2gpu_nofused","bench_1gpu"):
    d=json.loads(open(f"gpurun_out/r2t/{n}.json").read().strip().splitlines()[-1]); print(n, d["value"], d["ms_per_step"], d["test_rmse"], d["train_rmse"], d.get("rank_kernel_ms"), d["gpu_launches"])
PY
tail -n 3 $OUT/*.err
