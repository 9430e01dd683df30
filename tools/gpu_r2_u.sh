#!/usr/bin/env bash
# round 2, call U (8 GPUs): the scaling line at 8 and 4 ranks with per-rank kernel times, fused split rows on / off, e2e through
# doALS (CUMF_GPUS=8) and its phase lines
set -x
OUT=gpurun_out/r2u
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.err
CUMF_FUSED_SPLIT=0 timeout 600 $TR --nproc-per-node 8 --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e > $OUT/bench_8gpu_nofused.json 2> $OUT/bench_8gpu_nofused.err
timeout 600 $TR --nproc-per-node 8 --master-port 29553 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e > $OUT/bench_8gpu_again.json 2> $OUT/bench_8gpu_again.err
timeout 600 $TR --nproc-per-node 4 --master-port 29554 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e > $OUT/bench_4gpu.json 2> $OUT/bench_4gpu.err
CUMF_DEBUG=1 CUMF_GPUS=8 timeout 300 python tools/e2e_phases.py > $OUT/e2e_phases_8gpu.log 2>&1
grep -E "shard|setup|release|wall|download" $OUT/e2e_phases_8gpu.log | tail -n 24
python - <<'PY'
import json
for n in ("bench_8gpu","bench_8gpu_nofused","bench_8gpu_again","bench_4gpu"):
    try:
        d=json.loads(open(f"gpurun_out/r2u/{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"],1), round(d["ms_per_step"],3), d["test_rmse"], (d.get("e2e") or {}).get("value"), d["clocks"].get("sm_mhz"), d["clocks"].get("samples"))
        print("   ", d.get("rank_kernel_ms"))
    except Exception as e:
        print(n, "ERR", e)
PY
tail -n 3 $OUT/*.err
