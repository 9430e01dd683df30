#!/usr/bin/env bash
# round 2, call J: one device allocation per solver (arena) -- full GPU suite, same-device multi check with every shard's error,
# e2e numbers
set -x
OUT=gpurun_out/r2j
mkdir -p $OUT
timeout 600 python tools/multi_gpu_check.py 2 same > $OUT/multi_check_same.log 2>&1
grep -E "^\[|returned|error" $OUT/multi_check_same.log | cut -c1-400
timeout 1800 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -n 6 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -n 1 $OUT/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/bench_ours.json 2> $OUT/bench_ours.err; cut -c1-200 $OUT/bench_ours.json
CUMF_ARENA=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/bench_ours_noarena.json 2> $OUT/bench_ours_noarena.err
CUMF_DEBUG=1 CUMF_GPUS=1 timeout 300 python tools/e2e_phases.py > $OUT/e2e_phases.log 2>&1
python - <<'PY'
import json
for n in ("bench_ours","bench_ours_noarena"):
    d=json.loads(open(f"gpurun_out/r2j/{n}.json").read().strip().splitlines()[-1]); print(n, d["value"], d["e2e"])
PY
