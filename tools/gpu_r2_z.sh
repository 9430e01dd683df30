#!/usr/bin/env bash
# round 2, call Z: confirmation of the committed state -- every GPU test, smoke(), both bench arms
set -x
OUT=gpurun_out/r2z
mkdir -p $OUT
timeout 1800 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 $OUT/smoke.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-200 $OUT/bench_ref.json
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; cut -c1-300 $OUT/bench_ours.json
python - <<'PY'
import json
for n in ("bench_ours","bench_ref"):
    d=json.loads(open(f"gpurun_out/r2z/{n}.json").read().strip().splitlines()[-1]); print(n, d["value"], d["e2e"]["value"], d["e2e"].get("wall_s"), d["clocks"])
PY
