#!/usr/bin/env bash
# round 2, call X: both bench arms on one box with the final bench.py, and where the slow cudaFree of some boxes comes from
# (third destroy of the process, or the iteration count?)
set -x
OUT=gpurun_out/r2x
mkdir -p $OUT
E2E_ITERS=10,10,10,3,10 CUMF_DEBUG=1 timeout 300 python tools/e2e_phases.py > $OUT/e2e_phases_10_10_10_3_10.log 2>&1
grep -E "release|wall" $OUT/e2e_phases_10_10_10_3_10.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-200 $OUT/bench_ref.json
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; cut -c1-300 $OUT/bench_ours.json
python - <<'PY'
import json
for n in ("bench_ours","bench_ref"):
    d=json.loads(open(f"gpurun_out/r2x/{n}.json").read().strip().splitlines()[-1]); print(n, d["value"], d["e2e"]["value"], d["e2e"].get("wall_s"))
PY
