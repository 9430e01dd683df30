#!/usr/bin/env bash
# round 2, call Y: arena estimate that includes the split-row scratch (no allocation of a solver outside its arena but the two
# replicas) -- phase lines of five doALS calls, group / parity tests, bench line
set -x
OUT=gpurun_out/r2y
mkdir -p $OUT
E2E_ITERS=10,10,10,10,10 CUMF_DEBUG=1 timeout 300 python tools/e2e_phases.py > $OUT/e2e_phases_5x10.log 2>&1
grep -E "arena|release|wall" $OUT/e2e_phases_5x10.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_hugewiki_replica.py tests/test_gpu_parity.py tests/test_gpu_generic_f.py -q -m gpu > $OUT/pytest.log 2>&1; tail -n 4 $OUT/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/bench_ours.json 2> $OUT/bench_ours.err
timeout 600 python bench.py --workload yahoo --steps 5 --warmup 3 --no-cpu > $OUT/bench_yahoo.json 2> $OUT/bench_yahoo.err
python - <<'PY'
import json
for n in ("bench_ours","bench_yahoo"):
    d=json.loads(open(f"gpurun_out/r2y/{n}.json").read().strip().splitlines()[-1]); print(n, d["value"], d["e2e"]["value"], d["e2e"].get("wall_s"), d["clocks"])
PY
