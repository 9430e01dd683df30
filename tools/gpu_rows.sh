#!/usr/bin/env bash
# tools/gpu_rows.sh -- 32- vs 64-rating direct stages: bring-up parity, stress, parity tests, resident bench for both
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/rows
mkdir -p "$OUT"
timeout -s KILL 150 python tools/tc_bringup.py direct > "$OUT/bringup.log" 2>&1; echo "bring-up exit $?"; tail -n 1 "$OUT/bringup.log"
for rows in 32 64; do
  CUMF_TC_STAGE_ROWS=$rows timeout -s KILL 150 python tools/tc_bringup.py stress > "$OUT/stress_$rows.log" 2>&1; tail -n 1 "$OUT/stress_$rows.log"
  CUMF_TC_STAGE_ROWS=$rows CUMF_TC_CTAS=3 timeout -s KILL 150 python tools/tc_bringup.py stress >> "$OUT/stress_$rows.log" 2>&1; tail -n 1 "$OUT/stress_$rows.log"
  CUMF_TC_STAGE_ROWS=$rows timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > "$OUT/bench_$rows.json" 2> "$OUT/bench_$rows.err"
  python - "$OUT/bench_$rows.json" $rows <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print("stage rows", sys.argv[2], "it/s %.2f x %.2f theta %.2f rmse %.6f" % (d["value"], d["roofline"]["gram_x_ms"], d["roofline"]["gram_theta_ms"], d["test_rmse"]))
except Exception as e:
    print("stage rows", sys.argv[2], "FAILED", e); print(open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
CUMF_TC_STAGE_ROWS=64 timeout -s KILL 420 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > "$OUT/pytest_64.log" 2>&1; tail -n 3 "$OUT/pytest_64.log"
