#!/usr/bin/env bash
# tools/gpu_sweep.sh -- knob sweep of the resident bench (no e2e / cpu legs); logs to gpurun_out/sweep/
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/sweep
mkdir -p "$OUT"
for rc in ${SWEEP_VALUES:-0 32 64 96}; do
  CUMF_TC_ROW_COST=$rc timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > "$OUT/rowcost_$rc.json" 2> "$OUT/rowcost_$rc.err"
  python - "$OUT/rowcost_$rc.json" $rc <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print("row cost", sys.argv[2], "it/s %.2f x %.2f theta %.2f" % (d["value"], d["roofline"]["gram_x_ms"], d["roofline"]["gram_theta_ms"]))
PY
done
