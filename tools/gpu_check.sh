#!/usr/bin/env bash
# tools/gpu_check.sh -- one GPU-box round trip: fused-kernel bring-up parity, doALS phase times, the default bench line.
# Logs to gpurun_out/check/; nothing here is a bench value of record.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/check
mkdir -p "$OUT"
timeout -s KILL 150 python tools/tc_bringup.py direct > "$OUT/bringup.log" 2>&1; echo "bring-up exit $?"; tail -n 1 "$OUT/bringup.log"
timeout -s KILL 150 python tools/tc_bringup.py stress > "$OUT/stress.log" 2>&1; tail -n 1 "$OUT/stress.log"
timeout -s KILL 300 python tools/e2e_phases.py > "$OUT/e2e_phases.log" 2>&1; grep -E "setup|doALS wall|RMSE run|update X run|update theta run|download" "$OUT/e2e_phases.log" | head -n 22 | tail -n 14; grep -E "doALS wall|download" "$OUT/e2e_phases.log" | tail -n 3
timeout -s KILL 420 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > "$OUT/pytest_parity.log" 2>&1; tail -n 3 "$OUT/pytest_parity.log"
timeout -s KILL 600 python bench.py --no-cpu > "$OUT/bench.json" 2> "$OUT/bench.err"; cat "$OUT/bench.json"; tail -n 3 "$OUT/bench.err"
echo "== done"
