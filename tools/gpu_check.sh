#!/usr/bin/env bash
# tools/gpu_check.sh -- one GPU-box round trip: fused-kernel bring-up parity, doALS phase times, the default bench line.
# Logs to gpurun_out/check/; nothing here is a bench value of record.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/check
mkdir -p "$OUT"
timeout -s KILL 150 python tools/tc_bringup.py direct > "$OUT/bringup.log" 2>&1; echo "bring-up exit $?"; tail -n 1 "$OUT/bringup.log"
timeout -s KILL 150 python tools/tc_bringup.py stress > "$OUT/stress.log" 2>&1; tail -n 1 "$OUT/stress.log"
timeout -s KILL 300 python tools/e2e_phases.py > "$OUT/e2e_phases.log" 2>&1; grep -E "setup|doALS wall|RMSE run|update X run|update theta run" "$OUT/e2e_phases.log" | tail -n 14
timeout -s KILL 600 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; cat "$OUT/bench.json"; tail -n 3 "$OUT/bench.err"
echo "== done"
