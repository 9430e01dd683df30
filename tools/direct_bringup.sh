#!/usr/bin/env bash
# tools/direct_bringup.sh -- one GPU-box round trip for the direct-staging variant of the fused kernel:
# layout probe -> Gram parity (finds the working descriptor variant) -> stress -> bench A/B -> parity tests -> ncu.
# Every step runs under its own timeout and logs to gpurun_out/; nothing here is a bench value of record.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/direct
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > "$OUT/gpu.txt" 2>&1

echo "== bring-up (Gram parity, direct vs fp32 staging)"
CFG=""
if CUMF_TC_DIRECT=1 timeout -s KILL 150 python tools/tc_bringup.py direct > "$OUT/bringup.log" 2>&1; then CFG="CUMF_TC_DIRECT=1"; fi
tail -n 16 "$OUT/bringup.log"
echo "working config: ${CFG:-none}" | tee "$OUT/config.txt"

echo "== bench, fp32 staging (baseline of this box)"
CUMF_TC_DIRECT=0 timeout -s KILL 400 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > "$OUT/bench_conv.json" 2> "$OUT/bench_conv.err"
cat "$OUT/bench_conv.json"

if [ -n "$CFG" ]; then
  echo "== stress (many chunks per CTA, split rows): direct vs SIMT"
  env $CFG timeout -s KILL 150 python tools/tc_bringup.py stress > "$OUT/stress.log" 2>&1; echo "exit $?" >> "$OUT/stress.log"
  env $CFG CUMF_TC_CTAS=3 timeout -s KILL 150 python tools/tc_bringup.py stress >> "$OUT/stress.log" 2>&1; echo "exit $?" >> "$OUT/stress.log"
  cat "$OUT/stress.log"
  echo "== bench, direct staging"
  env $CFG timeout -s KILL 400 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > "$OUT/bench_direct.json" 2> "$OUT/bench_direct.err"
  cat "$OUT/bench_direct.json"; tail -n 5 "$OUT/bench_direct.err"
  echo "== parity tests with direct staging"
  env $CFG timeout -s KILL 420 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > "$OUT/pytest_direct.log" 2>&1
  tail -n 8 "$OUT/pytest_direct.log"
  echo "== ncu (direct)"
  env $CFG timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:als_fused -s 2 -c 2 -f \
      -o "$OUT/direct_full" python tools/profile_fused.py > "$OUT/ncu.log" 2>&1
  tail -n 3 "$OUT/ncu.log"
  ls -la "$OUT"
fi
echo "== done"
