#!/usr/bin/env bash
# tools/gpu_confirm.sh -- short re-check of the committed state: GPU tests, smoke(), both bench arms (stdout = one JSON line each)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/confirm
mkdir -p "$OUT"
timeout -s KILL 600 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?"; tail -n 2 "$OUT/pytest_gpu.log"
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke exit $?"; tail -n 1 "$OUT/smoke.log"
timeout -s KILL 400 python bench.py --impl reference --steps 5 --warmup 3 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "ref exit $? lines $(wc -l < "$OUT/bench_ref.json")"; cut -c1-200 "$OUT/bench_ref.json"
timeout -s KILL 400 python bench.py > "$OUT/bench_ours.json" 2> "$OUT/bench_ours.err"; echo "ours exit $? lines $(wc -l < "$OUT/bench_ours.json")"; cut -c1-260 "$OUT/bench_ours.json"
echo "== done"
