#!/usr/bin/env bash
# tools/gpu_final.sh -- the round-end evidence run on one B200: all GPU tests, smoke(), both bench arms, the ncu launch
# list of the bench command and one `--set full` capture of the fused kernel.  Logs to gpurun_out/final/.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/final
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,driver_version --format=csv > "$OUT/gpu.txt" 2>&1
timeout -s KILL 900 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?"; tail -n 3 "$OUT/pytest_gpu.log"
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke exit $?"; tail -n 1 "$OUT/smoke.log"
timeout -s KILL 400 python bench.py --impl reference > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "ref exit $?"; cut -c1-400 "$OUT/bench_ref.json"
timeout -s KILL 400 python bench.py > "$OUT/bench_ours.json" 2> "$OUT/bench_ours.err"; echo "ours exit $?"; cut -c1-300 "$OUT/bench_ours.json"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > "$OUT/launches_bench.json" 2> "$OUT/launches.err"; echo "launch list exit $?"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:als_fused -s 2 -c 2 -f \
    -o "$OUT/fused_full" python tools/profile_fused.py > "$OUT/ncu_full.log" 2>&1; echo "ncu full exit $?"; tail -n 2 "$OUT/ncu_full.log"
timeout -s KILL 200 python bench.py --workload ml10m --steps 5 --warmup 3 --no-e2e --no-cpu > "$OUT/bench_ml10m.json" 2> "$OUT/bench_ml10m.err"; cut -c1-260 "$OUT/bench_ml10m.json"
timeout -s KILL 300 python bench.py --workload netflix_f200 --steps 2 --warmup 1 --no-e2e --no-cpu > "$OUT/bench_f200.json" 2> "$OUT/bench_f200.err"; cut -c1-260 "$OUT/bench_f200.json"
ls -la "$OUT"
echo "== done"
