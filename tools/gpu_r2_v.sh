#!/usr/bin/env bash
# round 2, call V: the round-end evidence run on one B200 with the final kernels -- every GPU test, smoke(), both bench arms,
# the other configs with their reference arms, sanitizers on the generic kernel, the ncu launch list of the bench command and
# one `--set full` capture of the two default kernels
set -x
OUT=gpurun_out/r2v
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,driver_version --format=csv > $OUT/gpu.txt 2>&1
timeout 1800 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 6 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 $OUT/smoke.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-300 $OUT/bench_ref.json
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; cut -c1-300 $OUT/bench_ours.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/launches_bench.json 2> $OUT/launches.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_fused -s 2 -c 2 -f \
    -o $OUT/fused_full python tools/profile_fused.py > $OUT/ncu_full.log 2>&1; tail -n 2 $OUT/ncu_full.log
ncu -i $OUT/fused_full.ncu-rep --page raw --csv > $OUT/raw.csv 2> /dev/null
ncu -i $OUT/fused_full.ncu-rep --page source --csv > $OUT/source.csv 2> /dev/null
rm -f $OUT/fused_full.ncu-rep
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py 100 10 200 > $OUT/sanitizer_$tool.log 2>&1
  tail -n 2 $OUT/sanitizer_$tool.log
done
timeout 300 python bench.py --workload netflix_f200 --steps 5 --warmup 3 --no-cpu > $OUT/bench_f200.json 2> $OUT/bench_f200.err
timeout 300 python bench.py --workload ml10m --steps 10 --warmup 3 --no-cpu > $OUT/bench_ml10m.json 2> $OUT/bench_ml10m.err
timeout 600 python bench.py --workload yahoo --steps 5 --warmup 3 --no-cpu > $OUT/bench_yahoo.json 2> $OUT/bench_yahoo.err
cat $OUT/bench_f200.json $OUT/bench_ml10m.json $OUT/bench_yahoo.json | cut -c1-400
ls -la $OUT
