#!/usr/bin/env bash
# round 2, call A: live-reference parity at contract sizes, sanitizer logs, operand-window probes, reference arms for f=200 / f=10
set -x
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt
for args in "0 1024 4096 0 0 256" "0 1024 4096 16 0 256" "0 1024 4096 0 104 112" "0 1024 4096 104 104 112" "0 1024 4096 8 8 16" "0 1024 4096 128 208 48"; do
  timeout 60 tools/mn_major_probe $args >> $OUT/mn_major_probe.log 2>&1
done
timeout 60 tools/tmem_ld_probe > $OUT/tmem_ld_probe.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_contract_sizes.py -q -s -m gpu > $OUT/pytest_contract.log 2>&1
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_contract_sizes.py > $OUT/pytest_gpu.log 2>&1
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py 100 > $OUT/sanitizer_$tool.log 2>&1
done
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
python bench.py --steps 10 --warmup 3 > $OUT/bench_ours.json 2> $OUT/bench_ours.err
python bench.py --impl reference --workload netflix_f200 --steps 2 --warmup 1 > $OUT/bench_ref_f200.json 2> $OUT/bench_ref_f200.err
python bench.py --workload netflix_f200 --steps 3 --warmup 1 --no-cpu > $OUT/bench_ours_f200.json 2> $OUT/bench_ours_f200.err
python bench.py --impl reference --workload ml10m --ref-variant lu --steps 5 --warmup 1 > $OUT/bench_ref_ml10m_lu.json 2> $OUT/bench_ref_ml10m_lu.err
python bench.py --impl reference --workload netflix --ref-variant lu --steps 2 --warmup 1 > $OUT/bench_ref_netflix_lu.json 2> $OUT/bench_ref_netflix_lu.err
tail -5 $OUT/*.log
cat $OUT/*.json
