#!/usr/bin/env bash
# round 2, call B: bring-up + parity of the generic-f kernel, A/B against the round-1 f=100 kernel, f=200 / f=10 benches
set -x
OUT=gpurun_out/r2b
mkdir -p $OUT
for t in test_tc2_gram_vs_oracle test_tc2_gram_small_and_large_values test_tc2_half_step_vs_simt test_tc2_half_step_sym_variant_vs_simt \
         test_tc2_deterministic_and_partial_row_range test_tc2_doals_vs_oracle_and_simt test_tc2_plan_gram_ranges_vs_oracle; do
  timeout 600 python -m pytest tests/test_gpu_generic_f.py -q -s -m gpu -k $t > $OUT/pytest_$t.log 2>&1
done
CUMF_TC_IMPL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_f100_v1.json 2> $OUT/bench_f100_v1.err
CUMF_TC_IMPL=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_f100_v2.json 2> $OUT/bench_f100_v2.err
timeout 600 python bench.py --workload netflix_f200 --steps 5 --warmup 2 --no-cpu > $OUT/bench_f200.json 2> $OUT/bench_f200.err
timeout 600 python bench.py --workload ml10m --steps 10 --warmup 3 --no-cpu > $OUT/bench_ml10m.json 2> $OUT/bench_ml10m.err
CUMF_TC_IMPL=2 timeout 600 python bench.py --workload yahoo --steps 5 --warmup 2 --no-cpu --no-e2e > $OUT/bench_yahoo_v2.json 2> $OUT/bench_yahoo_v2.err
tail -4 $OUT/pytest_*.log
cat $OUT/*.json
