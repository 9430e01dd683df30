#!/usr/bin/env bash
# round 2, call C: where does the generic kernel spend its time (ncu full + source page, CG share), multi-GPU logic on one
# device, sanitizers on the generic kernel, updated tests
set -x
OUT=gpurun_out/r2c
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_generic_f.py tests/test_gpu_parity.py -q -m gpu > $OUT/pytest_gpu.log 2>&1
tail -n 15 $OUT/pytest_gpu.log
for impl in 1 2; do
  CUMF_TC_IMPL=$impl timeout 300 python tools/cg_share.py > $OUT/cg_share_v$impl.log 2>&1
done
cat $OUT/cg_share_v*.log
CUMF_TC_IMPL=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_fused2 -s 2 -c 2 -f \
    -o $OUT/fused2_full python tools/profile_fused.py > $OUT/ncu_full.log 2>&1
tail -n 3 $OUT/ncu_full.log
ncu -i $OUT/fused2_full.ncu-rep --page raw --csv > $OUT/raw.csv 2> /dev/null
ncu -i $OUT/fused2_full.ncu-rep --page source --csv > $OUT/source.csv 2> /dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_fused2 -s 2 -c 2 -f \
    -o $OUT/fused2_f200 python tools/profile_fused.py 1.0 auto netflix_f200 > $OUT/ncu_f200.log 2>&1
ncu -i $OUT/fused2_f200.ncu-rep --page raw --csv > $OUT/raw_f200.csv 2> /dev/null
ncu -i $OUT/fused2_f200.ncu-rep --page source --csv > $OUT/source_f200.csv 2> /dev/null
rm -f $OUT/fused2_f200.ncu-rep
for tool in memcheck synccheck racecheck; do
  CUMF_TC_IMPL=2 timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py 100 10 200 > $OUT/sanitizer_v2_$tool.log 2>&1
  tail -n 3 $OUT/sanitizer_v2_$tool.log
done
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_ours.json 2> $OUT/bench_ours.err
cat $OUT/bench_ours.json
ls -la $OUT
