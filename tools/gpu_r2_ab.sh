#!/usr/bin/env bash
# round 2: bring-up of the generic-f kernel first (cheap, decides what the rest measures), then call A, then the rest of call B
set -x
OUT=gpurun_out/r2b
mkdir -p $OUT
for f in 100 10 60 120 130 200; do
  timeout 120 python tools/tc2_bringup.py $f > $OUT/bringup_f$f.log 2>&1
done
timeout 120 python tools/tc2_bringup.py 100 1 > $OUT/bringup_f100_sym.log 2>&1
tail -2 $OUT/bringup_*.log
bash tools/gpu_r2_a.sh
bash tools/gpu_r2_b.sh
