#!/usr/bin/env bash
# round 2, call Z2: the live-reference whole-path test on its better-determined problem, five times (the reference's CG is not
# reproducible run to run), and the rest of the parity file once
set -x
OUT=gpurun_out/r2z2
mkdir -p $OUT
for k in 1 2 3 4 5; do
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k test_doals_vs_live_reference > $OUT/live_ref_$k.log 2>&1
  grep -E "reference vs itself|path=|passed|failed" $OUT/live_ref_$k.log
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_contract_sizes.py -q -m gpu > $OUT/pytest.log 2>&1; tail -n 3 $OUT/pytest.log
