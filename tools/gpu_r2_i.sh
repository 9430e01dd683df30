#!/usr/bin/env bash
# round 2, call I: A/B of the short-row CTA shape (workers x systems), wait hint, CTA partition row cost -- Netflix f=100 resident
set -x
OUT=gpurun_out/r2i
mkdir -p $OUT
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu"
timeout 300 $B > $OUT/bench_default.json 2> $OUT/bench_default.err
for v in w6s2 w6s3 wait2us; do
  CUMF_ALS_LIB=$PWD/cumf_als_b200/libcumf_als_b200_$v.so timeout 300 $B > $OUT/bench_$v.json 2> $OUT/bench_$v.err
done
for cost in 32 250 600; do
  CUMF_TC_ROW_COST=$cost timeout 300 $B > $OUT/bench_rowcost$cost.json 2> $OUT/bench_rowcost$cost.err
done
CUMF_ALS_LIB=$PWD/cumf_als_b200/libcumf_als_b200_w6s3.so CUMF_TC_ROW_COST=250 timeout 300 $B > $OUT/bench_w6s3_rowcost250.json 2> $OUT/bench_w6s3_rowcost250.err
CUMF_ALS_LIB=$PWD/cumf_als_b200/libcumf_als_b200_w6s3.so timeout 300 python bench.py --workload ml10m --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_ml10m_w6s3.json 2> $OUT/bench_ml10m_w6s3.err
CUMF_ALS_LIB=$PWD/cumf_als_b200/libcumf_als_b200_w6s3.so timeout 300 python bench.py --workload yahoo --steps 5 --warmup 3 --no-e2e --no-cpu > $OUT/bench_yahoo_w6s3.json 2> $OUT/bench_yahoo_w6s3.err
CUMF_ALS_LIB=$PWD/cumf_als_b200/libcumf_als_b200_w6s3.so CUMF_TC_IMPL=2 timeout 600 python -m pytest tests/test_gpu_generic_f.py -q -m gpu > $OUT/pytest_w6s3.log 2>&1
tail -n 3 $OUT/pytest_w6s3.log
python - <<'PY'
import json,glob
for n in sorted(glob.glob("gpurun_out/r2i/*.json")):
    try:
        d=json.loads(open(n).read().strip().splitlines()[-1]); print(n, round(d["value"],2), "it/s  x", round(d["x_ms"],2), "theta", round(d["theta_ms"],2))
    except Exception as e: print(n, "ERR", e)
PY
