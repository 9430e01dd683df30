"""Where does the whole-job time of doALS(host pointers) go?  (GPU box)  Prints the library's own
CUMF_DEBUG phase lines for a 3-iteration run on the Netflix-shaped workload."""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import cumf_als_b200 as c  # noqa: E402

w = bench.WORKLOADS["netflix"]
r, theta0, X0 = bench.make_inputs(w, float(sys.argv[1]) if len(sys.argv) > 1 else 1.0, "cuda")
os.environ["CUMF_DEBUG"] = "1"
for rep in range(2):
    th, X = theta0.copy(), X0.copy()
    t0 = time.perf_counter()
    fin = c.do_als(*r.doals_args(), th, X, r.test_row, r.test_col, r.test_val, r.m, r.n, w["f"], r.nnz, r.nnz_test, w["lam"], 3, 1, 1, 0)
    print(f"### doALS wall {time.perf_counter() - t0:.3f} s, final rmse {fin}", flush=True)
