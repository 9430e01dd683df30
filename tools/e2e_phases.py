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
import numpy as np  # noqa: E402
import torch  # noqa: E402
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # like main.cpp:50-69 / bench.py
for name in ("csr_indptr", "csr_indices", "csr_data", "csc_indptr", "csc_indices", "csc_data", "coo_row", "test_row", "test_col", "test_val"):
    setattr(r, name, pin(getattr(r, name)))
for rep, iters in enumerate(tuple(int(a) for a in os.environ.get("E2E_ITERS", "1,3,10").split(","))):
    th, X = pin(theta0), pin(X0)
    t0 = time.perf_counter()
    fin = c.do_als(*r.doals_args(), th, X, r.test_row, r.test_col, r.test_val, r.m, r.n, w["f"], r.nnz, r.nnz_test, w["lam"], iters, 1, 1, 0)
    print(f"### doALS wall {time.perf_counter() - t0:.3f} s for {iters} iterations, final rmse {fin}", flush=True)
