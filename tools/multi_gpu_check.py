"""Multi-GPU paths of the library, step by step with everything on stderr/stdout (a failing doALS exits the process like
the reference's cudacall macro, so pytest would only show a dead worker):
    python tools/multi_gpu_check.py [n_gpus] [same]     same = all shards on device 0 (CUMF_GROUP_SAME_DEVICE=1)
1. cumf_als_group on n devices == single solver (bit for bit), per-iteration RMSE
2. cumf_doALS under CUMF_GPUS=n == CUMF_GPUS=1"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
if len(sys.argv) > 2 and sys.argv[2] == "same":
    os.environ["CUMF_GROUP_SAME_DEVICE"] = "1"
import torch  # noqa: E402,F401

import cumf_als_b200 as c  # noqa: E402
from cumf_als_b200.data import init_factors, synth_ratings  # noqa: E402

f, lam, iters = 100, 0.048, 3
r = synth_ratings(3000, 12000, 1500000, 30000, seed=77)
theta0, X0 = init_factors(r.m, r.n, f, seed=4)
args = (r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, r.coo_row, r.test_row, r.test_col,
        r.test_val, r.m, r.n, f, lam)
for impl in ("1", "2"):
    os.environ["CUMF_TC_IMPL"] = impl
    s = c.AlsSolver(*args)
    s.set_factors(theta0, X0)
    s.collect_train_sse(True)
    want = []
    for _ in range(iters):
        s.iterate(1)
        want.append(s.rmse())
    th_w, X_w = s.get_factors()
    s.close()
    g = c.AlsGroup(*args, n_devices=n_gpus)
    g.set_factors(theta0, X0)
    g.collect_train_sse(True)
    got = []
    for _ in range(iters):
        ms = g.iterate(1)
        got.append(g.rmse())
    th_g, X_g = g.get_factors()
    g.close()
    print(f"[group impl={impl}] {n_gpus} shards: factors equal {np.array_equal(X_g, X_w) and np.array_equal(th_g, th_w)}; "
          f"rmse single {want[-1]} group {got[-1]}; last iteration {ms:.3f} ms", flush=True)

os.environ.pop("CUMF_QUIET", None)
out = {}
for gpus in ("1", str(n_gpus), str(n_gpus) + " debug"):
    os.environ["CUMF_DEBUG"] = "1" if gpus.endswith("debug") else "0"
    gpus = gpus.split()[0]
    os.environ["CUMF_GPUS"] = gpus
    th, X = theta0.copy(), X0.copy()
    print(f"---- doALS CUMF_GPUS={gpus}", flush=True)
    fin = c.do_als(*r.doals_args(), th, X, r.test_row, r.test_col, r.test_val, r.m, r.n, f, r.nnz, r.nnz_test, lam, iters, 1, 1, 0)
    out[gpus] = (fin, th, X)
    print(f"---- doALS CUMF_GPUS={gpus} returned {fin}", flush=True)
a, b = out["1"], out[str(n_gpus)]
print(f"[doALS] CUMF_GPUS={n_gpus} vs 1: factors equal {np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])}, rmse {a[0]} {b[0]}")
