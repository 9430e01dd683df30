"""Workload for `compute-sanitizer --tool {memcheck,racecheck,synccheck}`: the fused tcgen05 kernels (long-row symmetric
variant, short-row variant, split rows, materialising mode) and the unfused kernels on shapes small enough for the
sanitizer's ~100x slowdown.  SURVEY.md section 5 asked for exactly this on a kernel whose correctness hangs on mbarrier
parity bookkeeping.  Usage: compute-sanitizer --tool racecheck python tools/sanitize_small.py [f ...]"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import cumf_als_b200 as c  # noqa: E402
from cumf_als_b200.data import init_factors, synth_ratings  # noqa: E402

fs = [int(a) for a in sys.argv[1:]] or [100]
os.environ.setdefault("CUMF_SPLIT_NNZ", "600")          # a few rows split across CTAs
for f in fs:
    # m << n: the X side has long rows (symmetric variant), the theta side short ones
    r = synth_ratings(60, 900, 24000, 1000, seed=3)
    theta0, X0 = init_factors(r.m, r.n, f, seed=2)
    s = c.AlsSolver(r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, r.coo_row,
                    r.test_row, r.test_col, r.test_val, r.m, r.n, f, 0.05)
    s.set_factors(theta0, X0)
    s.collect_train_sse(True)
    ms = s.iterate(2)
    tr, te = s.rmse()
    s.close()
    print(f"sanitize_small: f={f} 2 iterations {ms:.1f} ms, rmse {tr:.5f} / {te:.5f}", flush=True)
    assert np.isfinite(tr) and np.isfinite(te)

if 100 in fs:
    # long X rows at f = 100: round-1 kernel on the X side, generic kernel on the theta side writing split rows into its table
    os.environ["CUMF_SPLIT_NNZ"] = "2200"
    r = synth_ratings(16, 4000, 32000, 1000, seed=4)
    theta0, X0 = init_factors(r.m, r.n, 100, seed=2)
    s = c.AlsSolver(r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, r.coo_row,
                    r.test_row, r.test_col, r.test_val, r.m, r.n, 100, 0.05)
    s.set_factors(theta0, X0)
    ms = s.iterate(3)
    tr, te = s.rmse()
    s.close()
    print(f"sanitize_small: f=100 long X rows, 3 iterations {ms:.1f} ms, rmse {tr:.5f} / {te:.5f}", flush=True)
    assert np.isfinite(tr) and np.isfinite(te)
