"""Bring-up of the generic-f kernel (gram_tc2.cuh): materialised [A|b] against the oracle for a few ranks, with a map of
where the error sits (64-column chunk / 16-column group / row block) when a rank is off.  One process per rank so that a
trapped launch does not poison the others:  python tools/tc2_bringup.py <f> [sym]"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
os.environ["CUMF_TC_IMPL"] = "2"
f = int(sys.argv[1])
if len(sys.argv) > 2:
    os.environ["CUMF_TC_SYM"] = sys.argv[2]
import torch  # noqa: E402

import cumf_als_b200 as c  # noqa: E402
from oracle import oracle as O  # noqa: E402
from test_gpu_parity import random_csr, run_gram  # noqa: E402

rng = np.random.default_rng(f)
lengths = [16, 1, 2, 15, 17, 33, 0, 100, 250, 1000, 3000, 48, 64, 65, 256, 257]
n = 5000
rowptr, colidx, val = random_csr(rng, lengths, n)
factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
tt, rhs = run_gram(torch, rowptr, colidx, val, factor, f, 0.05, path=c.PATH_TC)
ref = O.gram(rowptr, colidx, factor, f, 0.05)
ref_b = O.rhs(rowptr, colidx, val, factor, f)
bad = 0
for u, L in enumerate(lengths):
    scale = max(np.abs(ref[u]).max(), 1e-30)
    err = np.abs(tt[u] - ref[u]) / scale
    eb = np.abs(rhs[u] - ref_b[u]).max() / max(np.abs(ref_b[u]).max(), 1e-30)
    flag = "" if err.max() < 6e-6 and eb < 6e-6 and np.isfinite(err).all() else "   <-- OFF"
    print(f"f={f} row {u:2d} ({L:4d} ratings): A max err {err.max():.2e}  b {eb:.2e}  nan {int(np.isnan(tt[u]).sum())}{flag}")
    if flag:
        bad += 1
        rows_off = np.flatnonzero(~(err.max(axis=1) < 6e-6))
        cols_off = np.flatnonzero(~(err.max(axis=0) < 6e-6))
        print(f"     rows off: {rows_off[:12]}... ({rows_off.size}), cols off: {cols_off[:12]}... ({cols_off.size})")
        print(f"     got[0,:6] {tt[u][0,:6]}  want {ref[u][0,:6]}")
print(f"f={f}: {'OK' if bad == 0 else str(bad) + ' rows off'}")
