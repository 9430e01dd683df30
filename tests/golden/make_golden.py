"""Generate the golden vectors under tests/golden/ from the REAL reference.

Runs on a B200 (`gpurun -- python tests/golden/make_golden.py`): loads
oracle/_ref/libref_als_{cg,lu}.so -- the unmodified reference sources compiled for
sm_100a by oracle/build_ref.sh -- and records, on small seeded inputs, what the
reference's own kernels / doALS produce at each seam of the hot path:
  gram_*.npz   get_hermitian100 / get_hermitianT10 output `tt` (+ the csrmm2/geam RHS)
  cg_*.npz     updateXWithCGHost output
  lu_*.npz     updateX (cuBLAS LU, no pivot) output
  rmse.npz     RMSE kernel + Sasum, train-style and test-style launch
  doals_*.npz  full doALS: final factors, returned RMSE and the per-iteration
               Train/Test RMSE lines it prints
Inputs are stored next to the outputs so the CPU tests need nothing but numpy.
Outputs are written to gpurun_out/golden/ (brought back by gpurun) and, when run in a
writable checkout, to tests/golden/ directly.
"""
from __future__ import annotations

import ctypes
import os
import re
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from cumf_als_b200.data import synth_ratings  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT_DIRS = [ROOT / "gpurun_out" / "golden", ROOT / "tests" / "golden"]


def save(name, **arrays):
    for d in OUT_DIRS:
        d.mkdir(parents=True, exist_ok=True)
        np.savez_compressed(d / name, **arrays)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrays.items()})


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


class CaptureStdout:
    """Capture the C-level stdout of the reference's printf calls."""

    def __enter__(self):
        self.libc = ctypes.CDLL(None)
        sys.stdout.flush()
        self.libc.fflush(None)
        self.saved = os.dup(1)
        self.tmp = tempfile.TemporaryFile(mode="w+b")
        os.dup2(self.tmp.fileno(), 1)
        return self

    def __exit__(self, *exc):
        self.libc.fflush(None)
        os.dup2(self.saved, 1)
        os.close(self.saved)
        self.tmp.seek(0)
        self.text = self.tmp.read().decode(errors="replace")
        self.tmp.close()


def small_csr(lengths, n, rng):
    rowptr = np.zeros(len(lengths) + 1, np.int32)
    rowptr[1:] = np.cumsum(lengths)
    cols = np.concatenate([np.sort(rng.choice(n, size=k, replace=False)) for k in lengths] + [np.zeros(0, np.int64)])
    vals = rng.integers(1, 6, size=cols.size).astype(np.float32)
    return rowptr, cols.astype(np.int32), vals


def gram_case(name, f, lengths, n, lam, seed):
    rng = np.random.default_rng(seed)
    rowptr, cols, vals = small_csr(lengths, n, rng)
    factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
    m = len(lengths)
    lib = O.ref("cg")
    d_rowptr, d_cols, d_vals, d_factor = dev(rowptr), dev(cols), dev(vals), dev(factor)
    tt = torch.zeros((m, f, f), dtype=torch.float32, device="cuda")
    lib.ref_get_hermitian(0, m, tt.data_ptr(), d_rowptr.data_ptr(), d_cols.data_ptr(), lam, m, f, d_factor.data_ptr())
    rhs = torch.zeros((m, f), dtype=torch.float32, device="cuda")
    lib.ref_rhs(m, n, f, int(cols.size), d_vals.data_ptr(), d_rowptr.data_ptr(), d_cols.data_ptr(), d_factor.data_ptr(),
                rhs.data_ptr())
    torch.cuda.synchronize()
    # a second batch window: batch_offset = 2, batch_size = 3 (als.cu:768-777 style sub-batch)
    tt_sub = torch.zeros((3, f, f), dtype=torch.float32, device="cuda")
    lib.ref_get_hermitian(2, 3, tt_sub.data_ptr(), d_rowptr.data_ptr(), d_cols.data_ptr(), lam, m, f, d_factor.data_ptr())
    torch.cuda.synchronize()
    save(name, rowptr=rowptr, colidx=cols, val=vals, factor=factor, lam=np.float32(lam), f=np.int32(f),
         tt=tt.cpu().numpy(), rhs=rhs.cpu().numpy(), tt_sub=tt_sub.cpu().numpy())
    return tt.cpu().numpy(), rhs.cpu().numpy()


def solve_case(name, f, tt, seed):
    rng = np.random.default_rng(seed)
    batch = tt.shape[0]
    b = rng.standard_normal((batch, f)).astype(np.float32)
    x0 = (0.1 * rng.standard_normal((batch, f))).astype(np.float32)
    out = {}
    for it in (6.0, 2.0):
        dA, dx, db = dev(tt), dev(x0), dev(b)
        O.ref("cg").ref_cg(dA.data_ptr(), dx.data_ptr(), db.data_ptr(), batch, f, it)
        torch.cuda.synchronize()
        out[f"x_cg{int(it)}"] = dx.cpu().numpy()
    # LU oracle (updateX, als.cu:58-122)
    dA, db = dev(tt), dev(b)
    dX = torch.zeros((batch, f), dtype=torch.float32, device="cuda")
    O.ref("lu").ref_lu(batch, 0, db.data_ptr(), dA.data_ptr(), dX.data_ptr(), batch, 1, f, 1)
    torch.cuda.synchronize()
    save(name, A=tt, b=b, x0=x0, f=np.int32(f), x_lu=dX.cpu().numpy(), **out)


def rmse_case():
    rng = np.random.default_rng(77)
    f, m, n, cnt = 20, 50, 70, 1000
    theta = (0.5 * rng.standard_normal((n, f))).astype(np.float32)
    X = (0.5 * rng.standard_normal((m, f))).astype(np.float32)
    row = rng.integers(0, m, cnt).astype(np.int32)
    col = rng.integers(0, n, cnt).astype(np.int32)
    val = rng.integers(1, 6, cnt).astype(np.float32)
    lib = O.ref("cg")
    args = [dev(val), dev(row), dev(col), dev(theta), dev(X)]
    ptrs = [a.data_ptr() for a in args]
    train = lib.ref_rmse(*ptrs, cnt, f, 0)
    test = lib.ref_rmse(*ptrs, cnt, f, 1)
    save("rmse.npz", val=val, row=row, col=col, thetaT=theta, XT=X, f=np.int32(f), rmse_train=np.float32(train),
         rmse_test=np.float32(test))


def doals_case(name, m, n, nnz, nnz_test, f, lam, iters, seed, theta_batch=1):
    r = synth_ratings(m, n, nnz, nnz_test, seed=seed)
    rng = np.random.default_rng(seed + 1)
    theta0 = (0.2 * rng.random((n, f))).astype(np.float32)
    out = {}
    for variant in ("cg", "lu"):
        th, X = theta0.copy(), np.zeros((m, f), np.float32)
        with CaptureStdout() as cap:
            fin = O.ref_do_als(r, th, X, f, lam, iters, 1, theta_batch, variant)
        tr = [float(x) for x in re.findall(r"Train RMSE in iter \d+: ([0-9.naninf-]+)", cap.text)]
        te = [float(x) for x in re.findall(r"Test RMSE in iter \d+: ([0-9.naninf-]+)", cap.text)]
        out[f"theta_{variant}"] = th
        out[f"x_{variant}"] = X
        out[f"final_{variant}"] = np.float32(fin)
        out[f"train_{variant}"] = np.array(tr, np.float64)
        out[f"test_{variant}"] = np.array(te, np.float64)
        print(name, variant, "final", fin, "test", te)
    save(name, m=np.int32(m), n=np.int32(n), f=np.int32(f), lam=np.float32(lam), iters=np.int32(iters),
         csr_indptr=r.csr_indptr, csr_indices=r.csr_indices, csr_data=r.csr_data, csc_indptr=r.csc_indptr,
         csc_indices=r.csc_indices, csc_data=r.csc_data, coo_row=r.coo_row, test_row=r.test_row, test_col=r.test_col,
         test_val=r.test_val, theta0=theta0, **out)


def main():
    assert torch.cuda.is_available(), "golden vectors come from the reference kernels: needs a GPU"
    print("device:", torch.cuda.get_device_name(0))
    tt100, _ = gram_case("gram_f100.npz", 100, [1, 5, 27, 28, 29, 60, 0, 200], 300, 0.05, 11)
    tt20, _ = gram_case("gram_f20.npz", 20, [3, 1, 40, 28, 56, 57, 0, 9, 130, 2, 31, 64], 200, 0.048, 12)
    gram_case("gram_f200.npz", 200, [30, 3, 75], 120, 1.4, 13)
    keep100 = [0, 1, 2, 3, 4, 5, 7]            # drop the empty row (singular system)
    solve_case("solve_f100.npz", 100, tt100[keep100], 21)
    keep20 = [i for i in range(12) if i != 6]
    solve_case("solve_f20.npz", 20, tt20[keep20], 22)
    rmse_case()
    doals_case("doals_f20.npz", 120, 200, 6000, 700, 20, 0.05, 3, 3)
    doals_case("doals_f100.npz", 60, 90, 2500, 600, 100, 0.048, 2, 4)
    print("golden done")


if __name__ == "__main__":
    main()
