"""PyTorch-tensor front end over the same `doALS` entry.

`do_als_op` mirrors the reference's TensorFlow op `DoAls` (tensorflow/als_tf.cc:7-30, 42-137): the same 20 inputs in
the same order -- here CPU torch tensors (the TF kernel is registered for DEVICE_CPU and hands raw host pointers to
doALS, als_tf.cc:126-136) -- and the same three outputs: `thetat` float (f, n), `xt` float (f, m) and `rmse` float
(1, 1).  Like the TF kernel it allocates the outputs, initialises them itself (theta = 0.1 * rand()/RAND_MAX with glibc
rand() and no srand, X = 0; als_tf.cc:118-125) and lets doALS update them in place.  Quirk kept: the outputs are
*labelled* (f, n) / (f, m) but hold the row-major [n][f] / [m][f] factors doALS writes (als_tf.cc:109-112, als.cu:1024-1025).

`ratings_from_sparse` builds the ten arrays of the CLI / op from torch sparse matrices, the way
data/netflix/prepare_netflix_data.py:90-110 does it with scipy (`tocsr()` / `tocsc()` of the COO triplets: duplicates
summed, indices ascending inside a row / column, COO rows in CSR order).

No fallback: the compute is `cumf_doALS` of libcumf_als_b200.so, which needs a B200.
"""
from __future__ import annotations

import numpy as np
import torch

from .api import CumfError, _hp, load_library
from .data import Ratings


def _scalar(x) -> float:
    return float(x.reshape(-1)[0]) if isinstance(x, torch.Tensor) else float(x)


def _host_array(t: torch.Tensor, dtype: torch.dtype, what: str) -> np.ndarray:
    if not isinstance(t, torch.Tensor):
        raise CumfError(f"{what}: expected a torch tensor")
    if t.is_cuda:
        raise CumfError(f"{what}: DoAls takes host tensors (als_tf.cc:159 registers the op for DEVICE_CPU)")
    return t.detach().to(dtype).contiguous().reshape(-1).numpy()


def do_als_op(csrrow, csrcol, csrval, cscrow, csccol, cscval, coorow, coorowtest, coocoltest, coovaltest,
              m_t, n_t, f_t, nnz_t, nnz_test_t, lambda_t, iters_t, xbatch_t, thetabatch_t, deviceid_t):
    """(thetat, xt, rmse) = DoAls(...), argument for argument the op of als_tf.cc:7-30.

    csrrow: int32 [m+1] row pointer, csrcol / csrval: [nnz]; cscrow: int32 [nnz] row ids, csccol: int32 [n+1] column
    pointer, cscval: [nnz] (the reference's CSC argument order, main.cpp:99-101); coorow: int32 [nnz] row of every CSR
    entry; coorowtest / coocoltest / coovaltest: the test triplets.  Scalars may be Python numbers or 1-element tensors."""
    m, n, f = int(_scalar(m_t)), int(_scalar(n_t)), int(_scalar(f_t))
    nnz, nnz_test = int(_scalar(nnz_t)), int(_scalar(nnz_test_t))
    lam, iters = _scalar(lambda_t), int(_scalar(iters_t))
    arrs = [_host_array(csrrow, torch.int32, "csrrow"), _host_array(csrcol, torch.int32, "csrcol"),
            _host_array(csrval, torch.float32, "csrval"), _host_array(cscrow, torch.int32, "cscrow"),
            _host_array(csccol, torch.int32, "csccol"), _host_array(cscval, torch.float32, "cscval"),
            _host_array(coorow, torch.int32, "coorow")]
    tarrs = [_host_array(coorowtest, torch.int32, "coorowtest"), _host_array(coocoltest, torch.int32, "coocoltest"),
             _host_array(coovaltest, torch.float32, "coovaltest")]
    if arrs[0].size != m + 1 or arrs[4].size != n + 1 or arrs[1].size < nnz or arrs[3].size < nnz or tarrs[2].size < nnz_test:
        raise CumfError("DoAls: array sizes do not match m / n / nnz / nnz_test")
    lib = load_library()
    thetat = torch.empty((f, n), dtype=torch.float32)          # allocate_output(0, {f, n}), als_tf.cc:109
    xt = torch.empty((f, m), dtype=torch.float32)              # allocate_output(1, {f, m})
    th_np, x_np = thetat.numpy().reshape(-1), xt.numpy().reshape(-1)
    lib.cumf_init_factors(_hp(th_np), _hp(x_np), m, n, f, 0.1, -1)     # als_tf.cc:118-125: rand() unseeded, X = 0
    rmse = lib.cumf_doALS(*[_hp(a) for a in arrs], _hp(th_np), _hp(x_np), *[_hp(a) for a in tarrs], m, n, f, nnz, nnz_test,
                          lam, iters, int(_scalar(xbatch_t)), int(_scalar(thetabatch_t)), int(_scalar(deviceid_t)))
    return thetat, xt, torch.tensor([[rmse]], dtype=torch.float32)


def ratings_from_sparse(train: torch.Tensor, test: torch.Tensor) -> Ratings:
    """The CLI's / op's ten arrays from two torch sparse (COO or CSR) m x n rating matrices on the host."""
    if train.shape != test.shape or train.dim() != 2:
        raise CumfError("train and test must be 2-D sparse matrices of the same shape")
    m, n = int(train.shape[0]), int(train.shape[1])

    def triplets(a: torch.Tensor):
        a = a.to_sparse_coo().coalesce()                     # duplicates summed, (row, col) ascending: scipy's tocsr()
        idx = a.indices()
        return idx[0].to(torch.int64), idx[1].to(torch.int64), a.values().to(torch.float32)

    row, col, val = triplets(train.cpu())
    csr_indptr = torch.zeros(m + 1, dtype=torch.int64)
    csr_indptr[1:] = torch.cumsum(torch.bincount(row, minlength=m), 0)
    order = torch.argsort(col * m + row)                     # column-major order, rows ascending inside a column: tocsc()
    csc_indptr = torch.zeros(n + 1, dtype=torch.int64)
    csc_indptr[1:] = torch.cumsum(torch.bincount(col, minlength=n), 0)
    trow, tcol, tval = triplets(test.cpu())
    i32 = lambda t: t.to(torch.int32).contiguous().numpy()
    return Ratings(m=m, n=n, csr_indptr=i32(csr_indptr), csr_indices=i32(col), csr_data=val.contiguous().numpy(),
                   csc_indptr=i32(csc_indptr), csc_indices=i32(row[order]), csc_data=val[order].contiguous().numpy(),
                   coo_row=i32(row), test_row=i32(trow), test_col=i32(tcol), test_val=tval.contiguous().numpy())


def als_fit(train: torch.Tensor, test: torch.Tensor, f: int, lam: float, iters: int = 10, device: int = 0):
    """doALS on torch sparse matrices: returns (theta [n, f], X [m, f], final test RMSE).  Initialisation as the TF op."""
    r = ratings_from_sparse(train, test)
    t = torch.from_numpy
    thetat, xt, rmse = do_als_op(t(r.csr_indptr), t(r.csr_indices), t(r.csr_data), t(r.csc_indices), t(r.csc_indptr),
                                 t(r.csc_data), t(r.coo_row), t(r.test_row), t(r.test_col), t(r.test_val), r.m, r.n, f,
                                 r.nnz, r.nnz_test, lam, iters, 1, 1, device)
    return thetat.reshape(r.n, f), xt.reshape(r.m, f), float(rmse)
