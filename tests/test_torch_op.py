"""PyTorch-tensor front end (cumf_als_b200/torch_op.py): the DoAls op mirror of tensorflow/als_tf.cc and the
sparse-matrix -> ten-arrays conversion of data/netflix/prepare_netflix_data.py:90-110."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import cumf_als_b200 as c
from cumf_als_b200 import torch_op as T
from cumf_als_b200.data import synth_ratings


def test_ratings_from_sparse_matches_scipy_conversion():
    rng = np.random.default_rng(3)
    m, n, k = 40, 55, 600
    row, col = rng.integers(0, m, k), rng.integers(0, n, k)          # with duplicates: summed like scipy's tocsr()
    val = rng.integers(1, 6, k).astype(np.float32)
    trow, tcol = rng.integers(0, m, 90), rng.integers(0, n, 90)
    tval = rng.integers(1, 6, 90).astype(np.float32)
    train = torch.sparse_coo_tensor(torch.tensor(np.stack([row, col])), torch.tensor(val), (m, n))
    test = torch.sparse_coo_tensor(torch.tensor(np.stack([trow, tcol])), torch.tensor(tval), (m, n))
    r = T.ratings_from_sparse(train, test)
    coo = sp.coo_matrix((val, (row, col)), shape=(m, n))
    csr, csc = coo.tocsr(), coo.tocsc()
    csr.sort_indices(); csc.sort_indices()
    assert np.array_equal(r.csr_indptr, csr.indptr) and np.array_equal(r.csr_indices, csr.indices) and np.array_equal(r.csr_data, csr.data)
    assert np.array_equal(r.csc_indptr, csc.indptr) and np.array_equal(r.csc_indices, csc.indices) and np.array_equal(r.csc_data, csc.data)
    assert np.array_equal(r.coo_row, np.repeat(np.arange(m), np.diff(csr.indptr)))       # COO rows in CSR order
    tcoo = sp.coo_matrix((tval, (trow, tcol)), shape=(m, n)).tocsr().tocoo()
    assert np.array_equal(r.test_row, tcoo.row) and np.array_equal(r.test_col, tcoo.col) and np.array_equal(r.test_val, tcoo.data)
    # the same input as a CSR tensor
    r2 = T.ratings_from_sparse(train.coalesce().to_sparse_csr(), test)
    assert np.array_equal(r2.csr_indices, r.csr_indices) and np.array_equal(r2.csc_data, r.csc_data)


def test_init_factors_is_glibc_rand_like_the_front_ends():
    lib = c.load_library()
    libc = ctypes.CDLL(None)
    libc.rand.restype = ctypes.c_int
    m, n, f = 3, 5, 4
    th, x = np.full(n * f, 7.0, np.float32), np.full(m * f, 7.0, np.float32)
    as_p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.cumf_init_factors(as_p(th), as_p(x), m, n, f, 0.2, 0)                   # main.cpp:72-78: srand(0), 0.2 * rand()/RAND_MAX
    libc.srand(0)
    # the front ends multiply by the DOUBLE literal and round once (main.cpp:75: 0.2*((float)rand()/(float)RAND_MAX))
    want = np.array([np.float32(0.2 * float(np.float32(libc.rand()) / np.float32(2147483647))) for _ in range(n * f)], np.float32)
    assert np.array_equal(th, want) and not x.any()
    nxt = np.empty(n * f, np.float32)
    lib.cumf_init_factors(as_p(nxt), None, m, n, f, 0.1, -1)                   # als_tf.cc:118-123: no srand -> the sequence continues
    assert not np.array_equal(nxt, np.float32(0.5) * want) and (nxt >= 0).all() and (nxt <= 0.1).all()


def test_op_rejects_device_tensors_and_bad_sizes():
    r = synth_ratings(30, 40, 500, 60, seed=2)
    t = torch.from_numpy
    args = [t(r.csr_indptr), t(r.csr_indices), t(r.csr_data), t(r.csc_indices), t(r.csc_indptr), t(r.csc_data), t(r.coo_row),
            t(r.test_row), t(r.test_col), t(r.test_val)]
    with pytest.raises(c.CumfError):
        T.do_als_op(*args, r.m + 1, r.n, 10, r.nnz, r.nnz_test, 0.05, 1, 1, 1, 0)      # csrrow is not m+2 long


@pytest.mark.gpu
def test_op_equals_do_als_with_the_ops_initialisation():
    """DoAls mirror == doALS called the way als_tf.cc:118-136 calls it (same unseeded-rand initialisation replayed)."""
    assert torch.cuda.is_available()
    import os
    os.environ["CUMF_QUIET"] = "1"
    r = synth_ratings(300, 450, 20000, 1500, seed=9)
    f, lam, iters = 20, 0.05, 3
    libc = ctypes.CDLL(None)
    t = torch.from_numpy
    libc.srand(5)
    thetat, xt, rmse = T.do_als_op(t(r.csr_indptr), t(r.csr_indices), t(r.csr_data), t(r.csc_indices), t(r.csc_indptr), t(r.csc_data),
                                   t(r.coo_row), t(r.test_row), t(r.test_col), t(r.test_val), torch.tensor([r.m]), torch.tensor([r.n]),
                                   torch.tensor([f]), torch.tensor([r.nnz]), torch.tensor([r.nnz_test]), torch.tensor([lam]),
                                   torch.tensor([iters]), 1, 1, 0)
    assert tuple(thetat.shape) == (f, r.n) and tuple(xt.shape) == (f, r.m) and tuple(rmse.shape) == (1, 1)
    libc.srand(5)
    th = np.empty((r.n, f), np.float32)
    X = np.empty((r.m, f), np.float32)
    c.load_library().cumf_init_factors(th.ctypes.data_as(ctypes.c_void_p), X.ctypes.data_as(ctypes.c_void_p), r.m, r.n, f, 0.1, -1)
    fin = c.do_als(*r.doals_args(), th, X, r.test_row, r.test_col, r.test_val, r.m, r.n, f, r.nnz, r.nnz_test, lam, iters, 1, 1, 0)
    assert float(rmse) == fin
    assert np.array_equal(thetat.numpy().reshape(r.n, f), th) and np.array_equal(xt.numpy().reshape(r.m, f), X)
    # and through sparse tensors
    train = torch.sparse_csr_tensor(t(r.csr_indptr).long(), t(r.csr_indices).long(), t(r.csr_data), (r.m, r.n))
    test = torch.sparse_coo_tensor(torch.stack([t(r.test_row).long(), t(r.test_col).long()]), t(r.test_val), (r.m, r.n))
    libc.srand(5)
    theta2, X2, rmse2 = T.als_fit(train, test, f, lam, iters)
    rr = T.ratings_from_sparse(train, test)
    if rr.nnz_test == r.nnz_test:      # no duplicate test pairs were merged: the very same problem
        assert rmse2 == fin and np.array_equal(theta2.numpy(), th)
