"""BASELINE configs[4] (Hugewiki-scale synthetic, hugewiki.cu:27-42: m = 50 082 603, n = 39 780, 3.1 G ratings) as a 1/64
replica on one GPU: the matrix is generated shard by shard ON the device (csrc/synth.cu), the shards' solvers borrow those
slices (cumf_als_create_device: int64 pointers, no host copy) and run row-sharded with the peer stores and the device
barrier (two shards on device 0) -- and must reproduce, bit for bit, the GENERIC path: the same matrix downloaded to the
host, transposed there by scipy, and handed to cumf_als_create as the reference's ten arrays.  Run with -m gpu."""
import numpy as np
import pytest

import cumf_als_b200 as c

pytestmark = pytest.mark.gpu

M_FULL, N_FULL, NNZ_FULL = 50082603, 39780, 3101144313          # hugewiki.cu:27-38


def test_hugewiki_replica_device_shards_equal_generic_host_path(cuda, monkeypatch):
    import scipy.sparse as sp
    monkeypatch.setenv("CUMF_GROUP_SAME_DEVICE", "1")
    m, n, f, lam, seed, shards = M_FULL // 64, N_FULL, 100, 0.048, 2026, 2
    avg = NNZ_FULL / M_FULL                                       # 61.9 ratings per row
    g = c.AlsGroup.from_synth(m, n, avg, seed, test_per_shard=40000, f=f, lam=lam, n_devices=shards)
    assert abs(g.nnz / (NNZ_FULL / 64) - 1) < 0.02, g.nnz         # the degree law hits the requested density
    theta0, X0 = g.get_factors()
    assert not X0.any() and 0 <= theta0.min() and theta0.max() <= 0.2
    assert g.collect_train_sse(True)
    hist_g = []
    for _ in range(2):
        g.iterate(1)
        hist_g.append(g.rmse())
    th_g, X_g = g.get_factors()
    nnz = g.nnz
    g.close()

    # the generic path on the same matrix: one whole-matrix slice downloaded, CSC by scipy
    whole = c.SynthShard(m, n, avg, seed, (0, m), (0, n))
    ptr, col, val = whole.slice(0)
    cptr, crow, cval = whole.slice(1)
    assert ptr[-1] == nnz == cptr[-1] == whole.total_nnz
    assert (np.diff(ptr) >= 1).all()                              # every row rated (README.md:113)
    csr = sp.csr_matrix((val, col, ptr), shape=(m, n))
    assert csr.has_sorted_indices or (np.diff(col)[np.setdiff1d(np.arange(col.size - 1), ptr[1:-1] - 1)] > 0).all()
    csc = csr.tocsc()
    csc.sort_indices()
    # the generator's own CSC slice IS the transpose (same order: by column, rows ascending)
    assert np.array_equal(csc.indptr.astype(np.int64), cptr) and np.array_equal(csc.indices, crow) and np.array_equal(csc.data, cval)
    whole.close()
    coo_row = np.repeat(np.arange(m, dtype=np.int32), np.diff(ptr).astype(np.int64))
    tests = []
    for k in range(shards):
        sh = c.SynthShard(m, n, avg, seed, (m * k // shards, m * (k + 1) // shards), (0, 1), test_cnt=40000)
        tests.append(sh.test_samples())
        sh.close()
    t_row, t_col, t_val = (np.concatenate([t[i] for t in tests]) for i in range(3))
    s = c.AlsSolver(ptr.astype(np.int32), col, val, csc.indices.astype(np.int32), csc.indptr.astype(np.int32), csc.data.astype(np.float32),
                    coo_row, t_row, t_col, t_val, m, n, f, lam)
    s.set_factors(theta0, X0)
    s.collect_train_sse(True)
    hist_s = []
    for _ in range(2):
        s.iterate(1)
        hist_s.append(s.rmse())
    th_s, X_s = s.get_factors()
    s.close()
    assert np.array_equal(X_g, X_s) and np.array_equal(th_g, th_s)
    hist_g, hist_s = np.array(hist_g), np.array(hist_s)
    print(f"1/64 replica: m={m} n={n} nnz={nnz}; rmse device-shards {hist_g[-1]} generic {hist_s[-1]}")
    assert np.abs(hist_g[:, 0] - hist_s[:, 0]).max() < 1e-4 * hist_s[:, 0].max()
    assert np.abs(hist_g[:, 1] - hist_s[:, 1]).max() < 2e-3 * hist_s[:, 1].max()     # the generic path drops the tail block of test samples (als.cu:1006)


def test_csr_to_csc_device_equals_scipy(cuda):
    """cumf_csr_to_csc_device (radix sort of (column, row) keys) against scipy's tocsc on a ragged matrix with empty rows and
    columns, and shard loading from .bin files straight into a device-resident solver (no whole-matrix host arrays)."""
    import scipy.sparse as sp
    from cumf_als_b200.api import csr_to_csc_device
    from cumf_als_b200.data import synth_ratings
    rng = np.random.default_rng(5)
    m, n = 700, 1100
    dense = (rng.random((m, n)) < 0.02) * rng.integers(1, 6, (m, n))
    dense[13] = 0
    dense[:, 77] = 0
    csr = sp.csr_matrix(dense.astype(np.float32))
    dev = lambda a: cuda.from_numpy(np.ascontiguousarray(a)).cuda()
    colptr, rows, vals = csr_to_csc_device(m, n, dev(csr.indptr.astype(np.int64)), dev(csr.indices.astype(np.int32)), dev(csr.data))
    csc = csr.tocsc()
    csc.sort_indices()
    assert np.array_equal(colptr.cpu().numpy(), csc.indptr.astype(np.int64))
    assert np.array_equal(rows.cpu().numpy(), csc.indices) and np.array_equal(vals.cpu().numpy(), csc.data)
