"""BASELINE.json's full-size configuration (Netflix-shaped, m=17770, n=480189, nnz=99072112, f=100) on the GPU, checked
through properties that do not need an oracle run over 99 M ratings (north_star: "at BASELINE's full sizes through
size-independent properties"):
  * a sample of rows of each half-step against the CPU oracle's half_step on exactly those rows (short, long, split
    across CTAs) -- the fused tcgen05 path against the exact-fp32 restatement;
  * rows untouched by a partial plan stay bit-identical; the half-step is deterministic (two runs, same bits);
  * train SSE three ways (by-product of the theta half-step, chunked streaming walk, literal COO pairs);
  * sharded (two row ranges on one GPU) == unsharded, bit for bit.
Run with -m gpu; about half a minute on a B200."""
import numpy as np
import pytest

import cumf_als_b200 as c
from conftest import rel_fro
from oracle import oracle as O

pytestmark = pytest.mark.gpu

F, LAM = 100, 0.048


@pytest.fixture(scope="module")
def netflix(cuda):
    import bench
    w = bench.WORKLOADS["netflix"]
    r, theta0, X0 = bench.make_inputs(w, 1.0, "cuda")
    assert (r.m, r.n, r.nnz, r.nnz_test) == (17770, 480189, 99072112, 1408395)
    s = c.AlsSolver(r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, r.coo_row,
                    r.test_row, r.test_col, r.test_val, r.m, r.n, F, LAM)
    s.set_factors(theta0, X0)
    s.collect_train_sse(True)
    s.iterate(1)                       # X0 = 0 would make the first theta systems diagonal: start from one real iteration
    th1, X1 = s.get_factors()
    yield r, s, th1, X1
    s.close()


def _sample_rows(indptr, rng, k_each=6):
    n = np.diff(indptr.astype(np.int64))
    order = np.argsort(n)
    picks = np.concatenate([order[:k_each], order[-k_each:], rng.choice(order, k_each, replace=False),
                            order[np.searchsorted(n[order], 8192):][:k_each]])     # shortest, longest, random, just-split
    return np.unique(picks)


def _oracle_rows(indptr, indices, data, factor, x_before, rows):
    """oracle half_step on a compacted CSR holding only `rows`."""
    ip = np.zeros(rows.size + 1, np.int32)
    ip[1:] = np.cumsum(np.diff(indptr)[rows])
    idx = np.concatenate([indices[indptr[u]:indptr[u + 1]] for u in rows])
    val = np.concatenate([data[indptr[u]:indptr[u + 1]] for u in rows])
    out = np.ascontiguousarray(x_before[rows])
    O.half_step(ip, idx, val, factor, out, F, LAM)
    return out


def test_fullsize_half_steps_sampled_rows_vs_oracle_and_determinism(netflix):
    r, s, th1, X1 = netflix
    rng = np.random.default_rng(7)
    s.set_factors(th1, X1)
    s.update_x()
    _, X2 = s.get_factors()
    rows = _sample_rows(r.csr_indptr, rng)
    want = _oracle_rows(r.csr_indptr, r.csr_indices, r.csr_data, th1, X1, rows)
    err = np.linalg.norm(X2[rows] - want, axis=1) / np.linalg.norm(want, axis=1)
    # six unconverged CG steps amplify the 1.5e-6 difference of the split-fp16 Gram: 1e-3 per row is the
    # spread the reference shows between its own runs; the RMSE-level bar (1e-4) is checked in test_train_sse_* below
    assert err.max() < 2e-3 and np.median(err) < 2e-4, (err.max(), np.median(err))
    s.update_theta()
    th2, _ = s.get_factors()
    cols = _sample_rows(r.csc_indptr, rng)
    want_t = _oracle_rows(r.csc_indptr, r.csc_indices, r.csc_data, X2, th1, cols)
    err_t = np.linalg.norm(th2[cols] - want_t, axis=1) / np.linalg.norm(want_t, axis=1)
    assert err_t.max() < 2e-3 and np.median(err_t) < 2e-4, (err_t.max(), np.median(err_t))
    # determinism: same inputs, same bits (split rows are reduced in slot order, no atomics anywhere)
    s.set_factors(th1, X1)
    s.update_x()
    s.update_theta()
    th2b, X2b = s.get_factors()
    assert np.array_equal(X2, X2b) and np.array_equal(th2, th2b)


def test_fullsize_train_sse_three_walks_agree(netflix, monkeypatch):
    r, s, th1, X1 = netflix
    s.set_factors(th1, X1)
    s.update_x()
    s.update_theta()
    by_product, test_sse = s.sse()                       # theta half-step just ran, X untouched since
    th, X = s.get_factors()
    import torch
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    literal = c.rmse(dev(r.csr_data), dev(r.coo_row), dev(r.csr_indices), dev(th), dev(X), r.nnz, F)[1]
    s.set_factors(th, X)                                 # invalidates the by-product: the chunked streaming walk runs
    streamed, test_sse2 = s.sse()
    assert streamed == pytest.approx(literal, rel=1e-9)  # same per-sample arithmetic, different order of the double sum
    assert test_sse2 == test_sse
    assert by_product == pytest.approx(literal, rel=1e-4)
    assert np.sqrt(by_product / r.nnz) == pytest.approx(np.sqrt(literal / r.nnz), rel=5e-5)
    # host float64 on a 1 % sample of the ratings pins the kernels themselves
    rng = np.random.default_rng(3)
    pick = rng.choice(r.nnz, 1_000_000, replace=False)
    pred = np.einsum("ij,ij->i", th[r.csr_indices[pick]].astype(np.float64), X[r.coo_row[pick]].astype(np.float64))
    host = float(((r.csr_data[pick] - pred) ** 2).sum())
    dpick = torch.from_numpy(pick).cuda()
    gpu = c.rmse(dev(r.csr_data)[dpick].contiguous(), dev(r.coo_row)[dpick].contiguous(), dev(r.csr_indices)[dpick].contiguous(),
                 dev(th), dev(X), pick.size, F)[1]
    assert gpu == pytest.approx(host, rel=2e-6)


def test_fullsize_row_shards_equal_whole(netflix):
    r, s, th1, X1 = netflix
    s.set_factors(th1, X1)
    s.update_x()
    _, X_whole = s.get_factors()
    from cumf_als_b200.data import nnz_balanced_ranges
    X_parts = X1.copy()
    for lo, hi in nnz_balanced_ranges(r.csr_indptr, 2):
        p = c.AlsSolver(r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, r.coo_row,
                        r.test_row, r.test_col, r.test_val, r.m, r.n, F, LAM, x_range=(lo, hi), theta_range=(0, 1))
        p.set_factors(th1, X1)
        p.update_x()
        _, Xp = p.get_factors()
        assert np.array_equal(Xp[:lo], X1[:lo]) and np.array_equal(Xp[hi:], X1[hi:])     # other rows untouched
        X_parts[lo:hi] = Xp[lo:hi]
        p.close()
    assert np.array_equal(X_parts, X_whole)
