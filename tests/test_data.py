"""Synthetic data generator and the CLI's .bin format (host logic, CPU)."""
import numpy as np
import pytest
import scipy.sparse as sp

import cumf_als_b200 as c
from cumf_als_b200.data import Ratings, nnz_balanced_ranges, read_bin_dir, synth_ratings, write_bin_dir


@pytest.fixture(scope="module")
def ratings():
    return synth_ratings(150, 260, 7000, 900, seed=9)


def test_shape_and_coverage(ratings):
    r = ratings
    assert r.nnz == 7000 and r.nnz_test == 900
    assert r.csr_indptr[0] == 0 and r.csr_indptr[-1] == 7000
    assert (np.diff(r.csr_indptr) >= 1).all(), "every row needs a training rating (README.md:113)"
    assert (np.diff(r.csc_indptr) >= 1).all(), "every column needs a training rating"
    assert set(np.unique(r.csr_data)) <= {1.0, 2.0, 3.0, 4.0, 5.0}


def test_sorted_unique_columns(ratings):
    r = ratings
    for u in range(r.m):
        cols = r.csr_indices[r.csr_indptr[u]:r.csr_indptr[u + 1]]
        assert (np.diff(cols) > 0).all()


def test_csr_csc_coo_consistent(ratings):
    r = ratings
    A = sp.csr_matrix((r.csr_data, r.csr_indices, r.csr_indptr), shape=(r.m, r.n))
    B = sp.csc_matrix((r.csc_data, r.csc_indices, r.csc_indptr), shape=(r.m, r.n))
    assert abs(A - B).max() == 0
    # COO rows are the CSR expansion
    assert (r.coo_row == np.repeat(np.arange(r.m), np.diff(r.csr_indptr))).all()


def test_deterministic():
    a = synth_ratings(40, 50, 600, 80, seed=1)
    b = synth_ratings(40, 50, 600, 80, seed=1)
    assert (a.csr_indices == b.csr_indices).all() and (a.csr_data == b.csr_data).all()
    assert (a.test_row == b.test_row).all()


def test_bin_roundtrip_through_library_loaders(ratings, tmp_path, lib):
    """write_bin_dir -> the library's loaders (drop-in for host_utilities.cpp:19-98)."""
    r = ratings
    write_bin_dir(tmp_path, r)
    back = read_bin_dir(tmp_path, r.m, r.n)
    assert (back.csr_indices == r.csr_indices).all() and (back.csc_data == r.csc_data).all()
    import ctypes
    vp = ctypes.c_void_p
    data = np.zeros(r.nnz, np.float32)
    row = np.zeros(r.m + 1, np.int32)
    col = np.zeros(r.nnz, np.int32)
    p = lambda name: str(tmp_path / name).encode()
    rc = lib.cumf_load_csr_bin(p("R_train_csr.data.bin"), p("R_train_csr.indptr.bin"), p("R_train_csr.indices.bin"),
                               data.ctypes.data_as(vp), row.ctypes.data_as(vp), col.ctypes.data_as(vp), r.m, r.nnz)
    assert rc == 0
    assert (row == r.csr_indptr).all() and (col == r.csr_indices).all() and (data == r.csr_data).all()
    crow = np.zeros(r.nnz, np.int32)
    cptr = np.zeros(r.n + 1, np.int32)
    rc = lib.cumf_load_csc_bin(p("R_train_csc.data.bin"), p("R_train_csc.indices.bin"), p("R_train_csc.indptr.bin"),
                               data.ctypes.data_as(vp), crow.ctypes.data_as(vp), cptr.ctypes.data_as(vp), r.n, r.nnz)
    assert rc == 0 and (crow == r.csc_indices).all() and (cptr == r.csc_indptr).all()
    trow = np.zeros(r.nnz_test, np.int32)
    tcol = np.zeros(r.nnz_test, np.int32)
    tval = np.zeros(r.nnz_test, np.float32)
    rc = lib.cumf_load_coo_bin(p("R_test_coo.data.bin"), p("R_test_coo.row.bin"), p("R_test_coo.col.bin"),
                               tval.ctypes.data_as(vp), trow.ctypes.data_as(vp), tcol.ctypes.data_as(vp), r.nnz_test)
    assert rc == 0 and (trow == r.test_row).all() and (tval == r.test_val).all()
    # missing / short files are reported (the reference prints and carries on, host_utilities.cpp:27-31)
    assert lib.cumf_load_coo_row_bin(p("missing.bin"), trow.ctypes.data_as(vp), 5) == -1
    assert lib.cumf_load_coo_row_bin(p("R_test_coo.row.bin"), np.zeros(r.nnz_test + 8, np.int32).ctypes.data_as(vp),
                                     r.nnz_test + 8) == -1


def test_nnz_balanced_ranges(ratings):
    r = ratings
    for parts in (1, 2, 3, 8):
        ranges = nnz_balanced_ranges(r.csr_indptr, parts)
        assert ranges[0][0] == 0 and ranges[-1][1] == r.m
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        loads = [int(r.csr_indptr[hi] - r.csr_indptr[lo]) for lo, hi in ranges]
        assert sum(loads) == r.nnz
        if parts > 1:
            assert max(loads) <= r.nnz / parts + np.diff(r.csr_indptr).max()


def test_sharded_bin_loader_reads_exactly_the_slice(tmp_path):
    """cumf_load_csr_shard_bin: rows [b, e) of the CLI's file set, read with seeks (no whole-matrix pass)."""
    import cumf_als_b200 as c
    from cumf_als_b200.api import load_csr_shard
    r = synth_ratings(200, 300, 5000, 400, seed=21)
    write_bin_dir(tmp_path, r)
    for b, e in ((0, 200), (0, 0), (17, 93), (199, 200), (50, 50)):
        ptr, idx, val = load_csr_shard(tmp_path / "R_train_csr.data.bin", tmp_path / "R_train_csr.indptr.bin",
                                       tmp_path / "R_train_csr.indices.bin", r.m, b, e)
        lo, hi = int(r.csr_indptr[b]), int(r.csr_indptr[e])
        assert ptr.dtype == np.int64 and np.array_equal(ptr, r.csr_indptr[b:e + 1].astype(np.int64) - lo)
        assert np.array_equal(idx, r.csr_indices[lo:hi]) and np.array_equal(val, r.csr_data[lo:hi])
    # the CSC files are sliced the same way (columns instead of rows)
    ptr, idx, val = load_csr_shard(tmp_path / "R_train_csc.data.bin", tmp_path / "R_train_csc.indptr.bin",
                                   tmp_path / "R_train_csc.indices.bin", r.n, 100, 250)
    lo, hi = int(r.csc_indptr[100]), int(r.csc_indptr[250])
    assert np.array_equal(idx, r.csc_indices[lo:hi]) and np.array_equal(val, r.csc_data[lo:hi])
    with pytest.raises(c.CumfError):
        load_csr_shard(tmp_path / "R_train_csr.data.bin", tmp_path / "nope.bin", tmp_path / "R_train_csr.indices.bin", r.m, 0, 10)
