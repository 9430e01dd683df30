"""ctypes bindings of include/cumf_als.h.

Host-side mirror of the reference's interface for the hot path:
  do_als(...)          <-> float doALS(...)                  als.h:676-681
  gram(...)            <-> get_hermitian* launches            als.cu:804, 816
  cg(...)              <-> updateXWithCGHost                  cg.h:30
  lu(...)              <-> updateX / updateTheta (LU)         als.cu:58-189
  rmse(...)            <-> RMSE kernel + Sasum                als.cu:191-219, 979-1019
Device buffers are passed as objects exposing `.data_ptr()` (torch CUDA tensors --
torch is plumbing for device memory only) or as raw integer addresses.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

import numpy as np

SOLVER_CG, SOLVER_LU = 0, 1
PATH_AUTO, PATH_SIMT, PATH_TC = 0, 1, 2

_PKG = Path(__file__).resolve().parent
_LIB: Optional[C.CDLL] = None


class CumfError(RuntimeError):
    """A C-ABI call returned a negative status."""


def library_path() -> Path:
    return Path(os.environ.get("CUMF_ALS_LIB", _PKG / "libcumf_als_b200.so"))


_i32p = C.POINTER(C.c_int)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/cumf_als.h declares
SIGNATURES = {
    "cumf_last_error": (C.c_char_p, []),
    "cumf_version": (C.c_int, []),
    "cumf_doALS": (C.c_float, [_vp] * 12 + [C.c_int, C.c_int, C.c_int, C.c_long, C.c_long, C.c_float,
                                          C.c_int, C.c_int, C.c_int, C.c_int]),
    "cumf_load_csr_bin": (C.c_int, [C.c_char_p] * 3 + [_vp, _vp, _vp, C.c_int, C.c_long]),
    "cumf_load_csc_bin": (C.c_int, [C.c_char_p] * 3 + [_vp, _vp, _vp, C.c_int, C.c_long]),
    "cumf_load_coo_row_bin": (C.c_int, [C.c_char_p, _vp, C.c_long]),
    "cumf_load_coo_bin": (C.c_int, [C.c_char_p] * 3 + [_vp, _vp, _vp, C.c_long]),
    "cumf_gram": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_int, C.c_int, _vp,
                            C.c_int, _vp]),
    "cumf_cg": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_float, _vp]),
    "cumf_lu": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "cumf_rmse": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_long, C.c_int, C.c_int, _f32p, _f64p, _vp]),
    "cumf_plan_create": (C.c_int, [C.POINTER(_vp), _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cumf_plan_create_ranges": (C.c_int, [C.POINTER(_vp), _vp, _vp, C.c_int, C.c_int, C.c_int]),
    "cumf_plan_gram": (C.c_int, [_vp, _vp, _vp, _vp, C.c_float, _vp, _vp, _vp]),
    "cumf_plan_destroy": (C.c_int, [_vp]),
    "cumf_plan_last_launches": (C.c_int, [_vp]),
    "cumf_init_factors": (None, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_long]),
    "cumf_release_cached_memory": (C.c_int, []),
    "cumf_plan_set_factor_rows": (C.c_int, [_vp, C.c_int]),
    "cumf_update_factor": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_float, C.c_int, C.c_float, _vp]),
    "cumf_als_create": (C.c_int, [C.POINTER(_vp)] + [_vp] * 10 + [C.c_int, C.c_int, C.c_int, C.c_long, C.c_long,
                                                                   C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                                                   C.c_int, C.c_int, C.c_int]),
    "cumf_als_create64": (C.c_int, [C.POINTER(_vp)] + [_vp] * 10 + [C.c_int, C.c_int, C.c_int, C.c_long, C.c_long,
                                                                     C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                                                     C.c_int, C.c_int, C.c_int]),
    "cumf_als_create_device": (C.c_int, [C.POINTER(_vp)] + [_vp] * 9 + [C.c_long, C.c_int, C.c_int, C.c_int, C.c_long, C.c_long,
                                                                         C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                                                         C.c_int, C.c_int, C.c_int]),
    "cumf_plan_create64": (C.c_int, [C.POINTER(_vp), _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cumf_als_destroy": (C.c_int, [_vp]),
    "cumf_als_collect_train_sse": (C.c_int, [_vp, C.c_int]),
    "cumf_als_set_factors": (C.c_int, [_vp, _vp, _vp]),
    "cumf_als_get_factors": (C.c_int, [_vp, _vp, _vp]),
    "cumf_als_theta_ptr": (_vp, [_vp]),
    "cumf_als_x_ptr": (_vp, [_vp]),
    "cumf_als_update_x": (C.c_int, [_vp, _vp]),
    "cumf_als_update_theta": (C.c_int, [_vp, _vp]),
    "cumf_als_sse": (C.c_int, [_vp, _f64p, _f64p, _vp]),
    "cumf_als_iterate": (C.c_int, [_vp, C.c_int, _f32p, _vp]),
    "cumf_als_timers": (C.c_int, [_vp, _f64p, C.c_int]),
    "cumf_als_ipc_blob_bytes": (C.c_int, []),
    "cumf_als_ipc_export": (C.c_int, [_vp, _vp]),
    "cumf_als_ipc_import": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "cumf_als_peer_barrier": (C.c_int, [_vp, _vp]),
    "cumf_group_create": (C.c_int, [C.POINTER(_vp)] + [_vp] * 10 + [C.c_int, C.c_int, C.c_int, C.c_long, C.c_long,
                                                                     C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cumf_group_destroy": (C.c_int, [_vp]),
    "cumf_group_size": (C.c_int, [_vp]),
    "cumf_group_shard": (_vp, [_vp, C.c_int]),
    "cumf_group_set_factors": (C.c_int, [_vp, _vp, _vp]),
    "cumf_group_get_factors": (C.c_int, [_vp, _vp, _vp]),
    "cumf_group_iterate": (C.c_int, [_vp, C.c_int, _f32p]),
    "cumf_group_collect_train_sse": (C.c_int, [_vp, C.c_int]),
    "cumf_group_sse": (C.c_int, [_vp, _f64p, _f64p]),
    "cumf_als_shape": (C.c_int, [_vp, _i32p, _i32p, _i32p]),
    "cumf_synth_create": (C.c_int, [C.POINTER(_vp), C.c_longlong, C.c_int, C.c_float, C.c_ulonglong, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_long, C.c_int]),
    "cumf_synth_destroy": (C.c_int, [_vp]),
    "cumf_synth_slice": (C.c_longlong, [_vp, C.c_int, _vp]),
    "cumf_synth_total_nnz": (C.c_longlong, [_vp]),
    "cumf_synth_download": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "cumf_synth_download_test": (C.c_int, [_vp, _vp, _vp, _vp]),
    "cumf_synth_solver": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, C.c_float, C.c_long, C.c_int, C.c_int]),
    "cumf_als_init_factors_device": (C.c_int, [_vp, C.c_ulonglong, C.c_float]),
    "cumf_group_create_synth": (C.c_int, [C.POINTER(_vp), C.c_longlong, C.c_int, C.c_float, C.c_ulonglong, C.c_long, C.c_int,
                                          C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cumf_csr_to_csc_device": (C.c_int, [C.c_int, C.c_int, C.c_longlong, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cumf_bin_shard_extent": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "cumf_load_bin_slice": (C.c_int, [C.c_char_p, C.c_int, C.c_longlong, C.c_longlong, _vp]),
    "cumf_load_csr_shard_bin": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "cumf_group_nnz": (C.c_long, [_vp]),
    "cumf_group_nnz_test": (C.c_long, [_vp]),
}
# C++-linkage symbols the reference's main.cpp / als_tf.cc bind (als.h:676-681, host_utilities.h:31-40)
MANGLED_SYMBOLS = [
    "_Z5doALSPKiS0_PKfS0_S0_S2_S0_PfS3_S0_S0_S2_iiillfiiii",
    "_Z22loadCSRSparseMatrixBinPKcS0_S0_PfPiS2_il",
    "_Z22loadCSCSparseMatrixBinPKcS0_S0_PfPiS2_il",
    "_Z28loadCooSparseMatrixRowPtrBinPKcPil",
    "_Z22loadCooSparseMatrixBinPKcS0_S0_PfPiS2_l",
]


def load_library() -> C.CDLL:
    """Load libcumf_als_b200.so (built in-tree by cumf_als_b200.build).  Fails loudly."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not path.exists():
        raise CumfError(
            f"{path} is missing: build it with `python -m cumf_als_b200.build` "
            "(there is no CPU / PyTorch fallback for this path)")
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().cumf_last_error()
        raise CumfError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")


def _dptr(x) -> Optional[int]:
    """Device address of a torch CUDA tensor / raw int / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        if hasattr(x, "is_cuda") and not x.is_cuda:
            raise CumfError("expected a CUDA tensor")
        if hasattr(x, "is_contiguous") and not x.is_contiguous():
            raise CumfError("expected a contiguous tensor")
        return x.data_ptr()
    raise TypeError(f"cannot take a device pointer from {type(x)!r}")


def _stream_ptr(stream) -> Optional[int]:
    if stream is None:
        return None
    return stream if isinstance(stream, int) else stream.cuda_stream


def _host(a: np.ndarray, dtype) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def _hp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_vp)


# --------------------------------------------------------------------------------------------
# b1: doALS
# --------------------------------------------------------------------------------------------
def do_als(csr_indptr, csr_indices, csr_data, csc_indices, csc_indptr, csc_data, coo_row, thetaT, XT,
           test_row, test_col, test_val, m: int, n: int, f: int, nnz: int, nnz_test: int, lam: float,
           iters: int, x_batch: int = 1, theta_batch: int = 1, device: int = 0) -> float:
    """float doALS(...) (als.h:676-681), same argument order.  `thetaT` (n*f) and `XT` (m*f)
    are float32 numpy arrays updated IN PLACE; returns the last test RMSE.

    Note the reference's CSC argument order: row ids (nnz) first, then the column
    pointer array (n+1) (main.cpp:141-146)."""
    lib = load_library()
    arrs = [
        _host(csr_indptr, np.int32), _host(csr_indices, np.int32), _host(csr_data, np.float32),
        _host(csc_indices, np.int32), _host(csc_indptr, np.int32), _host(csc_data, np.float32),
        _host(coo_row, np.int32),
    ]
    if thetaT.dtype != np.float32 or XT.dtype != np.float32 or not thetaT.flags.c_contiguous or not XT.flags.c_contiguous:
        raise CumfError("thetaT and XT must be C-contiguous float32 arrays (they are written in place)")
    if thetaT.size != n * f or XT.size != m * f:
        raise CumfError("thetaT must hold n*f and XT m*f values")
    tarrs = [_host(test_row, np.int32), _host(test_col, np.int32), _host(test_val, np.float32)]
    return float(lib.cumf_doALS(*[_hp(a) for a in arrs], _hp(thetaT), _hp(XT), *[_hp(a) for a in tarrs],
                                m, n, f, nnz, nnz_test, lam, iters, x_batch, theta_batch, device))


# --------------------------------------------------------------------------------------------
# b4: stage seams on device pointers
# --------------------------------------------------------------------------------------------
def gram(batch_offset: int, batch_size: int, tt, rowptr, colidx, lam: float, m: int, f: int, factor,
         rhs=None, val=None, path: int = PATH_SIMT, stream=None) -> None:
    """get_hermitian*<<<batch_size,...>>>(batch_offset, tt, rowPtr, colIdx, lambda, m, F, factor)
    (als.cu:804, 816) and, when `rhs` is given, the RHS pass (als.cu:750-757)."""
    _check(load_library().cumf_gram(batch_offset, batch_size, _dptr(tt), _dptr(rhs), _dptr(rowptr), _dptr(colidx),
                                    _dptr(val), lam, m, f, _dptr(factor), path, _stream_ptr(stream)), "cumf_gram")


def cg(A, x, b, batch: int, f: int, cg_iter: float = 6.0, stream=None) -> None:
    """updateXWithCGHost(A, x, b, batchSize, f, cgIter) (cg.h:30); x is updated in place."""
    _check(load_library().cumf_cg(_dptr(A), _dptr(x), _dptr(b), batch, f, cg_iter, _stream_ptr(stream)), "cumf_cg")


def lu(A, x, b, batch: int, f: int, stream=None) -> None:
    """LU oracle (als.cu:58-122): A and b are overwritten, the solution is copied into x."""
    _check(load_library().cumf_lu(_dptr(A), _dptr(x), _dptr(b), batch, f, _stream_ptr(stream)), "cumf_lu")


def rmse(val, row, col, thetaT, XT, count: int, f: int, drop_tail: bool = False, stream=None):
    """RMSE over `count` samples (als.cu:191-219, 979-1019).  Returns (rmse, sse)."""
    r, s = C.c_float(0), C.c_double(0)
    _check(load_library().cumf_rmse(_dptr(val), _dptr(row), _dptr(col), _dptr(thetaT), _dptr(XT), count, f,
                                    int(drop_tail), C.byref(r), C.byref(s), _stream_ptr(stream)), "cumf_rmse")
    return r.value, s.value


class Plan:
    """Work decomposition of one half-step over rows [row_begin,row_end) (cumf_plan)."""

    def __init__(self, rowptr_host: np.ndarray, row_begin: int = 0, row_end: Optional[int] = None, f: int = 100,
                 path: int = PATH_AUTO):
        self._h = _vp()
        rp = _host(rowptr_host, np.int32)
        rows = rp.size - 1
        self.row_begin, self.row_end, self.f = row_begin, rows if row_end is None else row_end, f
        self.base = int(rp[row_begin])
        _check(load_library().cumf_plan_create(C.byref(self._h), _hp(rp), rows, row_begin, self.row_end, f, path),
               "cumf_plan_create")

    @classmethod
    def from_ranges(cls, begin: np.ndarray, end: np.ndarray, f: int = 100, path: int = PATH_AUTO) -> "Plan":
        """Partial-Gram plan (cumf_plan_create_ranges): row u covers ratings [begin[u], end[u])."""
        b, e = _host(begin, np.int64), _host(end, np.int64)
        if b.shape != e.shape or b.ndim != 1:
            raise ValueError("begin/end must be 1-D arrays of the same length")
        self = cls.__new__(cls)
        self._h = _vp()
        self.row_begin, self.row_end, self.f, self.base = 0, b.size, f, 0
        _check(load_library().cumf_plan_create_ranges(C.byref(self._h), _hp(b), _hp(e), b.size, f, path),
               "cumf_plan_create_ranges")
        return self

    def gram(self, colidx, val, factor, lam: float, tt, rhs, stream=None) -> None:
        """Partial [A|b] of every row over the plan's ranges into tt [rows,f*f] / rhs [rows,f]
        (cumf_plan_gram; asynchronous on the stream)."""
        _check(load_library().cumf_plan_gram(self._h, _dptr(colidx), _dptr(val), _dptr(factor), lam, _dptr(tt),
                                             _dptr(rhs), _stream_ptr(stream)), "cumf_plan_gram")

    @property
    def last_launches(self) -> int:
        return load_library().cumf_plan_last_launches(self._h)

    def set_factor_rows(self, rows: int) -> None:
        """Optional hint (cumf_plan_set_factor_rows): rows of the opposing factor, saves the first launch's index scan."""
        _check(load_library().cumf_plan_set_factor_rows(self._h, int(rows)), "cumf_plan_set_factor_rows")

    def close(self) -> None:
        if self._h:
            load_library().cumf_plan_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def update_factor(plan: Plan, colidx, val, factor, out, lam: float, solver: int = SOLVER_CG, cg_iter: float = 6.0,
                  stream=None) -> None:
    """One half-step (als.cu:727-853 / 858-961) for the plan's rows.  `colidx`/`val` are device
    tensors holding the FULL index/value arrays; the slice starting at the plan's first
    rating is passed down."""
    ci, va = _dptr(colidx), _dptr(val)
    _check(load_library().cumf_update_factor(plan._h, ci + 4 * plan.base, va + 4 * plan.base, _dptr(factor),
                                             _dptr(out), lam, solver, cg_iter, _stream_ptr(stream)),
           "cumf_update_factor")


# --------------------------------------------------------------------------------------------
# resident solver handle
# --------------------------------------------------------------------------------------------
class AlsSolver:
    """Resident ALS state on one GPU (cumf_als_solver): CSR/CSC/COO uploaded once, factors on
    the device, optional row shard for one-process-per-GPU runs."""

    def __init__(self, csr_indptr, csr_indices, csr_data, csc_indices, csc_indptr, csc_data, coo_row,
                 test_row, test_col, test_val, m: int, n: int, f: int, lam: float,
                 x_range=None, theta_range=None, device: int = 0, solver: int = SOLVER_CG, path: int = PATH_AUTO):
        lib = load_library()
        self.m, self.n, self.f, self.lam = m, n, f, lam
        self.nnz = int(np.asarray(csr_indptr)[-1])
        self.nnz_test = 0 if test_val is None else int(np.asarray(test_val).size)
        xb, xe = x_range if x_range is not None else (0, m)
        tb, te = theta_range if theta_range is not None else (0, n)
        self.x_range, self.theta_range = (xb, xe), (tb, te)
        arrs = [
            _host(csr_indptr, np.int32), _host(csr_indices, np.int32), _host(csr_data, np.float32),
            _host(csc_indices, np.int32), _host(csc_indptr, np.int32), _host(csc_data, np.float32),
            None if coo_row is None else _host(coo_row, np.int32),
            None if test_row is None else _host(test_row, np.int32),
            None if test_col is None else _host(test_col, np.int32),
            None if test_val is None else _host(test_val, np.float32),
        ]
        self._h = _vp()
        _check(lib.cumf_als_create(C.byref(self._h), *[_hp(a) for a in arrs], m, n, f, self.nnz, self.nnz_test, lam,
                                   xb, xe, tb, te, device, solver, path), "cumf_als_create")

    def collect_train_sse(self, on: bool = True) -> bool:
        """Train SSE as a by-product of update_theta (cumf_als_collect_train_sse); returns whether it is active."""
        return load_library().cumf_als_collect_train_sse(self._h, int(on)) == 1

    def set_factors(self, thetaT: np.ndarray, XT: np.ndarray) -> None:
        t, x = _host(thetaT, np.float32), _host(XT, np.float32)
        assert t.size == self.n * self.f and x.size == self.m * self.f
        _check(load_library().cumf_als_set_factors(self._h, _hp(t), _hp(x)), "cumf_als_set_factors")

    def get_factors(self):
        t = np.empty((self.n, self.f), np.float32)
        x = np.empty((self.m, self.f), np.float32)
        _check(load_library().cumf_als_get_factors(self._h, _hp(t), _hp(x)), "cumf_als_get_factors")
        return t, x

    @property
    def theta_ptr(self) -> int:
        return load_library().cumf_als_theta_ptr(self._h)

    @property
    def x_ptr(self) -> int:
        return load_library().cumf_als_x_ptr(self._h)

    def update_x(self, stream=None) -> None:
        _check(load_library().cumf_als_update_x(self._h, _stream_ptr(stream)), "cumf_als_update_x")

    def update_theta(self, stream=None) -> None:
        _check(load_library().cumf_als_update_theta(self._h, _stream_ptr(stream)), "cumf_als_update_theta")

    def sse(self, stream=None):
        a, b = C.c_double(0), C.c_double(0)
        _check(load_library().cumf_als_sse(self._h, C.byref(a), C.byref(b), _stream_ptr(stream)), "cumf_als_sse")
        return a.value, b.value

    def rmse(self):
        """(train_rmse, test_rmse) as the reference prints them (als.cu:991, 1018)."""
        tr, te = self.sse()
        f32 = np.float32
        train = float(np.sqrt(f32(tr) / f32(self.nnz))) if self.nnz else 0.0
        test = float(np.sqrt(f32(te) / f32(self.nnz_test))) if self.nnz_test else 0.0
        return train, test

    def iterate(self, iters: int = 1, stream=None) -> float:
        """`iters` ALS iterations; returns device-timed milliseconds (CUDA events)."""
        ms = C.c_float(0)
        _check(load_library().cumf_als_iterate(self._h, iters, C.byref(ms), _stream_ptr(stream)), "cumf_als_iterate")
        return ms.value

    def timers(self, reset: bool = False) -> dict:
        out = (C.c_double * 6)()
        _check(load_library().cumf_als_timers(self._h, out, int(reset)), "cumf_als_timers")
        keys = ["x_ms", "theta_ms", "gram_x_ms", "gram_theta_ms", "launches", "iterations"]
        return dict(zip(keys, list(out)))

    # ---- multi-GPU, one process per GPU: connect the ranks' replicas through CUDA IPC (include/cumf_als.h) ----
    def ipc_export(self) -> bytes:
        n = load_library().cumf_als_ipc_blob_bytes()
        buf = C.create_string_buffer(n)
        _check(load_library().cumf_als_ipc_export(self._h, buf), "cumf_als_ipc_export")
        return buf.raw

    def ipc_import(self, blobs, rank: int) -> None:
        """`blobs`: every rank's ipc_export() in rank order.  Afterwards update_x / update_theta also write the peers'
        replicas and iterate() ends every half-step with the device-side barrier."""
        joined = b"".join(blobs)
        _check(load_library().cumf_als_ipc_import(self._h, joined, len(blobs), rank), "cumf_als_ipc_import")

    def peer_barrier(self, stream=None) -> None:
        _check(load_library().cumf_als_peer_barrier(self._h, _stream_ptr(stream)), "cumf_als_peer_barrier")

    def close(self) -> None:
        if self._h:
            load_library().cumf_als_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_csr_shard(data_file, indptr_file, indices_file, rows: int, row_begin: int, row_end: int):
    """(ptr int64 rebased to 0, idx int32, val float32) of rows [row_begin, row_end) of a CSR (or columns of a CSC) .bin file
    set, read with seeks (cumf_load_csr_shard_bin)."""
    lib = load_library()
    first, count = C.c_longlong(0), C.c_longlong(0)
    if lib.cumf_bin_shard_extent(str(indptr_file).encode(), rows, row_begin, row_end, C.byref(first), C.byref(count)):
        raise CumfError(f"cannot read rows [{row_begin}, {row_end}) of {indptr_file}")
    ptr = np.empty(row_end - row_begin + 1, np.int64)
    idx, val = np.empty(count.value, np.int32), np.empty(count.value, np.float32)
    if lib.cumf_load_csr_shard_bin(str(data_file).encode(), str(indptr_file).encode(), str(indices_file).encode(), rows, row_begin,
                                   row_end, _hp(ptr), _hp(idx), _hp(val)):
        raise CumfError(f"cannot read the slice of {data_file} / {indices_file}")
    return ptr, idx, val


def csr_to_csc_device(rows: int, cols: int, rowptr, col, val, stream=None):
    """CSR -> CSC on the device (cumf_csr_to_csc_device); torch CUDA tensors in (rowptr int64), tensors out."""
    import torch
    nnz = int(col.numel())
    colptr = torch.empty(cols + 1, dtype=torch.int64, device=col.device)
    row_out = torch.empty(max(nnz, 1), dtype=torch.int32, device=col.device)[:nnz]
    val_out = torch.empty(max(nnz, 1), dtype=torch.float32, device=col.device)[:nnz]
    _check(load_library().cumf_csr_to_csc_device(rows, cols, nnz, _dptr(rowptr), _dptr(col), _dptr(val), _dptr(colptr), _dptr(row_out),
                                                 _dptr(val_out), _stream_ptr(stream)), "cumf_csr_to_csc_device")
    return colptr, row_out, val_out


class SynthShard:
    """cumf_synth_shard: CSR rows [x0,x1) and CSC columns [t0,t1) of the synthetic matrix (m, n, avg_deg, seed) on one device."""

    def __init__(self, m: int, n: int, avg_deg: float, seed: int, x_range, t_range, test_cnt: int = 0, device: int = 0):
        self._h = _vp()
        self.m, self.n, self.x_range, self.t_range, self.test_cnt = m, n, tuple(x_range), tuple(t_range), test_cnt
        _check(load_library().cumf_synth_create(C.byref(self._h), m, n, avg_deg, seed, x_range[0], x_range[1], t_range[0], t_range[1],
                                                test_cnt, device), "cumf_synth_create")

    @property
    def total_nnz(self) -> int:
        return int(load_library().cumf_synth_total_nnz(self._h))

    def slice(self, what: int):
        """(ptr int64 rebased, idx int32, val float32) of the CSR (0) / CSC (1) slice, on the host."""
        rows = (self.x_range[1] - self.x_range[0]) if what == 0 else (self.t_range[1] - self.t_range[0])
        ptr = np.empty(rows + 1, np.int64)
        cnt = int(load_library().cumf_synth_slice(self._h, what, _hp(ptr)))
        idx, val = np.empty(cnt, np.int32), np.empty(cnt, np.float32)
        _check(load_library().cumf_synth_download(self._h, what, _hp(idx), _hp(val)), "cumf_synth_download")
        return ptr, idx, val

    def test_samples(self):
        r, c_, v = np.empty(self.test_cnt, np.int32), np.empty(self.test_cnt, np.int32), np.empty(self.test_cnt, np.float32)
        _check(load_library().cumf_synth_download_test(self._h, _hp(r), _hp(c_), _hp(v)), "cumf_synth_download_test")
        return r, c_, v

    def close(self) -> None:
        if self._h:
            load_library().cumf_synth_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AlsGroup:
    """cumf_als_group: n row shards on n devices of this node, driven from one process (what cumf_doALS does under
    CUMF_GPUS=n).  Same host arrays as AlsSolver; factors are full replicas on every device."""

    def __init__(self, csr_indptr, csr_indices, csr_data, csc_indices, csc_indptr, csc_data, coo_row,
                 test_row, test_col, test_val, m: int, n: int, f: int, lam: float, n_devices: int, first_device: int = 0,
                 solver: int = SOLVER_CG, path: int = PATH_AUTO):
        lib = load_library()
        self.m, self.n, self.f, self.lam = m, n, f, lam
        self.nnz = int(np.asarray(csr_indptr)[-1])
        self.nnz_test = 0 if test_val is None else int(np.asarray(test_val).size)
        arrs = [
            _host(csr_indptr, np.int32), _host(csr_indices, np.int32), _host(csr_data, np.float32),
            _host(csc_indices, np.int32), _host(csc_indptr, np.int32), _host(csc_data, np.float32),
            None if coo_row is None else _host(coo_row, np.int32),
            None if test_row is None else _host(test_row, np.int32),
            None if test_col is None else _host(test_col, np.int32),
            None if test_val is None else _host(test_val, np.float32),
        ]
        self._h = _vp()
        _check(lib.cumf_group_create(C.byref(self._h), *[_hp(a) for a in arrs], m, n, f, self.nnz, self.nnz_test, lam,
                                     first_device, n_devices, solver, path), "cumf_group_create")

    @classmethod
    def from_synth(cls, m: int, n: int, avg_deg: float, seed: int, test_per_shard: int, f: int, lam: float, n_devices: int,
                   first_device: int = 0, solver: int = SOLVER_CG, path: int = PATH_AUTO) -> "AlsGroup":
        """cumf_group_create_synth: the matrix is generated shard by shard on the devices (Hugewiki-scale configuration)."""
        self = cls.__new__(cls)
        self._h = _vp()
        self.m, self.n, self.f, self.lam = m, n, f, lam
        _check(load_library().cumf_group_create_synth(C.byref(self._h), m, n, avg_deg, seed, test_per_shard, f, lam, first_device,
                                                      n_devices, solver, path), "cumf_group_create_synth")
        self.nnz = int(load_library().cumf_group_nnz(self._h))
        self.nnz_test = int(load_library().cumf_group_nnz_test(self._h))
        return self

    def set_factors(self, thetaT, XT) -> None:
        t, x = _host(thetaT, np.float32), _host(XT, np.float32)
        _check(load_library().cumf_group_set_factors(self._h, _hp(t), _hp(x)), "cumf_group_set_factors")

    def get_factors(self):
        t = np.empty((self.n, self.f), np.float32)
        x = np.empty((self.m, self.f), np.float32)
        _check(load_library().cumf_group_get_factors(self._h, _hp(t), _hp(x)), "cumf_group_get_factors")
        return t, x

    def collect_train_sse(self, on: bool = True) -> bool:
        return load_library().cumf_group_collect_train_sse(self._h, int(on)) == 1

    def iterate(self, iters: int = 1) -> float:
        ms = C.c_float(0)
        _check(load_library().cumf_group_iterate(self._h, iters, C.byref(ms)), "cumf_group_iterate")
        return ms.value

    def rmse(self):
        a, b = C.c_double(0), C.c_double(0)
        _check(load_library().cumf_group_sse(self._h, C.byref(a), C.byref(b)), "cumf_group_sse")
        f32 = np.float32
        return (float(np.sqrt(f32(a.value) / f32(self.nnz))) if self.nnz else 0.0,
                float(np.sqrt(f32(b.value) / f32(self.nnz_test))) if self.nnz_test else 0.0)

    def close(self) -> None:
        if self._h:
            load_library().cumf_group_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
