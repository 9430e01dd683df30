"""oracle/oracle.py -- loaders for the checker.  TEST INFRASTRUCTURE ONLY.

* `cpu()`  : ctypes view of liboracle_als.so (oracle/als_cpu.c, the CPU restatement).
* `ref(v)` : ctypes view of oracle/_ref/libref_als_{cg,lu}.so -- the REAL reference
             (unmodified sources, shim-compiled for sm_100a by oracle/build_ref.sh).
             Needs a GPU; device pointers are raw integers (e.g. torch .data_ptr()).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
CPU_LIB = HERE / "liboracle_als.so"
REF_DIR = HERE / "_ref"

_vp = C.c_void_p
_cpu = None
_ref = {}


def build_cpu(force: bool = False) -> Path:
    if force or not CPU_LIB.exists() or CPU_LIB.stat().st_mtime < (HERE / "als_cpu.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-B", "liboracle_als.so"], check=True, capture_output=True)
    return CPU_LIB


def cpu() -> C.CDLL:
    global _cpu
    if _cpu is None:
        build_cpu()
        lib = C.CDLL(str(CPU_LIB))
        lib.oracle_gram.argtypes = [C.c_int, C.c_int, _vp, _vp, _vp, C.c_float, C.c_int, C.c_int, _vp]
        lib.oracle_rhs.argtypes = [C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp]
        lib.oracle_cg.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, C.c_float, C.c_int]
        lib.oracle_lu.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int]
        lib.oracle_rmse.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_long, C.c_int, C.c_int]
        lib.oracle_rmse.restype = C.c_float
        lib.oracle_batch_range.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.oracle_half_step.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_float,
                                         C.c_int, C.c_float]
        lib.oracle_doALS.argtypes = [_vp] * 12 + [C.c_int, C.c_int, C.c_int, C.c_long, C.c_long, C.c_float, C.c_int,
                                                  C.c_int, _vp]
        lib.oracle_doALS.restype = C.c_float
        _cpu = lib
    return _cpu


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ---- numpy-level wrappers of the CPU restatement ------------------------------------------
def gram(rowptr, colidx, factor, f, lam, batch_offset=0, batch_size=None):
    rowptr, colidx, factor = _c(rowptr, np.int32), _c(colidx, np.int32), _c(factor, np.float32)
    m = rowptr.size - 1
    if batch_size is None:
        batch_size = m - batch_offset
    tt = np.zeros((batch_size, f, f), np.float32)
    cpu().oracle_gram(batch_offset, batch_size, _p(tt), _p(rowptr), _p(colidx), lam, m, f, _p(factor))
    return tt


def rhs(rowptr, colidx, val, factor, f):
    rowptr, colidx, val, factor = _c(rowptr, np.int32), _c(colidx, np.int32), _c(val, np.float32), _c(factor, np.float32)
    rows = rowptr.size - 1
    out = np.zeros((rows, f), np.float32)
    cpu().oracle_rhs(rows, _p(out), _p(rowptr), _p(colidx), _p(val), f, _p(factor))
    return out


def cg(A, x0, b, f, cg_iter=6.0, fused_fma=True):
    A, b = _c(A, np.float32), _c(b, np.float32)
    x = np.array(x0, dtype=np.float32, order="C", copy=True)
    batch = b.size // f
    cpu().oracle_cg(_p(A), _p(x), _p(b), batch, f, cg_iter, int(fused_fma))
    return x


def lu(A, b, f):
    A = np.array(A, dtype=np.float32, order="C", copy=True)
    b = _c(b, np.float32)
    x = np.zeros_like(b)
    cpu().oracle_lu(_p(A), _p(x), _p(b), b.size // f, f)
    return x


def rmse(val, row, col, thetaT, XT, f, drop_tail=False):
    val, row, col = _c(val, np.float32), _c(row, np.int32), _c(col, np.int32)
    thetaT, XT = _c(thetaT, np.float32), _c(XT, np.float32)
    return float(cpu().oracle_rmse(_p(val), _p(row), _p(col), _p(thetaT), _p(XT), val.size, f, int(drop_tail)))


def batch_range(rows, nbatch, batch_id):
    a, b = C.c_int(0), C.c_int(0)
    cpu().oracle_batch_range(rows, nbatch, batch_id, C.byref(a), C.byref(b))
    return a.value, b.value


def half_step(rowptr, colidx, val, factor, out, f, lam, solver=0, cg_iter=6.0, row_begin=0, row_end=None):
    """In place on `out` (float32, C-contiguous)."""
    rowptr, colidx, val, factor = _c(rowptr, np.int32), _c(colidx, np.int32), _c(val, np.float32), _c(factor, np.float32)
    rows = rowptr.size - 1
    if row_end is None:
        row_end = rows
    assert out.dtype == np.float32 and out.flags.c_contiguous
    cpu().oracle_half_step(_p(rowptr), _p(colidx), _p(val), rows, row_begin, row_end, _p(factor), _p(out), f, lam,
                           solver, cg_iter)


def do_als(r, thetaT, XT, f, lam, iters, solver=0):
    """r: cumf_als_b200.data.Ratings-like.  thetaT/XT updated in place.  Returns
    (final_test_rmse, rmse[iters][2] = (train, test))."""
    arrs = [_c(r.csr_indptr, np.int32), _c(r.csr_indices, np.int32), _c(r.csr_data, np.float32),
            _c(r.csc_indices, np.int32), _c(r.csc_indptr, np.int32), _c(r.csc_data, np.float32),
            _c(r.coo_row, np.int32)]
    tarrs = [_c(r.test_row, np.int32), _c(r.test_col, np.int32), _c(r.test_val, np.float32)]
    assert thetaT.dtype == np.float32 and XT.dtype == np.float32
    out = np.zeros((iters, 2), np.float32)
    fin = cpu().oracle_doALS(*[_p(a) for a in arrs], _p(thetaT), _p(XT), *[_p(a) for a in tarrs], r.m, r.n, f,
                             r.csr_indices.size, r.test_val.size, lam, iters, solver, _p(out))
    return float(fin), out


# ---- the real reference (GPU only) --------------------------------------------------------
DOALS_SYMBOL = "_Z5doALSPKiS0_PKfS0_S0_S2_S0_PfS3_S0_S0_S2_iiillfiiii"


def ref_available(variant: str = "cg") -> bool:
    return (REF_DIR / f"libref_als_{variant}.so").exists()


def ref(variant: str = "cg") -> C.CDLL:
    """The reference library: variant 'cg' (as shipped) or 'lu' (als.cu:28 disabled)."""
    if variant not in _ref:
        path = REF_DIR / f"libref_als_{variant}.so"
        if not path.exists():
            raise FileNotFoundError(f"{path} missing: run oracle/build_ref.sh where /root/reference exists")
        lib = C.CDLL(str(path))
        fn = getattr(lib, DOALS_SYMBOL)
        fn.restype = C.c_float
        fn.argtypes = [_vp] * 12 + [C.c_int, C.c_int, C.c_int, C.c_long, C.c_long, C.c_float, C.c_int, C.c_int,
                                    C.c_int, C.c_int]
        lib.ref_get_hermitian.argtypes = [C.c_int, C.c_int, _vp, _vp, _vp, C.c_float, C.c_int, C.c_int, _vp]
        lib.ref_rhs.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]
        lib.ref_cg.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, C.c_float]
        lib.ref_lu.argtypes = [C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.ref_rmse.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int]
        lib.ref_rmse.restype = C.c_float
        _ref[variant] = lib
    return _ref[variant]


def ref_do_als(r, thetaT, XT, f, lam, iters, x_batch=1, theta_batch=1, variant="cg", device=0):
    """Call the reference's own doALS (host pointers).  thetaT/XT updated in place."""
    arrs = [_c(r.csr_indptr, np.int32), _c(r.csr_indices, np.int32), _c(r.csr_data, np.float32),
            _c(r.csc_indices, np.int32), _c(r.csc_indptr, np.int32), _c(r.csc_data, np.float32),
            _c(r.coo_row, np.int32)]
    tarrs = [_c(r.test_row, np.int32), _c(r.test_col, np.int32), _c(r.test_val, np.float32)]
    assert thetaT.dtype == np.float32 and XT.dtype == np.float32
    fn = getattr(ref(variant), DOALS_SYMBOL)
    return float(fn(*[_p(a) for a in arrs], _p(thetaT), _p(XT), *[_p(a) for a in tarrs], r.m, r.n, f,
                    r.csr_indices.size, r.test_val.size, lam, iters, x_batch, theta_batch, device))
