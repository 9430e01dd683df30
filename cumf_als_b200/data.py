"""Synthetic rating matrices in the reference's on-disk/in-memory layout, and the .bin I/O.

No datasets can be downloaded here, so every configuration of BASELINE.json is a
*shaped* synthetic: exact (m, n, nnz, nnz_test), power-law row/column degrees, every
row and column has at least one training rating (the reference produces NaN factors
for empty rows, README.md:113), column ids inside a row unique and ascending (what
scipy's tocsr() gives prepare_netflix_data.py:98), ratings in {1..5} with low-rank
structure so ALS has something to fit, COO rows in CSR order.

torch is used as the array engine (CPU for tests, CUDA for Netflix-sized inputs);
results are deterministic for a given (seed, device type).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np
import torch

# BASELINE.json configs (m, n, nnz, nnz_test, f, lambda); seeds are 1000 + config id
CONFIGS = {
    "ml10m": dict(m=71567, n=65133, nnz=9000048, nnz_test=1000006, f=10, lam=0.05, seed=1001),
    "netflix": dict(m=17770, n=480189, nnz=99072112, nnz_test=1408395, f=100, lam=0.048, seed=1002),
    "netflix_f200": dict(m=17770, n=480189, nnz=99072112, nnz_test=1408395, f=200, lam=0.048, seed=1002),
    "yahoo": dict(m=1000990, n=624961, nnz=252800275, nnz_test=4003960, f=100, lam=1.4, seed=1004),
}


@dataclass
class Ratings:
    m: int
    n: int
    csr_indptr: np.ndarray    # int32 [m+1]
    csr_indices: np.ndarray   # int32 [nnz]  column ids
    csr_data: np.ndarray      # float32 [nnz]
    csc_indptr: np.ndarray    # int32 [n+1]
    csc_indices: np.ndarray   # int32 [nnz]  row ids
    csc_data: np.ndarray      # float32 [nnz]
    coo_row: np.ndarray       # int32 [nnz]  row id of every CSR entry
    test_row: np.ndarray      # int32 [nnz_test]
    test_col: np.ndarray      # int32 [nnz_test]
    test_val: np.ndarray      # float32 [nnz_test]

    @property
    def nnz(self) -> int:
        return int(self.csr_indices.size)

    @property
    def nnz_test(self) -> int:
        return int(self.test_val.size)

    def doals_args(self):
        """Positional host arrays in doALS order (als.h:676-681), factors excluded."""
        return (self.csr_indptr, self.csr_indices, self.csr_data, self.csc_indices, self.csc_indptr, self.csc_data,
                self.coo_row)


def _power_law_cdf(count: int, alpha: float, gen: torch.Generator, device) -> torch.Tensor:
    ranks = torch.arange(1, count + 1, dtype=torch.float64, device=device)
    w = ranks.pow(-alpha)
    perm = torch.randperm(count, generator=gen, device=device)
    w = w[perm]                      # popularity is not correlated with the id
    cdf = torch.cumsum(w, 0)
    return (cdf / cdf[-1]).to(torch.float64)


def _sample(cdf: torch.Tensor, count: int, gen: torch.Generator) -> torch.Tensor:
    u = torch.rand(count, generator=gen, device=cdf.device, dtype=torch.float64)
    return torch.searchsorted(cdf, u).clamp_(max=cdf.numel() - 1)


def _ratings_for(rows: torch.Tensor, cols: torch.Tensor, U: torch.Tensor, V: torch.Tensor, gen: torch.Generator):
    out = torch.empty(rows.numel(), dtype=torch.float32, device=rows.device)
    step = 1 << 24
    k = U.shape[1]
    for s in range(0, rows.numel(), step):
        r, c = rows[s:s + step], cols[s:s + step]
        score = (U[r] * V[c]).sum(1) / (k ** 0.5)
        noise = torch.randn(r.numel(), generator=gen, device=rows.device)
        out[s:s + step] = torch.clamp(torch.round(3.6 + 1.0 * score + 0.5 * noise), 1.0, 5.0)
    return out


def synth_ratings(m: int, n: int, nnz: int, nnz_test: int, seed: int = 0, device: str = "cpu",
                  alpha_row: float = 0.8, alpha_col: float = 0.6, rank: int = 8) -> Ratings:
    """Exactly `nnz` unique training entries and `nnz_test` test entries, see module docstring."""
    if nnz < max(m, n):
        raise ValueError("nnz must be >= max(m, n) so every row and column can hold a rating")
    if nnz > m * n:
        raise ValueError("nnz exceeds m*n")
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    row_cdf = _power_law_cdf(m, alpha_row, gen, dev)
    col_cdf = _power_law_cdf(n, alpha_col, gen, dev)

    # coverage: one entry per row and per column
    r0 = torch.arange(m, device=dev)
    c0 = _sample(col_cdf, m, gen)
    c1 = torch.arange(n, device=dev)
    r1 = _sample(row_cdf, n, gen)
    base = torch.unique(torch.cat([r0 * n + c0, r1 * n + c1]))
    need = nnz - base.numel()
    extra = torch.empty(0, dtype=torch.int64, device=dev)
    over = 1.12
    while need > 0:
        cnt = int(need * over) + 1024
        k = _sample(row_cdf, cnt, gen) * n + _sample(col_cdf, cnt, gen)
        k = torch.unique(torch.cat([extra, k]))
        k = k[~torch.isin(k, base)]
        if k.numel() >= need:
            pick = torch.randperm(k.numel(), generator=gen, device=dev)[:need]
            extra = k[pick]
            break
        extra = k
        over *= 1.5
        if over > 50:   # nearly dense: fill from the complement deterministically
            allk = torch.arange(m * n, device=dev)
            rest = allk[~torch.isin(allk, torch.cat([base, extra]))]
            extra = torch.cat([extra, rest[: need - extra.numel()]])
            break
    keys = torch.sort(torch.cat([base, extra[:need]] if need > 0 else [base]))[0]
    assert keys.numel() == nnz, (keys.numel(), nnz)
    rows = torch.div(keys, n, rounding_mode="floor")
    cols = keys - rows * n

    U = torch.randn(m, rank, generator=gen, device=dev)
    V = torch.randn(n, rank, generator=gen, device=dev)
    vals = _ratings_for(rows, cols, U, V, gen)

    csr_indptr = torch.zeros(m + 1, dtype=torch.int64, device=dev)
    csr_indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=m), 0)
    # CSC: stable order by (col, row)
    perm = torch.sort(cols * m + rows)[1]
    csc_indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    csc_indptr[1:] = torch.cumsum(torch.bincount(cols, minlength=n), 0)

    t_rows = _sample(row_cdf, nnz_test, gen)
    t_cols = _sample(col_cdf, nnz_test, gen)
    t_vals = _ratings_for(t_rows, t_cols, U, V, gen)

    def h(t, dt):
        return t.to(dt).cpu().numpy()

    return Ratings(
        m=m, n=n,
        csr_indptr=h(csr_indptr, torch.int32), csr_indices=h(cols, torch.int32), csr_data=h(vals, torch.float32),
        csc_indptr=h(csc_indptr, torch.int32), csc_indices=h(rows[perm], torch.int32),
        csc_data=h(vals[perm], torch.float32), coo_row=h(rows, torch.int32),
        test_row=h(t_rows, torch.int32), test_col=h(t_cols, torch.int32), test_val=h(t_vals, torch.float32),
    )


def init_factors(m: int, n: int, f: int, seed: int = 0, scale: float = 0.2):
    """theta0 uniform in [0, scale), X0 = 0 -- the shape of main.cpp:72-78 (which uses glibc
    rand(); the CLI reproduces that exactly, library callers bring their own theta0)."""
    rng = np.random.default_rng(seed)
    thetaT = (scale * rng.random((n, f), dtype=np.float32)).astype(np.float32)
    XT = np.zeros((m, f), dtype=np.float32)
    return thetaT, XT


# ---- the CLI's ten headerless little-endian files (main.cpp:91-103) ---------
_FILES = {
    "R_train_csr.indptr.bin": ("csr_indptr", "<i4"), "R_train_csr.indices.bin": ("csr_indices", "<i4"),
    "R_train_csr.data.bin": ("csr_data", "<f4"), "R_train_csc.indptr.bin": ("csc_indptr", "<i4"),
    "R_train_csc.indices.bin": ("csc_indices", "<i4"), "R_train_csc.data.bin": ("csc_data", "<f4"),
    "R_train_coo.row.bin": ("coo_row", "<i4"), "R_test_coo.row.bin": ("test_row", "<i4"),
    "R_test_coo.col.bin": ("test_col", "<i4"), "R_test_coo.data.bin": ("test_val", "<f4"),
}


def write_bin_dir(path, r: Ratings) -> None:
    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    for name, (attr, dt) in _FILES.items():
        getattr(r, attr).astype(dt).tofile(path / name)


def read_bin_dir(path, m: int, n: int) -> Ratings:
    path = Path(path)
    a = {attr: np.fromfile(path / name, dtype=dt) for name, (attr, dt) in _FILES.items()}
    if a["csr_indptr"].size != m + 1 or a["csc_indptr"].size != n + 1:
        raise ValueError("indptr files do not match m / n")
    return Ratings(m=m, n=n, **a)


def nnz_balanced_ranges(indptr: np.ndarray, parts: int):
    """Contiguous row ranges with (nearly) equal numbers of ratings: the deterministic
    replacement of hugewiki's dynamic batch queue (hugewiki.cu:2446-2496).  Integer path:
    range g is [lo_g, hi_g) with hi_g = first row whose prefix count reaches g+1 shares."""
    indptr = np.asarray(indptr, dtype=np.int64)
    rows = indptr.size - 1
    total = int(indptr[-1])
    bounds = [0]
    for g in range(1, parts):
        target = (total * g) // parts
        b = int(np.searchsorted(indptr, target, side="left"))
        b = min(max(b, bounds[-1]), rows)
        bounds.append(b)
    bounds.append(rows)
    return [(bounds[g], bounds[g + 1]) for g in range(parts)]
