"""Parity tests proper: the CUDA path through the C ABI (ctypes -> libcumf_als_b200.so)
against the oracle (CPU restatement), the golden vectors recorded from the reference, and
-- when oracle/_ref travelled to the box -- the live reference library.  Run with -m gpu."""
import os

import numpy as np
import pytest

import cumf_als_b200 as c
from conftest import golden, rel_fro
from cumf_als_b200.data import Ratings, init_factors, synth_ratings
from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-4   # north_star: factors and per-iteration RMSE within 1e-4 relative (fp32)


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run_gram(torch, rowptr, colidx, val, factor, f, lam, batch_offset=0, batch_size=None, path=c.PATH_SIMT):
    m = rowptr.size - 1
    bs = m - batch_offset if batch_size is None else batch_size
    tt = torch.full((bs, f, f), float("nan"), dtype=torch.float32, device="cuda")
    rhs = torch.full((bs, f), float("nan"), dtype=torch.float32, device="cuda")
    c.gram(batch_offset, bs, tt, dev(torch, rowptr), dev(torch, colidx), lam, m, f, dev(torch, factor), rhs=rhs,
           val=dev(torch, val), path=path)
    torch.cuda.synchronize()
    return tt.cpu().numpy(), rhs.cpu().numpy()


def random_csr(rng, lengths, n):
    rowptr = np.zeros(len(lengths) + 1, np.int32)
    rowptr[1:] = np.cumsum(lengths)
    cols = [np.sort(rng.choice(n, size=k, replace=False)) for k in lengths]
    colidx = np.concatenate(cols + [np.zeros(0, np.int64)]).astype(np.int32)
    val = rng.integers(1, 6, colidx.size).astype(np.float32)
    return rowptr, colidx, val


# ---- Gram + RHS -----------------------------------------------------------------------------
@pytest.mark.parametrize("f", [10, 20, 50, 100, 130, 200])
def test_gram_bit_exact_vs_oracle(cuda, f):
    rng = np.random.default_rng(f)
    lengths = [1, 2, 31, 32, 33, 64, 65, 0, 100, 7, 250, 3]     # ragged, empty, stage boundaries (KC = 32)
    n = 400
    rowptr, colidx, val = random_csr(rng, lengths, n)
    factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
    tt, rhs = run_gram(cuda, rowptr, colidx, val, factor, f, 0.05)
    ref = O.gram(rowptr, colidx, factor, f, 0.05)
    assert np.array_equal(tt, ref), f"max abs diff {np.abs(tt - ref).max()}"      # same FMA chain: bit-exact
    assert np.array_equal(rhs, O.rhs(rowptr, colidx, val, factor, f))


def test_gram_batch_window_and_rows_beyond_m(cuda):
    rng = np.random.default_rng(5)
    rowptr, colidx, val = random_csr(rng, [4, 9, 40, 2, 17, 33, 5], 90)
    factor = rng.standard_normal((90, 20)).astype(np.float32)
    full, _ = run_gram(cuda, rowptr, colidx, val, factor, 20, 0.1)
    sub, _ = run_gram(cuda, rowptr, colidx, val, factor, 20, 0.1, batch_offset=2, batch_size=3)
    assert np.array_equal(sub, full[2:5])
    # a window hanging over the end only touches rows < m (als.cu:449-450)
    over, _ = run_gram(cuda, rowptr, colidx, val, factor, 20, 0.1, batch_offset=5, batch_size=4)
    assert np.array_equal(over[:2], full[5:7]) and np.isnan(over[2:]).all()


def test_gram_split_rows_deterministic(cuda, monkeypatch):
    """Rows longer than the split threshold are cut up and reduced in a fixed order."""
    monkeypatch.setenv("CUMF_SPLIT_NNZ", "64")
    rng = np.random.default_rng(6)
    rowptr, colidx, val = random_csr(rng, [300, 64, 65, 1000, 5], 1200)
    factor = (0.3 * rng.standard_normal((1200, 100))).astype(np.float32)
    a, ra = run_gram(cuda, rowptr, colidx, val, factor, 100, 0.05)
    b, rb = run_gram(cuda, rowptr, colidx, val, factor, 100, 0.05)
    assert np.array_equal(a, b) and np.array_equal(ra, rb)          # run-to-run bit-exact
    ref = O.gram(rowptr, colidx, factor, 100, 0.05)
    assert rel_fro(a, ref) < 1e-6 and np.allclose(a, ref, rtol=1e-5, atol=1e-5)
    assert np.array_equal(a[[1, 4]], ref[[1, 4]])                    # unsplit rows stay bit-exact
    assert rel_fro(ra, O.rhs(rowptr, colidx, val, factor, 100)) < 1e-6


@pytest.mark.parametrize("name", ["gram_f100.npz", "gram_f20.npz", "gram_f200.npz"])
def test_gram_vs_reference_golden(cuda, name):
    g = golden(name)
    f, lam = int(g["f"]), float(g["lam"])
    tt, rhs = run_gram(cuda, g["rowptr"], g["colidx"], g["val"], g["factor"], f, lam)
    rows = [1, 2] if f == 200 else range(tt.shape[0])     # f=200 row 0: the reference kernel races (test_oracle.py)
    for u in rows:
        assert np.array_equal(tt[u], g["tt"][u]), f"row {u} max abs diff {np.abs(tt[u] - g['tt'][u]).max()}"
    assert np.allclose(rhs, g["rhs"], rtol=2e-5, atol=2e-5)


# ---- solvers --------------------------------------------------------------------------------
def spd_batch(rng, batch, f, k=None):
    k = k or 3 * f
    T = (0.3 * rng.standard_normal((batch, k, f))).astype(np.float32)
    A = np.einsum("bki,bkj->bij", T, T).astype(np.float32) + 0.05 * k * np.eye(f, dtype=np.float32)
    A = (A + A.transpose(0, 2, 1)) / 2
    return A.astype(np.float32)


@pytest.mark.parametrize("f", [10, 30, 100, 130, 200])
def test_cg_vs_oracle(cuda, f):
    rng = np.random.default_rng(100 + f)
    batch = 37
    A = spd_batch(rng, batch, f)
    b = rng.standard_normal((batch, f)).astype(np.float32)
    x0 = (0.1 * rng.standard_normal((batch, f))).astype(np.float32)
    for it in (6.0, 1.0, 0.0, 50.0):
        dA, dx, db = dev(cuda, A), dev(cuda, x0), dev(cuda, b)
        c.cg(dA, dx, db, batch, f, it)
        got = dx.cpu().numpy()
        want = O.cg(A, x0, b, f, it)
        assert rel_fro(got, want) < TOL, (f, it, rel_fro(got, want))
    assert np.array_equal(dA.cpu().numpy(), A)   # A is read-only for the CG


def test_cg_empty_row_gives_nan_like_reference(cuda):
    # A = 0, b = 0 -> alpha = 0/0 (cg.cu:128): NaN stays confined to that system
    f = 20
    A = np.zeros((2, f, f), np.float32)
    A[1] = np.eye(f)
    b = np.zeros((2, f), np.float32)
    b[1] = 1
    dx = dev(cuda, np.zeros((2, f), np.float32))
    c.cg(dev(cuda, A), dx, dev(cuda, b), 2, f, 6.0)
    x = dx.cpu().numpy()
    assert np.isnan(x[0]).all() and np.allclose(x[1], 1.0)


@pytest.mark.parametrize("name", ["solve_f100.npz", "solve_f20.npz"])
def test_solvers_vs_reference_golden(cuda, name):
    g = golden(name)
    f, batch = int(g["f"]), g["b"].shape[0]
    for it in (6, 2):
        dx = dev(cuda, g["x0"])
        c.cg(dev(cuda, g["A"]), dx, dev(cuda, g["b"]), batch, f, float(it))
        assert rel_fro(dx.cpu().numpy(), g[f"x_cg{it}"]) < TOL
    dx = dev(cuda, np.zeros_like(g["b"]))
    c.lu(dev(cuda, g["A"]), dx, dev(cuda, g["b"]), batch, f)
    assert rel_fro(dx.cpu().numpy(), g["x_lu"]) < TOL


# ---- RMSE -----------------------------------------------------------------------------------
def test_rmse_vs_oracle_and_golden(cuda):
    g = golden("rmse.npz")
    f, cnt = int(g["f"]), g["val"].size
    args = [dev(cuda, g[k]) for k in ("val", "row", "col", "thetaT", "XT")]
    tr, sse = c.rmse(*args, cnt, f, drop_tail=False)
    te, _ = c.rmse(*args, cnt, f, drop_tail=True)
    assert tr == pytest.approx(float(g["rmse_train"]), rel=1e-5)
    assert te == pytest.approx(float(g["rmse_test"]), rel=1e-5)
    assert tr == pytest.approx(O.rmse(g["val"], g["row"], g["col"], g["thetaT"], g["XT"], f, False), rel=1e-5)
    # the sample set is an integer path: 256*((count-1)/256) samples, bit-exact
    e = g["val"] - np.einsum("ij,ij->i", g["thetaT"][g["col"]].astype(np.float64), g["XT"][g["row"]].astype(np.float64))
    assert sse == pytest.approx(float((e ** 2).sum()), rel=1e-5)
    _, sse_t = c.rmse(*args, cnt, f, drop_tail=True)
    assert sse_t == pytest.approx(float((e[:256 * ((cnt - 1) // 256)] ** 2).sum()), rel=1e-5)


# ---- the whole path: doALS ------------------------------------------------------------------
def run_doals(r, theta0, f, lam, iters, solver="cg", path="auto"):
    os.environ["CUMF_SOLVER"], os.environ["CUMF_PATH"], os.environ["CUMF_QUIET"] = solver, path, "1"
    th, X = theta0.copy(), np.zeros((r.m, f), np.float32)
    fin = c.do_als(*r.doals_args(), th, X, r.test_row, r.test_col, r.test_val, r.m, r.n, f, r.nnz, r.nnz_test, lam,
                   iters, 1, 1, 0)
    return fin, th, X


def ratings_from(g):
    return Ratings(m=int(g["m"]), n=int(g["n"]),
                   **{k: g[k] for k in ("csr_indptr", "csr_indices", "csr_data", "csc_indptr", "csc_indices",
                                        "csc_data", "coo_row", "test_row", "test_col", "test_val")})


@pytest.mark.parametrize("name", ["doals_f20.npz", "doals_f100.npz"])
@pytest.mark.parametrize("solver", ["cg", "lu"])
def test_doals_vs_reference_golden(cuda, name, solver):
    g = golden(name)
    r = ratings_from(g)
    f, lam, iters = int(g["f"]), float(g["lam"]), int(g["iters"])
    fin, th, X = run_doals(r, g["theta0"], f, lam, iters, solver)
    ex, et = rel_fro(X, g[f"x_{solver}"]), rel_fro(th, g[f"theta_{solver}"])
    print(f"{name} {solver}: final {fin} vs ref {float(g['final_' + solver])}; rel X {ex:.2e} theta {et:.2e}")
    if solver == "lu":
        assert fin == pytest.approx(float(g[f"final_{solver}"]), rel=TOL) and ex < TOL and et < TOL
    elif f == 100:   # reference CG at f=100 reads non-existent lanes: see tests/test_oracle.py
        assert fin == pytest.approx(float(g[f"final_{solver}"]), rel=5 * TOL) and ex < 2e-2 and et < 2e-2
    else:            # unconverged CG: rounding-noise floor ~1e-3 on the factors
        assert fin == pytest.approx(float(g[f"final_{solver}"]), rel=TOL) and ex < 1e-3 and et < 1e-3


@pytest.mark.parametrize("f,solver", [(10, "lu"), (10, "cg"), (100, "cg"), (200, "cg")])
def test_doals_vs_oracle(cuda, f, solver):
    r = synth_ratings(300, 420, 16000, 1500, seed=40 + f)
    theta0, _ = init_factors(r.m, r.n, f, seed=f)
    fin, th, X = run_doals(r, theta0, f, 0.05, 3, solver)
    th_o, X_o = theta0.copy(), np.zeros((r.m, f), np.float32)
    fin_o, hist = O.do_als(r, th_o, X_o, f, 0.05, 3, 0 if solver == "cg" else 1)
    assert fin == pytest.approx(fin_o, rel=TOL)
    tol = TOL if solver == "lu" else (5e-3 if f <= 100 else 2e-2)   # CG: noise floor of the 6-step solve
    assert rel_fro(X, X_o) < tol and rel_fro(th, th_o) < tol, (rel_fro(X, X_o), rel_fro(th, th_o))


def test_doals_batches_are_advisory(cuda):
    """Results do not depend on X_BATCH / THETA_BATCH (rows are independent)."""
    r = synth_ratings(200, 300, 9000, 800, seed=77)
    theta0, _ = init_factors(r.m, r.n, 20, seed=1)
    os.environ["CUMF_QUIET"] = "1"
    outs = []
    for xb, tb in ((1, 1), (3, 7)):
        th, X = theta0.copy(), np.zeros((r.m, 20), np.float32)
        c.do_als(*r.doals_args(), th, X, r.test_row, r.test_col, r.test_val, r.m, r.n, 20, r.nnz, r.nnz_test, 0.05, 2,
                 xb, tb, 0)
        outs.append((th, X))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_row_shard_equals_full_update(cuda):
    """E1 sharding: updating a row range touches exactly those rows and gives the
    same bits as the full update."""
    r = synth_ratings(240, 360, 12000, 900, seed=12)
    f = 100
    theta0, X0 = init_factors(r.m, r.n, f, seed=3)
    args = (r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, r.coo_row, r.test_row,
            r.test_col, r.test_val, r.m, r.n, f, 0.05)
    full = c.AlsSolver(*args, path=c.PATH_SIMT)
    full.set_factors(theta0, X0)
    full.update_x()
    _, X_full = full.get_factors()
    lo, hi = 60, 150
    part = c.AlsSolver(*args, x_range=(lo, hi), theta_range=(0, r.n), path=c.PATH_SIMT)
    part.set_factors(theta0, X0)
    part.update_x()
    _, X_part = part.get_factors()
    assert np.array_equal(X_part[lo:hi], X_full[lo:hi])
    assert np.array_equal(X_part[:lo], X0[:lo]) and np.array_equal(X_part[hi:], X0[hi:])
    # sse over shards adds up to the whole (integer partition of the sample set)
    tr_full, te_full = full.sse()
    part2 = c.AlsSolver(*args, x_range=(0, lo), theta_range=(0, r.n), path=c.PATH_SIMT)
    part3 = c.AlsSolver(*args, x_range=(hi, r.m), theta_range=(0, r.n), path=c.PATH_SIMT)
    th_f, X_f = full.get_factors()
    sums = np.zeros(2)
    for p in (part, part2, part3):
        p.set_factors(th_f, X_f)
        sums += np.array(p.sse())
    assert sums[0] == pytest.approx(tr_full, rel=1e-9) and sums[1] == pytest.approx(te_full, rel=1e-9)


# ---- live reference (oracle/_ref shipped to the box) ---------------------------------------
@pytest.mark.parametrize("f,theta_batch", [(100, 1), (20, 2)])
def test_doals_vs_live_reference(cuda, f, theta_batch):
    """Whole path against the reference library itself, same inputs, same box.  The reference's CG is
    not reproducible run to run at f=100 (its block sums add warp partials with atomicAdd in arrival
    order, device_utilities.h:36-48), so the bar is: our distance to the reference is of the order of
    the reference's distance to itself."""
    if not O.ref_available("cg"):
        pytest.skip("oracle/_ref not present")
    # 480 ratings per X row, 277 per theta row: with the 120 / 69 of this test's first version the f = 100 systems were so poorly
    # determined that the reference's final RMSE moved by 1.9e-4 between boxes (0.6857539 ... 0.6858321) while three runs on one box
    # sometimes agreed to 1.5e-5 -- and our (deterministic) 0.6858819 then missed a bar derived from that spread
    r = synth_ratings(1500, 2600, 720000, 20000, seed=90 + f)
    theta0, _ = init_factors(r.m, r.n, f, seed=5)
    iters = 3
    refs = []
    for _ in range(3):
        th_r, X_r = theta0.copy(), np.zeros((r.m, f), np.float32)
        refs.append((O.ref_do_als(r, th_r, X_r, f, 0.048, iters, 1, theta_batch, "cg"), th_r, X_r))
    spread_rmse = max(abs(a[0] - refs[0][0]) / refs[0][0] for a in refs[1:])
    spread_th = max(rel_fro(a[1], refs[0][1]) for a in refs[1:])
    print(f"f={f}: reference vs itself: rmse spread {spread_rmse:.2e}, theta {spread_th:.2e}")
    for path in ("simt", "auto"):
        fin, th, X = run_doals(r, theta0, f, 0.048, iters, "cg", path)
        d_rmse = min(abs(fin - a[0]) / a[0] for a in refs)
        d_th = min(rel_fro(th, a[1]) for a in refs)
        print(f"f={f} path={path}: final rmse {fin:.7f} (ref {refs[0][0]:.7f}) rel {d_rmse:.2e}; theta rel {d_th:.2e}")
        assert d_rmse < max(3 * spread_rmse, TOL)
        assert d_th < max(3 * spread_th, 5e-3)


# ---- fused tcgen05 path (f = 100) -----------------------------------------------------------
TC_LENGTHS = [16, 1, 2, 15, 17, 31, 32, 33, 0, 100, 250, 1000, 3000, 48, 5]


def test_tc_gram_vs_oracle(cuda, monkeypatch):
    """A materialised through the fused TMA + tcgen05 kernel (split-fp16 operands, fp32 TMEM
    accumulation) against the exact-fp32 restatement: error far below the 1e-4 parity bar."""
    monkeypatch.setenv("CUMF_TC_IMPL", "1")          # the round-1 f = 100 kernel (gram_tc.cu); the generic one: test_gpu_generic_f.py
    rng = np.random.default_rng(1)
    n, f, lam = 5000, 100, 0.05
    rowptr, colidx, val = random_csr(rng, TC_LENGTHS, n)
    factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
    tt, rhs = run_gram(cuda, rowptr, colidx, val, factor, f, lam, path=c.PATH_TC)
    ref = O.gram(rowptr, colidx, factor, f, lam)
    for u in range(len(TC_LENGTHS)):
        scale = max(np.abs(ref[u]).max(), 1e-30)
        assert np.abs(tt[u] - ref[u]).max() / scale < 3e-6, (u, TC_LENGTHS[u], np.abs(tt[u] - ref[u]).max() / scale)
    assert np.allclose(rhs, O.rhs(rowptr, colidx, val, factor, f), rtol=1e-5, atol=1e-4)     # fp32 FMAs, two partial sums


def test_tc_gram_small_and_large_values(cuda, monkeypatch):
    """The hi/lo split keeps ~22 mantissa bits across magnitudes (values from 1e-3 to 30)."""
    monkeypatch.setenv("CUMF_TC_IMPL", "1")
    rng = np.random.default_rng(2)
    n, f = 800, 100
    rowptr, colidx, val = random_csr(rng, [200, 40, 333], n)
    for scale in (1e-3, 1.0, 30.0):
        factor = (scale * rng.standard_normal((n, f))).astype(np.float32)
        tt, _ = run_gram(cuda, rowptr, colidx, val, factor, f, 0.048, path=c.PATH_TC)
        ref = O.gram(rowptr, colidx, factor, f, 0.048)
        assert rel_fro(tt, ref) < 2e-6, scale


@pytest.mark.parametrize("split", [None, "64"])
def test_tc_half_step_vs_simt(cuda, monkeypatch, split):
    """One half-step (Gram + RHS + CG) fused vs unfused, including rows split across CTAs."""
    monkeypatch.setenv("CUMF_TC_IMPL", "1")
    if split:
        monkeypatch.setenv("CUMF_SPLIT_NNZ", split)
    rng = np.random.default_rng(3)
    n, f, lam = 5000, 100, 0.05
    lengths = [l for l in TC_LENGTHS if l > 0] * 12        # > 148 rows: every CTA gets work
    rowptr, colidx, val = random_csr(rng, lengths, n)
    factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
    x0 = (0.1 * rng.standard_normal((len(lengths), f))).astype(np.float32)
    outs = {}
    for name, path in (("simt", c.PATH_SIMT), ("tc", c.PATH_TC)):
        plan = c.Plan(rowptr, 0, len(lengths), f, path)
        x = dev(cuda, x0)
        c.update_factor(plan, dev(cuda, colidx), dev(cuda, val), dev(cuda, factor), x, lam)
        cuda.cuda.synchronize()
        outs[name] = x.cpu().numpy()
        if name == "tc":
            # fused kernel (+ split-row reduce + CG); direct staging adds the index scan (first call) and the pre-split pass
            base = 1 if not split else 3
            assert plan.last_launches in (base, base + 2)
        plan.close()
    rows = np.linalg.norm(outs["tc"].astype(np.float64) - outs["simt"], axis=1) / np.linalg.norm(outs["simt"].astype(np.float64), axis=1)
    assert np.median(rows) < 1e-5 and rel_fro(outs["tc"], outs["simt"]) < TOL, (np.median(rows), rows.max())
    # oracle agrees too
    want = x0.copy()
    O.half_step(rowptr, colidx, val, factor, want, f, lam)
    assert rel_fro(outs["tc"], want) < TOL


def test_tc_partial_row_range(cuda):
    """A plan over a row sub-range updates exactly those rows (sharding seam for E1)."""
    rng = np.random.default_rng(4)
    n, f = 2000, 100
    lengths = list(rng.integers(1, 400, 300))
    rowptr, colidx, val = random_csr(rng, lengths, n)
    factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
    x0 = np.zeros((300, f), np.float32)
    full = dev(cuda, x0)
    p = c.Plan(rowptr, 0, 300, f, c.PATH_TC)
    c.update_factor(p, dev(cuda, colidx), dev(cuda, val), dev(cuda, factor), full, 0.05)
    part = dev(cuda, x0)
    p2 = c.Plan(rowptr, 100, 220, f, c.PATH_TC)
    c.update_factor(p2, dev(cuda, colidx), dev(cuda, val), dev(cuda, factor), part, 0.05)
    cuda.cuda.synchronize()
    full, part = full.cpu().numpy(), part.cpu().numpy()
    assert np.array_equal(part[100:220], full[100:220])
    assert not part[:100].any() and not part[220:].any()


def test_doals_fused_vs_oracle_and_reference_lu(cuda):
    """The default (fused) path end to end: per-iteration RMSE within 1e-4 of the oracle; against the
    reference's LU build (golden) the RMSE agrees to the CG-vs-LU gap the reference itself shows."""
    g = golden("doals_f100.npz")
    r = ratings_from(g)
    f, lam, iters = int(g["f"]), float(g["lam"]), int(g["iters"])
    fin, th, X = run_doals(r, g["theta0"], f, lam, iters, "cg", "tc")
    th_o, X_o = g["theta0"].copy(), np.zeros((r.m, f), np.float32)
    fin_o, hist = O.do_als(r, th_o, X_o, f, lam, iters, 0)
    # 60 x 90 problem, 2 iterations: the reference's own CG runs differ by 1.6e-4 here (tools/ref_cg_ub_probe.py)
    assert fin == pytest.approx(fin_o, rel=5 * TOL)
    assert abs(fin - float(g["final_lu"])) < 2e-3


def test_doals_fused_midsize_vs_simt(cuda):
    r = synth_ratings(3000, 5000, 400000, 20000, seed=8)
    theta0, _ = init_factors(r.m, r.n, 100, seed=8)
    fin_tc, th_tc, X_tc = run_doals(r, theta0, 100, 0.048, 3, "cg", "tc")
    fin_si, th_si, X_si = run_doals(r, theta0, 100, 0.048, 3, "cg", "simt")
    assert fin_tc == pytest.approx(fin_si, rel=TOL)
    rows = np.linalg.norm(th_tc.astype(np.float64) - th_si, axis=1) / np.linalg.norm(th_si.astype(np.float64), axis=1)
    print("fused vs simt after 3 iterations: median row rel", np.median(rows), "fro", rel_fro(th_tc, th_si))
    assert np.median(rows) < 2e-3


# ---- partial Gram over per-row rating ranges (multi-GPU form, hugewiki.cu:1675-1678, 2629-2696) -----------------
@pytest.mark.parametrize("f,path", [(20, c.PATH_SIMT), (100, c.PATH_SIMT), (100, c.PATH_TC)])
def test_plan_gram_ranges_vs_oracle(cuda, monkeypatch, f, path):
    """cumf_plan_create_ranges + cumf_plan_gram: [A|b] over a sub-range of every row, lambda * local count."""
    monkeypatch.setenv("CUMF_TC_IMPL", "1")
    rng = np.random.default_rng(11)
    n, lam = 3000, 0.05
    lengths = [0, 1, 16, 17, 40, 333, 5, 64, 1200, 2, 90, 31] * 14
    rowptr, colidx, val = random_csr(rng, lengths, n)
    factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
    rows = len(lengths)
    # columns [t0, t1): the share of one of three "GPUs"; some rows get nothing
    from cumf_als_b200.dist import compact_share, local_share
    t0, t1 = 1000, 2100
    begin, end = local_share(rowptr, colidx, t0, t1)
    assert (end - begin).min() == 0 and (end - begin).max() > 300
    plan = c.Plan.from_ranges(begin, end, f, path)
    tt = cuda.full((rows, f * f), float("nan"), dtype=cuda.float32, device="cuda")
    rhs = cuda.full((rows, f), float("nan"), dtype=cuda.float32, device="cuda")
    plan.gram(dev(cuda, colidx), dev(cuda, val), dev(cuda, factor), lam, tt, rhs)
    cuda.cuda.synchronize()
    tt, rhs = tt.cpu().numpy().reshape(rows, f, f), rhs.cpu().numpy()
    ip, ccol, cval = compact_share(rowptr, colidx, val, t0, t1)
    ref = O.gram(ip.astype(np.int32), ccol, factor, f, lam)
    ref_b = O.rhs(ip.astype(np.int32), ccol, cval, factor, f)
    if path == c.PATH_SIMT:
        assert np.array_equal(tt, ref) and np.array_equal(rhs, ref_b)          # same FMA chain: bit-exact
    else:
        for u in range(rows):
            scale = max(np.abs(ref[u]).max(), 1e-30)
            assert np.abs(tt[u] - ref[u]).max() / scale < 3e-6, (u, lengths[u])
        assert np.allclose(rhs, ref_b, rtol=1e-5, atol=1e-4)
    empty = np.flatnonzero(end == begin)
    assert not tt[empty].any() and not rhs[empty].any()                          # lambda * 0: an all-zero partial
    plan.close()


@pytest.mark.parametrize("f,path", [(20, c.PATH_SIMT), (100, c.PATH_TC)])
def test_partial_gram_engines_two_shares_on_one_gpu(cuda, f, path):
    """The E2 iteration with two theta shards living on one GPU (the all-reduce replaced by an explicit sum):
    X replicated without exchange, theta rank-local, result == the unsharded oracle run to fp32 tolerance."""
    from cumf_als_b200.data import nnz_balanced_ranges
    from cumf_als_b200.dist import GpuPartialGramEngine
    r = synth_ratings(700, 1500, 60000, 3000, seed=21)
    lam, iters = 0.05, 2
    theta0, x0 = init_factors(r.m, r.n, f, seed=6)
    tr = nnz_balanced_ranges(r.csc_indptr, 2)
    engs = [GpuPartialGramEngine(r, f, lam, theta0, x0, t, 0, path=path, cap_bytes=4 * f * (f + 1) * 300) for t in tr]
    assert len(engs[0].batches) == 3 and engs[0].local_nnz + engs[1].local_nnz == r.nnz
    for _ in range(iters):
        for b in range(len(engs[0].batches)):
            parts = [e.partial_gram(b) for e in engs]
            tt, rhs = parts[0][0] + parts[1][0], parts[0][1] + parts[1][1]
            for e in engs:
                e.solve_x(b, tt.clone(), rhs.clone())
        for e in engs:
            e.update_theta()
    cuda.cuda.synchronize()
    assert cuda.equal(engs[0].x, engs[1].x)
    theta = engs[0].theta.clone()
    theta[tr[1][0]:tr[1][1]] = engs[1].theta[tr[1][0]:tr[1][1]]
    sse = np.sum([e.sse() for e in engs], axis=0)
    th_o, X_o = theta0.copy(), x0.copy()
    fin_o, hist = O.do_als(r, th_o, X_o, f, lam, iters, 0)
    eff = 256 * ((r.nnz_test - 1) // 256)
    assert np.sqrt(sse[0] / r.nnz) == pytest.approx(float(hist[-1, 0]), rel=TOL)
    # the oracle divides by nnz_test although it sums `eff` samples (als.cu:1006-1018)
    assert np.sqrt(sse[1] / r.nnz_test) == pytest.approx(float(hist[-1, 1]), rel=TOL)
    assert eff <= r.nnz_test
    # factors: six unconverged CG steps amplify last-bit differences of A (summation order, split-fp16 products) to
    # ~1e-3 per row after two iterations -- the spread the reference shows between its own runs (test_oracle.py)
    assert rel_fro(engs[0].x.cpu().numpy(), X_o) < 30 * TOL
    assert rel_fro(theta.cpu().numpy(), th_o) < 30 * TOL


# ---- train RMSE: by-product of the theta half-step, chunked streaming walk, literal COO pairs --------------------
def _direct_sse(r, theta, X, row=None):
    """sum (val - <theta[col], X[row]>)^2 over the CSR entries, float64 on the host (row defaults to the CSR rows)."""
    row = r.coo_row if row is None else row
    pred = np.einsum("ij,ij->i", theta[r.csr_indices].astype(np.float64), X[row].astype(np.float64))
    return float(((r.csr_data - pred) ** 2).sum())


def test_train_sse_byproduct_vs_streaming_vs_host(cuda, monkeypatch):
    """f = 100 fused path: after a theta half-step the train SSE comes from x^T b + x^T r + reg x^T x collected in the
    CG epilogue (no pass over the ratings); after an X half-step it is the chunked streaming kernel.  Both against a
    float64 evaluation on the host; includes columns split across CTAs (CUMF_SPLIT_NNZ)."""
    monkeypatch.setenv("CUMF_SPLIT_NNZ", "256")
    r = synth_ratings(900, 1400, 150000, 4000, seed=33)
    f, lam = 100, 0.05
    theta0, X0 = init_factors(r.m, r.n, f, seed=5)
    args = (r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, r.coo_row, r.test_row,
            r.test_col, r.test_val, r.m, r.n, f, lam)
    s = c.AlsSolver(*args, path=c.PATH_TC)
    s.set_factors(theta0, X0)
    assert s.collect_train_sse(True)
    for it in range(3):
        s.update_x()
        tr_stream, _ = s.sse()                      # X changed since the last theta half-step: streaming walk
        th, X = s.get_factors()
        assert tr_stream == pytest.approx(_direct_sse(r, th, X), rel=2e-6)
        s.update_theta()
        tr_by, te = s.sse()                         # by-product
        th, X = s.get_factors()
        want = _direct_sse(r, th, X)
        # measured on the Netflix-shaped workload: -1e-5 .. -3e-5 relative (fp16-split Gram, fp32 row terms)
        assert tr_by == pytest.approx(want, rel=1e-4), (it, tr_by, want)
        assert np.sqrt(tr_by / r.nnz) == pytest.approx(np.sqrt(want / r.nnz), rel=TOL)
    s.close()
    monkeypatch.setenv("CUMF_SSE_DIRECT", "1")      # same run with the streaming kernel only
    d = c.AlsSolver(*args, path=c.PATH_TC)
    d.set_factors(theta0, X0)
    d.iterate(3)
    tr_d, te_d = d.sse()
    assert tr_d == pytest.approx(want, rel=2e-6) and te_d == pytest.approx(te, rel=1e-9)
    d.close()


def test_train_sse_keeps_literal_pairs_when_coo_disagrees_with_csr(cuda):
    """als.cu:979-980 pairs cooRowIndex[i] with csrColIndex[i]/csrVal[i]; a cooRowIndex that is not the CSR row
    expansion must be honoured literally (no row/column walk, no by-product)."""
    r = synth_ratings(300, 500, 20000, 800, seed=44)
    f, lam = 100, 0.05
    theta0, X0 = init_factors(r.m, r.n, f, seed=6)
    coo = r.coo_row.copy()
    coo[::97] = (coo[::97] + 7) % r.m
    s = c.AlsSolver(r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, coo, r.test_row,
                    r.test_col, r.test_val, r.m, r.n, f, lam, path=c.PATH_TC)
    s.set_factors(theta0, X0)
    s.iterate(1)
    tr, _ = s.sse()
    th, X = s.get_factors()
    assert tr == pytest.approx(_direct_sse(r, th, X, row=coo), rel=2e-6)
    assert abs(tr - _direct_sse(r, th, X)) > 1e-3 * tr          # and it is not the matrix walk
    s.close()


# ---- direct staging (pre-split fp16 table, swizzled MN-major gather) vs the fp32 staging ring ---------------
def test_tc_direct_staging_bit_identical_to_fp32_staging(cuda, monkeypatch):
    """Both stagings hand the tensor core the same operands in the same order (16-rating k-groups, accumulation
    chains cut every 256 ratings), so the materialised [A|b] must agree bit for bit -- including ragged stages,
    empty rows and multi-tile rows."""
    monkeypatch.setenv("CUMF_TC_IMPL", "1")          # both stagings belong to the round-1 kernel
    rng = np.random.default_rng(11)
    n, f, lam = 6000, 100, 0.05
    lengths = TC_LENGTHS + [64, 65, 255, 256, 257, 511, 513]
    rowptr, colidx, val = random_csr(rng, lengths, n)
    factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("CUMF_TC_DIRECT", mode)
        out[mode] = run_gram(cuda, rowptr, colidx, val, factor, f, lam, path=c.PATH_TC)
    assert np.array_equal(out["0"][0], out["1"][0]) and np.array_equal(out["0"][1], out["1"][1])


@pytest.mark.parametrize("direct", ["0", "1"])
def test_tc_half_step_both_stagings_vs_simt(cuda, monkeypatch, direct):
    """Fused half-step (long-row and short-row kernel variants of either staging) against the exact-fp32 unfused path."""
    monkeypatch.setenv("CUMF_TC_DIRECT", direct)
    monkeypatch.setenv("CUMF_TC_IMPL", "1")
    rng = np.random.default_rng(12)
    n, f, lam = 9000, 100, 0.048
    for lengths in ([int(x) for x in rng.integers(1, 120, 600)],            # short rows: two-MMA variant (three solver warpgroups)
                    [int(x) for x in rng.integers(1500, 4000, 40)] + [0]):  # long rows: symmetric single-MMA variant
        rowptr, colidx, val = random_csr(rng, lengths, n)
        factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
        x0 = (0.1 * rng.standard_normal((len(lengths), f))).astype(np.float32)
        outs = {}
        for name, path in (("simt", c.PATH_SIMT), ("tc", c.PATH_TC)):
            plan = c.Plan(rowptr, 0, len(lengths), f, path)
            x = dev(cuda, x0)
            c.update_factor(plan, dev(cuda, colidx), dev(cuda, val), dev(cuda, factor), x, lam)
            cuda.cuda.synchronize()
            outs[name] = x.cpu().numpy()
            plan.close()
        ok = np.isfinite(outs["simt"]).all(axis=1)
        assert ok.sum() >= len(lengths) - 1
        assert rel_fro(outs["tc"][ok], outs["simt"][ok]) < TOL


def test_plan_factor_rows_hint_saves_the_index_scan(cuda, monkeypatch):
    """cumf_plan_set_factor_rows: same result and no stream synchronisation on the first call -- the scan of the column ids
    still runs (it validates the hint) but asynchronously, its verdict is read by the plan's next launch."""
    monkeypatch.setenv("CUMF_TC_DIRECT", "1")
    rng = np.random.default_rng(13)
    n, f, lam = 3000, 100, 0.05
    lengths = [int(x) for x in rng.integers(1, 300, 200)]
    rowptr, colidx, val = random_csr(rng, lengths, n)
    factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
    x0 = (0.1 * rng.standard_normal((len(lengths), f))).astype(np.float32)
    res, launches = [], []
    for hint in (False, True):
        plan = c.Plan(rowptr, 0, len(lengths), f, c.PATH_TC)
        if hint:
            plan.set_factor_rows(n)
        x = dev(cuda, x0)
        c.update_factor(plan, dev(cuda, colidx), dev(cuda, val), dev(cuda, factor), x, lam)
        cuda.cuda.synchronize()
        res.append(x.cpu().numpy())
        launches.append(plan.last_launches)
        plan.close()
    assert np.array_equal(res[0], res[1])
    assert launches[1] <= launches[0]


def test_plan_factor_rows_hint_too_small_is_reported(cuda, monkeypatch):
    """A hint smaller than the largest column id + 1 would gather out-of-bounds rows as zeros: the asynchronous validation
    turns that into CUMF_EINVAL on the plan's next launch instead of a silently wrong Gram."""
    rng = np.random.default_rng(14)
    n, f = 3000, 100
    lengths = [int(x) for x in rng.integers(1, 300, 200)]
    rowptr, colidx, val = random_csr(rng, lengths, n)
    factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
    plan = c.Plan(rowptr, 0, len(lengths), f, c.PATH_TC)
    plan.set_factor_rows(int(colidx.max()))            # one too few
    x = dev(cuda, np.zeros((len(lengths), f), np.float32))
    args = (plan, dev(cuda, colidx), dev(cuda, val), dev(cuda, factor), x, 0.05)
    c.update_factor(*args)
    cuda.cuda.synchronize()
    with pytest.raises(c.CumfError, match="exceeds"):
        c.update_factor(*args)
    plan.close()


def test_doals_frees_its_device_memory_by_default(cuda, monkeypatch):
    """Like the reference (als.cu:1026-1033), doALS returns with its device buffers released: the buffer cache is opt-in."""
    monkeypatch.delenv("CUMF_CACHE_MB", raising=False)
    monkeypatch.setenv("CUMF_QUIET", "1")
    f, lam = 100, 0.048
    r = synth_ratings(3000, 5000, 400000, 20000, seed=5)
    theta0, X0 = init_factors(r.m, r.n, f, seed=2)
    c.load_library().cumf_release_cached_memory()
    call = lambda: c.do_als(*r.doals_args(), theta0.copy(), X0.copy(), r.test_row, r.test_col, r.test_val, r.m, r.n, f, r.nnz,
                            r.nnz_test, lam, 1, 1, 1, 0)
    call()                                  # context-level allocations (module load, cuBLAS-free path) happen once
    cuda.cuda.synchronize()
    free0 = cuda.cuda.mem_get_info()[0]
    call()
    cuda.cuda.synchronize()
    assert cuda.cuda.mem_get_info()[0] >= free0 - (1 << 20), "doALS kept device memory after returning"


def test_doals_twice_reuses_cached_buffers_and_matches(cuda, monkeypatch):
    """CUMF_CACHE_MB=<n> (opt-in): cumf_als_destroy keeps the device buffers for the next solver; a second doALS on the same
    inputs must give the same factors and RMSE, and cumf_release_cached_memory must hand the memory back."""
    monkeypatch.setenv("CUMF_CACHE_MB", "4096")
    f, lam, iters = 100, 0.048, 2
    r = synth_ratings(300, 500, 20000, 2000, seed=5)
    theta0, X0 = init_factors(r.m, r.n, f, seed=2)
    os.environ["CUMF_QUIET"] = "1"
    runs = []
    for _ in range(2):
        th, X = theta0.copy(), X0.copy()
        fin = c.do_als(*r.doals_args(), th, X, r.test_row, r.test_col, r.test_val, r.m, r.n, f, r.nnz, r.nnz_test, lam, iters, 1, 1, 0)
        runs.append((fin, th, X))
    assert runs[0][0] == runs[1][0]
    assert np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][2], runs[1][2])
    free0 = cuda.cuda.mem_get_info()[0]
    assert c.load_library().cumf_release_cached_memory() == 0
    assert cuda.cuda.mem_get_info()[0] >= free0
