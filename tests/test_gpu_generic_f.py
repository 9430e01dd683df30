"""The generic-f fused tcgen05 kernel (cumf_als_b200/csrc/gram_tc2.cuh) through the C ABI, against the oracle and the
exact-fp32 SIMT path: every accumulator geometry (one chunk per region f <= 62, two f <= 126, f > 127 with two 128-lane row
blocks and two warpgroups per system), both launch variants (short rows / long rows "sym"), rows split across CTAs, the
materialising (partial-Gram) mode, value magnitudes.  f = 100 is run through this kernel with CUMF_TC_IMPL=2.  -m gpu."""
import numpy as np
import pytest

import cumf_als_b200 as c
from conftest import rel_fro
from cumf_als_b200.data import init_factors, synth_ratings
from oracle import oracle as O
from test_gpu_parity import dev, random_csr, run_gram

pytestmark = pytest.mark.gpu

TOL = 1e-4
LENGTHS = [16, 1, 2, 15, 17, 31, 32, 33, 0, 100, 250, 1000, 3000, 48, 5, 64, 65, 255, 256, 257, 513]
FS = [10, 20, 40, 60, 70, 100, 120, 130, 160, 200]


@pytest.fixture(autouse=True)
def _generic_kernel(monkeypatch):
    monkeypatch.setenv("CUMF_TC_IMPL", "2")


@pytest.mark.parametrize("f", FS)
def test_tc2_gram_vs_oracle(cuda, f):
    """[A|b] materialised through the generic kernel (every chunk stores its partial, then the deterministic reduce) against
    the exact-fp32 restatement.  Three truncating tensor-core accumulations per 16 ratings, chains cut every 256 ratings."""
    rng = np.random.default_rng(f)
    n, lam = 5000, 0.05
    rowptr, colidx, val = random_csr(rng, LENGTHS, n)
    factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
    tt, rhs = run_gram(cuda, rowptr, colidx, val, factor, f, lam, path=c.PATH_TC)
    ref = O.gram(rowptr, colidx, factor, f, lam)
    ref_b = O.rhs(rowptr, colidx, val, factor, f)
    worst = 0.0
    for u in range(len(LENGTHS)):
        scale = max(np.abs(ref[u]).max(), 1e-30)
        worst = max(worst, np.abs(tt[u] - ref[u]).max() / scale)
        # (i, j) and (j, i) add the same three products in a different order (hi_i lo_j before lo_i hi_j): symmetric to an ulp
        assert np.abs(tt[u] - tt[u].T).max() <= 1e-6 * scale
    bscale = max(np.abs(ref_b).max(), 1e-30)
    print(f"f={f}: max element error of A / row max {worst:.2e}, of b {np.abs(rhs - ref_b).max() / bscale:.2e}")
    assert worst < 6e-6
    assert np.abs(rhs - ref_b).max() / bscale < 6e-6


@pytest.mark.parametrize("f", [10, 100, 200])
def test_tc2_gram_small_and_large_values(cuda, f):
    """One power-of-two scale per half-step keeps ~22 mantissa bits whatever the magnitude of the factor; inside one factor,
    entries down to 2^-17 of the largest keep them too (smaller ones degrade gradually, absolute error ~ max * 2^-39)."""
    rng = np.random.default_rng(2)
    n = 800
    rowptr, colidx, val = random_csr(rng, [200, 40, 333], n)
    for scale in (1e-6, 1e-3, 1.0, 30.0, 3e4):
        factor = (scale * rng.standard_normal((n, f))).astype(np.float32)
        tt, _ = run_gram(cuda, rowptr, colidx, val, factor, f, 0.048, path=c.PATH_TC)
        ref = O.gram(rowptr, colidx, factor, f, 0.048)
        assert rel_fro(tt, ref) < 4e-6, (scale, rel_fro(tt, ref))
    # mixed magnitudes inside one factor: columns scaled by 2^0 .. 2^-16
    factor = (rng.standard_normal((n, f)) * np.exp2(-rng.integers(0, 17, f))[None, :]).astype(np.float32)
    tt, _ = run_gram(cuda, rowptr, colidx, val, factor, f, 0.0, path=c.PATH_TC)
    ref = O.gram(rowptr, colidx, factor, f, 0.0).astype(np.float64)
    d = np.sqrt(np.abs(np.einsum("uii->ui", ref)))
    rel = np.abs(tt - ref) / np.maximum(d[:, :, None] * d[:, None, :], 1e-300)          # error relative to |a_i| |a_j|
    assert rel.max() < 1e-5, rel.max()


@pytest.mark.parametrize("f", FS)
@pytest.mark.parametrize("split", [None, "64"])
def test_tc2_half_step_vs_simt(cuda, monkeypatch, f, split):
    """One half-step (Gram + RHS + CG) fused vs unfused, short-row variant; with CUMF_SPLIT_NNZ=64 most rows are split."""
    if split:
        monkeypatch.setenv("CUMF_SPLIT_NNZ", split)
    monkeypatch.setenv("CUMF_TC_SYM", "0")
    _half_step_case(cuda, f, seed=3)


@pytest.mark.parametrize("f", [10, 50, 100])
def test_tc2_half_step_sym_variant_vs_simt(cuda, monkeypatch, f):
    """The long-row variant (two MMAs per k-group, G + G^T through shared memory) forced on the same ragged rows."""
    monkeypatch.setenv("CUMF_TC_SYM", "1")
    _half_step_case(cuda, f, seed=4)


def _half_step_case(cuda, f, seed):
    rng = np.random.default_rng(seed)
    n, lam = 5000, 0.05
    lengths = [l for l in LENGTHS if l > 0] * 9        # > 148 rows: every CTA gets work
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
    rows = np.linalg.norm(outs["tc"].astype(np.float64) - outs["simt"], axis=1) / np.linalg.norm(outs["simt"].astype(np.float64), axis=1)
    print(f"f={f}: fused vs exact-fp32 path per row: median {np.median(rows):.2e} max {rows.max():.2e}")
    assert np.median(rows) < 2e-5 and rel_fro(outs["tc"], outs["simt"]) < TOL, (np.median(rows), rows.max())
    want = x0.copy()
    O.half_step(rowptr, colidx, val, factor, want, f, lam)
    assert rel_fro(outs["tc"], want) < TOL


@pytest.mark.parametrize("f", [10, 100, 200])
def test_tc2_deterministic_and_partial_row_range(cuda, f):
    rng = np.random.default_rng(5)
    n = 2000
    lengths = list(rng.integers(1, 400, 300))
    rowptr, colidx, val = random_csr(rng, lengths, n)
    factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
    x0 = np.zeros((300, f), np.float32)
    runs = []
    for _ in range(2):
        full = dev(cuda, x0)
        p = c.Plan(rowptr, 0, 300, f, c.PATH_TC)
        c.update_factor(p, dev(cuda, colidx), dev(cuda, val), dev(cuda, factor), full, 0.05)
        cuda.cuda.synchronize()
        runs.append(full.cpu().numpy())
        p.close()
    assert np.array_equal(runs[0], runs[1])
    part = dev(cuda, x0)
    p2 = c.Plan(rowptr, 100, 220, f, c.PATH_TC)
    c.update_factor(p2, dev(cuda, colidx), dev(cuda, val), dev(cuda, factor), part, 0.05)
    cuda.cuda.synchronize()
    part = part.cpu().numpy()
    assert np.array_equal(part[100:220], runs[0][100:220])
    assert not part[:100].any() and not part[220:].any()


@pytest.mark.parametrize("f", [10, 60, 100, 130, 200])
def test_tc2_doals_vs_oracle_and_simt(cuda, monkeypatch, f):
    """Whole path (resident solver, fused kernel both sides incl. the long-row variant where it applies, by-product train
    RMSE) against the exact-fp32 path and the oracle: per-iteration RMSE within 1e-4."""
    r = synth_ratings(700, 4000, 160000, 8000, seed=20 + f)        # X side: 230 ratings per row, theta side: 40
    theta0, X0 = init_factors(r.m, r.n, f, seed=3)
    res = {}
    for path in (c.PATH_SIMT, c.PATH_TC):
        s = c.AlsSolver(r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, r.coo_row,
                        r.test_row, r.test_col, r.test_val, r.m, r.n, f, 0.048, path=path)
        s.set_factors(theta0, X0)
        s.collect_train_sse(True)
        hist = []
        for _ in range(3):
            s.iterate(1)
            hist.append(s.rmse())
        res[path] = (np.array(hist), *s.get_factors())
        s.close()
    th_o, X_o = theta0.copy(), X0.copy()
    _, hist_o = O.do_als(r, th_o, X_o, f, 0.048, 3, 0)
    h_tc, h_si = res[c.PATH_TC][0], res[c.PATH_SIMT][0]
    print(f"f={f}: rmse tc {h_tc[-1]} simt {h_si[-1]} oracle {hist_o[-1]}; theta rel tc-vs-simt {rel_fro(res[c.PATH_TC][1], res[c.PATH_SIMT][1]):.2e}")
    # f >= 130 on this small problem: 230 ratings for 200 unknowns per X row, six unconverged CG steps on nearly singular
    # systems amplify the 3e-6 Gram difference; the contract-size run (test_gpu_contract_sizes.py, C3) is the bar there
    tol = TOL if f <= 100 else 5 * TOL
    assert np.abs(h_tc - h_si).max() / h_si.min() < tol
    assert np.abs(h_tc - hist_o).max() / hist_o.min() < tol


def test_tc2_plan_gram_ranges_vs_oracle(cuda):
    """cumf_plan_create_ranges + cumf_plan_gram through the generic kernel (E2's partial Gram), f = 100 and f = 200."""
    from cumf_als_b200.dist import compact_share, local_share
    for f in (100, 200):
        rng = np.random.default_rng(11)
        n, lam = 3000, 0.05
        lengths = [0, 1, 16, 17, 40, 333, 5, 64, 1200, 2, 90, 31] * 14
        rowptr, colidx, val = random_csr(rng, lengths, n)
        factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
        rows = len(lengths)
        begin, end = local_share(rowptr, colidx, 1000, 2100)
        plan = c.Plan.from_ranges(begin, end, f, c.PATH_TC)
        tt = cuda.full((rows, f * f), float("nan"), dtype=cuda.float32, device="cuda")
        rhs = cuda.full((rows, f), float("nan"), dtype=cuda.float32, device="cuda")
        plan.gram(dev(cuda, colidx), dev(cuda, val), dev(cuda, factor), lam, tt, rhs)
        cuda.cuda.synchronize()
        tt, rhs = tt.cpu().numpy().reshape(rows, f, f), rhs.cpu().numpy()
        ip, ccol, cval = compact_share(rowptr, colidx, val, 1000, 2100)
        ref = O.gram(ip.astype(np.int32), ccol, factor, f, lam)
        ref_b = O.rhs(ip.astype(np.int32), ccol, cval, factor, f)
        for u in range(rows):
            scale = max(np.abs(ref[u]).max(), 1e-30)
            assert np.abs(tt[u] - ref[u]).max() / scale < 6e-6, (f, u, lengths[u])
        assert np.allclose(rhs, ref_b, rtol=2e-5, atol=1e-4)
        plan.close()


def test_tc2_reduced_precision_mode(cuda, monkeypatch):
    """CUMF_TT_FP16=1, the run-time form of the reference's CUMF_TT_FP16 / CUMF_XX_FP16 storage flags (als.cu:30-31, 335-441;
    SURVEY.md 8f f3): only the 11-bit hi halves of the factor are gathered and multiplied.  Own tolerance: the Gram is
    accurate to 2^-11 per operand (measured ~3e-4 relative here), the per-iteration RMSE to 1e-3; off by default."""
    rng = np.random.default_rng(8)
    n, f, lam = 5000, 100, 0.05
    rowptr, colidx, val = random_csr(rng, LENGTHS, n)
    factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
    ref = O.gram(rowptr, colidx, factor, f, lam)
    exact, _ = run_gram(cuda, rowptr, colidx, val, factor, f, lam, path=c.PATH_TC)
    monkeypatch.setenv("CUMF_TT_FP16", "1")
    lossy, rhs = run_gram(cuda, rowptr, colidx, val, factor, f, lam, path=c.PATH_TC)
    e_exact, e_lossy = rel_fro(exact, ref), rel_fro(lossy, ref)
    print(f"Gram rel error: split-fp16 {e_exact:.2e}, hi only {e_lossy:.2e}")
    assert e_exact < 4e-6 and 2e-5 < e_lossy < 2e-3             # the switch is live, and as lossy as 11 bits say
    assert rel_fro(rhs, O.rhs(rowptr, colidx, val, factor, f)) < 2e-3
    r = synth_ratings(700, 4000, 160000, 8000, seed=31)
    theta0, X0 = init_factors(r.m, r.n, f, seed=3)
    hist = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("CUMF_TT_FP16", mode)
        s = c.AlsSolver(r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, r.coo_row,
                        r.test_row, r.test_col, r.test_val, r.m, r.n, f, 0.048, path=c.PATH_TC)
        s.set_factors(theta0, X0)
        h = []
        for _ in range(3):
            s.iterate(1)
            h.append(s.rmse())
        hist[mode] = np.array(h)
        s.close()
    d = np.abs(hist["1"] - hist["0"]).max() / hist["0"].min()
    print(f"per-iteration RMSE, hi only vs split-fp16: max rel diff {d:.2e}")
    assert d < 3e-3
