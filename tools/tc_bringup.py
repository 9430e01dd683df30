"""Bring-up diagnostics for the fused tcgen05 kernel (run on the GPU box).

Materialises A through the fused kernel (cumf_gram, path=TC) on a few ragged rows and prints
the error against the CPU oracle, for both orderings of the UMMA descriptor strides, then a
single half-step (Gram + CG) against the SIMT path."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import cumf_als_b200 as c  # noqa: E402
from oracle import oracle as O  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def csr(rng, lengths, n):
    rowptr = np.zeros(len(lengths) + 1, np.int32)
    rowptr[1:] = np.cumsum(lengths)
    cols = np.concatenate([np.sort(rng.choice(n, size=k, replace=False)) for k in lengths] + [np.zeros(0, np.int64)])
    return rowptr, cols.astype(np.int32), rng.integers(1, 6, cols.size).astype(np.float32)


def main():
    f, lam = 100, 0.05
    rng = np.random.default_rng(1)
    which = sys.argv[1] if len(sys.argv) > 1 else "gram"
    lengths = [16, 1, 2, 15, 17, 31, 32, 33, 0, 100, 250, 1000, 3000]
    n = 5000
    rowptr, colidx, val = csr(rng, lengths, n)
    factor = (0.3 * rng.standard_normal((n, f))).astype(np.float32)
    m = len(lengths)
    ref = O.gram(rowptr, colidx, factor, f, lam)
    rhs_ref = O.rhs(rowptr, colidx, val, factor, f)
    if which == "gram":
        tt = torch.full((m, f, f), float("nan"), device="cuda")
        rhs = torch.full((m, f), float("nan"), device="cuda")
        c.gram(0, m, tt, dev(rowptr), dev(colidx), lam, m, f, dev(factor), rhs=rhs, val=dev(val), path=c.PATH_TC)
        torch.cuda.synchronize()
        tt, rhs = tt.cpu().numpy(), rhs.cpu().numpy()
        print("swap =", os.environ.get("CUMF_TC_SWAP_LBO_SBO", "0"))
        for u in range(m):
            scale = max(np.abs(ref[u]).max(), 1e-30)
            err = np.abs(tt[u] - ref[u]).max() / scale
            sym = np.abs(tt[u] - tt[u].T).max() / scale
            print(f"  row {u:2d} len {lengths[u]:5d}: max rel err {err:.3e}  asym {sym:.2e}  "
                  f"rhs err {np.abs(rhs[u] - rhs_ref[u]).max() / max(np.abs(rhs_ref[u]).max(), 1e-30):.1e}  nan {np.isnan(tt[u]).any()}")
        if np.nanmax(np.abs(tt - ref)) > 1e-2:
            u = 9
            print("  row 9 ref[0,:6]", ref[u][0, :6], "\n  row 9 got[0,:6]", tt[u][0, :6])
            print("  row 9 ref[1,:6]", ref[u][1, :6], "\n  row 9 got[1,:6]", tt[u][1, :6])
    elif which == "direct":
        # direct staging (pre-split fp16 table, swizzled MN-major gather; 32- or 64-rating stages) against the fp32-staging
        # kernel: all feed the tensor core identical operands in the same order, so materialised [A|b] must agree bit for bit
        lengths += [64, 65, 127, 128, 129, 255, 256, 257, 511, 513]
        rowptr, colidx, val = csr(rng, lengths, n)
        m = len(lengths)
        ref = O.gram(rowptr, colidx, factor, f, lam)
        rhs_ref = O.rhs(rowptr, colidx, val, factor, f)
        modes = {"fp32": {"CUMF_TC_DIRECT": "0"}, "direct32": {"CUMF_TC_DIRECT": "1", "CUMF_TC_STAGE_ROWS": "32"},
                 "direct64": {"CUMF_TC_DIRECT": "1", "CUMF_TC_STAGE_ROWS": "64"}}
        outs = {}
        for name, env in modes.items():
            os.environ.update(env)
            tt = torch.full((m, f, f), float("nan"), device="cuda")
            rhs = torch.full((m, f), float("nan"), device="cuda")
            c.gram(0, m, tt, dev(rowptr), dev(colidx), lam, m, f, dev(factor), rhs=rhs, val=dev(val), path=c.PATH_TC)
            torch.cuda.synchronize()
            outs[name] = (tt.cpu().numpy(), rhs.cpu().numpy())
        worst, all_same = 0.0, True
        for u in range(m):
            scale = max(np.abs(ref[u]).max(), 1e-30)
            errs = {k: np.abs(v[0][u] - ref[u]).max() / scale for k, v in outs.items()}
            same = all(np.array_equal(outs["fp32"][0][u], v[0][u]) and np.array_equal(outs["fp32"][1][u], v[1][u]) for v in outs.values())
            all_same &= same
            rb = max(np.abs(v[1][u] - rhs_ref[u]).max() / max(np.abs(rhs_ref[u]).max(), 1e-30) for v in outs.values())
            worst = max([worst, rb if np.isfinite(rb) else 1e9] + [e if np.isfinite(e) else 1e9 for e in errs.values()])
            print(f"  row {u:2d} len {lengths[u]:5d}: err vs oracle " + "  ".join(f"{k} {e:.3e}" for k, e in errs.items()) + f"  rhs {rb:.1e}  bit-identical {same}")
        print("DIRECT GRAM", "OK" if worst < 1e-5 and all_same else "WRONG", f"(worst {worst:.3e}, bit-identical {all_same})")
        if not (worst < 1e-5 and all_same):
            sys.exit(3)
    elif which == "stress":
        # many chunks per CTA (CUMF_TC_CTAS small), multi-tile rows, split rows: fused vs SIMT
        rng = np.random.default_rng(7)
        lens = [int(x) for x in rng.integers(1, 2600, 400)] + [0, 1, 9000, 20000]
        rowptr, colidx, val = csr(rng, lens, 24000)
        factor = (0.3 * rng.standard_normal((24000, f))).astype(np.float32)
        x0 = (0.1 * rng.standard_normal((len(lens), f))).astype(np.float32)
        outs = {}
        for name, path in (("simt", c.PATH_SIMT), ("tc", c.PATH_TC)):
            plan = c.Plan(rowptr, 0, len(lens), f, path)
            x = dev(x0)
            for _ in range(2):
                c.update_factor(plan, dev(colidx), dev(val), dev(factor), x, lam)
            torch.cuda.synchronize()
            outs[name] = x.cpu().numpy()
        a, b = outs["tc"].astype(np.float64), outs["simt"].astype(np.float64)
        ok = np.isfinite(b).all(axis=1)
        rows = np.linalg.norm(a[ok] - b[ok], axis=1) / np.maximum(np.linalg.norm(b[ok], axis=1), 1e-30)
        print(f"stress CTAS={os.environ.get('CUMF_TC_CTAS')}: {len(lens)} rows, median row rel diff {np.median(rows):.2e}, max {rows.max():.2e}")
        assert rows.max() < 1e-3, "fused and SIMT half-steps disagree"
    else:
        # one half-step through plans: fused vs SIMT
        x0 = (0.1 * rng.standard_normal((m, f))).astype(np.float32)
        outs = {}
        for name, path in (("simt", c.PATH_SIMT), ("tc", c.PATH_TC)):
            plan = c.Plan(rowptr, 0, m, f, path)
            x = dev(x0)
            c.update_factor(plan, dev(colidx), dev(val), dev(factor), x, lam)
            torch.cuda.synchronize()
            outs[name] = x.cpu().numpy()
            print(name, "launches", plan.last_launches)
        for u in range(m):
            a, b = outs["tc"][u].astype(np.float64), outs["simt"][u].astype(np.float64)
            print(f"  row {u:2d} len {lengths[u]:5d}: rel diff {np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30):.3e}")


if __name__ == "__main__":
    main()
