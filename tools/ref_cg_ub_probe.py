"""Does the reference CG (blockDim = f = 100) depend on what ran before it?  (GPU box)

updateXWithCGKernel reduces with __shfl_down over a last warp that has only 4 live lanes
(device_utilities.h:9-13, cg.cu:683).  This probe runs the reference CG on the SAME systems
(a) right after a kernel that leaves zeros in the register file and (b) right after the
reference's own get_hermitian100 (which leaves live data), and prints how far the outputs move."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402


def main():
    g = dict(np.load(ROOT / "tests" / "golden" / "solve_f100.npz"))
    f = 100
    reps = 400                                  # enough systems to cover every SM several times
    A = np.tile(g["A"], (reps, 1, 1)); b = np.tile(g["b"], (reps, 1)); x0 = np.tile(g["x0"], (reps, 1))
    batch = A.shape[0]
    lib = O.ref("cg")
    dA, db = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    want = O.cg(g["A"], g["x0"], g["b"], f, 6.0)

    def run(prep):
        dx = torch.from_numpy(x0).cuda()
        prep()
        torch.cuda.synchronize()
        lib.ref_cg(dA.data_ptr(), dx.data_ptr(), db.data_ptr(), batch, f, 6.0)
        torch.cuda.synchronize()
        x = dx.cpu().numpy().reshape(reps, -1, f)
        err = np.linalg.norm(x - want[None], axis=2) / np.linalg.norm(want[None], axis=2)
        return err

    def zeros():
        torch.zeros(64 << 20, device="cuda").mul_(0.0)

    # dirty the register file with the reference's own Gram kernel, as doALS does (als.cu:804 then :831)
    rng = np.random.default_rng(0)
    n, m = 4000, 2000
    lengths = rng.integers(50, 400, m)
    rowptr = np.zeros(m + 1, np.int32); rowptr[1:] = np.cumsum(lengths)
    cols = rng.integers(0, n, rowptr[-1]).astype(np.int32)
    fac = torch.from_numpy(rng.standard_normal((n, f)).astype(np.float32) * 50).cuda()
    d_rowptr, d_cols = torch.from_numpy(rowptr).cuda(), torch.from_numpy(cols).cuda()
    tt = torch.empty((m, f, f), device="cuda")

    def gram():
        lib.ref_get_hermitian(0, m, tt.data_ptr(), d_rowptr.data_ptr(), d_cols.data_ptr(), 0.05, m, f, fac.data_ptr())

    for name, prep in (("after zero-fill kernel", zeros), ("after get_hermitian100", gram), ("after zero-fill kernel", zeros)):
        e = run(prep)
        print(f"{name:28s}: rel err vs oracle  median {np.median(e):.2e}  max {e.max():.2e}  "
              f"systems off by >1e-4: {(e > 1e-4).sum()} / {e.size}")


def doals_repeat():
    """Is the reference's own doALS (CG, f=100) reproducible run to run?  Its block sums add the
    four warp partials with shared-memory atomicAdd in arrival order (device_utilities.h:36-48)."""
    from cumf_als_b200.data import Ratings
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    from make_golden import CaptureStdout
    g = dict(np.load(ROOT / "tests" / "golden" / "doals_f100.npz"))
    m, n, f, lam, iters = int(g["m"]), int(g["n"]), int(g["f"]), float(g["lam"]), int(g["iters"])
    r = Ratings(m=m, n=n, **{k: g[k] for k in ("csr_indptr", "csr_indices", "csr_data", "csc_indptr", "csc_indices",
                                                "csc_data", "coo_row", "test_row", "test_col", "test_val")})
    outs = []
    for rep in range(4):
        th, X = g["theta0"].copy(), np.zeros((m, f), np.float32)
        with CaptureStdout():
            fin = O.ref_do_als(r, th, X, f, lam, iters, 1, 1, "cg")
        outs.append((fin, th, X))
    rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))
    print("reference doALS (cg, f=100) repeated on identical inputs:")
    for k in range(1, 4):
        print(f"  run {k} vs run 0: final rmse {outs[k][0]:.7f} vs {outs[0][0]:.7f}; rel X {rel(outs[k][2], outs[0][2]):.2e} "
              f"theta {rel(outs[k][1], outs[0][1]):.2e}")
    print(f"  run 0 vs committed golden: rel X {rel(outs[0][2], g['x_cg']):.2e} theta {rel(outs[0][1], g['theta_cg']):.2e}")


if __name__ == "__main__":
    main()
    doals_repeat()
