"""Workload for ncu captures of the hot kernels (GPU box):
    ncu --set full --import-source on -k regex:als_fused -s 2 -c 2 -o gpurun_out/prof python tools/profile_fused.py
One warm-up iteration, then one profiled ALS iteration (X side launch, theta side launch) on
the Netflix-shaped workload (optionally scaled: argv[1] = scale, argv[2] = path auto|simt|tc, argv[3] = bench workload)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import cumf_als_b200 as c  # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    path = {"auto": c.PATH_AUTO, "simt": c.PATH_SIMT, "tc": c.PATH_TC}[sys.argv[2] if len(sys.argv) > 2 else "auto"]
    w = bench.WORKLOADS[sys.argv[3] if len(sys.argv) > 3 else "netflix"]
    r, theta0, X0 = bench.make_inputs(w, scale, "cuda")
    s = c.AlsSolver(r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, r.coo_row,
                    r.test_row, r.test_col, r.test_val, r.m, r.n, w["f"], w["lam"], path=path)
    s.set_factors(theta0, X0)
    ms1 = s.iterate(1)
    torch.cuda.synchronize()
    ms2 = s.iterate(1)
    print(f"iteration ms: warm-up {ms1:.2f}, profiled {ms2:.2f}; rmse {s.rmse()}")
    s.close()


if __name__ == "__main__":
    main()
