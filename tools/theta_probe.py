"""Which sub-pipeline bounds a fused half-step?  (GPU box)  Times cumf_update_factor of both sides on the Netflix-shaped
workload with the library named by CUMF_ALS_LIB -- the shipped one, or experiment builds (tools/build_variant.sh exp_<X>
-DCUMF_TC2_EXP_<X>) that leave out the MMAs, the gathers or the solver.  Factors after two real iterations are computed once
with the shipped library and cached under /tmp so that every variant starts from the same, finite inputs.

    python tools/theta_probe.py prepare            # shipped library: writes /tmp/theta_probe_{theta,X}.npy
    CUMF_ALS_LIB=... python tools/theta_probe.py   # prints the best-of-5 time of each side"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import cumf_als_b200 as c  # noqa: E402

w = bench.WORKLOADS[os.environ.get("PROBE_WORKLOAD", "netflix")]
r, theta0, X0 = bench.make_inputs(w, 1.0, "cuda")
f, lam = w["f"], w["lam"]
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
if len(sys.argv) > 1 and sys.argv[1] == "prepare":
    theta, X = dev(theta0), dev(X0)
else:
    theta, X = dev(np.load("/tmp/theta_probe_theta.npy")), dev(np.load("/tmp/theta_probe_X.npy"))
sides = {
    "X": (c.Plan(r.csr_indptr, 0, r.m, f, c.PATH_TC), dev(r.csr_indices), dev(r.csr_data), theta, X),
    "theta": (c.Plan(r.csc_indptr, 0, r.n, f, c.PATH_TC), dev(r.csc_indices), dev(r.csc_data), X, theta),
}
if len(sys.argv) > 1 and sys.argv[1] == "prepare":
    for _ in range(2):
        for name in ("X", "theta"):
            plan, idx, val, fac, out = sides[name]
            c.update_factor(plan, idx, val, fac, out, lam, c.SOLVER_CG, 6.0)
    torch.cuda.synchronize()
    np.save("/tmp/theta_probe_theta.npy", theta.cpu().numpy())
    np.save("/tmp/theta_probe_X.npy", X.cpu().numpy())
    sys.exit(0)
res = {}
for name, (plan, idx, val, fac, out) in sides.items():
    keep = out.clone()
    fkeep = fac.clone()
    ts = []
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        c.update_factor(plan, idx, val, fac, out, lam, c.SOLVER_CG, float(os.environ.get("PROBE_CG", "6")))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        out.copy_(keep)
        fac.copy_(fkeep)
    res[name] = min(ts[1:])
print(f"{os.environ.get('PROBE_TAG', Path(os.environ.get('CUMF_ALS_LIB', 'shipped')).stem):40s} X {res['X']:7.3f} ms   theta {res['theta']:7.3f} ms", flush=True)
