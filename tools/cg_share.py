"""How much of each fused half-step is the in-kernel CG?  (GPU box)  Times cumf_update_factor on the Netflix-shaped
workload with cgIter = 6 (the product), 3 and 0 (drain + initial residual only): the cgIter = 0 time is the floor set
by gather + staging + MMA + drain, the difference is what the two solver warpgroups cost."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import cumf_als_b200 as c  # noqa: E402

w = bench.WORKLOADS["netflix"]
r, theta0, X0 = bench.make_inputs(w, float(sys.argv[1]) if len(sys.argv) > 1 else 1.0, "cuda")
f, lam = w["f"], w["lam"]
dev = lambda a: torch.from_numpy(a).cuda()
theta, X = dev(theta0), dev(X0)
sides = {
    "X": (c.Plan(r.csr_indptr, 0, r.m, f, c.PATH_TC), dev(r.csr_indices), dev(r.csr_data), theta, X),
    "theta": (c.Plan(r.csc_indptr, 0, r.n, f, c.PATH_TC), dev(r.csc_indices), dev(r.csc_data), X, theta),
}
for _ in range(2):      # two real iterations first: X0 = 0 would make the theta systems diagonal (CG exits after one step)
    for name in ("X", "theta"):
        plan, idx, val, fac, out = sides[name]
        c.update_factor(plan, idx, val, fac, out, lam, c.SOLVER_CG, 6.0)
torch.cuda.synchronize()
import os  # noqa: E402
if os.environ.get("SWEEP_ROW_COST"):
    # per-row cost (in ratings) used to cut the chunk list into per-CTA ranges: the slowest CTA sets the kernel time
    for cost in os.environ["SWEEP_ROW_COST"].split(","):
        os.environ["CUMF_TC_ROW_COST"] = cost
        for name, (rp, rows, idx, val, fac, out) in {"X": (r.csr_indptr, r.m, sides["X"][1], sides["X"][2], theta, X),
                                                     "theta": (r.csc_indptr, r.n, sides["theta"][1], sides["theta"][2], X, theta)}.items():
            plan = c.Plan(rp, 0, rows, f, c.PATH_TC)
            keep = out.clone()
            best = 1e9
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                c.update_factor(plan, idx, val, fac, out, lam, c.SOLVER_CG, 6.0)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
                out.copy_(keep)
            print(f"row cost {cost:>5s}  {name:6s} {best:7.3f} ms", flush=True)
            plan.close()
    sys.exit(0)
for name, (plan, idx, val, fac, out) in sides.items():
    for cg in (6.0, 3.0, 0.0, 6.0):
        keep = out.clone()
        c.update_factor(plan, idx, val, fac, out, lam, c.SOLVER_CG, cg)      # warm
        out.copy_(keep)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        c.update_factor(plan, idx, val, fac, out, lam, c.SOLVER_CG, cg)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:6s} cgIter {cg:.0f}: {e0.elapsed_time(e1):7.3f} ms", flush=True)
        out.copy_(keep)
