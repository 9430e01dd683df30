#!/usr/bin/env python
"""bench.py -- ALS iterations/sec on the BASELINE.json workload (Netflix-shaped, f=100).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload netflix]

One "step" = one ALS iteration = update-X half-step + update-theta half-step (Gram + RHS
formation and the batched CG solve for every row and every column), the span the reference's
`update X run` + `update theta run` timers cover (als.cu:730-850, 858-963).

Prints ONE JSON line (rank 0):
  value      iterations/sec with CSR/CSC/factors already resident in HBM, timed with CUDA events
             (max over ranks), W warm-up iterations first; inputs (theta 192 MB, ratings 1.2 GB
             per orientation) exceed the 126 MB L2, so no flush is needed between iterations.
  e2e        the same metric through the reference-facing call doALS(host pointers): uploads,
             K iterations, per-iteration train/test RMSE, factor download -- wall clock / K.
  roofline   Gram(+fused solve) kernel: algorithmic bytes / CUDA-event kernel time.
  cpu_baseline  the CPU oracle (oracle/als_cpu.c port) timed on a bounded row sample, scaled.
`--impl reference` runs the UNMODIFIED reference (oracle/_ref, its Kepler-era kernels recompiled
for sm_100a + cuBLAS/cuSPARSE) through its own doALS on the same inputs.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "ALS iterations/sec ({workload}-shaped synthetic, f={f}, CG solver; BASELINE.json metric (i))"

WORKLOADS = {
    # name: m, n, nnz, nnz_test, f, lambda, seed, reference (X_BATCH, THETA_BATCH) (test_als.sh:5-28)
    "netflix": dict(m=17770, n=480189, nnz=99072112, nnz_test=1408395, f=100, lam=0.048, seed=1002, ref_batches=(1, 3)),
    "netflix_f200": dict(m=17770, n=480189, nnz=99072112, nnz_test=1408395, f=200, lam=0.048, seed=1002, ref_batches=(1, 10)),
    "ml10m": dict(m=71567, n=65133, nnz=9000048, nnz_test=1000006, f=10, lam=0.05, seed=1001, ref_batches=(1, 1)),
    # BASELINE configs[3] (README.md:77-79: ./main 1000990 624961 100 252800275 4003960 1.4 6 3)
    "yahoo": dict(m=1000990, n=624961, nnz=252800275, nnz_test=4003960, f=100, lam=1.4, seed=1004, ref_batches=(6, 3)),
    "tiny": dict(m=2000, n=5000, nnz=400000, nnz_test=20000, f=100, lam=0.048, seed=1, ref_batches=(1, 1)),
}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def make_inputs(w, scale: float, device: str):
    from cumf_als_b200.data import init_factors, synth_ratings
    m, n = w["m"], w["n"]
    nnz, nnz_test = int(w["nnz"] * scale), max(1024, int(w["nnz_test"] * scale))
    nnz = max(nnz, m + n)
    r = synth_ratings(m, n, nnz, nnz_test, seed=w["seed"], device=device, alpha_row=0.5, alpha_col=0.6)
    theta0, X0 = init_factors(m, n, w["f"], seed=w["seed"])
    return r, theta0, X0


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.tmp = gpu_index, None, None

    def __enter__(self):
        try:
            self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.tmp is None:
            return out
        try:
            self.tmp.flush()
            rows = [l.split(",") for l in open(self.tmp.name).read().strip().splitlines() if l.strip()]
            os.unlink(self.tmp.name)
        except Exception:
            return out
        sm = [float(r[1]) for r in rows if len(r) >= 9]
        if not sm:
            return out
        out["samples"] = len(sm)
        out["sm_mhz"] = float(np.median(sm))
        out["sm_max_mhz"] = float(rows[0][2])
        out["power_w_max"] = max(float(r[3]) for r in rows if len(r) >= 9)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, nm in enumerate(names):
            if any("Active" in r[5 + i] and "Not" not in r[5 + i] for r in rows if len(r) >= 9):
                out["reasons"].append(nm)
        return out


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload: str, path: str):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    p = ROOT / "profiles" / "traffic.json"
    if not p.exists():
        return None
    return json.loads(p.read_text()).get(f"{workload}:{path}")


def gram_bytes(rows, nnz, f, fused: bool):
    """Algorithmic bytes of one half-step's Gram formation."""
    if fused:
        return nnz * (4 * f + 4 + 4) + (rows + 1) * 4 + 2 * rows * f * 4
    return nnz * (4 * f + 4) + (rows + 1) * 4 + rows * f * f * 4


def cpu_baseline(r, theta0, X0, w, budget_s: float = 15.0):
    """CPU oracle on a bounded sample of rows of both half-steps, scaled by rating share."""
    from oracle import oracle as O
    f, lam = w["f"], w["lam"]
    cores = os.cpu_count() or 1
    est = {}
    sample_desc = []
    for side, (ptr, idx, val, fac, out, rows) in {
        "x": (r.csr_indptr, r.csr_indices, r.csr_data, theta0, X0.copy(), r.m),
        "theta": (r.csc_indptr, r.csc_indices, r.csc_data, np.ascontiguousarray(
            np.random.default_rng(0).random((r.m, f), dtype=np.float32) * 0.2), theta0.copy(), r.n),
    }.items():
        # grow the sample until it costs ~budget/2 seconds
        share_rows = max(8, rows // 2000)
        t, k = 0.0, 0
        while True:
            lo = (rows // 3)
            hi = min(rows, lo + share_rows)
            t0 = time.perf_counter()
            O.half_step(ptr, idx, val, fac, out, f, lam, 0, 6.0, lo, hi)
            t = time.perf_counter() - t0
            k = int(ptr[hi] - ptr[lo])
            if t > budget_s / 4 or hi - lo >= rows - lo:
                break
            share_rows = int(share_rows * max(2.0, min(16.0, (budget_s / 2) / max(t, 1e-3))))
        est[side] = t * (int(ptr[-1]) / max(k, 1))
        sample_desc.append(f"{side}: rows [{lo},{hi}) = {k} ratings in {t:.2f} s")
    per_iter = est["x"] + est["theta"]
    return {"value": 1.0 / per_iter, "unit": "iterations/s", "cores": cores, "kind": "port",
            "sample": "oracle/als_cpu.c (OpenMP) on " + "; ".join(sample_desc) + "; scaled by rating share"}


def e2e_sharded(args, r, theta0, X0, f, lam, path, rank, local_rank, world, barrier):
    """e2e at N > 1: cumf_als_b200.dist.ShardedAls over one AlsSolver per rank, from pinned HOST buffers -- every rank
    uploads its CSR/CSC shard + the factors, runs K iterations with the row-block exchange and the per-iteration
    all-reduced RMSE doALS prints (als.cu:991, 1018), and downloads the factors; max wall clock over ranks."""
    import torch
    import torch.distributed as dist
    import cumf_als_b200 as c
    from cumf_als_b200.data import nnz_balanced_ranges
    from cumf_als_b200.dist import GpuEngine, ShardedAls

    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    for name in ("csr_indptr", "csr_indices", "csr_data", "csc_indptr", "csc_indices", "csc_data", "coo_row",
                 "test_row", "test_col", "test_val"):
        setattr(r, name, pin(getattr(r, name)))
    th0, x0 = pin(theta0), pin(X0)
    x_ranges = nnz_balanced_ranges(r.csr_indptr, world)
    t_ranges = nnz_balanced_ranges(r.csc_indptr, world)

    def call(iters):
        s2 = c.AlsSolver(r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, r.coo_row,
                         r.test_row, r.test_col, r.test_val, r.m, r.n, f, lam, x_range=x_ranges[rank],
                         theta_range=t_ranges[rank], device=local_rank, path=path)
        s2.set_factors(th0, x0)
        sh = ShardedAls(GpuEngine(s2, local_rank), x_ranges, t_ranges, r.nnz, r.nnz_test)
        fin = None
        for _ in range(iters):
            sh.step()
            fin = sh.rmse()
        out = s2.get_factors()
        s2.close()
        return fin, out

    if args.warmup > 0:
        call(1)
    barrier()
    t0 = time.perf_counter()
    fin, _ = call(args.steps)
    barrier()
    wall = time.perf_counter() - t0
    # bytes this rank moved: its rating shards + both full factors in, both full factors out; summed over ranks
    xr0, xr1 = x_ranges[rank]
    tr0, tr1 = t_ranges[rank]
    xn = int(r.csr_indptr[xr1] - r.csr_indptr[xr0])      # CSR shard: column ids + values + COO rows (12 B per rating)
    tn = int(r.csc_indptr[tr1] - r.csc_indptr[tr0])      # CSC shard: row ids + values (8 B per rating)
    eff = ((r.nnz_test - 1) // 256) * 256                # test samples the reference's launch covers (als.cu:1006)
    test_share = eff * xr1 // r.m - eff * xr0 // r.m
    stats = [wall, float(xn * 12 + tn * 8 + test_share * 12 + th0.nbytes + x0.nbytes), float(th0.nbytes + x0.nbytes)]
    if world > 1:
        wt = torch.tensor([stats[0]], device="cuda", dtype=torch.float64)
        dist.all_reduce(wt, op=dist.ReduceOp.MAX)
        bt = torch.tensor(stats[1:], device="cuda", dtype=torch.float64)
        dist.all_reduce(bt)
        stats = [float(wt.item()), float(bt[0].item()), float(bt[1].item())]
    return {"value": args.steps / stats[0], "unit": "iterations/s", "h2d_bytes_per_step": int(stats[1]) // args.steps,
            "d2h_bytes_per_step": int(stats[2]) // args.steps, "wall_s": stats[0],
            "final_test_rmse": fin[1] if fin else None,
            "what": "per rank: AlsSolver(host shard) + set_factors + K x (ShardedAls.step + all-reduced RMSE) + get_factors, "
                    "pinned host buffers; max wall clock over ranks; one untimed 1-iteration call first"}


def run_ours(args, w):
    import torch
    import torch.distributed as dist
    import cumf_als_b200 as c
    from cumf_als_b200.data import nnz_balanced_ranges

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    r, theta0, X0 = make_inputs(w, args.scale, "cuda")
    f, lam = w["f"], w["lam"]
    path = {"auto": c.PATH_AUTO, "simt": c.PATH_SIMT, "tc": c.PATH_TC}[args.path]
    os.environ["CUMF_TIME_KERNELS"] = "1"

    if world == 1:
        xr, tr_ = (0, r.m), (0, r.n)
    else:
        xr = nnz_balanced_ranges(r.csr_indptr, world)[rank]
        tr_ = nnz_balanced_ranges(r.csc_indptr, world)[rank]
    solver = c.AlsSolver(r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, r.coo_row,
                         r.test_row, r.test_col, r.test_val, r.m, r.n, f, lam, x_range=xr, theta_range=tr_,
                         device=local_rank, path=path)
    solver.set_factors(theta0, X0)

    if world > 1:
        from cumf_als_b200.dist import GpuEngine, ShardedAls
        x_ranges = nnz_balanced_ranges(r.csr_indptr, world)
        t_ranges = nnz_balanced_ranges(r.csc_indptr, world)
        sharded = ShardedAls(GpuEngine(solver, local_rank), x_ranges, t_ranges, r.nnz, r.nnz_test)
        step = sharded.iterate
    else:
        step = lambda k: solver.iterate(k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step(args.warmup)
    solver.timers(reset=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
        ms = step(args.steps)
        barrier()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    tm = solver.timers()
    train_rmse, test_rmse = (sharded.rmse() if world > 1 else solver.rmse())
    iters_per_s = args.steps / (ms / 1e3)

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        fused = args.path != "simt" and _tc_active(c, f)
        xs, ts = xr[1] - xr[0], tr_[1] - tr_[0]
        nnz_x = int(r.csr_indptr[xr[1]] - r.csr_indptr[xr[0]])
        nnz_t = int(r.csc_indptr[tr_[1]] - r.csc_indptr[tr_[0]])
        gb = (gram_bytes(xs, nnz_x, f, fused) + gram_bytes(ts, nnz_t, f, fused)) / 1e9   # per iteration, this rank
        iters_done = tm["iterations"] if tm["iterations"] > 0 else args.steps      # sharded runs drive the half-steps directly
        tm["iterations"] = iters_done
        if world > 1:
            tm["x_ms"], tm["theta_ms"] = tm["gram_x_ms"], tm["gram_theta_ms"]     # per-rank kernel time (rank 0)
        gram_ms = (tm["gram_x_ms"] + tm["gram_theta_ms"]) / max(iters_done, 1)
        achieved = gb / (gram_ms / 1e3) if gram_ms > 0 else None
        line = {
            "metric": METRIC.format(workload=args.workload, f=f), "value": iters_per_s,
            "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "m": r.m, "n": r.n, "nnz": r.nnz, "nnz_test": r.nnz_test, "f": f,
                       "lambda": lam, "solver": "cg6", "path": "tcgen05-fused" if fused else "simt-unfused",
                       "sharding": "single GPU" if world == 1 else f"rows nnz-balanced over {world} ranks, all-gather",
                       "l2": f"inputs exceed L2 (factors {(r.m + r.n) * f * 4 / 1e6:.0f} MB, ratings {r.nnz * 12 / 1e9:.1f} GB per "
                             f"orientation): no flush"},
            "x_ms": tm["x_ms"] / max(tm["iterations"], 1), "theta_ms": tm["theta_ms"] / max(tm["iterations"], 1),
            "train_rmse": train_rmse, "test_rmse": test_rmse,
            "gpu_launches": int(tm["launches"]),
            "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": ncu_traffic(args.workload, "fused" if fused else "simt") if world == 1 and args.scale == 1.0 else None,
                         "kernel": "gram (+rhs" + (" + fused CG)" if fused else ")") + " X side + theta side",
                         "bytes_per_iteration_gb": gb, "kernel_ms_per_iteration": gram_ms,
                         "gram_x_ms": tm["gram_x_ms"] / max(tm["iterations"], 1),
                         "gram_theta_ms": tm["gram_theta_ms"] / max(tm["iterations"], 1),
                         "formula": "B_gram_fused" if fused else "B_gram (materialised A)", "peak_source": peak_src},
        }
    solver.close()

    # e2e: the reference-facing call with host buffers (rank 0 of a single-GPU run; under
    # torchrun every rank would repeat the same whole-job call, so it is only timed at N=1)
    if rank == 0 and world == 1 and not args.no_e2e:
        os.environ["CUMF_QUIET"] = "1"
        os.environ["CUMF_PATH"] = args.path
        # the reference's CLI keeps every input in pinned host memory (cudaMallocHost, main.cpp:50-69): same here
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        for name in ("csr_indptr", "csr_indices", "csr_data", "csc_indptr", "csc_indices", "csc_data", "coo_row",
                     "test_row", "test_col", "test_val"):
            setattr(r, name, pin(getattr(r, name)))
        # pinned host -> device copy rate of this box (the e2e number moves with it: 2.2 GB of ratings per call)
        probe = torch.from_numpy(r.csr_data).cuda()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        src = torch.from_numpy(r.csr_data)
        probe.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0.record()
        probe.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        h2d_gbs = r.csr_data.nbytes / 1e9 / (e0.elapsed_time(e1) / 1e3)
        del probe
        if args.warmup > 0:   # allocator / driver warm-up outside the timed call, as in the reference arm
            c.do_als(*r.doals_args(), pin(theta0), pin(X0), r.test_row, r.test_col, r.test_val, r.m, r.n, f, r.nnz,
                     r.nnz_test, lam, 1, 1, 1, local_rank)
        th, X = pin(theta0), pin(X0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fin = c.do_als(*r.doals_args(), th, X, r.test_row, r.test_col, r.test_val, r.m, r.n, f, r.nnz, r.nnz_test, lam,
                       args.steps, 1, 1, local_rank)
        wall = time.perf_counter() - t0
        h2d = (r.csr_indices.nbytes + r.csr_data.nbytes + r.csc_indices.nbytes + r.csc_data.nbytes + r.coo_row.nbytes +
               r.test_row.nbytes + r.test_col.nbytes + r.test_val.nbytes + th.nbytes + X.nbytes)
        line["e2e"] = {"value": args.steps / wall, "unit": "iterations/s", "h2d_bytes_per_step": h2d // args.steps,
                       "d2h_bytes_per_step": (th.nbytes + X.nbytes) // args.steps, "wall_s": wall,
                       "pinned_h2d_gbs_this_box": h2d_gbs,
                       "final_test_rmse": fin,
                       "what": "doALS(host pointers, ITERS=steps): upload + iterations + per-iteration RMSE + download; "
                               "one untimed 1-iteration call first (both arms)"}
    # e2e at N > 1: the sharded public API from pinned host buffers (CUMF_BENCH_SHARDED_E2E=1 exercises it on one GPU)
    force_sharded = os.environ.get("CUMF_BENCH_SHARDED_E2E") == "1"
    if (world > 1 or force_sharded) and not args.no_e2e:
        try:
            e2e = e2e_sharded(args, r, theta0, X0, f, lam, path, rank, local_rank, world, barrier)
        except Exception as exc:        # the resident numbers above stay valid; every rank fails the same way
            e2e = {"error": f"{type(exc).__name__}: {exc}"}
        if rank == 0:
            line["e2e_sharded" if world == 1 else "e2e"] = e2e
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(r, theta0, X0, w)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_ours_partial_gram(args, w):
    """N ranks, theta sharded (E2): partial [A|b] of every X row per rank -> NCCL all-reduce -> replicated CG;
    theta-step rank-local.  Same line format; phases timed with CUDA events on the launching stream."""
    import torch
    import torch.distributed as dist
    import cumf_als_b200 as c
    from cumf_als_b200.data import nnz_balanced_ranges
    from cumf_als_b200.dist import GpuPartialGramEngine, PartialGramAls

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    r, theta0, X0 = make_inputs(w, args.scale, "cuda")
    f, lam = w["f"], w["lam"]
    path = {"auto": c.PATH_AUTO, "simt": c.PATH_SIMT, "tc": c.PATH_TC}[args.path]
    t_ranges = nnz_balanced_ranges(r.csc_indptr, world)
    eng = GpuPartialGramEngine(r, f, lam, theta0, X0, t_ranges[rank], local_rank, path=path)
    drv = PartialGramAls(eng, r.nnz, r.nnz_test)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    drv.iterate(args.warmup)
    eng.launches = 0
    drv.allreduce_bytes = 0
    barrier()
    with ClockSampler(local_rank) as clocks:
        ms = drv.iterate(args.steps)
        barrier()
    # one more, instrumented, iteration for the phase split (outside the timed region)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    phases = None
    if len(eng.batches) == 1:
        barrier()
        ev[0].record()
        tt, rhs = eng.partial_gram(0)
        ev[1].record()
        if world > 1:
            dist.all_reduce(tt)
            dist.all_reduce(rhs)
        ev[2].record()
        eng.solve_x(0, tt, rhs)
        ev[3].record()
        eng.update_theta()
        ev[4].record()
        torch.cuda.synchronize()
        phases = {k: ev[i].elapsed_time(ev[i + 1]) for i, k in enumerate(["partial_gram_ms", "allreduce_ms", "cg_x_ms", "theta_ms"])}
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    train_rmse, test_rmse = drv.rmse()
    if rank == 0:
        peak, peak_src = measured_peaks()
        fused = args.path != "simt" and _tc_active(c, f)
        t0, t1 = t_ranges[rank]
        nnz_t = int(r.csc_indptr[t1] - r.csc_indptr[t0])
        gb = (gram_bytes(r.m, eng.local_nnz, f, False) + gram_bytes(t1 - t0, nnz_t, f, fused)) / 1e9
        kernel_ms = (phases["partial_gram_ms"] + phases["theta_ms"]) if phases else None
        achieved = gb / (kernel_ms / 1e3) if kernel_ms else None
        print(json.dumps({
            "metric": METRIC.format(workload=args.workload, f=f), "value": args.steps / (ms / 1e3),
            "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "m": r.m, "n": r.n, "nnz": r.nnz, "nnz_test": r.nnz_test, "f": f,
                       "lambda": lam, "solver": "cg6", "path": "tcgen05" if fused else "simt",
                       "sharding": f"theta rows nnz-balanced over {world} ranks; X-step = partial [A|b] + NCCL all-reduce "
                                   f"+ replicated CG; theta-step rank-local (no factor exchange)",
                       "l2": "inputs exceed L2: no flush"},
            "phases_ms": phases, "allreduce_bytes_per_iteration": drv.allreduce_bytes // max(args.steps, 1),
            "train_rmse": train_rmse, "test_rmse": test_rmse, "gpu_launches": int(eng.launches),
            "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": None,
                         "kernel": "partial gram (materialised) X side + fused theta side, rank 0",
                         "bytes_per_iteration_gb": gb, "kernel_ms_per_iteration": kernel_ms, "peak_source": peak_src},
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _tc_active(c, f):
    try:
        import ctypes
        h = ctypes.c_void_p()
        rp = np.array([0, 1], np.int32)
        rc = c.api.load_library().cumf_plan_create(ctypes.byref(h), rp.ctypes.data_as(ctypes.c_void_p), 1, 0, 1, f, c.PATH_TC)
        if rc == 0:
            c.api.load_library().cumf_plan_destroy(h)
        return rc == 0
    except Exception:
        return False


def run_reference(args, w):
    """The unmodified reference (oracle/_ref) through its own doALS, rank 0 only."""
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    import torch
    from oracle import oracle as O
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    torch.cuda.set_device(local_rank)
    r, theta0, X0 = make_inputs(w, args.scale, "cuda")
    f, lam = w["f"], w["lam"]
    xb, tb = w["ref_batches"]
    variant = args.ref_variant
    if not O.ref_available(variant):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_als_%s.so not built" % variant}))
        return
    from make_golden import CaptureStdout
    # pinned host inputs, like the reference's own CLI (cudaMallocHost, main.cpp:50-69): its per-iteration
    # CSR re-upload (als.cu:734-739) then runs at full PCIe speed
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    for name in ("csr_indptr", "csr_indices", "csr_data", "csc_indptr", "csc_indices", "csc_data", "coo_row",
                 "test_row", "test_col", "test_val"):
        setattr(r, name, pin(getattr(r, name)))
    theta0, X0 = pin(theta0), pin(X0)
    if args.warmup > 0:   # context / cuBLAS / cuSPARSE initialisation outside the timed call
        with CaptureStdout():
            O.ref_do_als(r, theta0.copy(), X0.copy(), f, lam, 1, xb, tb, variant, local_rank)
    th, X = pin(theta0), pin(X0)
    iters = args.steps
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        with CaptureStdout() as cap:
            fin = O.ref_do_als(r, th, X, f, lam, iters, xb, tb, variant, local_rank)
        wall = time.perf_counter() - t0
    xs = [float(v) for v in re.findall(r"^update X run ([0-9.]+) seconds", cap.text, flags=re.M)]
    ts = [float(v) for v in re.findall(r"^update theta run ([0-9.]+) seconds", cap.text, flags=re.M)]
    kx = [float(v) for v in re.findall(r"update X kernel run ([0-9.]+) seconds", cap.text)]
    kt = [float(v) for v in re.findall(r"update Theta kernel run ([0-9.]+) seconds", cap.text)]
    sv = [float(v) for v in re.findall(r"solver run seconds: ([0-9.]+)", cap.text)]
    als_s = sum(xs) + sum(ts)
    value = iters / als_s if als_s > 0 else iters / wall
    peak, peak_src = measured_peaks()
    gb = (gram_bytes(r.m, r.nnz, f, False) + gram_bytes(r.n, r.nnz, f, False)) / 1e9
    gram_s = (sum(kx) + sum(kt)) / iters if kx else None
    line = {
        "impl": "reference", "metric": METRIC.format(workload=args.workload, f=f), "value": value,
        "unit": "iterations/s", "n_gpus": 1, "steps": iters, "warmup": 0, "ms_per_step": 1e3 * als_s / iters,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "m": r.m, "n": r.n, "nnz": r.nnz, "nnz_test": r.nnz_test, "f": f,
                   "lambda": lam, "solver": "cg6" if variant == "cg" else "cublas-lu", "X_BATCH": xb, "THETA_BATCH": tb,
                   "what": "unmodified reference sources (oracle/_ref), Kepler-era kernels recompiled for sm_100a; value = "
                           "iterations / sum of its own `update X run`+`update theta run` timers (als.cu:850, 963)"},
        "e2e": {"value": iters / wall, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "wall_s": wall},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": 1, "kind": "reference",
                         "sample": f"reference doALS ({variant}) single host thread driving GPU kernels, {iters} iterations"},
        "final_test_rmse": fin, "clocks": clocks.summary(), "gpu_launches": 0,
        "roofline": {"bound": "hbm", "achieved": (gb / gram_s) if gram_s else None, "peak": peak, "unit": "GB/s",
                     "frac": (gb / gram_s / peak) if gram_s else None, "traffic": None,
                     "kernel": "get_hermitian100 X+theta (reference timers)", "peak_source": peak_src},
        "ref_timers": {"update_x_s": sum(xs) / max(len(xs), 1), "update_theta_s": sum(ts) / max(len(ts), 1),
                       "gram_kernels_s_per_iter": gram_s, "solver_s_per_iter": sum(sv) / iters if sv else None},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="netflix", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink nnz (debug only; 1.0 = the BASELINE workload)")
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--ref-variant", default="cg", choices=["cg", "lu"])
    ap.add_argument("--sharding", default="rows", choices=["rows", "partial-gram"],
                    help="N>1: 'rows' = each rank updates its rows from full factors and broadcasts them (E1); "
                         "'partial-gram' = theta sharded, partial [A|b] all-reduced (E2, hugewiki.cu:2629-2827)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    # stdout carries exactly one JSON line: everything libraries print meanwhile ("NCCL version ...", the reference's
    # progress lines) goes to stderr; print() below is pointed at the real stdout again just for the result
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global print

    def print(*a, **k):          # noqa: A001 -- the result line(s) of this script
        os.write(real_stdout, (" ".join(str(x) for x in a) + "\n").encode())

    if args.impl == "reference":
        run_reference(args, w)
    elif args.sharding == "partial-gram":
        run_ours_partial_gram(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
