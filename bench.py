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
  roofline   Gram(+fused solve) kernel: algorithmic bytes / CUDA-event kernel time, for the iteration and per
             launch (`launches.x_side`, `launches.theta_side`: each with its own bound -- the theta-side launch of the Netflix
             shape gathers from an L2-resident 7 MB table and is bound on chip, not by HBM).
  cpu_baseline  the CPU oracle (oracle/als_cpu.c port) timed on a bounded row sample, scaled: 1 thread (north_star) and all cores.
N > 1 (torchrun, one process per GPU): rows sharded by rating count; every rank's solver epilogues store the updated rows into
all ranks' factor replicas (CUDA IPC peer mappings) and a flag-barrier kernel ends each half-step -- no collective on the data
path; e2e = cumf_doALS(host pointers) on rank 0 with CUMF_GPUS=N (one process driving the N GPUs).
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


DTYPE = "f32 (Gram: split-fp16 tcgen05 MMA, hi + lo, fp32 accumulate in TMEM; CG and RMSE in fp32)"


def common_config(args, r, w):
    """Identical keys and values in both arms (ours / --impl reference): the driver compares them."""
    f = w["f"]
    return {"workload": args.workload, "m": r.m, "n": r.n, "nnz": r.nnz, "nnz_test": r.nnz_test, "f": f, "lambda": w["lam"],
            "solver": "cg6",
            "l2": f"inputs exceed L2 (factors {(r.m + r.n) * f * 4 / 1e6:.0f} MB, ratings {r.nnz * 12 / 1e9:.1f} GB per "
                  f"orientation): no flush"}


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


_CLOCK_HELPER = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
sel, path = sys.argv[1], sys.argv[2]
try:
    h = nv.nvmlDeviceGetHandleByUUID(sel) if sel.startswith("GPU-") else nv.nvmlDeviceGetHandleByIndex(int(sel))
except Exception:
    h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[3]))
mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
gate = path + ".gate"
import os
parent = os.getppid()
with open(path, "w", buffering=1) as f:
    while True:
        if os.getppid() != parent:        # the bench process is gone (killed before its atexit hook): do not linger
            break
        if not os.path.exists(gate):      # NVML is only queried inside a timed region (its queries contend with the driver
            time.sleep(0.001)             # calls of everything else on the box: allocation, peer access, IPC)
            continue
        try:
            period = float(open(gate).read() or 4) / 1000.0
        except Exception:
            period = 0.004
        try:
            f.write("%.6f,%d,%d,%.1f,%d\n" % (time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mx,
                                              nv.nvmlDeviceGetPowerUsage(h) / 1000.0, nv.nvmlDeviceGetCurrentClocksEventReasons(h)))
        except Exception:
            pass
        time.sleep(period)
"""


class ClockSampler:
    """SM clock, power and throttle reasons of one GPU DURING the timed region.  A helper PROCESS is started when the sampler
    is first constructed (process start-up and nvmlInit take longer than an 8-GPU timed region of 45 ms) and polls NVML every
    4 ms while a `with` block holds its gate file -- only then: polled all the time, it slowed the driver calls of the e2e
    leg (allocation, peer access) several-fold.  summary() keeps the samples between the block's wall-clock time stamps.
    Falls back to `nvidia-smi -lms` without pynvml."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    _helpers = {}       # gpu index -> (Popen, path)

    def __init__(self, gpu_index: int, period_ms: float = 4.0):
        self.gpu, self.proc, self.tmp = gpu_index, None, None
        self.t0 = self.t1 = None
        self.period_ms = period_ms
        self.helper = ClockSampler._helpers.get(gpu_index)
        if self.helper is None:
            try:
                import pynvml  # noqa: F401  (only: is it there?)
                sel = str(gpu_index)
                try:
                    import torch
                    sel = "GPU-" + str(torch.cuda.get_device_properties(gpu_index).uuid)
                except Exception:
                    pass
                tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
                tmp.close()
                proc = subprocess.Popen([sys.executable, "-c", _CLOCK_HELPER, sel, tmp.name, str(gpu_index)],
                                        stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                import atexit
                atexit.register(lambda: (proc.terminate(), [os.path.exists(q) and os.unlink(q) for q in (tmp.name, tmp.name + ".gate")]))
                self.helper = ClockSampler._helpers[gpu_index] = (proc, tmp.name)
            except Exception:
                self.helper = None

    def __enter__(self):
        self.t0 = time.time()
        if self.helper is not None and self.helper[0].poll() is None:
            with open(self.helper[1] + ".gate", "w") as g:     # the helper polls NVML while this file exists
                g.write(str(self.period_ms))
            return self
        self.helper = None                 # the helper died (no NVML): nvidia-smi
        try:
            self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        self.t1 = time.time()
        if self.helper is not None:
            try:
                os.unlink(self.helper[1] + ".gate")
            except OSError:
                pass
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.helper is not None:
            try:
                rows = [l.split(",") for l in open(self.helper[1]).read().strip().splitlines() if l.count(",") == 4]
                rows = [(float(a), float(b), float(c), float(d), int(e)) for a, b, c, d, e in rows]
            except Exception:
                rows = []
            inside = [x for x in rows if self.t0 <= x[0] <= self.t1]
            if not inside:
                return out
            out["samples"] = len(inside)
            out["sm_mhz"] = float(np.median([x[1] for x in inside]))
            out["sm_max_mhz"] = inside[0][2]
            out["power_w_max"] = max(x[3] for x in inside)
            bits = 0
            for x in inside:
                bits |= x[4]
            # nvml.h nvmlClocksEventReason*: SwPowerCap 0x4, HwSlowdown 0x8, SwThermalSlowdown 0x20, HwThermalSlowdown 0x40
            for nm, mask in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)):
                if bits & mask:
                    out["reasons"].append(nm)
            out["source"] = "nvml helper process, samples inside the timed region"
            return out
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
        out["source"] = "nvidia-smi"
        return out


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def tensor_peak():
    """Dense bf16/fp16 TFLOP/s, the sustained figure (the kernel is timed inside a long step)."""
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md ~1.4 PFLOP/s sustained)"


def launch_roofline(rows, nnz, f, fused, ms, ncu):
    """One half-step launch: algorithmic HBM rate and useful tensor rate, with the bound the committed ncu capture shows
    (profiles/traffic.json: DRAM bytes, tensor-pipe %, L2 hit rate of that launch)."""
    if not ms or ms <= 0:
        return None
    hbm, _ = measured_peaks()
    tf, _ = tensor_peak()
    b = gram_bytes(rows, nnz, f, fused)
    flops = 2.0 * f * f * nnz
    out = {"ms": ms, "algorithmic_bytes": b, "algorithmic_gbs": b / 1e9 / (ms / 1e3), "frac_hbm": b / 1e9 / (ms / 1e3) / hbm,
           "useful_tflops": flops / 1e12 / (ms / 1e3), "frac_tensor": flops / 1e12 / (ms / 1e3) / tf}
    if ncu:
        out.update({k: ncu[k] for k in ("dram_bytes", "tensor_pipe_pct", "l2_hit_pct", "smem_data_pipe_pct") if k in ncu})
        if "dram_bytes" in ncu:
            out["dram_gbs"] = ncu["dram_bytes"] / 1e9 / (ms / 1e3)
            if ncu["dram_bytes"] > 0.25 * b:
                out["bound"] = "hbm"
            elif sum((ncu.get("smem_data_pipe_pct") or {}).values()) > 80:
                out["bound"] = ("shared-memory data pipe (tensor-core operand reads + the CG's p broadcasts); the gather source is "
                                "L2-resident, so neither the HBM nor the tensor fraction is this launch's roof")
            else:
                out["bound"] = "tensor+l2 (gather source is L2-resident)"
    return out


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


def _omp_set_threads(n: int) -> bool:
    import ctypes
    for name in ("libgomp.so.1", "libomp.so", "libiomp5.so"):
        try:
            ctypes.CDLL(name).omp_set_num_threads(int(n))
            return True
        except (OSError, AttributeError):
            continue
    return False


def _cpu_sample(r, theta0, X0, w, budget_s):
    """CPU oracle on a bounded sample of rows of both half-steps, scaled by rating share."""
    from oracle import oracle as O
    f, lam = w["f"], w["lam"]
    est = {}
    sample_desc = []
    for side, (ptr, idx, val, fac, out, rows) in {
        "x": (r.csr_indptr, r.csr_indices, r.csr_data, theta0, X0.copy(), r.m),
        "theta": (r.csc_indptr, r.csc_indices, r.csc_data, np.ascontiguousarray(
            np.random.default_rng(0).random((r.m, f), dtype=np.float32) * 0.2), theta0.copy(), r.n),
    }.items():
        # grow the sample until it costs ~budget/4 seconds
        share_rows = max(8, rows // 20000)
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
    return 1.0 / (est["x"] + est["theta"]), "; ".join(sample_desc)


def cpu_baseline(r, theta0, X0, w, budget_s: float = 12.0):
    """north_star: the CPU path on ONE thread (count stated); the all-core figure rides along."""
    cores = os.cpu_count() or 1
    out = {"unit": "iterations/s", "kind": "port"}
    if _omp_set_threads(1):
        v1, d1 = _cpu_sample(r, theta0, X0, w, budget_s)
        out.update({"value": v1, "cores": 1, "sample": "oracle/als_cpu.c, 1 OpenMP thread, on " + d1 + "; scaled by rating share"})
        _omp_set_threads(cores)
        vn, dn = _cpu_sample(r, theta0, X0, w, budget_s)
        out["all_cores"] = {"value": vn, "cores": cores, "sample": dn}
    else:
        vn, dn = _cpu_sample(r, theta0, X0, w, budget_s)
        out.update({"value": vn, "cores": cores, "sample": "oracle/als_cpu.c (OpenMP, all cores: omp_set_num_threads unavailable) on " + dn})
    return out


def pin_inputs(r, theta0, X0):
    """The reference's CLI keeps every input in pinned host memory (cudaMallocHost, main.cpp:50-69): same here."""
    import torch
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    for name in ("csr_indptr", "csr_indices", "csr_data", "csc_indptr", "csc_indices", "csc_data", "coo_row",
                 "test_row", "test_col", "test_val"):
        setattr(r, name, pin(getattr(r, name)))
    return pin(theta0), pin(X0), pin


def e2e_doals(args, r, theta0, X0, f, lam, gpus: int, device: int):
    """The reference-facing call with HOST buffers: cumf_doALS(host pointers, ITERS = steps) -- uploads, K iterations with the
    per-iteration train/test RMSE doALS prints, factor download, release of every device buffer -- wall clock / K.  gpus > 1:
    the same call with CUMF_GPUS=gpus (one process driving the GPUs, rows sharded inside the library)."""
    import torch
    import cumf_als_b200 as c
    os.environ["CUMF_QUIET"] = "1"
    os.environ["CUMF_PATH"] = args.path
    os.environ["CUMF_GPUS"] = str(gpus)
    th0, x0, pin = pin_inputs(r, theta0, X0)
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
    torch.cuda.empty_cache()
    call = lambda th, X, iters: c.do_als(*r.doals_args(), th, X, r.test_row, r.test_col, r.test_val, r.m, r.n, f, r.nnz,
                                         r.nnz_test, lam, iters, 1, 1, device)
    if args.warmup > 0:   # context / allocator / driver warm-up outside the timed call, as in the reference arm
        call(pin(theta0), pin(X0), 1)
    th, X = pin(theta0), pin(X0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fin = call(th, X, args.steps)
    wall = time.perf_counter() - t0
    os.environ["CUMF_GPUS"] = "1"
    ratings = (r.csr_indices.nbytes + r.csr_data.nbytes + r.csc_indices.nbytes + r.csc_data.nbytes + r.coo_row.nbytes +
               r.test_row.nbytes + r.test_col.nbytes + r.test_val.nbytes)
    return {"value": args.steps / wall, "unit": "iterations/s",
            "h2d_bytes_per_step": (ratings + gpus * (th.nbytes + X.nbytes)) // args.steps,
            "d2h_bytes_per_step": (th.nbytes + X.nbytes) // args.steps, "wall_s": wall, "pinned_h2d_gbs_this_box": h2d_gbs,
            "final_test_rmse": fin, "gpus": gpus,
            "what": "cumf_doALS(host pointers, ITERS=steps" + (f", CUMF_GPUS={gpus}" if gpus > 1 else "") + "): every rating "
                    "slice uploaded once, both factors to every GPU, iterations with per-iteration train/test RMSE, ONE factor "
                    "replica downloaded, all device buffers freed; one untimed 1-iteration call first (both arms)"}


def run_ours(args, w):
    import torch
    import torch.distributed as dist
    import cumf_als_b200 as c
    from cumf_als_b200.data import nnz_balanced_ranges

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    ClockSampler(local_rank)        # starts the NVML helper process now: it is polling long before the timed region
    host_pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_pg = dist.new_group(backend="gloo")        # host-side barriers that do not occupy the GPUs
    r, theta0, X0 = make_inputs(w, args.scale, "cuda")
    f, lam = w["f"], w["lam"]
    path = {"auto": c.PATH_AUTO, "simt": c.PATH_SIMT, "tc": c.PATH_TC}[args.path]
    os.environ["CUMF_TIME_KERNELS"] = "1"

    x_ranges = nnz_balanced_ranges(r.csr_indptr, world)
    t_ranges = nnz_balanced_ranges(r.csc_indptr, world)
    xr, tr_ = x_ranges[rank], t_ranges[rank]
    solver = c.AlsSolver(r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, r.coo_row,
                         r.test_row, r.test_col, r.test_val, r.m, r.n, f, lam, x_range=xr, theta_range=tr_,
                         device=local_rank, path=path)
    solver.set_factors(theta0, X0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:
        # connect the ranks' factor replicas (CUDA IPC): from here on every half-step stores its rows into all of them and
        # ends with the device-side barrier -- the library's own multi-GPU path, no collective on the data path
        blobs = [None] * world
        dist.all_gather_object(blobs, solver.ipc_export())
        solver.ipc_import(blobs, rank)
        barrier()

    def rmse():
        tr, te = solver.sse()
        if world > 1:
            t = torch.tensor([tr, te], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            tr, te = (float(v) for v in t.cpu())
        f32 = np.float32
        return float(np.sqrt(f32(tr) / f32(r.nnz))), float(np.sqrt(f32(te) / f32(r.nnz_test)))

    solver.iterate(args.warmup)
    solver.timers(reset=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
        ms = solver.iterate(args.steps)
        barrier()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    tm = solver.timers()
    rank_kernel_ms = None
    if world > 1:      # dominant-kernel time of every rank (the slowest one sets the step: the others wait in the barrier kernel)
        it = max(tm["iterations"], 1)
        mine = torch.tensor([tm["gram_x_ms"] / it, tm["gram_theta_ms"] / it], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_kernel_ms = [[round(float(v), 4) for v in t.cpu()] for t in allr]
    train_rmse, test_rmse = rmse()
    iters_per_s = args.steps / (ms / 1e3)

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        fused = args.path != "simt" and _tc_active(c, f)
        xs, ts = xr[1] - xr[0], tr_[1] - tr_[0]
        nnz_x = int(r.csr_indptr[xr[1]] - r.csr_indptr[xr[0]])
        nnz_t = int(r.csc_indptr[tr_[1]] - r.csc_indptr[tr_[0]])
        gb = (gram_bytes(xs, nnz_x, f, fused) + gram_bytes(ts, nnz_t, f, fused)) / 1e9   # per iteration, this rank
        iters_done = max(tm["iterations"], 1)
        gx, gt = tm["gram_x_ms"] / iters_done, tm["gram_theta_ms"] / iters_done
        gram_ms = gx + gt
        achieved = gb / (gram_ms / 1e3) if gram_ms > 0 else None
        ncu = ncu_traffic(args.workload, "fused" if fused else "simt") if world == 1 and args.scale == 1.0 else None
        impl = os.environ.get("CUMF_TC_IMPL", "")
        line = {
            "metric": METRIC.format(workload=args.workload, f=f), "value": iters_per_s,
            "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": DTYPE if fused else "f32", "data": "synthetic",
            "config": common_config(args, r, w),
            "impl_detail": {"path": ("tcgen05-fused" + (f" (CUMF_TC_IMPL={impl})" if impl else "")) if fused else "simt-unfused",
                            "sharding": "single GPU" if world == 1 else
                            f"rows rating-balanced over {world} ranks; solver epilogues store rows into every rank's replica "
                            f"(CUDA IPC peer pointers), flag-barrier kernel per half-step"},
            "x_ms": tm["x_ms"] / iters_done, "theta_ms": tm["theta_ms"] / iters_done,
            "train_rmse": train_rmse, "test_rmse": test_rmse,
            "gpu_launches": int(tm["launches"]),
            "clocks": clocks.summary(),
            **({"rank_kernel_ms": {"x_side_theta_side_per_rank": rank_kernel_ms,
                                   "slowest_sum": max(a + b for a, b in rank_kernel_ms),
                                   "sum_of_per_side_maxima": max(a for a, _ in rank_kernel_ms) + max(b for _, b in rank_kernel_ms)}}
               if rank_kernel_ms else {}),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": (sum(v.get("dram_bytes", 0) for v in ncu.values() if isinstance(v, dict)) or None) if ncu else None,
                         "kernel": "gram (+rhs" + (" + fused CG)" if fused else ")") + ", X-side launch + theta-side launch (rank 0)",
                         "bytes_per_iteration_gb": gb, "kernel_ms_per_iteration": gram_ms,
                         "formula": "B_gram_fused" if fused else "B_gram (materialised A)", "peak_source": peak_src,
                         "traffic_source": (ncu or {}).get("source"),
                         "launches": {"x_side": launch_roofline(xs, nnz_x, f, fused, gx, (ncu or {}).get("x_side")),
                                      "theta_side": launch_roofline(ts, nnz_t, f, fused, gt, (ncu or {}).get("theta_side")),
                                      "tensor_peak_source": tensor_peak()[1]}},
        }
    solver.close()
    torch.cuda.empty_cache()

    # e2e: the reference-facing call with host buffers.  N > 1: rank 0 alone calls cumf_doALS with CUMF_GPUS=N (the library
    # shards inside the call); the other ranks have released their shards and wait on a host-side (gloo) barrier.
    if not args.no_e2e:
        if world > 1:
            dist.barrier(group=host_pg)
        if rank == 0:
            try:
                line["e2e"] = e2e_doals(args, r, theta0, X0, f, lam, world, local_rank)
            except Exception as exc:        # the resident numbers above stay valid
                line["e2e"] = {"error": f"{type(exc).__name__}: {exc}"}
        if world > 1:
            dist.barrier(group=host_pg)
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
    ClockSampler(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    r, theta0, X0 = make_inputs(w, args.scale, "cuda")
    f, lam = w["f"], w["lam"]
    path = {"auto": c.PATH_AUTO, "simt": c.PATH_SIMT, "tc": c.PATH_TC}[args.path]
    t_ranges = nnz_balanced_ranges(r.csc_indptr, world)
    if args.free_sms > 0:      # leave SMs to the collective's kernels: the persistent Gram kernel otherwise owns every SM
        os.environ["CUMF_TC_CTAS"] = str(max(1, torch.cuda.get_device_properties(local_rank).multi_processor_count - args.free_sms))
    eng = GpuPartialGramEngine(r, f, lam, theta0, X0, t_ranges[rank], local_rank, path=path, cap_bytes=args.e2_batch_mb << 20)
    drv = PartialGramAls(eng, r.nnz, r.nnz_test)
    drv.overlap = args.e2_overlap

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
    # one more, instrumented and serial, iteration for the phase split (outside the timed region): per phase summed over the
    # row batches
    phases = {"partial_gram_ms": 0.0, "allreduce_ms": 0.0, "cg_x_ms": 0.0, "theta_ms": 0.0}
    barrier()
    for b in range(len(eng.batches)):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        tt, rhs = eng.partial_gram(b)
        ev[1].record()
        if world > 1:
            dist.all_reduce(tt)
            dist.all_reduce(rhs)
        ev[2].record()
        eng.solve_x(b, tt, rhs)
        ev[3].record()
        torch.cuda.synchronize()
        for i, k in enumerate(["partial_gram_ms", "allreduce_ms", "cg_x_ms"]):
            phases[k] += ev[i].elapsed_time(ev[i + 1])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.update_theta()
    e1.record()
    torch.cuda.synchronize()
    phases["theta_ms"] = e0.elapsed_time(e1)
    phases["serial_sum_ms"] = sum(phases.values())
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
        # how much of the all-reduce the timed (possibly overlapped) iteration hides: serial phase sum vs measured step
        if phases and phases["allreduce_ms"] > 0:
            phases["allreduce_hidden_frac"] = max(0.0, min(1.0, (phases["serial_sum_ms"] - ms / args.steps) / phases["allreduce_ms"]))
        achieved = gb / (kernel_ms / 1e3) if kernel_ms else None
        print(json.dumps({
            "metric": METRIC.format(workload=args.workload, f=f), "value": args.steps / (ms / 1e3),
            "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": DTYPE if fused else "f32", "data": "synthetic",
            "config": common_config(args, r, w),
            "impl_detail": {"path": "tcgen05" if fused else "simt",
                            "sharding": f"theta rows nnz-balanced over {world} ranks; X-step = partial [A|b] + NCCL all-reduce "
                                        f"+ replicated CG; theta-step rank-local (no factor exchange)",
                            "x_batches": len(eng.batches), "overlap": bool(args.e2_overlap and len(eng.batches) > 1),
                            "free_sms_for_the_collective": args.free_sms},
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
    ClockSampler(local_rank)
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
    if args.warmup > 0:   # context / cuBLAS / cuSPARSE initialisation outside the timed call: W iterations, like the other arm
        with CaptureStdout():
            O.ref_do_als(r, theta0.copy(), X0.copy(), f, lam, args.warmup, xb, tb, variant, local_rank)
    th, X = pin(theta0), pin(X0)
    iters = args.steps
    with ClockSampler(local_rank, period_ms=20.0) as clocks:      # (the reference makes driver calls inside its timed call)
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
        "unit": "iterations/s", "n_gpus": 1, "steps": iters, "warmup": args.warmup, "ms_per_step": 1e3 * als_s / iters,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": common_config(args, r, w),
        "impl_detail": {"solver": "cg6" if variant == "cg" else "cublas-lu (getrfBatched, no pivoting)", "X_BATCH": xb, "THETA_BATCH": tb,
                        "what": "unmodified reference sources (oracle/_ref), Kepler-era kernels recompiled for sm_100a; value = "
                                "iterations / sum of its own `update X run`+`update theta run` timers (als.cu:850, 963); "
                                "one untimed call of `warmup` iterations first"},
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
    ap.add_argument("--e2-overlap", action="store_true", help="partial-gram: all-reduce of batch k under the Gram of batch k+1")
    ap.add_argument("--e2-batch-mb", type=int, default=8192, help="partial-gram: workspace per X-row batch (MiB)")
    ap.add_argument("--free-sms", type=int, default=0, help="partial-gram: SMs the Gram kernel leaves to the collective")
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
