"""One-process-per-GPU sharding of the ALS half-steps.

Replaces the X_BATCH / THETA_BATCH loop (als.cu:768-777, 881-890) and hugewiki's dynamic batch
queue (hugewiki.cu:2446-2496): rank g owns a contiguous, rating-balanced range of X rows and of
theta rows, updates exactly those rows from a full replica of the opposing factor (no
collective inside the half-step), then the updated row blocks are exchanged so every rank holds
both full factors again.  torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo in
the CPU tests); the per-rank compute is the C-ABI library through an `engine`.

An engine exposes: m, n, f, x (torch tensor [m, f]), theta ([n, f]) -- views of the buffers the
engine updates in place --, update_x(), update_theta(), sse() -> (train_sse, test_sse).
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist

from .data import nnz_balanced_ranges


class _DevArray:
    """Expose a raw device pointer to torch through the CUDA array interface (no copy)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class GpuEngine:
    """cumf_als_solver (include/cumf_als.h) as a sharded engine."""

    def __init__(self, solver, device: int):
        self.solver = solver
        self.m, self.n, self.f = solver.m, solver.n, solver.f
        dev = torch.device("cuda", device)
        self.x = torch.as_tensor(_DevArray(solver.x_ptr, (self.m, self.f)), device=dev)
        self.theta = torch.as_tensor(_DevArray(solver.theta_ptr, (self.n, self.f)), device=dev)

    def update_x(self):
        self.solver.update_x()

    def update_theta(self):
        self.solver.update_theta()

    def sse(self):
        return self.solver.sse()


def shard_ranges(csr_indptr, csc_indptr, world: int):
    """(x_ranges, theta_ranges): rating-balanced contiguous row ranges, one per rank (integer path)."""
    return nnz_balanced_ranges(csr_indptr, world), nnz_balanced_ranges(csc_indptr, world)


def exchange_plan(rank: int, ranges):
    """Point-to-point operations of `rank` for the row-block exchange: ("send", peer, lo, hi) of its own block to every
    other rank, ("recv", peer, lo, hi) of every other rank's block; empty blocks move nothing on either side, so every send
    has exactly one matching receive (integer path, tested for every world size in tests/test_dist_gloo.py)."""
    lo_me, hi_me = ranges[rank]
    plan = []
    for r, (lo, hi) in enumerate(ranges):
        if r == rank:
            continue
        if hi_me > lo_me:
            plan.append(("send", r, lo_me, hi_me))
        if hi > lo:
            plan.append(("recv", r, lo, hi))
    return plan


class ShardedAls:
    def __init__(self, engine, x_ranges, theta_ranges, nnz: int, nnz_test: int, group=None):
        self.e = engine
        self.x_ranges, self.theta_ranges = list(x_ranges), list(theta_ranges)
        self.nnz, self.nnz_test = nnz, nnz_test
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        assert len(self.x_ranges) == self.world and len(self.theta_ranges) == self.world
        import os
        self._p2p = (dist.is_initialized() and dist.get_backend(group) == "nccl" and
                     os.environ.get("CUMF_EXCHANGE", "p2p") != "broadcast")

    def _exchange(self, full: torch.Tensor, ranges):
        """Every rank hands the row block it owns to every other rank (uneven blocks: rating-balanced, not
        row-balanced, so this is not an ncclAllGather).  NCCL: one grouped launch of point-to-point sends/receives
        (`batch_isend_irecv`: all blocks move at once over NVSwitch, 2 launches per iteration instead of 2 N broadcasts);
        gloo (the CPU tests), or CUMF_EXCHANGE=broadcast: one broadcast per owner.  Same bytes land in the same places."""
        if self.world == 1:
            return
        peer = (lambda r: dist.get_global_rank(self.group, r)) if self.group else (lambda r: r)
        if self._p2p and full.is_cuda:
            ops = [dist.P2POp(dist.isend if kind == "send" else dist.irecv, full[lo:hi], peer(r), self.group)
                   for kind, r, lo, hi in exchange_plan(self.rank, ranges)]
            if ops:
                for work in dist.batch_isend_irecv(ops):
                    work.wait()          # stream-ordered on NCCL: the next half-step's launch queues behind the exchange
            return
        for r, (lo, hi) in enumerate(ranges):
            if hi > lo:
                dist.broadcast(full[lo:hi], src=peer(r), group=self.group)

    def step(self):
        """One ALS iteration: als.cu:727-961 with both half-steps sharded."""
        self.e.update_x()
        self._exchange(self.e.x, self.x_ranges)
        self.e.update_theta()
        self._exchange(self.e.theta, self.theta_ranges)

    def iterate(self, iters: int) -> float:
        """`iters` iterations; returns elapsed milliseconds on this rank (CUDA events on GPU)."""
        on_gpu = self.e.x.is_cuda
        if on_gpu:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
        else:
            t0 = time.perf_counter()
        for _ in range(iters):
            self.step()
        if on_gpu:
            e1.record()
            e1.synchronize()
            return float(e0.elapsed_time(e1))
        return 1e3 * (time.perf_counter() - t0)

    def rmse(self):
        """(train, test) RMSE: per-shard sums of squared errors, all-reduced, then sqrt(S / count) in fp32
        like als.cu:991 / 1018."""
        tr, te = self.e.sse()
        t = torch.tensor([tr, te], dtype=torch.float64, device=self.e.x.device)
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        tr, te = (float(v) for v in t.cpu())
        f32 = np.float32
        return (float(np.sqrt(f32(tr) / f32(self.nnz))) if self.nnz else 0.0,
                float(np.sqrt(f32(te) / f32(self.nnz_test))) if self.nnz_test else 0.0)


# ================================================================================================
# E2: data-parallel partial Gram + all-reduce (hugewiki.cu:2629-2827)
# ================================================================================================
# The long side (theta, n rows) is sharded: rank g owns theta rows [t0, t1) and only the ratings of
# those columns.  X-step: every rank forms the PARTIAL [A_u | b_u] of every X row over its own
# ratings (lambda scaled by the local count, hugewiki.cu:1675-1678), the partials are summed over
# the ranks (the reference: peer copies + reduce kernels, hugewiki.cu:2769-2827; here one
# all-reduce on the same buffers) and every rank solves the same m systems, so X stays replicated
# without an exchange.  Theta-step: rank-local (its rows only need its own ratings and the
# replicated X); theta is never gathered.


def local_share(csr_indptr: np.ndarray, csr_indices: np.ndarray, t0: int, t1: int):
    """Positions of rank-local ratings inside each CSR row: (begin, end) int64 arrays, row u's ratings
    with column id in [t0, t1) are exactly positions [begin[u], end[u]) because column ids ascend
    inside a row.  Integer path (bit-exact against a mask-based count in the tests)."""
    indptr = np.asarray(csr_indptr, dtype=np.int64)
    col = np.asarray(csr_indices, dtype=np.int64)
    m = indptr.size - 1
    n_bound = int(col.max()) + 1 if col.size else 1
    n_bound = max(n_bound, t1, 1)
    row_of = np.repeat(np.arange(m, dtype=np.int64), np.diff(indptr))
    keys = row_of * n_bound + col                         # ascending: CSR order with sorted columns
    rows = np.arange(m, dtype=np.int64) * n_bound
    begin = np.searchsorted(keys, rows + t0, side="left")
    end = np.searchsorted(keys, rows + t1, side="left")
    return begin.astype(np.int64), end.astype(np.int64)


def compact_share(csr_indptr, csr_indices, csr_data, t0: int, t1: int):
    """The rank-local CSR (indptr int64, indices int32, data float32): every row keeps only its ratings with
    column id in [t0, t1).  What one rank uploads for the X-step."""
    begin, end = local_share(csr_indptr, csr_indices, t0, t1)
    cnt = end - begin
    indptr = np.zeros(cnt.size + 1, dtype=np.int64)
    np.cumsum(cnt, out=indptr[1:])
    col = np.asarray(csr_indices)
    keep = (col >= t0) & (col < t1)
    return indptr, np.ascontiguousarray(col[keep], dtype=np.int32), np.ascontiguousarray(np.asarray(csr_data)[keep], dtype=np.float32)


def row_batches(rows: int, f: int, cap_bytes: int):
    """Split [0, rows) into batches whose materialised [A|b] fits `cap_bytes` (the reference's X_BATCH,
    als.cu:768-777, chosen from memory instead of by the user)."""
    per = max(1, cap_bytes // (4 * f * (f + 1)))
    return [(b, min(rows, b + per)) for b in range(0, max(rows, 1), per) if b < rows]


class GpuPartialGramEngine:
    """Per-rank state of the E2 scheme on one B200, driven through the C ABI (cumf_plan_create_ranges,
    cumf_plan_gram, cumf_cg, cumf_plan_create / cumf_update_factor, cumf_rmse)."""

    def __init__(self, r, f: int, lam: float, theta0: np.ndarray, x0: np.ndarray, t_range, device: int,
                 path: int = 0, cg_iter: float = 6.0, cap_bytes: int = 8 << 30):
        from . import api
        self.api, self.f, self.lam, self.cg_iter = api, f, lam, cg_iter
        self.m, self.n = r.m, r.n
        self.t0, self.t1 = t_range
        dev = self.dev = torch.device("cuda", device)
        torch.cuda.set_device(dev)
        t0, t1 = self.t0, self.t1
        # X-step data: the rank-local CSR
        indptr, col, val = compact_share(r.csr_indptr, r.csr_indices, r.csr_data, t0, t1)
        self.local_nnz = int(indptr[-1])
        self.x_col = torch.from_numpy(col).to(dev)
        self.x_val = torch.from_numpy(val).to(dev)
        self.batches = row_batches(self.m, f, cap_bytes)
        self.x_plans = [api.Plan.from_ranges(indptr[b0:b1], indptr[b0 + 1:b1 + 1], f, path) for b0, b1 in self.batches]
        rows_max = max((b1 - b0 for b0, b1 in self.batches), default=1)
        self.tt = torch.empty((rows_max, f * f), dtype=torch.float32, device=dev)
        self.rhs = torch.empty((rows_max, f), dtype=torch.float32, device=dev)
        self.tt2 = self.rhs2 = None         # second workspace, allocated by the overlapped driver
        # theta-step data: the owned CSC columns, rebased
        cp = np.asarray(r.csc_indptr, dtype=np.int64)
        lo, hi = int(cp[t0]), int(cp[t1])
        self.t_indptr = (cp[t0:t1 + 1] - lo).astype(np.int32)
        self.t_row = torch.from_numpy(np.ascontiguousarray(r.csc_indices[lo:hi], dtype=np.int32)).to(dev)
        self.t_val = torch.from_numpy(np.ascontiguousarray(r.csc_data[lo:hi], dtype=np.float32)).to(dev)
        self.t_plan = api.Plan(self.t_indptr, 0, t1 - t0, f, path) if t1 > t0 else None
        # RMSE samples of the owned users: train = the CSC slice, test = the launched test samples (als.cu:1006)
        self.t_col = torch.from_numpy(np.repeat(np.arange(t0, t1, dtype=np.int32), np.diff(self.t_indptr))).to(dev)
        eff = 256 * ((r.nnz_test - 1) // 256) if r.nnz_test > 0 else 0
        tc = np.asarray(r.test_col[:eff])
        keep = (tc >= t0) & (tc < t1)
        self.test_row = torch.from_numpy(np.ascontiguousarray(np.asarray(r.test_row[:eff])[keep], dtype=np.int32)).to(dev)
        self.test_col = torch.from_numpy(np.ascontiguousarray(tc[keep], dtype=np.int32)).to(dev)
        self.test_val = torch.from_numpy(np.ascontiguousarray(np.asarray(r.test_val[:eff])[keep], dtype=np.float32)).to(dev)
        # factors: X replicated, theta valid on the owned rows only
        self.x = torch.from_numpy(np.ascontiguousarray(x0, dtype=np.float32)).to(dev)
        self.theta = torch.from_numpy(np.ascontiguousarray(theta0, dtype=np.float32)).to(dev)
        self.launches = 0

    def partial_gram(self, batch: int, buf: int = 0, stream=None):
        """Partial [A|b] of the batch's rows over this rank's ratings into workspace `buf` (two workspaces when the
        all-reduce of one batch overlaps the Gram of the next)."""
        b0, b1 = self.batches[batch]
        if buf == 1 and self.tt2 is None:
            self.tt2, self.rhs2 = torch.empty_like(self.tt), torch.empty_like(self.rhs)
        tt, rhs = (self.tt, self.rhs) if buf == 0 else (self.tt2, self.rhs2)
        tt, rhs = tt[: b1 - b0], rhs[: b1 - b0]
        p = self.x_plans[batch]
        p.gram(self.x_col, self.x_val, self.theta, self.lam, tt, rhs, stream=stream)
        self.launches += p.last_launches
        return tt, rhs

    def solve_x(self, batch: int, tt, rhs, stream=None):
        b0, b1 = self.batches[batch]
        self.api.cg(tt, self.x[b0:b1], rhs, b1 - b0, self.f, self.cg_iter, stream=stream)
        self.launches += 1

    def update_theta(self):
        if self.t_plan is None:
            return
        self.api.update_factor(self.t_plan, self.t_row, self.t_val, self.x, self.theta[self.t0:], self.lam,
                               self.api.SOLVER_CG, self.cg_iter)
        self.launches += self.t_plan.last_launches

    def sse(self):
        tr = self.api.rmse(self.t_val, self.t_row, self.t_col, self.theta, self.x, int(self.t_val.numel()), self.f)[1] \
            if self.t_val.numel() else 0.0
        te = self.api.rmse(self.test_val, self.test_row, self.test_col, self.theta, self.x, int(self.test_val.numel()),
                           self.f)[1] if self.test_val.numel() else 0.0
        return tr, te


class PartialGramAls:
    """ALS iterations under the E2 scheme.  `engine` exposes: batches, x, theta, partial_gram(batch) ->
    (tt, rhs) tensors holding this rank's partial systems, solve_x(batch, tt, rhs), update_theta(), sse()."""

    def __init__(self, engine, nnz: int, nnz_test: int, group=None):
        self.e, self.nnz, self.nnz_test, self.group = engine, nnz, nnz_test, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.allreduce_bytes = 0
        self.overlap = False            # double-buffered batches: all-reduce(k) under Gram(k+1)
        self._comm = None

    def step(self):
        if self.overlap and self.world > 1 and len(self.e.batches) > 1 and self.e.x.is_cuda:
            return self._step_overlapped()
        e = self.e
        for b in range(len(e.batches)):
            tt, rhs = e.partial_gram(b)
            if self.world > 1:
                dist.all_reduce(tt, group=self.group)
                dist.all_reduce(rhs, group=self.group)
                self.allreduce_bytes += 4 * (tt.numel() + rhs.numel())
            e.solve_x(b, tt, rhs)
        e.update_theta()

    def _step_overlapped(self):
        """Batches double-buffered over two streams: the all-reduce of batch k (communication stream) runs while the compute
        stream forms the partial Gram of batch k+1; the replicated CG of batch k follows on the compute stream once its sum
        has arrived.  compute: G0 G1 S0 G2 S1 ...   comm: A0 A1 A2 ...   (hugewiki.cu:2703-2745 does reduce + broadcast
        serially after each batch).  The Gram kernel must leave a few SMs free for the collective (CUMF_TC_CTAS)."""
        e = self.e
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=e.x.device)
        comp = torch.cuda.current_stream(e.x.device)
        nb = len(e.batches)
        done = [None] * nb          # all-reduce of batch b finished
        bufs = [None] * nb

        def gram(b):
            bufs[b] = e.partial_gram(b, buf=b % 2)
            ev = torch.cuda.Event()
            ev.record(comp)
            with torch.cuda.stream(self._comm):
                self._comm.wait_event(ev)
                dist.all_reduce(bufs[b][0], group=self.group)
                dist.all_reduce(bufs[b][1], group=self.group)
                done[b] = torch.cuda.Event()
                done[b].record(self._comm)
            self.allreduce_bytes += 4 * (bufs[b][0].numel() + bufs[b][1].numel())

        def solve(b):
            comp.wait_event(done[b])
            e.solve_x(b, *bufs[b])

        gram(0)
        for b in range(1, nb):
            gram(b)             # workspace b % 2 was released by solve(b - 2), issued earlier on the compute stream
            solve(b - 1)
        solve(nb - 1)
        e.update_theta()

    iterate = ShardedAls.iterate

    def rmse(self):
        tr, te = self.e.sse()
        t = torch.tensor([tr, te], dtype=torch.float64, device=self.e.x.device)
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        tr, te = (float(v) for v in t.cpu())
        f32 = np.float32
        return (float(np.sqrt(f32(tr) / f32(self.nnz))) if self.nnz else 0.0,
                float(np.sqrt(f32(te) / f32(self.nnz_test))) if self.nnz_test else 0.0)

    def gather_theta(self, t_ranges) -> torch.Tensor:
        """Full theta on every rank (for export / comparison only; the iterations never need it)."""
        if self.world > 1:
            for r, (lo, hi) in enumerate(t_ranges):
                if hi > lo:
                    dist.broadcast(self.e.theta[lo:hi], src=dist.get_global_rank(self.group, r) if self.group else r,
                                   group=self.group)
        return self.e.theta
