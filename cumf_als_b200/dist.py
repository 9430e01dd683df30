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


class ShardedAls:
    def __init__(self, engine, x_ranges, theta_ranges, nnz: int, nnz_test: int, group=None):
        self.e = engine
        self.x_ranges, self.theta_ranges = list(x_ranges), list(theta_ranges)
        self.nnz, self.nnz_test = nnz, nnz_test
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        assert len(self.x_ranges) == self.world and len(self.theta_ranges) == self.world

    def _exchange(self, full: torch.Tensor, ranges):
        """Every rank broadcasts the row block it owns (uneven blocks: rating-balanced, not row-balanced)."""
        if self.world == 1:
            return
        for r, (lo, hi) in enumerate(ranges):
            if hi > lo:
                dist.broadcast(full[lo:hi], src=dist.get_global_rank(self.group, r) if self.group else r, group=self.group)

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
