"""The N>1 host logic on CPU: world_size-2 gloo run of the sharded driver (cumf_als_b200.dist) with
the oracle standing in for the per-rank GPU engine.  Sharded result == single-process result,
bit for bit (rows are independent given the opposing factor)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


class OracleEngine:
    """Same interface as cumf_als_b200.dist.GpuEngine, computed by the CPU oracle on a row shard."""

    def __init__(self, r, f, lam, theta0, x_range, theta_range, test_slice):
        from oracle import oracle as O
        self.O, self.r, self.f, self.lam = O, r, f, lam
        self.m, self.n = r.m, r.n
        self.x = torch.zeros((r.m, f), dtype=torch.float32)
        self.theta = torch.from_numpy(theta0.copy())
        self.x_range, self.theta_range, self.test_slice = x_range, theta_range, test_slice

    def update_x(self):
        out = self.x.numpy()
        self.O.half_step(self.r.csr_indptr, self.r.csr_indices, self.r.csr_data, self.theta.numpy(), out, self.f, self.lam,
                         0, 6.0, *self.x_range)

    def update_theta(self):
        out = self.theta.numpy()
        self.O.half_step(self.r.csc_indptr, self.r.csc_indices, self.r.csc_data, self.x.numpy(), out, self.f, self.lam,
                         0, 6.0, *self.theta_range)

    def sse(self):
        r, th, X = self.r, self.theta.numpy().astype(np.float64), self.x.numpy().astype(np.float64)
        lo, hi = r.csr_indptr[self.x_range[0]], r.csr_indptr[self.x_range[1]]
        e = r.csr_data[lo:hi] - np.einsum("ij,ij->i", th[r.csr_indices[lo:hi]], X[r.coo_row[lo:hi]])
        t0, t1 = self.test_slice
        et = r.test_val[t0:t1] - np.einsum("ij,ij->i", th[r.test_col[t0:t1]], X[r.test_row[t0:t1]])
        return float((e ** 2).sum()), float((et ** 2).sum())


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cumf_als_b200.data import init_factors, synth_ratings
    from cumf_als_b200.dist import ShardedAls, shard_ranges
    r = synth_ratings(160, 230, 7000, 900, seed=31)
    f, lam = 20, 0.05
    theta0, _ = init_factors(r.m, r.n, f, seed=4)
    xr, tr = shard_ranges(r.csr_indptr, r.csc_indptr, world)
    eff = 256 * ((r.nnz_test - 1) // 256)                      # the reference's test launch (als.cu:1006)
    test_slice = (eff * xr[rank][0] // r.m, eff * xr[rank][1] // r.m)
    eng = OracleEngine(r, f, lam, theta0, xr[rank], tr[rank], test_slice)
    sh = ShardedAls(eng, xr, tr, r.nnz, r.nnz_test)
    sh.iterate(2)
    train, test = sh.rmse()
    np.savez(Path(out_dir) / f"rank{rank}.npz", x=eng.x.numpy(), theta=eng.theta.numpy(), train=train, test=test)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_equals_single_process(tmp_path):
    world, port = 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["theta"], b["theta"])     # replicas agree
    # single process oracle
    from cumf_als_b200.data import init_factors, synth_ratings
    from oracle import oracle as O
    r = synth_ratings(160, 230, 7000, 900, seed=31)
    theta0, X0 = init_factors(r.m, r.n, 20, seed=4)
    th, X = theta0.copy(), X0.copy()
    fin, hist = O.do_als(r, th, X, 20, 0.05, 2, 0)
    assert np.array_equal(a["x"], X) and np.array_equal(a["theta"], th)                 # bit-exact vs unsharded
    assert float(a["train"]) == pytest.approx(float(hist[-1, 0]), rel=1e-5)
    assert float(a["test"]) == pytest.approx(float(hist[-1, 1]), rel=1e-5)
    assert float(a["test"]) == pytest.approx(float(b["test"]), rel=1e-12)


def test_shard_ranges_partition_every_row():
    from cumf_als_b200.data import synth_ratings
    from cumf_als_b200.dist import shard_ranges
    r = synth_ratings(300, 500, 20000, 100, seed=2)
    for world in (1, 2, 4, 8):
        xr, tr = shard_ranges(r.csr_indptr, r.csc_indptr, world)
        for ranges, rows in ((xr, r.m), (tr, r.n)):
            assert ranges[0][0] == 0 and ranges[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
