"""The N>1 host logic on CPU: world_size-2 (and 4) gloo runs of the sharded driver (cumf_als_b200.dist) with
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


@pytest.mark.timeout(300)
def test_four_rank_sharded_equals_single_process(tmp_path):
    """Same driver on 4 ranks (uneven rating-balanced blocks, one of them much shorter than the others): every replica
    equals the unsharded oracle bit for bit -- the exchange covers every row exactly once at any world size."""
    world, port = 4, 33000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from cumf_als_b200.data import init_factors, synth_ratings
    from oracle import oracle as O
    r = synth_ratings(160, 230, 7000, 900, seed=31)
    theta0, X0 = init_factors(r.m, r.n, 20, seed=4)
    th, X = theta0.copy(), X0.copy()
    fin, hist = O.do_als(r, th, X, 20, 0.05, 2, 0)
    tests = []
    for rank in range(world):
        a = np.load(tmp_path / f"rank{rank}.npz")
        assert np.array_equal(a["x"], X) and np.array_equal(a["theta"], th)
        tests.append(float(a["test"]))
    assert tests[0] == pytest.approx(float(hist[-1, 1]), rel=1e-5) and max(tests) == min(tests)


def test_exchange_plan_sends_and_receives_pair_up():
    """The NCCL row-block exchange is one grouped launch of sends and receives per rank: every send must meet exactly one
    receive of the same rows on the peer (an unmatched one would hang the group), and together the receives of a rank
    cover every row it does not own exactly once -- at every world size, with empty blocks in the partition."""
    from cumf_als_b200.dist import exchange_plan
    rng = np.random.default_rng(5)
    for world in (2, 3, 4, 8):
        for trial in range(20):
            rows = int(rng.integers(world, 400))
            cuts = np.sort(rng.integers(0, rows + 1, world - 1))          # repeated cut points -> empty blocks
            bounds = [0, *cuts.tolist(), rows]
            ranges = [(bounds[i], bounds[i + 1]) for i in range(world)]
            plans = [exchange_plan(r, ranges) for r in range(world)]
            for a in range(world):
                sends = [(p, lo, hi) for k, p, lo, hi in plans[a] if k == "send"]
                recvs = [(p, lo, hi) for k, p, lo, hi in plans[a] if k == "recv"]
                for peer, lo, hi in sends:
                    assert (lo, hi) == ranges[a] and hi > lo
                    assert [(p, l, h) for k, p, l, h in plans[peer] if k == "recv" and p == a] == [(a, lo, hi)]
                for peer, lo, hi in recvs:
                    assert (lo, hi) == ranges[peer] and ("send", a, lo, hi) in plans[peer]
                covered = np.zeros(rows, np.int32)
                for _, lo, hi in recvs:
                    covered[lo:hi] += 1
                covered[ranges[a][0]:ranges[a][1]] += 1
                assert (covered == 1).all()


def test_shard_ranges_partition_every_row():
    from cumf_als_b200.data import synth_ratings
    from cumf_als_b200.dist import shard_ranges
    r = synth_ratings(300, 500, 20000, 100, seed=2)
    for world in (1, 2, 4, 8):
        xr, tr = shard_ranges(r.csr_indptr, r.csc_indptr, world)
        for ranges, rows in ((xr, r.m), (tr, r.n)):
            assert ranges[0][0] == 0 and ranges[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))


# ---- E2: partial Gram + all-reduce (hugewiki.cu:2629-2827) -------------------------------------------------------
class OraclePartialGramEngine:
    """Same interface as cumf_als_b200.dist.GpuPartialGramEngine, computed by the CPU oracle."""

    def __init__(self, r, f, lam, theta0, x0, t_range, cap_bytes):
        from cumf_als_b200.dist import compact_share, row_batches
        from oracle import oracle as O
        self.O, self.r, self.f, self.lam = O, r, f, lam
        self.t0, self.t1 = t_range
        self.indptr, self.col, self.val = compact_share(r.csr_indptr, r.csr_indices, r.csr_data, *t_range)
        self.batches = row_batches(r.m, f, cap_bytes)
        self.x = torch.from_numpy(x0.copy())
        self.theta = torch.from_numpy(theta0.copy())
        # rows this rank does not own are never read: poison them to prove it
        own = np.zeros(r.n, bool)
        own[self.t0:self.t1] = True
        self.theta[torch.from_numpy(~own)] = float("nan")

    def partial_gram(self, b):
        b0, b1 = self.batches[b]
        ip = (self.indptr[b0:b1 + 1] - self.indptr[b0]).astype(np.int32)
        lo, hi = int(self.indptr[b0]), int(self.indptr[b1])
        th = self.theta.numpy()
        tt = self.O.gram(ip, self.col[lo:hi], th, self.f, self.lam).reshape(b1 - b0, -1)
        rhs = self.O.rhs(ip, self.col[lo:hi], self.val[lo:hi], th, self.f)
        return torch.from_numpy(tt), torch.from_numpy(rhs)

    def solve_x(self, b, tt, rhs):
        b0, b1 = self.batches[b]
        self.x[b0:b1] = torch.from_numpy(self.O.cg(tt.numpy(), self.x[b0:b1].numpy(), rhs.numpy(), self.f))

    def update_theta(self):
        out = self.theta.numpy()
        self.O.half_step(self.r.csc_indptr, self.r.csc_indices, self.r.csc_data, self.x.numpy(), out, self.f, self.lam,
                         0, 6.0, self.t0, self.t1)

    def sse(self):
        r, th, X = self.r, self.theta.numpy().astype(np.float64), self.x.numpy().astype(np.float64)
        lo, hi = r.csc_indptr[self.t0], r.csc_indptr[self.t1]
        cols = np.repeat(np.arange(self.t0, self.t1), np.diff(r.csc_indptr[self.t0:self.t1 + 1]))
        e = r.csc_data[lo:hi] - np.einsum("ij,ij->i", th[cols], X[r.csc_indices[lo:hi]])
        eff = 256 * ((r.nnz_test - 1) // 256)
        keep = (r.test_col[:eff] >= self.t0) & (r.test_col[:eff] < self.t1)
        et = r.test_val[:eff][keep] - np.einsum("ij,ij->i", th[r.test_col[:eff][keep]], X[r.test_row[:eff][keep]])
        return float((e ** 2).sum()), float((et ** 2).sum())


def _worker_e2(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cumf_als_b200.data import init_factors, nnz_balanced_ranges, synth_ratings
    from cumf_als_b200.dist import PartialGramAls
    r = synth_ratings(160, 230, 7000, 900, seed=31)
    f, lam = 20, 0.05
    theta0, x0 = init_factors(r.m, r.n, f, seed=4)
    tr = nnz_balanced_ranges(r.csc_indptr, world)
    eng = OraclePartialGramEngine(r, f, lam, theta0, x0, tr[rank], cap_bytes=4 * f * (f + 1) * 64)   # 3 row batches
    assert len(eng.batches) == 3
    sh = PartialGramAls(eng, r.nnz, r.nnz_test)
    sh.iterate(2)
    train, test = sh.rmse()
    own_theta = eng.theta[tr[rank][0]:tr[rank][1]].numpy().copy()
    theta = sh.gather_theta(tr).numpy()
    np.savez(Path(out_dir) / f"rank{rank}.npz", x=eng.x.numpy(), theta=theta, own=own_theta, train=train, test=test,
             allreduce_bytes=sh.allreduce_bytes)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_partial_gram_allreduce_matches_single_process(tmp_path):
    world, port = 2, 31000 + os.getpid() % 2000
    mp.spawn(_worker_e2, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(a["x"], b["x"])                      # X stays replicated without an exchange
    assert np.array_equal(a["theta"], b["theta"]) and np.isfinite(a["theta"]).all()
    assert int(a["allreduce_bytes"]) == 2 * 4 * 160 * (20 * 20 + 20)
    from cumf_als_b200.data import init_factors, synth_ratings
    from oracle import oracle as O
    r = synth_ratings(160, 230, 7000, 900, seed=31)
    theta0, X0 = init_factors(r.m, r.n, 20, seed=4)
    th, X = theta0.copy(), X0.copy()
    fin, hist = O.do_als(r, th, X, 20, 0.05, 2, 0)
    # the sum of per-rank partials differs from the sequential sum in the last bits and six CG steps amplify that
    # (same sensitivity as the reference's own run-to-run spread, tests/test_oracle.py): fp32 tolerance, not bit-exact
    assert np.abs(a["x"] - X).max() <= 1e-3 * np.abs(X).max()
    assert np.abs(a["theta"] - th).max() <= 1e-3 * np.abs(th).max()
    assert float(a["train"]) == pytest.approx(float(hist[-1, 0]), rel=1e-4)
    assert float(a["test"]) == pytest.approx(float(hist[-1, 1]), rel=1e-4)


def test_local_share_is_the_column_range_of_every_row():
    from cumf_als_b200.data import nnz_balanced_ranges, synth_ratings
    from cumf_als_b200.dist import compact_share, local_share, row_batches
    r = synth_ratings(90, 400, 6000, 50, seed=9)
    for world in (1, 2, 3, 8):
        tr = nnz_balanced_ranges(r.csc_indptr, world)
        prev_end = r.csr_indptr[:-1].astype(np.int64)
        total = 0
        for t0, t1 in tr:
            b, e = local_share(r.csr_indptr, r.csr_indices, t0, t1)
            assert np.array_equal(b, prev_end)                 # shares tile each row left to right
            for u in (0, 17, 89):
                cols = r.csr_indices[b[u]:e[u]]
                assert ((cols >= t0) & (cols < t1)).all()
            ip, col, val = compact_share(r.csr_indptr, r.csr_indices, r.csr_data, t0, t1)
            assert ip[-1] == col.size == val.size == int((e - b).sum())
            total += col.size
            prev_end = e
        assert np.array_equal(prev_end, r.csr_indptr[1:].astype(np.int64)) and total == r.nnz
    assert row_batches(10, 4, 4 * 4 * 5 * 3) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert row_batches(0, 4, 1 << 20) == []


def test_partials_sum_to_the_full_system():
    """sum_g [A|b]_g (lambda * local count each) == the unsharded [A|b] (lambda * n_u), to fp32 summation order."""
    from cumf_als_b200.data import init_factors, nnz_balanced_ranges, synth_ratings
    from cumf_als_b200.dist import compact_share
    from oracle import oracle as O
    r = synth_ratings(60, 300, 5000, 50, seed=12)
    f, lam = 30, 0.05
    theta, _ = init_factors(r.m, r.n, f, seed=2)
    full_A = O.gram(r.csr_indptr, r.csr_indices, theta, f, lam)
    full_b = O.rhs(r.csr_indptr, r.csr_indices, r.csr_data, theta, f)
    for world in (2, 5):
        A, b = np.zeros_like(full_A), np.zeros_like(full_b)
        for t0, t1 in nnz_balanced_ranges(r.csc_indptr, world):
            ip, col, val = compact_share(r.csr_indptr, r.csr_indices, r.csr_data, t0, t1)
            A += O.gram(ip.astype(np.int32), col, theta, f, lam)
            b += O.rhs(ip.astype(np.int32), col, val, theta, f)
        assert np.abs(A - full_A).max() <= 2e-6 * np.abs(full_A).max()
        assert np.abs(b - full_b).max() <= 2e-6 * np.abs(full_b).max()
