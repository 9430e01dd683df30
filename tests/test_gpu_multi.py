"""Multi-GPU row sharding inside the library (include/cumf_als.h, "multi-GPU row sharding"), exercised on ONE GPU: with
CUMF_GROUP_SAME_DEVICE=1 every shard of a cumf_als_group lives on device 0 with its own stream, so the peer stores of the
solver epilogues, the per-shard by-product train RMSE and doALS under CUMF_GPUS are all run for real; on n GPUs the addresses
differ and the half-steps are ordered by the device-side flag barrier instead of events (als_api.cu, SameDeviceSync).  Sharded == unsharded bit for bit (every row's arithmetic is independent of
the partition).  The two-process CUDA IPC path needs two GPUs: tools/multi_gpu_check.py.  Run with -m gpu."""
import numpy as np
import pytest

import cumf_als_b200 as c
from cumf_als_b200.data import init_factors, synth_ratings

pytestmark = pytest.mark.gpu


def _solver_args(r, f, lam):
    return (r.csr_indptr, r.csr_indices, r.csr_data, r.csc_indices, r.csc_indptr, r.csc_data, r.coo_row, r.test_row, r.test_col,
            r.test_val, r.m, r.n, f, lam)


@pytest.mark.parametrize("f,impl,shards", [(100, "1", 2), (100, "2", 3), (10, "2", 4), (200, "2", 2)])
def test_group_on_one_device_equals_single_solver(cuda, monkeypatch, f, impl, shards):
    monkeypatch.setenv("CUMF_GROUP_SAME_DEVICE", "1")
    monkeypatch.setenv("CUMF_TC_IMPL", impl)
    monkeypatch.setenv("CUMF_SPLIT_NNZ", "700")            # a few rows split across CTAs: their CG tail pushes to the peers too
    r = synth_ratings(600, 3000, 150000, 7000, seed=40 + f)
    theta0, X0 = init_factors(r.m, r.n, f, seed=9)
    lam, iters = 0.048, 3
    s = c.AlsSolver(*_solver_args(r, f, lam))
    s.set_factors(theta0, X0)
    s.collect_train_sse(True)
    want = []
    for _ in range(iters):
        s.iterate(1)
        want.append(s.rmse())
    th_w, X_w = s.get_factors()
    s.close()
    g = c.AlsGroup(*_solver_args(r, f, lam), n_devices=shards)
    g.set_factors(theta0, X0)
    assert g.collect_train_sse(True)
    got = []
    for _ in range(iters):
        g.iterate(1)
        got.append(g.rmse())
    th_g, X_g = g.get_factors()
    g.close()
    assert np.array_equal(X_g, X_w) and np.array_equal(th_g, th_w)
    want, got = np.array(want), np.array(got)
    print(f"f={f} impl={impl} shards={shards}: rmse single {want[-1]} group {got[-1]}")
    assert np.abs(got[:, 1] - want[:, 1]).max() < 1e-6 * want[:, 1].max()          # test RMSE: same terms, other summation order
    assert np.abs(got[:, 0] - want[:, 0]).max() < 1e-4 * want[:, 0].max()          # train RMSE: per-shard by-products


def test_doals_cumf_gpus_equals_single(cuda):
    """doALS(host pointers) under CUMF_GPUS=2: what the reference's main.cpp gets by exporting one variable.  Run in a child
    process: doALS ends the process on an error like the reference's cudacall macro (als.h:628-640)."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    p = subprocess.run([sys.executable, str(root / "tools" / "multi_gpu_check.py"), "2", "same"], capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, CUMF_QUIET="0"))
    print(p.stdout[-3000:], p.stderr[-3000:])
    assert p.returncode == 0
    assert "[doALS] CUMF_GPUS=2 vs 1: factors equal True" in p.stdout
    assert p.stdout.count("factors equal True") == 3          # group with both kernels, then doALS


@pytest.mark.parametrize("shards", [1, 2, 3])
def test_theta_half_step_keeps_the_x_side_gather_table_current(cuda, monkeypatch, shards):
    """f = 100, long X rows (round-1 kernel, 512-byte [hi | lo'] gather table of theta) and short theta rows (generic kernel): the
    theta half-step's solver epilogues write the split form of every row they solve into that table -- on every shard -- and the
    next X half-step skips its split pass over all of theta.  Same table bits, so the factors equal those of
    CUMF_FUSED_SPLIT=0 (every X half-step re-splits) bit for bit, with one launch less per iteration and shard."""
    monkeypatch.setenv("CUMF_GROUP_SAME_DEVICE", "1")
    monkeypatch.delenv("CUMF_TC_IMPL", raising=False)
    monkeypatch.setenv("CUMF_SPLIT_NNZ", "4000")           # some X rows split across CTAs, a few theta rows too (CG tail writes split rows)
    f, lam, iters = 100, 0.048, 3
    r = synth_ratings(900, 14000, 1400000, 20000, seed=123)
    theta0, X0 = init_factors(r.m, r.n, f, seed=5)

    def run(fused):
        monkeypatch.setenv("CUMF_FUSED_SPLIT", "1" if fused else "0")
        if shards == 1:
            s = c.AlsSolver(*_solver_args(r, f, lam))
        else:
            s = c.AlsGroup(*_solver_args(r, f, lam), n_devices=shards)
        s.set_factors(theta0, X0)
        s.iterate(iters)
        out = s.get_factors()
        n_launch = int(s.timers()["launches"]) if shards == 1 else None
        s.close()
        return out, n_launch

    (th_a, X_a), la = run(False)
    (th_b, X_b), lb = run(True)
    assert np.array_equal(th_a, th_b) and np.array_equal(X_a, X_b)
    if shards == 1:
        print(f"launches: re-split every X half-step {la}, fused {lb}")
        assert lb == la - (iters - 1)


def test_raw_theta_pointer_disables_the_table_shortcut(cuda, monkeypatch):
    """A caller that took cumf_als_theta_ptr may rewrite theta behind the solver's back: after that the X side re-splits."""
    monkeypatch.delenv("CUMF_TC_IMPL", raising=False)
    f, lam = 100, 0.048
    r = synth_ratings(900, 14000, 1400000, 20000, seed=124)
    theta0, X0 = init_factors(r.m, r.n, f, seed=6)
    s = c.AlsSolver(*_solver_args(r, f, lam))
    s.set_factors(theta0, X0)
    launches = lambda: int(s.timers()["launches"])
    s.iterate(1)
    l0 = launches()
    s.iterate(1)
    per_iter_fused = launches() - l0
    assert s.theta_ptr != 0
    l1 = launches()
    s.iterate(1)
    assert launches() - l1 == per_iter_fused + 1
    s.close()
