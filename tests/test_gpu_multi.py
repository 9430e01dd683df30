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
