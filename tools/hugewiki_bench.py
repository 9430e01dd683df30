"""BASELINE configs[4]: Hugewiki-scale synthetic (m = 50 082 603, n = 39 780, ~3.1 G ratings, f = 100, hugewiki.cu:27-42) on
the GPUs of one box, matrix generated shard by shard on the devices, rows sharded inside the library (one process).
    python tools/hugewiki_bench.py [n_gpus=8] [scale=1.0] [iters=3]
Prints one JSON line: iterations/s (device time of the slowest shard), setup seconds, RMSE."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402,F401

import cumf_als_b200 as c  # noqa: E402

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
M, N, NNZ = 50082603, 39780, 3101144313
m = int(M * scale)
t0 = time.perf_counter()
g = c.AlsGroup.from_synth(m, N, NNZ / M, 2026, test_per_shard=500000, f=100, lam=0.048, n_devices=n_gpus)
setup = time.perf_counter() - t0
g.collect_train_sse(True)
warm = g.iterate(1)
ms = g.iterate(iters)
tr, te = g.rmse()
print(json.dumps({"metric": "ALS iterations/sec (Hugewiki-scale synthetic, f=100, CG solver; BASELINE configs[4])",
                  "value": iters / (ms / 1e3), "unit": "iterations/s", "n_gpus": n_gpus, "steps": iters, "ms_per_step": ms / iters,
                  "config": {"workload": "hugewiki-scale synthetic", "m": m, "n": N, "nnz": g.nnz, "f": 100, "lambda": 0.048,
                             "sharding": f"rows split evenly over {n_gpus} GPUs, matrix generated on the devices (csrc/synth.cu), "
                                         f"int64 pointers, peer stores + device barrier"},
                  "setup_s": setup, "first_iteration_ms": warm, "train_rmse": tr, "test_rmse": te}))
g.close()
