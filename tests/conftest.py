import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def golden(name: str):
    """Load a golden fixture produced by tests/golden/make_golden.py from the real reference."""
    path = GOLDEN / name
    if not path.exists():
        pytest.skip(f"golden fixture {name} not generated yet (tests/golden/make_golden.py on a B200)")
    return dict(np.load(path))


def rel_fro(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope="session")
def lib():
    import cumf_als_b200 as c
    return c.load_library()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch
