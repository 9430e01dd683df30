"""cumf_als_b200 -- B200-native ALS factor-update path behind cuMF's doALS interface.

The product is the C-ABI shared library `libcumf_als_b200.so` (hand-written sm_100a
CUDA, see include/cumf_als.h); this package is the thin host-side mirror of the
reference's interface for that path (doALS, the .bin loaders, the stage seams).
There is no CPU or PyTorch fallback: importing the bindings without the built
library, or calling them without a B200, raises.
"""
from .api import (  # noqa: F401
    AlsGroup,
    SynthShard,
    AlsSolver,
    CumfError,
    PATH_AUTO,
    PATH_SIMT,
    PATH_TC,
    SOLVER_CG,
    SOLVER_LU,
    cg,
    do_als,
    gram,
    library_path,
    load_library,
    lu,
    rmse,
    update_factor,
    Plan,
)

__all__ = [
    "AlsGroup", "SynthShard", "AlsSolver", "CumfError", "Plan", "PATH_AUTO", "PATH_SIMT", "PATH_TC", "SOLVER_CG", "SOLVER_LU",
    "cg", "do_als", "gram", "library_path", "load_library", "lu", "rmse", "update_factor",
]
