"""tiger_hlm_gpu_b200 — B200-native batched RK45 for hillslope-link runoff ODEs.

Python mirror of the reference's operator surface for this path (the product itself is the C-ABI
library ``libhlm_b200.so`` declared in ``include/hlm_b200.h`` and the header-only C++ shims in
``include/hlm_b200/rk45_api.hpp``).  There is no CPU fallback: importing works anywhere (so the
build can be checked on a box without a GPU), but every compute entry point raises if the CUDA
library is missing or no sm_100-class device is present.
"""
from .api import (  # noqa: F401
    ABI_VERSION,
    DummyModel,
    HlmError,
    Model200,
    Model204,
    Parameters,
    SPATIAL_PARAMS_DTYPE,
    Solver,
    lib_path,
    load_library,
    model_info,
    run_rk45,
    setModelParameters,
)

__all__ = [
    "ABI_VERSION", "DummyModel", "HlmError", "Model200", "Model204", "Parameters", "SPATIAL_PARAMS_DTYPE", "Solver",
    "lib_path", "load_library", "model_info", "run_rk45", "setModelParameters",
]
