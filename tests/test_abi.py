"""CPU-side checks of the drop-in boundary: the library loads and exports what the header declares."""
import ctypes
import os
import re

import pytest

import tiger_hlm_gpu_b200 as hlm
from tiger_hlm_gpu_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "hlm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hlm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_symbol_the_header_declares():
    names = header_functions()
    assert len(names) >= 25
    lib = ctypes.CDLL(hlm.lib_path())
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/hlm_b200.h but not exported: {missing}"


def test_binding_covers_the_header():
    assert sorted(api.SIGNATURES) == header_functions()


def test_abi_version_and_model_registry():
    lib = hlm.load_library()
    assert lib.hlm_abi_version() == hlm.ABI_VERSION
    assert hlm.model_info(204) == (5, 15, 2)   # Model204::N_EQ = 5 (models/model_204.hpp:19)
    assert hlm.model_info(0) == (5, 0, 0)
    assert hlm.model_info(200) == (5, 15, 2)   # project-defined hillslope-link model (README.md:95 names it only)
    with pytest.raises(hlm.HlmError, match="unknown model uid"):
        hlm.model_info(190)


def test_reference_trait_constants():
    assert hlm.Model204.UID == 204 and hlm.Model204.N_EQ == 5
    p = hlm.Parameters()
    assert (p.initialStep, p.rtol, p.atol, p.safety, p.minScale, p.maxScale) == (0.01, 1e-6, 1e-9, 0.9, 0.2, 10.0)
    assert hlm.SPATIAL_PARAMS_DTYPE.itemsize == 136


def test_no_cpu_fallback():
    """Without a usable B200 the compute entry points must fail loudly, never compute on the host."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present: the loud-failure path cannot be exercised")
    with pytest.raises(hlm.HlmError, match="hlm status -2"):
        hlm.Solver(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tiger_hlm_gpu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("oracle/oracle_rk45.c:rhs_dummy", ""), f"{f} mentions the oracle"
