"""The C-ABI boundary: libmpc_b200.so loads and exports every symbol include/mpc_b200.h declares,
the Python binding table covers the header, and the product refuses to run without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mpc_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpc_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from model_predictive_control_b200 import _build, _lib
    if not os.path.exists(_lib.LIB_PATH):
        _build.build_library()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_declares_the_path():
    fns = header_functions()
    for required in ("mpc_riccati", "mpc_lq_rollout", "mpc_lq_solve", "mpc_last_error", "mpc_version"):
        assert required in fns


def test_library_exports_every_declared_symbol(built_lib):
    for name in header_functions():
        assert hasattr(built_lib, name), f"{name} declared in include/mpc_b200.h but not exported"


def test_binding_table_covers_header():
    from model_predictive_control_b200 import _lib
    import importlib
    import pkgutil
    import model_predictive_control_b200 as pkg
    for mod in pkgutil.iter_modules(pkg.__path__):  # modules register their entry points on import
        importlib.import_module(f"model_predictive_control_b200.{mod.name}")
    missing = set(header_functions()) - set(_lib.SIGNATURES)
    assert not missing, f"no ctypes signature for {sorted(missing)}"


def test_version_and_error_string(built_lib):
    built_lib.mpc_version.restype = ctypes.c_int
    built_lib.mpc_last_error.restype = ctypes.c_char_p
    assert built_lib.mpc_version() == 100
    assert isinstance(built_lib.mpc_last_error(), bytes)


def test_argument_errors_need_no_gpu(built_lib):
    """Invalid arguments are rejected before any CUDA call (negative mpc_error codes)."""
    from model_predictive_control_b200 import _lib
    L = _lib.lib()
    assert L.mpc_riccati(None, 0, None, 0, None, 0, None, 0, None, 0, None, None, 0, 4, 2, 1, 5, 0, None) == -1
    assert b"null" in L.mpc_last_error()
    one = ctypes.c_void_p(64)
    assert L.mpc_riccati(one, 0, one, 0, one, 0, one, 0, one, 0, one, one, 0, 4, 99, 1, 5, 0, None) == -2
    assert L.mpc_riccati(one, 0, one, 0, one, 0, one, 0, one, 0, one, one, 0, 4, 2, 1, 5, 7, None) == -3
    odd = ctypes.c_void_p(66)
    assert L.mpc_riccati(odd, 0, one, 0, one, 0, one, 0, one, 0, one, one, 0, 4, 2, 1, 5, 0, None) == -4
    assert L.mpc_lq_solve(one, 0, one, 0, one, 0, one, 0, one, 0, one, one, one, one, None, None, 4, 12, 4, 5, 0,
                          None) == -5
    # empty batch is a no-op
    assert L.mpc_riccati(one, 0, one, 0, one, 0, one, 0, one, 0, one, one, 0, 0, 2, 1, 5, 0, None) == 0


def test_no_cpu_fallback():
    """numpy inputs without a CUDA device raise; nothing is computed on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from model_predictive_control_b200 import FHC
    A, B = FHC.get_dynamics_discrete(0.5)
    with pytest.raises(RuntimeError, match="CUDA"):
        FHC.ricatti_recursion(A, B, np.eye(2), np.array([0.1]), np.eye(2), 3)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "model_predictive_control_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
