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


def test_workspace_sizes_are_whole_tiles(built_lib):
    """The box-QP / RTI workspaces are laid out in tiles of 32 lanes (csrc/boxqp_core.cuh): their size depends on the
    batch only through the number of tiles, grows linearly in tiles and in the horizon, and the float32 product needs
    no more than the float64 one (no GPU needed: pure host arithmetic of the library)."""
    I64 = ctypes.c_int64
    f = built_lib.mpc_boxqp_workspace_bytes
    f.restype = I64
    f.argtypes = [I64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    g = built_lib.mpc_boxqp_rows_workspace_bytes
    g.restype = I64
    g.argtypes = [I64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    r = built_lib.mpc_rti_workspace_bytes
    r.restype = I64
    r.argtypes = [I64, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    header = 256
    from model_predictive_control_b200 import _lib
    F64, F32 = _lib.MPC_F64, _lib.MPC_F32
    for n, m in ((2, 1), (4, 1), (4, 2)):
        for N in (5, 30):
            one = f(1, n, m, N, F64) - header
            assert one > 0 and f(32, n, m, N, F64) - header == one and f(33, n, m, N, F64) - header == 2 * one
            assert f(64 * 1000, n, m, N, F64) - header == 2000 * one
            assert f(32, n, m, 2 * N, F64) - header == 2 * one
            assert f(32, n, m, N, F32) <= f(32, n, m, N, F64)
            # all-float64 layout: 9 (n+m) + m*n + m*m + m doubles per stage and lane (z, 4 slack/multiplier rows, dz_aff,
            # dz, e, g; gains K, S^-1, d)
            d = n + m
            assert one == N * 32 * 8 * (9 * d + m * n + m * m + m)
    assert g(32, 4, 2, 10, 9, F64) - header == 10 * 32 * 8 * (9 * 6 + 8 + 4 + 2 + 3 * 9)
    assert g(1, 4, 2, 10, 9, F64) == g(32, 4, 2, 10, 9, F64) < g(33, 4, 2, 10, 9, F64)
    assert r(64, 20, 0, F64) > r(32, 20, 0, F64) > 0 and r(32, 20, 9, F64) > r(32, 20, 0, F64)
    assert f(0, 2, 1, 30, F64) == header   # empty batch: the header only
