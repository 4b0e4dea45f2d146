"""ctypes binding of libmpc_b200.so (the C ABI declared in include/mpc_b200.h).

There is no CPU execution path: every wrapper here needs the built library and CUDA tensors.
If the library is missing the import of any compute entry point raises ``MpcLibraryMissing``
with the build command -- nothing falls back to numpy / torch.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_void_p, POINTER

PKG = os.path.dirname(os.path.abspath(__file__))
# MPC_B200_LIB: load another build of the same library (ablation builds under lib/variants, tools/prof)
LIB_PATH = os.environ.get("MPC_B200_LIB") or os.path.join(PKG, "lib", "libmpc_b200.so")

MPC_F64, MPC_F32 = 0, 1
MPC_SOLVED, MPC_MAX_ITER, MPC_INFEASIBLE, MPC_UNSOLVED = 1, 2, 3, 0
MAX_NX, MAX_NU = 32, 16


class MpcError(RuntimeError):
    """Non-zero return of a libmpc_b200 entry point (negative: argument error, positive: CUDA)."""

    def __init__(self, code, msg):
        super().__init__(f"libmpc_b200 error {code}: {msg}")
        self.code = code


class MpcLibraryMissing(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); kept in one table so that tests can check it against the header.
SIGNATURES = {
    "mpc_version": (c_int, []),
    "mpc_last_error": (c_char_p, []),
    "mpc_riccati": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                            c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int,
                            c_int, c_void_p]),
    "mpc_lq_rollout": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int,
                               c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_double, c_int64, c_int, c_int, c_int, c_int,
                               c_void_p]),
    "mpc_linear_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                c_int, c_void_p]),
    "mpc_lq_solve": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                             c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p]),
    "mpc_lq_solve_variant": (c_int, [c_int, c_int, c_int, c_int]),
    "mpc_fma_peak_probe": (c_int, [c_int, POINTER(c_double)]),
}


def register(name, restype, argtypes):
    """Used by sibling modules that bind further entry points."""
    SIGNATURES[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype, fn.argtypes = restype, argtypes


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MpcLibraryMissing(
                f"{LIB_PATH} not built; run `python -m model_predictive_control_b200._build` "
                "(needs nvcc).  There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            try:
                fn = getattr(L, name)
            except AttributeError:
                if os.environ.get("MPC_B200_LIB"):   # an older ablation build may lack newer entry points
                    continue
                raise
            fn.restype, fn.argtypes = restype, argtypes
        _lib = L
    return _lib


def check(code):
    if code != 0:
        raise MpcError(code, lib().mpc_last_error().decode(errors="replace"))


def dtype_enum(t):
    import torch
    if t.dtype == torch.float64:
        return MPC_F64
    if t.dtype == torch.float32:
        return MPC_F32
    raise ValueError(f"unsupported dtype {t.dtype}: libmpc_b200 computes in float64 or float32")


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def stream(device=None):
    import torch
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ValueError("libmpc_b200 operates on CUDA tensors only (no CPU path)")
