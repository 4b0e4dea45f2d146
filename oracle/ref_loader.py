"""Import the reference's own session-1 modules (oracle, test infra).

Only works where /root/reference is mounted (this container; never the GPU box).
``FHC.py`` imports casadi, rcracers and matplotlib at module top
(/root/reference/session_1/FHC.py:1-17) and ``session1_sol.py`` exits without
matplotlib (:4-8); none of them is used by the numeric functions, so they are
replaced by MagicMock modules before the import.
"""
from __future__ import annotations

import importlib
import os
import sys
from unittest import mock

REFERENCE_ROOT = "/root/reference"
_STUBS = ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "casadi", "rcracers")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "session_1", "FHC.py"))


def load_session1():
    """Return (FHC, LinearSystem, session1_sol) modules of the unmodified reference."""
    if not available():
        raise RuntimeError("reference not mounted at " + REFERENCE_ROOT)
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    path = os.path.join(REFERENCE_ROOT, "session_1")
    # the reference's module names (FHC, LinearSystem) collide with nothing of ours:
    # the product lives in the model_predictive_control_b200 package.
    sys.path.insert(0, path)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mods = tuple(importlib.import_module(n) for n in ("FHC", "LinearSystem", "session1_sol"))
    finally:
        sys.path.remove(path)
    for mod in mods:
        if not mod.__file__.startswith(REFERENCE_ROOT):
            raise RuntimeError(f"{mod.__name__} resolved to {mod.__file__}, not the reference")
    return mods
