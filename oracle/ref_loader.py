"""Import the reference's own session-1 modules (oracle, test infra).

Works where /root/reference is mounted (this container) or where ``stage_reference()`` has staged
the three session-1 files into the git-ignored ``oracle/_ref/session_1`` (the staging runs in
``__graft_entry__.build()`` here; ``oracle/_ref`` is not gpurun-ignored, so the staged files travel to the
GPU box and ``bench.py --impl reference`` times the reference's OWN code there).
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
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_SESSION1_FILES = ("FHC.py", "LinearSystem.py", "session1_sol.py")
_STUBS = ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "casadi", "rcracers")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "session_1", "FHC.py"))


def staged() -> bool:
    return all(os.path.isfile(os.path.join(STAGED_ROOT, "session_1", f)) for f in _SESSION1_FILES)


def stage_reference() -> bool:
    """Stage the reference's session-1 files, byte for byte, into oracle/_ref/session_1 (git-ignored: never part of
    the history; shipped to the GPU box by gpurun).  No-op without /root/reference.  Returns whether a staged copy
    exists afterwards."""
    if available():
        import shutil
        dst = os.path.join(STAGED_ROOT, "session_1")
        os.makedirs(dst, exist_ok=True)
        for f in _SESSION1_FILES:
            shutil.copyfile(os.path.join(REFERENCE_ROOT, "session_1", f), os.path.join(dst, f))
    return staged()


def session1_root():
    """Directory the reference's session-1 modules are imported from: the mounted reference, else the staged copy."""
    if available():
        return REFERENCE_ROOT
    if staged():
        return STAGED_ROOT
    return None


def load_session1():
    """Return (FHC, LinearSystem, session1_sol) modules of the unmodified reference."""
    root = session1_root()
    if root is None:
        raise RuntimeError("reference neither mounted at " + REFERENCE_ROOT + " nor staged in " + STAGED_ROOT)
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    path = os.path.join(root, "session_1")
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
        if not mod.__file__.startswith(root):
            raise RuntimeError(f"{mod.__name__} resolved to {mod.__file__}, not the reference")
    return mods


# ------------------------------------------------------------------------------------------------
# sessions 2-4: the reference's DATA and OCP construction, run unmodified under numeric stand-ins
# ------------------------------------------------------------------------------------------------
def _permissive_dataclass(cls=None, **_kw):
    """Stand-in for dataclasses.dataclass while importing session_{2,3}/problem.py: the stdlib decorator
    rejects the reference's ``Q: np.ndarray = np.diag([10, 1])`` defaults on Python >= 3.11
    (ValueError: mutable default).  Same observable behaviour for these classes: keyword constructor over the
    annotated fields with the class-level defaults, then ``__post_init__``."""
    def wrap(c):
        ann = dict(c.__dict__.get("__annotations__", {}))
        defaults = {k: c.__dict__[k] for k in ann if k in c.__dict__}

        def __init__(self, **kw):
            unknown = set(kw) - set(ann)
            if unknown:
                raise TypeError(f"unexpected fields {sorted(unknown)}")
            for k in ann:
                setattr(self, k, kw[k] if k in kw else defaults.get(k))
            if hasattr(self, "__post_init__"):
                self.__post_init__()
        c.__init__ = __init__
        return c
    return wrap if cls is None else wrap(cls)


def _import_file(path, name, extra_path=None):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    before = list(sys.path)
    if extra_path:
        sys.path.insert(0, extra_path)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(mod)
    finally:
        sys.path[:] = before
    return mod


def load_problem(session: int):
    """The ``Problem`` class of /root/reference/session_{2,3}/problem.py (source unmodified; only the stdlib
    ``dataclass`` decorator is replaced while the module body runs, see _permissive_dataclass)."""
    if not available():
        raise RuntimeError("reference not mounted at " + REFERENCE_ROOT)
    import dataclasses
    path = os.path.join(REFERENCE_ROOT, f"session_{session}", "problem.py")
    with mock.patch.object(dataclasses, "dataclass", _permissive_dataclass):
        mod = _import_file(path, f"_ref_problem_s{session}")
    return mod.Problem


def load_parameters():
    """``VehicleParameters`` of /root/reference/session_4/parameters.py (imports as is)."""
    if not available():
        raise RuntimeError("reference not mounted at " + REFERENCE_ROOT)
    return _import_file(os.path.join(REFERENCE_ROOT, "session_4", "parameters.py"), "_ref_parameters").VehicleParameters


class NumericCasadi:
    """Stand-in for the ``casadi`` module that makes the reference's OCP construction
    (session4_sol.py:131-217, main.py:41-113) evaluate NUMERICALLY: ``SX.sym(name, shape)`` returns the numpy value
    registered for ``name`` in ``values``, so after ``MPCController(...)`` the nlp dict holds the reference's own
    cost f(U; x0), its constraint vector g(U; x0) and its bound vectors for that (x0, U) -- computed by the
    reference's code (weights, shooting order, integrator, collision geometry), not by a restatement."""

    def __init__(self):
        import numpy as np
        self.values = {}
        self._np = np
        outer = self

        class SX:
            @staticmethod
            def sym(name, shape=(1, 1)):
                shape = shape if isinstance(shape, tuple) else (shape, 1)
                return np.array(outer.values[name], dtype=float).reshape(shape)

            @staticmethod
            def zeros(r, c=1):
                return np.zeros((r, c))
        self.SX = SX
        self.cos, self.sin = np.cos, np.sin

    def diagcat(self, *v):
        return self._np.diag(self._np.array(v, dtype=float))

    def vertcat(self, *xs):
        np = self._np
        if not xs:
            return np.zeros((0, 1))
        return np.concatenate([np.asarray(x, dtype=float).reshape(-1, 1) for x in xs])

    def norm_2(self, v):
        return self._np.linalg.norm(self._np.asarray(v, dtype=float))

    def nlpsol(self, name, solver, nlp, opts=None):
        def _no_solver(**_kw):
            raise NotImplementedError("IPOPT is not available; only the OCP construction is evaluated")
        _no_solver.nlp = nlp
        return _no_solver


def load_session4(which: str = "session4_sol"):
    """(module, casadi_stub) for /root/reference/session_4/{session4_sol,main}.py, source unmodified.
    casadi -> NumericCasadi; rcracers (absent, unpinned) -> ``KinematicBicycle`` = the bicycle ODE of
    oracle/bicycle.py (OUR definition) and a plain ``simulate`` loop; matplotlib/animation/plotting ->
    mocks.  Everything else -- integrators, weights, bounds, shooting order, collision geometry -- is the
    reference's own code."""
    if not available():
        raise RuntimeError("reference not mounted at " + REFERENCE_ROOT)
    import types
    import numpy as np
    from . import bicycle as obc

    cs = NumericCasadi()

    class KinematicBicycle:
        def __init__(self, params, symbolic=False):
            self.params = params

        def __call__(self, x, u):
            p = self.params
            xa = np.asarray(x, dtype=float)
            f = obc.bicycle_f(xa.reshape(-1), np.asarray(u, dtype=float).reshape(-1), p.axis_rear, p.axis_front,
                              p.friction, p.acceleration)
            return f.reshape(xa.shape)

    def simulate(x0, dynamics, n_steps, policy=None):
        xs = [np.asarray(x0, dtype=float)]
        for t in range(n_steps):
            xs.append(np.asarray(dynamics(xs[-1], policy(xs[-1], t))))
        return np.array(xs)

    rc = types.ModuleType("rcracers"); rcs = types.ModuleType("rcracers.simulator")
    rcd = types.ModuleType("rcracers.simulator.dynamics")
    rcd.KinematicBicycle = KinematicBicycle; rcs.simulate = simulate; rcs.dynamics = rcd; rc.simulator = rcs
    plotting = types.ModuleType("plotting")
    plotting.plot_state_trajectory = plotting.plot_input_sequence = mock.MagicMock()
    animation = types.ModuleType("animation"); animation.AnimateParking = mock.MagicMock()
    stubs = {"casadi": cs, "rcracers": rc, "rcracers.simulator": rcs, "rcracers.simulator.dynamics": rcd,
             "plotting": plotting, "animation": animation, "matplotlib": mock.MagicMock(),
             "matplotlib.pyplot": mock.MagicMock(), "matplotlib.patches": mock.MagicMock()}
    path = os.path.join(REFERENCE_ROOT, "session_4", which + ".py")
    with mock.patch.dict(sys.modules, stubs):
        sys.modules.pop("parameters", None)
        mod = _import_file(path, "_ref_s4_" + which, extra_path=os.path.join(REFERENCE_ROOT, "session_4"))
        sys.modules.pop("parameters", None)
    return mod, cs


def load_log(session: int = 2):
    """``ControllerLog`` of /root/reference/session_{2,3}/log.py (source unmodified; rcracers' absent
    ``BaseControllerLog`` is replaced by an empty dataclass base)."""
    if not available():
        raise RuntimeError("reference not mounted at " + REFERENCE_ROOT)
    import dataclasses
    import types
    core = types.ModuleType("rcracers.simulator.core")
    core.BaseControllerLog = dataclasses.make_dataclass("BaseControllerLog", [])
    stubs = {"rcracers": types.ModuleType("rcracers"), "rcracers.simulator": types.ModuleType("rcracers.simulator"),
             "rcracers.simulator.core": core}
    with mock.patch.dict(sys.modules, stubs):
        mod = _import_file(os.path.join(REFERENCE_ROOT, f"session_{session}", "log.py"), f"_ref_log_s{session}")
    return mod.ControllerLog
