"""Drop-in for the reference's ``session_1/session1_sol.py`` numeric functions.

``riccati_recursion(A, B, R, Q, Pf, N)`` (R before Q, reference session1_sol.py:44-65) and the
generic loop ``simulate(x0, f, policy, steps)`` (:68-91: t = 0..steps-1, returns ``steps+1``
states and the ``|x| > 100`` instability flag).  When ``f`` comes from :func:`linear_dynamics`
and ``policy`` from :func:`feedback_policy` the whole loop is one CUDA kernel; any other
callables run through the same per-step loop as in the reference.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np
import torch

from . import _interop as io
from . import lq
from .FHC import _recursion, get_dynamics_continuous, get_dynamics_discrete  # noqa: F401


def riccati_recursion(A, B, R, Q, Pf, N: int):
    """Same recursion as FHC.ricatti_recursion, instructor-solution argument order."""
    return _recursion(A, B, Q, R, Pf, N)


class linear_dynamics:
    """f(x, u) = A x + B u as a callable that ``simulate`` can fuse (reference :159-160)."""

    def __init__(self, A, B):
        self.A, self.B = A, B

    def __call__(self, x, u):
        return self.A @ x + self.B @ u


class feedback_policy:
    """policy(x, t) = gains[0] @ x (receding horizon, reference :107-109) or gains[t] @ x
    (open-loop prediction, reference :121-123)."""

    def __init__(self, gains, receding=True):
        self.gains, self.receding = gains, bool(receding)

    def __call__(self, x, t):
        return self.gains[0 if self.receding else t] @ x


def simulate(x0, f: Callable, policy: Callable, steps: int) -> Tuple[np.ndarray, bool]:
    """States (steps+1, n) -- or (steps+1, batch, n) for batched x0 (batch, n) -- and the flag."""
    steps = int(steps)
    if isinstance(f, linear_dynamics) and isinstance(policy, feedback_policy):
        as_np = not io.is_tensor(x0)
        dt = io.pick_dtype(x0, f.A)
        xd = io.to_dev(x0, dt)
        single = xd.dim() == 1
        x0cols = xd[:, None] if single else xd.t()
        K = torch.stack([io.to_dev(g, dt) for g in policy.gains], dim=0)
        res = lq.lq_rollout(io.to_dev(f.A, dt), io.to_dev(f.B, dt), K, x0cols, steps + 1,
                            gain_offset=0, gain_step=0 if policy.receding else 1,
                            want_unstable=True, norm_limit=100.0)
        X = res["X"]  # [steps+1, n, batch]
        X = X[:, :, 0] if single else X.permute(0, 2, 1)
        flag = res["unstable"].bool()
        flag_out = bool(flag[0].item()) if single else io.back(flag, as_np)
        return io.back(X, as_np), flag_out
    # generic callables: the reference loop itself (session1_sol.py:79-91)
    instability_occured = False
    x = [x0]
    for t in range(steps):
        xt = x[-1]
        xnext = f(xt, policy(xt, t))
        x.append(xnext)
        nrm = torch.linalg.norm(xnext) if io.is_tensor(xnext) else np.linalg.norm(xnext)
        if nrm > 100 and not instability_occured:
            instability_occured = True
    if io.is_tensor(x0):
        return torch.stack(x), instability_occured
    return np.array(x), instability_occured


def is_stable(A, B, gains) -> bool:
    """The exact test the reference's ``plot_ex4`` leaves as an exercise (session1_sol.py:114-116): the receding-horizon
    closed loop x+ = (A + B gains[0]) x is stable iff its spectral radius is below one."""
    from .FHC import closed_loop_spectral_radius
    K0 = gains[0] if isinstance(gains, (list, tuple)) else gains
    rho = closed_loop_spectral_radius(A, B, K0)
    return bool(rho < 1.0) if np.ndim(rho) == 0 else rho < 1.0


def setup():
    """Problem data of the exercise (reference session1_sol.py:136-144)."""
    ts = 0.5
    C = np.array([[1, -2.0 / 3]])
    Q = C.T @ C + 1e-3 * np.eye(2)
    R = np.array([[0.1]])
    A, B = get_dynamics_discrete(ts)
    return A, B, Q, R
