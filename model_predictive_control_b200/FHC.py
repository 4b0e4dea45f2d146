"""Drop-in for the reference's ``session_1/FHC.py``: finite-horizon LQ controller.

``ricatti_recursion`` (sic, one "c") keeps the reference signature and conventions
(/root/reference/session_1/FHC.py:51-61): argument order (A, B, Q, R, P_f, N), gains include the
minus sign (u = +K x), both lists are returned first-stage-first, P is not symmetrised, and R
may be the 1-D array of FHC.py:141.  The recursion runs on the GPU (K1, ``mpc_riccati``); any
of the matrices may carry a leading batch dimension, in which case every returned entry does too.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import _interop as io
from . import lq
from .LinearSystem import LinearSystem


def _prep_R(R, m, dt):
    Rd = io.to_dev(R, dt)
    if Rd.dim() == 0:
        Rd = Rd.reshape(1)
    if Rd.dim() == 1:
        # numpy broadcasting of `R + B'PB` with a 1-D R (reference FHC.py:56 with R of FHC.py:141)
        if Rd.shape[0] not in (1, m):
            raise ValueError(f"operands could not be broadcast together with shapes ({Rd.shape[0]},) ({m},{m})")
        Rd = Rd.expand(m, m).contiguous() if Rd.shape[0] == m else Rd.expand(m, m).contiguous()
    return Rd


def _recursion(A, B, Q, R, P_f, N):
    as_np = not io.any_tensor(A, B, Q, R, P_f)
    dt = io.pick_dtype(A, B, Q, R, P_f)
    Ad, Bd, Qd, Pd = (io.to_dev(M, dt) for M in (A, B, Q, P_f))
    if Bd.dim() < 2:
        raise ValueError("B must be (n, m)")
    m = Bd.shape[-1]
    Rd = _prep_R(R, m, dt)
    batched = any(M.dim() == 3 for M in (Ad, Bd, Qd, Rd, Pd))
    K, P = lq.riccati(Ad, Bd, Qd, Rd, Pd, int(N), all_P=True)
    if not batched:
        K, P = K[:, 0], P[:, 0]
    if as_np:
        Kn, Pn = K.cpu().numpy(), P.cpu().numpy()
        return [Pn[k] for k in range(Pn.shape[0])], [Kn[k] for k in range(Kn.shape[0])]
    return list(P.unbind(0)), list(K.unbind(0))


def ricatti_recursion(A, B, Q, R, P_f, N: int):
    """(P, K): P[0] = cost-to-go at stage 0, ..., P[N] = P_f;  K[0] = first-stage gain."""
    return _recursion(A, B, Q, R, P_f, N)


class AutoCruising(LinearSystem):
    """Policy holder of the reference (FHC.py:20-29)."""

    def set_opti_gain(self, gains) -> None:
        self.gains = gains
        self._gain_cache = {}

    def _gains_tensor(self, dt):
        cache = self.__dict__.setdefault("_gain_cache", {})
        if dt not in cache:
            cache[dt] = torch.stack([io.to_dev(g, dt) for g in self.gains], dim=0)
        return cache[dt]

    def control_law(self, x, t):
        return self.gains[0] @ x

    def pred(self, x, t):
        return self.gains[t] @ x


AutoCruising.control_law._mpc_gain_policy = "receding"
AutoCruising.pred._mpc_gain_policy = "time_varying"


def get_dynamics_continuous() -> Tuple[np.ndarray]:
    """Double integrator with input -a (reference FHC.py:32-41)."""
    return np.array([[0.0, 1.0], [0.0, 0.0]]), np.array([[0], [-1]])


def get_dynamics_discrete(ts: float) -> Tuple[np.ndarray]:
    """Forward-Euler discretisation (reference FHC.py:44-48)."""
    A, B = get_dynamics_continuous()
    return np.eye(2) + A * ts, B * ts


def terminal_cost_sweep(A, B, Q, R, P_f, x0, horizons=range(1, 10)):
    """Numeric part of ``compare_term_cost`` (reference FHC.py:117-127): V_N(x0) = x0' P_N[0] x0 for
    each horizon, plus the infinite-horizon cost from the DARE (one-off host call, scipy, as in
    the reference).  Returns (list of V_N, V_inf)."""
    from scipy import linalg
    V = []
    x0n = np.asarray(io.back(x0, False).cpu() if io.is_tensor(x0) else x0, dtype=np.float64)
    for N in horizons:
        P, _ = ricatti_recursion(A, B, Q, R, P_f, N)
        P0 = P[0].cpu().numpy() if io.is_tensor(P[0]) else P[0]
        V.append(float(np.squeeze(x0n.T @ P0 @ x0n)))
    P_inf = linalg.solve_discrete_are(np.asarray(A, float), np.asarray(B, float), np.asarray(Q, float),
                                      np.asarray(R, float))
    return V, float(np.squeeze(x0n.T @ P_inf @ x0n))


def compare_term_cost(A, B, Q, R, P_f, x0):
    """Reference FHC.py:117-131 (the plot needs matplotlib; the numbers do not)."""
    V, V_inf = terminal_cost_sweep(A, B, Q, R, P_f, x0)
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        return V, V_inf
    plt.plot(np.arange(1, 10), np.array(V), marker="x", linestyle="--", markersize=8)
    plt.hlines(V_inf, 1, 9, colors="#20A0EA", linestyles="dashdot")
    plt.show()
    return V, V_inf
