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


def infinite_horizon(A, B, Q, R, tol=None, max_horizon=1 << 16):
    """(P_inf, K_inf) of the reference's infinite-horizon comparison (FHC.py:97-98,126), computed on
    the GPU by running the Riccati recursion (K1) with doubling horizons until P stops changing --
    the fixed point of the recursion is the stabilising DARE solution the reference gets from
    scipy.linalg.solve_discrete_are.  K_inf = -(R + B'P B)^-1 B'P A is the converged first-stage gain.
    ``tol``: relative change of P that counts as converged (default 1e-13 in float64, 1e-5 in float32: what the
    arithmetic can resolve); a RuntimeWarning is issued when ``max_horizon`` is reached first."""
    as_np = not io.any_tensor(A, B, Q, R)
    dt = io.pick_dtype(A, B, Q, R)
    if tol is None:
        tol = 1e-13 if dt == torch.float64 else 1e-5
    Ad, Bd, Qd = (io.to_dev(M, dt) for M in (A, B, Q))
    Rd = _prep_R(R, Bd.shape[-1], dt)
    P, N = Qd, 64
    while True:
        K, Pn = lq.riccati(Ad, Bd, Qd, Rd, P, N, all_P=False)
        P_new = Pn[0] if Ad.dim() == 2 else Pn
        done = bool((P_new - P).abs().max() <= tol * P_new.abs().max())
        P = P_new
        if done:
            break
        if N >= max_horizon:
            import warnings
            warnings.warn(f"infinite_horizon: P still changes by more than {tol:g} (relative) at horizon {N}; "
                          "returning the last iterate (unstabilisable model, or tol below the arithmetic's resolution)",
                          RuntimeWarning, stacklevel=2)
            break
        N *= 2
    K0 = K[0, 0] if Ad.dim() == 2 else K[0]
    return io.back(P, as_np), io.back(K0, as_np)


def closed_loop_with_predictions(A, B, Q, R, P_f, x0, N, n_steps=30, gains=None):
    """Numeric part of one panel of ``run_and_plot_traj`` (reference FHC.py:70-91): gains for horizon
    N, the closed loop under gains[0] (``n_steps`` states, (n, batch, n_steps)) and, for every
    closed-loop state x_t, the open-loop prediction of horizon N through ``LinearSystem.prediction``
    (the reference loops over t; here all n_steps predictions are ONE batched rollout).
    Returns (gains, x, bundle) with bundle[t] = prediction from x_t, shape (n_steps, n, batch, N)."""
    if gains is None:
        _, gains = ricatti_recursion(A, B, Q, R, P_f, N)
    sys = AutoCruising(A, B)
    sys.set_opti_gain(gains)
    sys.simulate(x0, sys.control_law, n_steps)
    x = sys.x
    n, batch = x.shape[0], x.shape[1]
    if io.is_tensor(x):
        starts = x.permute(0, 2, 1).reshape(n, n_steps * batch)
    else:
        starts = np.ascontiguousarray(np.transpose(x, (0, 2, 1)).reshape(n, n_steps * batch))
    pred = sys.prediction(starts, sys.pred, N)               # (n, n_steps*batch, N)
    bundle = pred.reshape(n, n_steps, batch, pred.shape[-1])
    bundle = bundle.permute(1, 0, 2, 3) if io.is_tensor(bundle) else np.transpose(bundle, (1, 0, 2, 3))
    return gains, x, bundle


def closed_loop_spectral_radius(A, B, K):
    """rho(A + B K): the exact stability test of the receding-horizon closed loop u = gains[0] x that the reference
    hints at (session_1/session1_sol.py:114-116) -- stable iff < 1.  ``K`` is ``gains[0]`` (or any (m, n) gain);
    A, B, K may carry a leading batch axis.  numpy in -> numpy out."""
    as_np = not io.any_tensor(A, B, K)
    dt = io.pick_dtype(A, B, K)
    rho = lq.spectral_radius(io.to_dev(A, dt), io.to_dev(B, dt), io.to_dev(K, dt))
    batched = any(np.ndim(M) == 3 if not io.is_tensor(M) else M.dim() == 3 for M in (A, B, K))
    out = rho if batched else rho[0]
    return io.back(out, as_np) if batched else (float(out.item()) if as_np else out)


def run_and_plot_traj(A, B, Q, R, P_f, x0):
    """Reference FHC.py:64-114: closed loops and prediction bundles for N in (4, 6, 10) and for the
    infinite-horizon gain.  Returns the numbers ({N: (x, bundle)}, (x_inf, bundle_inf)); the figures
    are drawn only when matplotlib is importable."""
    out = {}
    for N in (4, 6, 10):
        _, x, bundle = closed_loop_with_predictions(A, B, Q, R, P_f, x0, N)
        out[N] = (x, bundle)
    _, K_inf = infinite_horizon(A, B, Q, R)
    _, x_inf, b_inf = closed_loop_with_predictions(A, B, Q, R, P_f, x0, 10, gains=[K_inf] * 10)
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        return out, (x_inf, b_inf)
    for idx, N in enumerate((4, 6, 10)):
        x, bundle = out[N]
        plt.figure(1)
        plt.subplot(1, 3, idx + 1)
        plt.plot(x[0, 0, :], x[1, 0, :], "x", linestyle="--")
        for t in range(bundle.shape[0]):
            plt.plot(bundle[t, 0, 0, :], bundle[t, 1, 0, :], "o", linestyle=":")
    plt.show()
    return out, (x_inf, b_inf)


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
