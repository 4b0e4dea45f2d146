"""Drop-in for the reference's ``session_1/LinearSystem.py`` (class ``LinearSystem``).

Same constructor, methods, argument meaning and tensor layouts as the reference
(/root/reference/session_1/LinearSystem.py:7-46); the closed loop and the open-loop prediction
run as one batched CUDA kernel (K2, ``mpc_lq_rollout``) when the policy is an ``AutoCruising``
gain policy, and as a per-step device loop (``mpc_linear_step``) for arbitrary callables.
Column-batched initial states ``(n, batch)`` work exactly as they do in the reference's numpy code.
"""
from __future__ import annotations

from typing import Callable

import torch

from . import _interop as io
from . import lq


def _gain_policy(law):
    """(owner, gain_offset, gain_step) when ``law`` is a bound AutoCruising policy, else None."""
    owner = getattr(law, "__self__", None)
    func = getattr(law, "__func__", None)
    if owner is None or func is None or not hasattr(owner, "gains"):
        return None
    kind = getattr(func, "_mpc_gain_policy", None)
    if kind == "receding":      # u = gains[0] @ x       (reference FHC.py:25-26)
        return owner, 0, 0
    if kind == "time_varying":  # u = gains[t] @ x       (reference FHC.py:28-29)
        return owner, None, 1
    return None


class LinearSystem:
    def __init__(self, A, B) -> None:
        self.A = A
        self.B = B

    def set_output_eq(self, C, D) -> None:
        self.C = C
        self.D = D

    # -- x+ = A x + B u (reference LinearSystem.py:16-18)
    def f(self, x, u):
        as_np = not io.any_tensor(x, u)
        dt = io.pick_dtype(x, u, self.A)
        xd, ud = io.to_dev(x, dt), io.to_dev(u, dt)
        squeeze = xd.dim() == 1
        if squeeze:
            xd, ud = xd[:, None], ud[:, None]
        xn = lq.linear_step(io.to_dev(self.A, dt), io.to_dev(self.B, dt), xd, ud)
        return io.back(xn[:, 0] if squeeze else xn, as_np)

    def _rollout(self, x0, law: Callable, T: int, t_first: int):
        """States x_0 .. x_{T-1} with u = law(x, t), t = t_first, t_first+1, ...; (n, batch, T)."""
        as_np = not io.is_tensor(x0)
        dt = io.pick_dtype(x0, self.A)
        x0d = io.to_dev(x0, dt)
        if x0d.dim() != 2:
            raise ValueError("x0 must be a column (n, 1) or column-batched (n, batch) array, as in the reference")
        A, B = io.to_dev(self.A, dt), io.to_dev(self.B, dt)
        T = int(T)
        fused = _gain_policy(law)
        if fused is not None and T >= 1:
            owner, off, step = fused
            K = owner._gains_tensor(dt)
            if off is None:
                off = t_first
            res = lq.lq_rollout(A, B, K, x0d, T, gain_offset=off, gain_step=step)
            X = res["X"].permute(1, 2, 0)  # [T, n, batch] -> (n, batch, T) view, batch-contiguous storage
            return io.back(X, as_np)
        # arbitrary policy: one device step per call of the user's callable
        xs = [x0d]
        for i in range(max(T - 1, 0)):
            xt = xs[-1]
            u = law(io.back(xt, as_np), t_first + i)
            xs.append(lq.linear_step(A, B, xt, io.to_dev(u, dt).reshape(-1, xt.shape[1])))
        return io.back(torch.stack(xs, dim=2), as_np)

    # -- reference LinearSystem.py:20-26: `steps` states including x0, t = 1 .. steps-1, result in self.x
    def simulate(self, x0, control_law: Callable, steps: int) -> None:
        self.x = self._rollout(x0, control_law, max(int(steps), 1), 1)

    # -- reference LinearSystem.py:28-35: t = 1 .. horizon-1 (gains[0] is never applied)
    def prediction(self, xt, pred_law: Callable, horizon: int):
        return self._rollout(xt, pred_law, max(int(horizon), 1), 1)

    # -- presentation (reference LinearSystem.py:37-46); matplotlib is optional
    def plot_traj(self) -> None:
        import matplotlib.pyplot as plt
        x = io.back(self.x, False)
        x = x.detach().cpu().numpy() if io.is_tensor(x) else x
        plt.plot(x[0, 0, :], x[1, 0, :], "x", linestyle="--", color="#685BF5", label="Trajectory")
        plt.legend()

    def plot_cost(self, P_N) -> None:
        pass

    def plot_pred(self, horizon: int) -> None:
        pass
