"""Drop-in for the reference's ``session_2/problem.py`` and ``session_3/problem.py``.

Same dataclass, field names, defaults, ``__post_init__`` and properties as the reference
(/root/reference/session_2/problem.py:4-32); the ndarray defaults use ``default_factory`` because
the reference's literal defaults raise ``ValueError`` on Python >= 3.11.  ``Problem()`` ,
``Problem(N=30)`` and ``Problem(Ts=...)`` behave as in the reference.

The reference has no solver for this data; :class:`LinearMPC` poses the QP (stage cost
x'Qx + u'Ru over k < N, terminal x_N'Q x_N, box bounds on position, velocity and input) and
solves it for a whole batch of initial states on the GPU (K4).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import _interop as io
from . import boxqp, lq
from .log import ControllerLog


@dataclass
class Problem:
    """Convenience class representing the problem data for session 2."""

    Ts: float = 0.3
    Q: np.ndarray = field(default_factory=lambda: np.diag([10, 1]))
    R: np.ndarray = field(default_factory=lambda: np.diag([0.01]))
    p_min: float = -150  # Minimal position
    p_max: float = 1.0   # Maximal position
    v_min: float = -20   # Minimal velocity
    v_max: float = 25.0  # Maximal velocity
    u_min: float = -20.0
    u_max: float = 10.0
    N: int = 5

    A: np.ndarray = None
    B: np.ndarray = None

    def __post_init__(self):
        self.A = np.array([[1.0, self.Ts], [0, 1.0]])
        self.B = np.array([[0], [self.Ts]])

    @property
    def n_state(self):
        return self.A.shape[0]

    @property
    def n_input(self):
        return self.B.shape[1]


@dataclass
class Problem3(Problem):
    """Session 3 variant (/root/reference/session_3/problem.py:15,17).  ``problem3.Problem`` is this class under the
    reference's own name, for code that does ``from problem import Problem`` inside session_3."""
    p_min: float = -120
    v_min: float = -50


class LinearMPC:
    """Batched MPC policy for a :class:`Problem`.

    ``solve(x0)`` with x0 (n,) or (batch, n) returns a :class:`boxqp.BoxQpResult` (fields
    ``solver_success``, ``state_prediction`` (batch, N+1, n), ``input_prediction`` (batch, N, m) as
    in the reference's ControllerLog; the tensors alias a workspace that the next ``solve`` reuses).
    ``__call__(y, log=None)`` is the policy the course's
    simulator calls every step: returns u_0 and appends the three log entries.
    """

    def __init__(self, problem: Problem, terminal_weight=None, max_iter=60, eps=1e-9, dtype=torch.float64):
        self.problem = problem
        self.Pf = problem.Q if terminal_weight is None else terminal_weight
        self.max_iter, self.eps = max_iter, eps
        self.dtype = dtype   # torch.float64 (reference arithmetic) or torch.float32 arrays (1e-4 tolerance class)
        self._ws = None

    def bounds(self):
        p = self.problem
        return [p.u_min], [p.u_max], [p.p_min, p.v_min], [p.p_max, p.v_max]

    def solve(self, x0, warm_U=None):
        p = self.problem
        dt = self.dtype
        x = io.to_dev(x0, dt)
        if x.dim() == 1:
            x = x[None, :]
        if x.shape[1] != p.n_state:
            raise ValueError(f"x0 must be (batch, {p.n_state})")
        xT = x.t().contiguous()
        dev = xT.device
        shape = (xT.shape[1], p.n_state, p.n_input, int(p.N))
        if self._ws is None or self._ws.shape != shape or self._ws.U.device != dev or self._ws.dtype != dt:
            self._ws = boxqp.BoxQpWorkspace(*shape, dev, dtype=dt)
        u_lo, u_hi, x_lo, x_hi = self.bounds()
        A, B = io.to_dev(p.A, dt, dev), io.to_dev(p.B, dt, dev)
        Q, R, Pf = (io.to_dev(np.asarray(M, dtype=np.float64) if not io.is_tensor(M) else M, dt, dev)
                    for M in (p.Q, p.R, self.Pf))
        return boxqp.solve(A, B, Q, R, Pf, p.N, xT, u_lo, u_hi, x_lo, x_hi, warm_U=warm_U,
                           max_iter=self.max_iter, eps=self.eps, workspace=self._ws)

    def __call__(self, y, log: ControllerLog = None):
        as_np = not io.is_tensor(y)
        single = (np.ndim(y) if as_np else y.dim()) == 1
        res = self.solve(y)
        # the result aliases the reusable workspace: what leaves this call is copied
        keep = (lambda t: io.back(t, True)) if as_np else (lambda t: t.clone())
        if log is not None:
            log.solver_success.append(keep(res.solver_success[0] if single else res.solver_success))
            log.state_prediction.append(keep(res.state_prediction[0] if single else res.state_prediction))
            log.input_prediction.append(keep(res.input_prediction[0] if single else res.input_prediction))
        u0 = res.U[0].t()  # (batch, m)
        return keep(u0[0] if single else u0)


def closed_loop(problem: Problem, x0, n_steps: int, controller: LinearMPC = None, log: ControllerLog = None):
    """Closed loop x+ = A x + B u0(x) for ``n_steps`` steps (the course's
    ``simulate(x0, dynamics, n_steps, policy=controller, log=log)``, batched).  Returns states
    (batch, n_steps+1, n) and inputs (batch, n_steps, m); scenarios whose QP becomes infeasible keep
    being simulated with the solver's last (bound-clamped) input and are flagged in ``log``."""
    controller = controller or LinearMPC(problem)
    as_np = not io.is_tensor(x0)
    x = io.to_dev(x0, controller.dtype)
    if x.dim() == 1:
        x = x[None, :]
    A, B = io.to_dev(problem.A, controller.dtype, x.device), io.to_dev(problem.B, controller.dtype, x.device)
    xs, us = [x], []
    for _ in range(int(n_steps)):
        u = controller(xs[-1], log)
        xn = lq.linear_step(A, B, xs[-1].t().contiguous(), u.to(xs[-1].dtype).t().contiguous()).t()
        xs.append(xn)
        us.append(u)
    return io.back(torch.stack(xs, dim=1), as_np), io.back(torch.stack(us, dim=1), as_np)
