"""Drop-in for the hot path of the reference's ``session_4/session4_sol.py`` (variants:
``template.py``): integrators, ``MPCController`` and the closed-loop ``simulate`` driver, batched
over scenarios on the GPU.

Call surface kept from the reference:
    forward_euler(f, ts), runge_kutta4(f, ts), exact_integration(f, ts)   (session4_sol.py:22-56)
    MPCController(N, ts, *, params) with .N, .ts, .bounds, solve(x) -> {"x": U},
        reshape_input(sol) -> (N, 2), __call__(y) -> u_0                   (session4_sol.py:113-230)
    simulate(x0, dynamics, n_steps, policy=controller)                      (call sites :458, :465)
    KinematicBicycle(params, symbolic=False)  -- the model the reference imports from rcracers.

Deliberate differences: the nonlinear OCP is not handed to IPOPT; one linearised
QP is solved per control step (real-time iteration), warm-started by the shifted previous plan --
the reference cold-starts every step (session4_sol.py:129-130).  ``exact_integration`` is RK4 with
fixed sub-steps instead of scipy ``odeint``.  The bicycle ODE is our definition (rcracers is not
available), csrc/bicycle_core.cuh.
"""
from __future__ import annotations

from ctypes import c_double, c_int, c_int64, c_void_p
from dataclasses import dataclass
from typing import Callable

import numpy as np
import torch

from . import _interop as io
from . import _lib, boxqp
from .parameters import VehicleParameters

_lib.register("mpc_bicycle_rti_prepare", c_int, [c_double] * 5 + [c_int, c_void_p, c_void_p, c_int] + [c_void_p] * 4 +
              [c_int64, c_int, c_int, c_void_p])
_lib.register("mpc_bicycle_plant_step", c_int, [c_double] * 4 + [c_void_p, c_int64, c_int] + [c_void_p] * 3 +
              [c_int64, c_int, c_void_p])
_lib.register("mpc_rti_workspace_bytes", c_int64, [c_int64, c_int, c_int])
_lib.register("mpc_rti_closed_loop", c_int, [c_double] * 5 + [c_int, c_void_p, c_int, c_int] + [c_void_p] * 21 +
              [c_int64, c_int64, c_int, c_int, c_double, c_int, c_void_p])

_lib.register("mpc_bicycle_rti_prepare_obstacle", c_int, [c_double] * 5 + [c_int, c_double, c_double, c_void_p, c_void_p,
              c_void_p, c_int] + [c_void_p] * 6 + [c_int64, c_int, c_int, c_void_p])

F64 = torch.float64


# ------------------------------------------------------------------------------------------------
# model and integrators
# ------------------------------------------------------------------------------------------------
class KinematicBicycle:
    """Continuous-time model (x, u) -> xdot, state [p_x, p_y, psi, v], input [a, delta].
    ``symbolic`` is accepted for signature compatibility (the reference builds CasADi expressions
    with it, session4_sol.py:191); the GPU path differentiates the model analytically."""

    def __init__(self, params: VehicleParameters = None, symbolic: bool = False):
        self.params = params if params is not None else VehicleParameters()
        self.symbolic = symbolic

    def __call__(self, x, u):
        p = self.params
        xp = torch if io.is_tensor(x) else np
        psi, v = x[..., 2], x[..., 3]
        beta = xp.arctan(p.axis_rear * xp.tan(u[..., 1]) / (p.axis_rear + p.axis_front))
        return xp.stack([v * xp.cos(psi + beta), v * xp.sin(psi + beta), v * xp.sin(beta) / p.axis_rear,
                         p.acceleration * u[..., 0] - p.friction * v], -1)


class _Discrete:
    """Discrete-time dynamics (x, u) -> x+ produced by one of the integrator factories.  When ``f`` is
    a :class:`KinematicBicycle` the step runs as a CUDA kernel and ``simulate`` can fuse it."""

    def __init__(self, f, ts, kind, substeps):
        self.f, self.ts, self.kind, self.substeps = f, float(ts), kind, substeps

    @property
    def fusable(self):
        return isinstance(self.f, KinematicBicycle)

    def __call__(self, x, u):
        if self.fusable:
            as_np = not io.any_tensor(x, u)
            xd, ud = io.to_dev(x, F64), io.to_dev(u, F64)
            single = xd.dim() == 1
            xT = (xd[None, :] if single else xd).t().contiguous()
            uT = (ud[None, :] if single else ud).t().contiguous()
            xn = plant_step(self.f.params, self.ts, xT, uT, substeps=self.substeps).t()
            return io.back(xn[0] if single else xn, as_np)
        f, ts = self.f, self.ts
        if self.kind == "euler":
            return x + f(x, u) * ts
        h = ts / max(self.substeps, 1)
        for _ in range(max(self.substeps, 1)):
            s1 = f(x, u)
            s2 = f(x + 0.5 * h * s1, u)
            s3 = f(x + 0.5 * h * s2, u)
            s4 = f(x + h * s3, u)
            x = x + h / 6.0 * (s1 + 2 * s2 + 2 * s3 + s4)
        return x


def forward_euler(f, ts) -> Callable:
    return _Discrete(f, ts, "euler", 0)


def runge_kutta4(f, ts) -> Callable:
    return _Discrete(f, ts, "rk4", 1)


def exact_integration(f, ts, substeps: int = 4, adaptive: bool = False, tol_exp: int = 10) -> Callable:
    """Accurate plant integration (the reference uses scipy odeint, session4_sol.py:37-56): RK4 with
    ``substeps`` sub-steps, or -- ``adaptive=True``, bicycle dynamics only -- the Dormand-Prince 5(4)
    pair with rtol = atol = 10^-tol_exp on the GPU."""
    if adaptive:
        if not isinstance(f, KinematicBicycle):
            raise ValueError("the adaptive integrator runs on the GPU and needs KinematicBicycle dynamics")
        return _Discrete(f, ts, "rk4", -int(tol_exp))
    return _Discrete(f, ts, "rk4", int(substeps))


def plant_step(params, ts, x, u, friction=None, substeps=4):
    """x [4, batch], u [2, batch] -> x+ [4, batch] on the GPU.  ``friction``: None (params.friction),
    scalar or [batch] tensor.  substeps = 0: forward Euler; > 0: RK4 sub-steps."""
    _lib.require_cuda(x, u)
    batch = x.shape[1]
    if friction is None:
        friction = params.friction
    fr = friction if io.is_tensor(friction) else torch.full((1,), float(friction), dtype=F64, device=x.device)
    fr = fr.to(device=x.device, dtype=F64).contiguous()
    sfr = 1 if fr.numel() == batch and batch > 1 else (1 if fr.numel() == batch else 0)
    if fr.numel() not in (1, batch):
        raise ValueError("friction must be a scalar or one value per scenario")
    if fr.numel() == 1:
        sfr = 0
    xn = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().mpc_bicycle_plant_step(params.axis_rear, params.axis_front, params.acceleration, float(ts),
                                                     _lib.ptr(fr), sfr, int(substeps), _lib.ptr(x.contiguous()),
                                                     _lib.ptr(u.contiguous()), _lib.ptr(xn), batch, _lib.MPC_F64,
                                                     _lib.stream(x.device)))
    return xn


# ------------------------------------------------------------------------------------------------
# controller
# ------------------------------------------------------------------------------------------------
@dataclass
class RtiClosedLoopResult:
    X: torch.Tensor            # [steps+1, 4, batch] closed-loop states
    U: torch.Tensor            # [steps, 2, batch] applied inputs
    cost: torch.Tensor         # [batch] sum_t x_t'Q x_t + u_t'R u_t
    violation: torch.Tensor    # [batch] max state-bound violation along the closed loop
    n_saturated: torch.Tensor  # [batch] applied inputs on a bound
    n_failed: torch.Tensor     # [batch] steps whose QP did not report success
    iters: torch.Tensor        # [batch] interior-point iterations over all steps
    last_status: torch.Tensor  # [batch]
    X_bundle: torch.Tensor = None  # [steps, N+1, 4, batch] state prediction of every control step (keep_predictions)
    U_bundle: torch.Tensor = None  # [steps, N, 2, batch]

    def bundle(self, scenario: int = 0):
        """(time steps x horizon x states) array of one scenario, the layout AnimateParking.bundle takes
        (reference session_4/animation.py:75-83)."""
        return self.X_bundle[:, :, :, scenario]

    @property
    def states(self):          # (batch, steps+1, 4) view
        return self.X.permute(2, 0, 1)

    @property
    def inputs(self):          # (batch, steps, 2) view
        return self.U.permute(2, 0, 1)


class MPCController:
    """RTI counterpart of the reference controller (session4_sol.py:113-230).

    Weights Q = diag(1, 3, .1, .01), Q_T = ``terminal_scale`` * Q (10 in session4_sol.py:167, 5 in
    template.py:136), R = diag(1, .01); input box as variable bounds and state box on x_1..x_N from
    ``params`` (session4_sol.py:176-181).  ``integrator``: "euler" (session4_sol.py:192) or "rk4"
    (template.py:141).  The controller keeps one input plan per scenario between calls (warm start).
    """

    def __init__(self, N: int, ts: float, *, params: VehicleParameters, integrator: str = "euler",
                 terminal_scale: float = 10.0, max_iter: int = 60, eps: float = 1e-9):
        if integrator not in ("euler", "rk4"):
            raise ValueError("integrator must be 'euler' or 'rk4'")
        self.N, self.ts = int(N), float(ts)
        self.params = params
        self.integrator = integrator
        self.max_iter, self.eps = max_iter, eps
        self.Q = np.diag([1.0, 3.0, 0.1, 0.01])
        self.QT = terminal_scale * self.Q
        self.R = np.diag([1.0, 1e-2])
        self.bounds = self.build_bounds(params)
        self._plan = None  # [N, 2, batch]
        self._qp_ws = None

    # -- the bounds dictionary of the reference's build_ocp (session4_sol.py:206-212)
    def build_bounds(self, params):
        s_lb = np.array([params.min_pos_x, params.min_pos_y, params.min_heading, params.min_vel])
        s_ub = np.array([params.max_pos_x, params.max_pos_y, params.max_heading, params.max_vel])
        i_lb = np.array([params.min_drive, -params.max_steer])
        i_ub = np.array([params.max_drive, params.max_steer])
        self._boxes = (i_lb, i_ub, s_lb, s_ub)
        return {"lbx": np.tile(i_lb, self.N), "ubx": np.tile(i_ub, self.N),
                "lbg": np.tile(s_lb, self.N), "ubg": np.tile(s_ub, self.N)}

    def reset(self):
        self._plan = None

    def _model_args(self):
        p = self.params
        return (p.axis_rear, p.axis_front, p.acceleration, float(p.friction), self.ts, 1 if self.integrator == "rk4" else 0)

    def _solve_dev(self, yT):
        """yT [4, batch] on the device -> BoxQpResult (aliases the controller's workspace)."""
        batch, N = yT.shape[1], self.N
        dev = yT.device
        first = self._plan is None or self._plan.shape[2] != batch or self._plan.device != dev
        if first:
            self._plan = torch.zeros((N, 2, batch), dtype=F64, device=dev)
            self._lin = [torch.empty((N, k, batch), dtype=F64, device=dev) for k in (2, 16, 8, 4)]
            self._qp_ws = boxqp.BoxQpWorkspace(batch, 4, 2, N, dev)
        warm, A, B, c = self._lin
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mpc_bicycle_rti_prepare(*self._model_args(), _lib.ptr(yT), _lib.ptr(self._plan),
                                                          1 if first else 0, _lib.ptr(warm), _lib.ptr(A), _lib.ptr(B),
                                                          _lib.ptr(c), batch, N, _lib.MPC_F64, _lib.stream(dev)))
        i_lb, i_ub, s_lb, s_ub = self._boxes
        Q, R, QT = (torch.as_tensor(M, dtype=F64, device=dev) for M in (self.Q, self.R, self.QT))
        res = boxqp.solve(A, B, Q, R, QT, N, yT, i_lb, i_ub, s_lb, s_ub, c=c, warm_U=warm, max_iter=self.max_iter,
                          eps=self.eps, workspace=self._qp_ws)
        self._plan.copy_(res.U)
        return res

    def solve(self, x) -> dict:
        """{"x": stacked inputs}: (2N,) for one state, (batch, 2N) for a batch -- the layout of the
        reference's ``sol["x"]`` (u_0, u_1, ... concatenated, session4_sol.py:206)."""
        as_np = not io.is_tensor(x)
        xd = io.to_dev(x, F64)
        single = xd.dim() == 1
        yT = (xd[None, :] if single else xd).t().contiguous()
        res = self._solve_dev(yT)
        U = res.U.permute(2, 0, 1).reshape(yT.shape[1], 2 * self.N)
        out = {"x": io.back(U[0] if single else U, True) if as_np else (U[0] if single else U).clone(),
               "success": io.back(res.solver_success[0] if single else res.solver_success, as_np),
               "state_prediction": io.back(res.state_prediction[0] if single else res.state_prediction.clone(), as_np)}
        return out

    def reshape_input(self, sol):
        x = sol["x"]
        if io.is_tensor(x):
            return x.reshape(x.shape[:-1] + (-1, 2)) if x.dim() > 1 else x.reshape(-1, 2)
        return np.reshape(x, x.shape[:-1] + (-1, 2)) if np.ndim(x) > 1 else np.reshape(x, (-1, 2))

    def __call__(self, y):
        u = self.reshape_input(self.solve(y))
        return u[0] if (u.dim() if io.is_tensor(u) else u.ndim) == 2 else u[:, 0]

    # -- fused closed loop: `steps` x (prepare, QP, plant) in one kernel
    def closed_loop(self, x0, n_steps, plant: _Discrete = None, friction_plant=None,
                    keep_predictions: bool = False) -> RtiClosedLoopResult:
        xd = io.to_dev(x0, F64)
        if xd.dim() == 1:
            xd = xd[None, :]
        x0T = xd.t().contiguous()
        batch, N, dev = x0T.shape[1], self.N, x0T.device
        pp = self.params if plant is None or not plant.fusable else plant.f.params
        substeps = 4 if plant is None else (0 if plant.kind == "euler" else (plant.substeps if plant.substeps < 0 else max(plant.substeps, 1)))
        if plant is not None and abs(plant.ts - self.ts) > 1e-15:
            raise ValueError("plant and controller sampling times differ")
        if friction_plant is None:
            fr = torch.full((batch,), float(pp.friction), dtype=F64, device=dev)
        else:
            fr = io.to_dev(friction_plant, F64).reshape(-1).expand(batch).contiguous() if not io.is_tensor(friction_plant) \
                else friction_plant.to(device=dev, dtype=F64).reshape(-1).expand(batch).contiguous()
        plan = torch.zeros((N, 2, batch), dtype=F64, device=dev)
        Xp = torch.empty((N + 1, 4, batch), dtype=F64, device=dev)
        Xc = torch.empty((n_steps + 1, 4, batch), dtype=F64, device=dev)
        Uc = torch.empty((n_steps, 2, batch), dtype=F64, device=dev)
        cost = torch.empty(batch, dtype=F64, device=dev)
        viol = torch.empty(batch, dtype=F64, device=dev)
        ints = [torch.empty(batch, dtype=torch.int32, device=dev) for _ in range(4)]
        Xb = torch.empty((n_steps, N + 1, 4, batch), dtype=F64, device=dev) if keep_predictions else None
        Ub = torch.empty((n_steps, N, 2, batch), dtype=F64, device=dev) if keep_predictions else None
        nbytes = _lib.lib().mpc_rti_workspace_bytes(batch, N, _lib.MPC_F64)
        ws = torch.empty(max(nbytes // 8, 1), dtype=F64, device=dev)
        i_lb, i_ub, s_lb, s_ub = (torch.as_tensor(v, dtype=F64, device=dev).contiguous() for v in self._boxes)
        Q, R, QT = (torch.as_tensor(M, dtype=F64, device=dev).contiguous() for M in (self.Q, self.R, self.QT))
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mpc_rti_closed_loop(
                *self._model_args(), _lib.ptr(fr), int(substeps), int(n_steps), _lib.ptr(Q), _lib.ptr(R), _lib.ptr(QT),
                _lib.ptr(i_lb), _lib.ptr(i_ub), _lib.ptr(s_lb), _lib.ptr(s_ub), _lib.ptr(x0T), _lib.ptr(plan), _lib.ptr(Xp),
                _lib.ptr(Xc), _lib.ptr(Uc), _lib.ptr(cost), _lib.ptr(viol), *[_lib.ptr(t) for t in ints], _lib.ptr(Xb), _lib.ptr(Ub),
                _lib.ptr(ws),
                nbytes, batch, N, int(self.max_iter), float(self.eps), _lib.MPC_F64, _lib.stream(dev)))
        self._plan = plan
        return RtiClosedLoopResult(Xc, Uc, cost, viol, ints[0], ints[1], ints[2], ints[3], Xb, Ub)


# ------------------------------------------------------------------------------------------------
# obstacle-avoidance controller (reference session_4/main.py)
# ------------------------------------------------------------------------------------------------
def x2T(x, symbolic: bool = False):
    """Homogeneous transform of a pose [p_x, p_y, psi, ...] (reference main.py:173-188)."""
    c, s_ = np.cos(x[2]), np.sin(x[2])
    return np.array([[c, -s_, x[0]], [s_, c, x[1]], [0.0, 0.0, 0.0]])


def create_cover_circles(l, w, n_c: int):
    """Centres (homogeneous, vehicle frame) and radius of the n_c covering circles (reference main.py:191-200)."""
    d = l / (2 * n_c)
    r = np.sqrt(d ** 2 + (w ** 2) / 4)
    return [np.array([(2 * k + 1) * d - l / 2, 0, 1]) for k in range(n_c)], r


class ObstacleMPCController(MPCController):
    """RTI counterpart of the obstacle-avoidance controller of the reference's ``session_4/main.py``
    (:29-129): same constructor arguments ``(N, ts, params, model, x_obs)``, weights
    Q = diag(1, 6, .2, .05), Q_N = 100 Q, R = diag(1, .01) (:72-74), forward-Euler prediction model
    (:76), input box, state box, and the nine collision constraints between the three covering
    circles of the vehicle and of the obstacle parked at ``x_obs`` (:49-56, :95-104).  Every control
    step linearises the collision constraints along the rolled-out plan and solves ONE QP with
    polytopic stage constraints on the GPU (K4 with general rows)."""

    def __init__(self, N: int, ts: float, params: VehicleParameters, model=None, x_obs=None, max_iter: int = 60,
                 eps: float = 1e-9):
        if x_obs is None:
            raise ValueError("x_obs (pose of the parked obstacle) is required")
        super().__init__(N, ts, params=params, integrator="euler", max_iter=max_iter, eps=eps)
        self.model = model
        self.x_obs = np.asarray(x_obs, dtype=np.float64).reshape(-1)
        self.Q = np.diag([1.0, 6.0, 0.2, 0.05])
        self.QT = 100.0 * self.Q
        self.R = np.diag([1.0, 0.01])
        n_c = 3
        self.bounds["lbg"] = np.tile(np.concatenate([self._boxes[2], np.full(n_c * n_c, self.collision_radius2())]), self.N)
        self.bounds["ubg"] = np.tile(np.concatenate([self._boxes[3], np.full(n_c * n_c, np.inf)]), self.N)

    def collision_radius2(self):
        _, r = create_cover_circles(self.params.length, self.params.width, 3)
        return float((2 * r) ** 2)

    def _solve_dev(self, yT):
        batch, N = yT.shape[1], self.N
        dev = yT.device
        first = self._plan is None or self._plan.shape[2] != batch or self._plan.device != dev
        if first:
            self._plan = torch.zeros((N, 2, batch), dtype=F64, device=dev)
            self._lin = [torch.empty((N, k, batch), dtype=F64, device=dev) for k in (2, 16, 8, 4, 36, 9)]
            self._qp_ws = boxqp.BoxQpWorkspace(batch, 4, 2, N, dev, nc=9)
        warm, A, B, c, Cg, hg = self._lin
        from ctypes import POINTER
        xo = (c_double * 4)(*[float(v) for v in self.x_obs[:4]])
        p = self.params
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mpc_bicycle_rti_prepare_obstacle(
                *self._model_args(), float(p.length), float(p.width), xo, _lib.ptr(yT), _lib.ptr(self._plan),
                1 if first else 0, _lib.ptr(warm), _lib.ptr(A), _lib.ptr(B), _lib.ptr(c), _lib.ptr(Cg), _lib.ptr(hg), batch, N,
                _lib.MPC_F64, _lib.stream(dev)))
        i_lb, i_ub, s_lb, s_ub = self._boxes
        Q, R, QT = (torch.as_tensor(M, dtype=F64, device=dev) for M in (self.Q, self.R, self.QT))
        res = boxqp.solve(A, B, Q, R, QT, N, yT, i_lb, i_ub, s_lb, s_ub, c=c, warm_U=warm, max_iter=self.max_iter,
                          eps=self.eps, workspace=self._qp_ws, Cg=Cg, hg=hg)
        self._plan.copy_(res.U)
        return res

    def closed_loop(self, *args, **kwargs):
        raise NotImplementedError("the fused closed-loop kernel covers the box-constrained controller; drive this "
                                  "controller step by step (simulate(x0, dynamics, n_steps, policy=controller))")


# ------------------------------------------------------------------------------------------------
# closed-loop driver (the reference imports it from rcracers.simulator)
# ------------------------------------------------------------------------------------------------
def simulate(x0, dynamics: Callable, n_steps: int, policy=None, friction_plant=None):
    """States of the closed loop: (n_steps+1, 4) for one x0, (batch, n_steps+1, 4) for a batch.
    With an :class:`MPCController` policy and bicycle dynamics from one of the integrator factories
    the whole loop is one fused kernel; any other callables run step by step.  The policy is called
    as ``policy(y)`` or ``policy(y, t)`` depending on its signature, as the course simulator does."""
    as_np = not io.is_tensor(x0)
    if (isinstance(policy, MPCController) and not isinstance(policy, ObstacleMPCController)
            and isinstance(dynamics, _Discrete) and dynamics.fusable):
        policy.reset()
        res = policy.closed_loop(x0, int(n_steps), plant=dynamics, friction_plant=friction_plant)
        X = res.states
        single = (np.ndim(x0) if as_np else x0.dim()) == 1
        return io.back(X[0] if single else X, as_np)
    import inspect
    if isinstance(policy, MPCController):
        policy.reset()
    takes_t = policy is not None and len(inspect.signature(policy).parameters) >= 2
    x = x0
    xs = [x]
    for t in range(int(n_steps)):
        u = policy(x, t) if takes_t else policy(x)
        x = dynamics(x, u)
        xs.append(x)
    if as_np:
        out = np.array(xs)
        return out if out.ndim == 2 else np.swapaxes(out, 0, 1)
    out = torch.stack(xs)
    return out if out.dim() == 2 else out.transpose(0, 1)


def build_test_policy():
    """Open-loop test input of the reference (session4_sol.py:58-62)."""
    acceleration = 1
    return lambda y, t: np.array([acceleration, 0.1 * np.sin(t)])


def compare_open_loop(ts: float, x0, steps: int, params: VehicleParameters = None):
    """Numeric part of the reference's integrator comparison (session4_sol.py:65-104): open-loop
    trajectories under the test policy with forward Euler, RK4 and the accurate integrator, and the
    error norms against the accurate one.  Returns (results, errors) dictionaries keyed as in the
    reference ("Forward Euler", "RK 4", "Ground truth")."""
    bike = KinematicBicycle(params or VehicleParameters())
    schemes = {"Forward Euler": forward_euler(bike, ts), "RK 4": runge_kutta4(bike, ts),
               "Ground truth": exact_integration(bike, ts, adaptive=True)}
    policy = build_test_policy()
    results = {name: simulate(np.asarray(x0, dtype=float), dyn, steps, policy=policy) for name, dyn in schemes.items()}
    errors = {name: np.linalg.norm(results["Ground truth"] - results[name], axis=1) for name in ("Forward Euler", "RK 4")}
    return results, errors
