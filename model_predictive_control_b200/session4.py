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
_lib.register("mpc_rti_workspace_bytes", c_int64, [c_int64, c_int, c_int, c_int])
_lib.register("mpc_rti_closed_loop", c_int,
              [c_double] * 5 + [c_int] + [c_double] * 3 + [c_void_p, c_int, c_int, c_int, c_double] + [c_void_p] * 7 +
              [c_int, c_double, c_double, c_void_p] + [c_void_p] * 15 + [c_int64, c_int64, c_int, c_int, c_double, c_int, c_void_p])

_lib.register("mpc_bicycle_sqp_linesearch", c_int, [c_double] * 5 + [c_int] + [c_void_p] * 5 + [c_int, c_double, c_double] +
              [c_void_p] * 6 + [c_int64, c_int, c_int, c_void_p])
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
            dt = io.pick_dtype(x, u)
            xd, ud = io.to_dev(x, dt), io.to_dev(u, dt)
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
    if u.dtype != x.dtype:
        raise ValueError("x and u must share one dtype")
    fr = friction if io.is_tensor(friction) else torch.full((1,), float(friction), dtype=x.dtype, device=x.device)
    fr = fr.to(device=x.device, dtype=x.dtype).contiguous()
    sfr = 1 if fr.numel() == batch and batch > 1 else (1 if fr.numel() == batch else 0)
    if fr.numel() not in (1, batch):
        raise ValueError("friction must be a scalar or one value per scenario")
    if fr.numel() == 1:
        sfr = 0
    xn = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().mpc_bicycle_plant_step(params.axis_rear, params.axis_front, params.acceleration, float(ts),
                                                     _lib.ptr(fr), sfr, int(substeps), _lib.ptr(x.contiguous()),
                                                     _lib.ptr(u.contiguous()), _lib.ptr(xn), batch, _lib.dtype_enum(x),
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
    clearance: torch.Tensor = None  # [batch] obstacle controller: min |c_i - o_j|^2 - (2r)^2 along the loop

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

    ``sqp_iters``: linearise-and-solve rounds per call / control step.  1 (default) is the real-time
    iteration; k > 1 re-linearises at the new plan k - 1 more times (full-step SQP), which converges to the
    solution of the nonlinear OCP that the reference's IPOPT call returns (session4_sol.py:126-130);
    ``sqp_tol`` > 0 ends the rounds of a scenario once its plan moves by less than sqp_tol * max(1, |U|).
    ``swap_state_bounds``: reproduce template.py:132-133, which lists the state bounds in the order
    [x, y, vel, heading] against the state order [x, y, heading, vel] (default: the corrected order of
    session4_sol.py:176-177).  ``dtype``: torch.float64 (reference arithmetic) or torch.float32 arrays.
    """

    fusable_loop = True
    _nc = 0

    def __init__(self, N: int, ts: float, *, params: VehicleParameters, integrator: str = "euler",
                 terminal_scale: float = 10.0, max_iter: int = 60, eps: float = 1e-9, sqp_iters: int = 1,
                 sqp_tol: float = 0.0, swap_state_bounds: bool = False, dtype=torch.float64):
        if integrator not in ("euler", "rk4"):
            raise ValueError("integrator must be 'euler' or 'rk4'")
        if int(sqp_iters) < 1 or sqp_tol < 0:
            raise ValueError("sqp_iters must be >= 1 and sqp_tol >= 0")
        if dtype not in (torch.float64, torch.float32):
            raise ValueError("dtype must be torch.float64 or torch.float32")
        self.N, self.ts = int(N), float(ts)
        self.params = params
        self.integrator = integrator
        self.max_iter, self.eps = max_iter, eps
        self.sqp_iters, self.sqp_tol = int(sqp_iters), float(sqp_tol)
        self.swap_state_bounds = bool(swap_state_bounds)
        self.dtype = dtype
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
        if self.swap_state_bounds:   # template.py:132-133: [x, y, vel, heading] applied to [x, y, heading, vel]
            s_lb, s_ub = s_lb[[0, 1, 3, 2]], s_ub[[0, 1, 3, 2]]
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

    def _obstacle_args(self):
        return 0, 0.0, 0.0, None

    def _prepare(self, yT, first, bufs):
        warm, A, B, c = bufs[:4]
        dev = yT.device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mpc_bicycle_rti_prepare(*self._model_args(), _lib.ptr(yT), _lib.ptr(self._plan),
                                                          1 if first else 0, _lib.ptr(warm), _lib.ptr(A), _lib.ptr(B),
                                                          _lib.ptr(c), yT.shape[1], self.N, _lib.dtype_enum(yT),
                                                          _lib.stream(dev)))
        return {}

    _lin_sizes = (2, 16, 8, 4)

    def _solve_dev(self, yT):
        """yT [4, batch] on the device -> BoxQpResult (aliases the controller's workspace)."""
        batch, N = yT.shape[1], self.N
        dev, dt = yT.device, yT.dtype
        first = (self._plan is None or self._plan.shape[2] != batch or self._plan.device != dev or self._plan.dtype != dt)
        if first:
            self._plan = torch.zeros((N, 2, batch), dtype=dt, device=dev)
            self._lin = [torch.empty((N, k, batch), dtype=dt, device=dev) for k in self._lin_sizes]
            self._qp_ws = boxqp.BoxQpWorkspace(batch, 4, 2, N, dev, nc=self._nc, dtype=dt)
        warm, A, B, c = self._lin[:4]
        i_lb, i_ub, s_lb, s_ub = self._boxes
        Q, R, QT = (torch.as_tensor(M, dtype=dt, device=dev) for M in (self.Q, self.R, self.QT))
        res = None
        for rnd in range(self.sqp_iters):
            rows = self._prepare(yT, first or rnd > 0, self._lin)
            res = boxqp.solve(A, B, Q, R, QT, N, yT, i_lb, i_ub, s_lb, s_ub, c=c, warm_U=warm, max_iter=self.max_iter,
                              eps=self.eps, workspace=self._qp_ws, **rows)
            if self.sqp_iters > 1:
                # globalised round: backtracking on the l1 merit of the nonlinear OCP (the fused loop does the same
                # inside its kernel); full Gauss-Newton steps cycle on this OCP
                nc, length, width, xo = self._obstacle_args()
                xlo, xhi = (torch.as_tensor(v, dtype=dt, device=dev).contiguous() for v in (s_lb, s_ub))
                with torch.cuda.device(dev):
                    _lib.check(_lib.lib().mpc_bicycle_sqp_linesearch(
                        *self._model_args(), _lib.ptr(Q), _lib.ptr(R), _lib.ptr(QT), _lib.ptr(xlo), _lib.ptr(xhi), nc, length,
                        width, xo, _lib.ptr(yT), _lib.ptr(warm), _lib.ptr(res.U), _lib.ptr(res.status), None, batch, N,
                        _lib.dtype_enum(yT), _lib.stream(dev)))
            moved = None
            if self.sqp_iters > 1 and rnd + 1 < self.sqp_iters:
                moved = float(((res.U - warm).abs().amax(dim=(0, 1)) / res.U.abs().amax(dim=(0, 1)).clamp(min=1.0)).max())
            self._plan.copy_(res.U)
            if moved is not None and moved <= self.sqp_tol:
                break
        return res

    def solve(self, x) -> dict:
        """{"x": stacked inputs}: (2N,) for one state, (batch, 2N) for a batch -- the layout of the
        reference's ``sol["x"]`` (u_0, u_1, ... concatenated, session4_sol.py:206)."""
        as_np = not io.is_tensor(x)
        xd = io.to_dev(x, self.dtype)
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

    # -- fused closed loop: `steps` x (sqp_iters x (prepare, QP), plant) in one kernel
    def closed_loop(self, x0, n_steps, plant: _Discrete = None, friction_plant=None,
                    keep_predictions: bool = False, order="auto") -> RtiClosedLoopResult:
        """Closed loop of this controller against a bicycle plant in ONE kernel launch.  ``plant``: dynamics from
        forward_euler / runge_kutta4 / exact_integration over a :class:`KinematicBicycle` (default: RK4 x 4 sub-steps of
        the controller's own parameters); its axle distances, acceleration gain and friction are the PLANT's, the
        controller predicts with its own ``params`` (the reference's mismatch study, session4_sol.py:461-465).
        ``friction_plant`` overrides the plant friction per scenario.  ``order``: "auto" (batches of at least
        ``boxqp.ORDER_MIN_BATCH`` scenarios run in Morton order of their initial states), None, or an int32 permutation."""
        dt = self.dtype
        xd = io.to_dev(x0, dt)
        if xd.dim() == 1:
            xd = xd[None, :]
        x0T = xd.t().contiguous()
        batch, N, dev = x0T.shape[1], self.N, x0T.device
        if plant is not None and not (isinstance(plant, _Discrete) and plant.fusable):
            raise ValueError("closed_loop fuses bicycle plants only (forward_euler / runge_kutta4 / exact_integration of a "
                             "KinematicBicycle); drive other dynamics with simulate(x0, dynamics, n_steps, policy=controller)")
        pp = self.params if plant is None else plant.f.params
        substeps = 4 if plant is None else (0 if plant.kind == "euler" else (plant.substeps if plant.substeps < 0 else max(plant.substeps, 1)))
        if plant is not None and abs(plant.ts - self.ts) > 1e-15:
            raise ValueError("plant and controller sampling times differ")
        if friction_plant is None:
            fr = torch.full((batch,), float(pp.friction), dtype=dt, device=dev)
        else:
            fr = io.to_dev(friction_plant, dt).reshape(-1).expand(batch).contiguous() if not io.is_tensor(friction_plant) \
                else friction_plant.to(device=dev, dtype=dt).reshape(-1).expand(batch).contiguous()
        # scenarios ordered along the Morton curve of their initial states: neighbours stay neighbours along the closed
        # loop, so the lanes of a warp need similar iteration counts at every control step (results do not depend on
        # the order; they are returned in the caller's order)
        inv = None
        if order == "auto":
            order = boxqp.state_order(x0T) if batch >= boxqp.ORDER_MIN_BATCH else None
        if order is not None:
            ol = order.long()
            x0T, fr = x0T[:, ol].contiguous(), fr[ol].contiguous()
            inv = torch.empty_like(ol)
            inv[ol] = torch.arange(batch, device=dev)
        plan = torch.zeros((N, 2, batch), dtype=dt, device=dev)
        Xp = torch.empty((N + 1, 4, batch), dtype=dt, device=dev)
        Xc = torch.empty((n_steps + 1, 4, batch), dtype=dt, device=dev)
        Uc = torch.empty((n_steps, 2, batch), dtype=dt, device=dev)
        cost = torch.empty(batch, dtype=dt, device=dev)
        viol = torch.empty(batch, dtype=dt, device=dev)
        nc, length, width, xo = self._obstacle_args()
        clear = torch.empty(batch, dtype=dt, device=dev) if nc else None
        ints = [torch.empty(batch, dtype=torch.int32, device=dev) for _ in range(4)]
        Xb = torch.empty((n_steps, N + 1, 4, batch), dtype=dt, device=dev) if keep_predictions else None
        Ub = torch.empty((n_steps, N, 2, batch), dtype=dt, device=dev) if keep_predictions else None
        en = _lib.MPC_F64 if dt == torch.float64 else _lib.MPC_F32
        nbytes = _lib.lib().mpc_rti_workspace_bytes(batch, N, nc, en)
        ws = torch.empty(max(nbytes // 8, 1) + 2, dtype=torch.float64, device=dev)
        i_lb, i_ub, s_lb, s_ub = (torch.as_tensor(v, dtype=dt, device=dev).contiguous() for v in self._boxes)
        Q, R, QT = (torch.as_tensor(M, dtype=dt, device=dev).contiguous() for M in (self.Q, self.R, self.QT))
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mpc_rti_closed_loop(
                *self._model_args(), float(pp.axis_rear), float(pp.axis_front), float(pp.acceleration), _lib.ptr(fr),
                int(substeps), int(n_steps), self.sqp_iters, self.sqp_tol, _lib.ptr(Q), _lib.ptr(R), _lib.ptr(QT),
                _lib.ptr(i_lb), _lib.ptr(i_ub), _lib.ptr(s_lb), _lib.ptr(s_ub), nc, length, width, xo, _lib.ptr(x0T),
                _lib.ptr(plan), _lib.ptr(Xp), _lib.ptr(Xc), _lib.ptr(Uc), _lib.ptr(cost), _lib.ptr(viol), _lib.ptr(clear),
                *[_lib.ptr(t) for t in ints], _lib.ptr(Xb), _lib.ptr(Ub), _lib.ptr(ws), nbytes, batch, N, int(self.max_iter),
                float(self.eps), en, _lib.stream(dev)))
        if inv is not None:
            back = lambda t: None if t is None else t.index_select(t.dim() - 1, inv)
            plan, Xc, Uc, cost, viol, clear, Xb, Ub = (back(t) for t in (plan, Xc, Uc, cost, viol, clear, Xb, Ub))
            ints = [back(t) for t in ints]
        self._plan = plan
        return RtiClosedLoopResult(Xc, Uc, cost, viol, ints[0], ints[1], ints[2], ints[3], Xb, Ub, clear)


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
    polytopic stage constraints on the GPU (K4 with general rows); ``closed_loop`` /
    ``simulate(x0, exact_integration(bicycle, ts), n_steps, policy=controller)`` (main.py:269-271) run the
    whole loop in one kernel."""

    _nc = 9
    _lin_sizes = (2, 16, 8, 4, 36, 9)

    def __init__(self, N: int, ts: float, params: VehicleParameters, model=None, x_obs=None, max_iter: int = 60,
                 eps: float = 1e-9, sqp_iters: int = 1, sqp_tol: float = 0.0, dtype=torch.float64):
        if x_obs is None:
            raise ValueError("x_obs (pose of the parked obstacle) is required")
        super().__init__(N, ts, params=params, integrator="euler", max_iter=max_iter, eps=eps, sqp_iters=sqp_iters,
                         sqp_tol=sqp_tol, dtype=dtype)
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

    def _obstacle_args(self):
        p = self.params
        return 9, float(p.length), float(p.width), (c_double * 4)(*[float(v) for v in self.x_obs[:4]])

    def _prepare(self, yT, first, bufs):
        warm, A, B, c, Cg, hg = bufs
        dev = yT.device
        _, length, width, xo = self._obstacle_args()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mpc_bicycle_rti_prepare_obstacle(
                *self._model_args(), length, width, xo, _lib.ptr(yT), _lib.ptr(self._plan),
                1 if first else 0, _lib.ptr(warm), _lib.ptr(A), _lib.ptr(B), _lib.ptr(c), _lib.ptr(Cg), _lib.ptr(hg),
                yT.shape[1], self.N, _lib.dtype_enum(yT), _lib.stream(dev)))
        return {"Cg": Cg, "hg": hg}


# ------------------------------------------------------------------------------------------------
# closed-loop driver (the reference imports it from rcracers.simulator)
# ------------------------------------------------------------------------------------------------
def _policy_caller(policy):
    """How the course simulator (rcracers.simulator.simulate; call sites session4_sol.py:361,366,407,416,458,465,
    template.py:391,396, main.py:270) hands arguments to a policy: by PARAMETER NAME.  A policy may name any of ``y``
    (the measured state), ``t`` (the step index) and ``log`` (the controller log of sessions 2/3,
    session_2/log.py:8-12) and receives exactly those: ``open_loop_policy(t)`` (session4_sol.py:353-354) gets the step
    index, ``lambda y, t:`` (:61) both, ``MPCController.__call__(y)`` (:222) the state, a sessions-2/3 controller
    ``__call__(y, log)`` the state and the log.  Parameters with other names are filled positionally from
    (y, t) so that ``lambda x: ...`` and ``lambda x, k: ...`` keep working."""
    import inspect
    try:
        params = [q for q in inspect.signature(policy).parameters.values()
                  if q.kind in (q.POSITIONAL_ONLY, q.POSITIONAL_OR_KEYWORD, q.KEYWORD_ONLY)]
    except (TypeError, ValueError):   # builtins / C callables: call with the measurement only
        return lambda y, t, log: policy(y)
    names = [q.name for q in params]
    known = {"y", "t", "log"}
    if names and all(nm in known for nm in names):
        def call(y, t, log):
            kw = {"y": y, "t": t, "log": log}
            return policy(**{nm: kw[nm] for nm in names})
        return call
    # unknown names: (state, step index) in order; a parameter called `log` still receives the log
    free = [q for q in params if q.name != "log" and q.kind != q.KEYWORD_ONLY and q.default is q.empty]
    takes_log = "log" in names

    def call(y, t, log):
        args = (y, t)[:max(1, min(2, len(free)))]
        return policy(*args, log=log) if takes_log else policy(*args)
    return call


def _same_bicycle(p, q):
    return all(float(getattr(p, k)) == float(getattr(q, k)) for k in ("axis_rear", "axis_front", "acceleration"))


def simulate(x0, dynamics: Callable, n_steps: int, policy=None, friction_plant=None, log=None):
    """States of the closed loop: (n_steps+1, n) for one x0, (batch, n_steps+1, n) for a batch -- the call
    ``simulate(x0, dynamics, n_steps, policy=..., log=...)`` the reference imports from rcracers.
    With an :class:`MPCController` policy and bicycle dynamics from one of the integrator factories
    the whole loop is one fused kernel (prediction model = the controller's parameters, plant = the
    dynamics' parameters); any other callables run step by step.  The policy receives ``y`` / ``t`` /
    ``log`` according to its parameter NAMES (see :func:`_policy_caller`)."""
    as_np = not io.is_tensor(x0)
    if (isinstance(policy, MPCController) and policy.fusable_loop and log is None
            and isinstance(dynamics, _Discrete) and dynamics.fusable):
        policy.reset()
        res = policy.closed_loop(x0, int(n_steps), plant=dynamics, friction_plant=friction_plant)
        X = res.states
        single = (np.ndim(x0) if as_np else x0.dim()) == 1
        return io.back(X[0] if single else X, as_np)
    if policy is None:
        raise ValueError("simulate needs a policy")
    if isinstance(policy, MPCController):
        policy.reset()
    call = _policy_caller(policy)
    x = x0
    xs = [x]
    for t in range(int(n_steps)):
        u = call(x, t, log)
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
