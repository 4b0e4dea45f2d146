"""CPU oracle for the session-4 path (TEST INFRASTRUCTURE): kinematic bicycle, integrators, the
real-time-iteration (RTI) step and the closed-loop driver.

OCP DATA AND INTEGRATORS PINNED, ODE AND SOLVER PARITY UNPINNED BY THE REFERENCE.  The reference
solves the nonlinear OCP with CasADi + IPOPT (/root/reference/session_4/session4_sol.py:113-230) and
takes the vehicle model from ``rcracers`` (``KinematicBicycle``, call sites session4_sol.py:11,74,191,452);
neither dependency is vendored, pinned or installed, so the converged NLP solution and the ODE cannot be
checked against reference outputs.  Everything else is: tests/golden/session234.json holds outputs of
the reference's own parameters.py, integrators (session4_sol.py:22-56) and of MPCController.build_ocp of
session4_sol.py and main.py evaluated numerically (cost, constraint vector, bounds) with this module's
ODE plugged in for rcracers (oracle/ref_loader.py::load_session4, NumericCasadi), and
tests/test_oracle_golden_s234.py checks this restatement against them.  What follows the
reference: state order [p_x, p_y, psi, v] and input order [a, delta] (session4_sol.py:176-181),
weights Q = diag(1, 3, .1, .01), Q_T = 10 Q, R = diag(1, .01) (:166-169), bounds from
VehicleParameters (:176-181, parameters.py:17-29), N = 50, ts = 0.05, x0 = [.6, -.25, 0, 0]
(:445-447), Euler / RK4 integrators (:22-34), plant mismatch friction *= 0.8 (:461-463).

OUR DEFINITION of the bicycle ODE (an assumption, the rcracers source is not available):
    beta = atan(l_r tan(delta) / (l_r + l_f))
    p_x' = v cos(psi + beta),  p_y' = v sin(psi + beta),  psi' = v sin(beta) / l_r,
    v'   = acceleration * a - friction * v
with l_r = axis_rear, l_f = axis_front, friction, acceleration from VehicleParameters
(parameters.py:7-8,47-48).

The north star replaces the converged IPOPT solve by ONE linearised QP per control step (RTI):
shift the previous input plan, roll the nonlinear model out from the measured state, linearise
along that trajectory, solve the LTV box QP, apply u_0.  The restatement here is that algorithm;
the QP is solved by oracle.boxqp (exact active-set solver or the numpy port of the GPU method).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import boxqp as bq


@dataclass
class VehicleParameters:
    """Subset of /root/reference/session_4/parameters.py:4-54 that the kinematic model uses."""
    axis_front: float = 0.047
    axis_rear: float = 0.05
    max_steer: float = 0.384
    max_drive: float = 1.0
    min_drive: float = -1.0
    min_pos_x: float = -3.0
    max_pos_x: float = 3.0
    min_pos_y: float = -2.0
    max_pos_y: float = 2.0
    min_vel: float = -0.5
    max_vel: float = 0.5
    max_heading: float = 2 * np.pi
    min_heading: float = -2 * np.pi
    friction: float = 1
    acceleration: float = 2


def weights(variant="sol"):
    """(Q, Q_T, R): session4_sol.py:166-169; template.py:135-137 (Q_N = 5 Q)."""
    Q = np.diag([1.0, 3.0, 0.1, 0.01])
    R = np.diag([1.0, 1e-2])
    return Q, (10.0 if variant == "sol" else 5.0) * Q, R


def bounds(p: VehicleParameters):
    """Input and state boxes (session4_sol.py:176-181)."""
    return (np.array([p.min_drive, -p.max_steer]), np.array([p.max_drive, p.max_steer]),
            np.array([p.min_pos_x, p.min_pos_y, p.min_heading, p.min_vel]),
            np.array([p.max_pos_x, p.max_pos_y, p.max_heading, p.max_vel]))


# ------------------------------------------------------------------------------------------------
# model (batched over the leading axis: x [batch, 4], u [batch, 2], friction scalar or [batch])
# ------------------------------------------------------------------------------------------------
def bicycle_f(x, u, lr, lf, friction, accel):
    psi, v = x[..., 2], x[..., 3]
    a, delta = u[..., 0], u[..., 1]
    beta = np.arctan(lr * np.tan(delta) / (lr + lf))
    return np.stack([v * np.cos(psi + beta), v * np.sin(psi + beta), v * np.sin(beta) / lr,
                     accel * a - friction * v], axis=-1)


def bicycle_jac(x, u, lr, lf, friction, accel):
    """(f, df/dx [...,4,4], df/du [...,4,2])."""
    psi, v = x[..., 2], x[..., 3]
    delta = u[..., 1]
    kap = lr / (lr + lf)
    td = np.tan(delta)
    beta = np.arctan(kap * td)
    dbeta = kap * (1 + td * td) / (1 + kap * kap * td * td)
    c, s = np.cos(psi + beta), np.sin(psi + beta)
    Jx = np.zeros(x.shape[:-1] + (4, 4)); Ju = np.zeros(x.shape[:-1] + (4, 2))
    Jx[..., 0, 2] = -v * s; Jx[..., 0, 3] = c
    Jx[..., 1, 2] = v * c; Jx[..., 1, 3] = s
    Jx[..., 2, 3] = np.sin(beta) / lr
    Jx[..., 3, 3] = -friction
    Ju[..., 0, 1] = -v * s * dbeta
    Ju[..., 1, 1] = v * c * dbeta
    Ju[..., 2, 1] = v * np.cos(beta) * dbeta / lr
    Ju[..., 3, 0] = accel
    return bicycle_f(x, u, lr, lf, friction, accel), Jx, Ju


def forward_euler(f, ts):
    """session4_sol.py:22-25."""
    return lambda x, u: x + f(x, u) * ts


def runge_kutta4(f, ts):
    """session4_sol.py:27-34."""
    def rk4(x, u):
        s1 = f(x, u)
        s2 = f(x + 0.5 * ts * s1, u)
        s3 = f(x + 0.5 * ts * s2, u)
        s4 = f(x + ts * s3, u)
        return x + ts / 6.0 * (s1 + 2 * s2 + 2 * s3 + s4)
    return rk4


def discretize(x, u, ts, par, friction, method="euler"):
    """One step x+ = f_d(x, u) with its Jacobians A = d f_d/dx, B = d f_d/du (exact chain rule)."""
    lr, lf, acc = par.axis_rear, par.axis_front, par.acceleration
    I = np.eye(4)
    if method == "euler":
        f, Jx, Ju = bicycle_jac(x, u, lr, lf, friction, acc)
        return x + ts * f, I + ts * Jx, ts * Ju
    k1, J1x, J1u = bicycle_jac(x, u, lr, lf, friction, acc)
    k2, J2x, J2u = bicycle_jac(x + 0.5 * ts * k1, u, lr, lf, friction, acc)
    D2x = J2x @ (I + 0.5 * ts * J1x); D2u = J2x @ (0.5 * ts * J1u) + J2u
    k3, J3x, J3u = bicycle_jac(x + 0.5 * ts * k2, u, lr, lf, friction, acc)
    D3x = J3x @ (I + 0.5 * ts * D2x); D3u = J3x @ (0.5 * ts * D2u) + J3u
    k4, J4x, J4u = bicycle_jac(x + ts * k3, u, lr, lf, friction, acc)
    D4x = J4x @ (I + ts * D3x); D4u = J4x @ (ts * D3u) + J4u
    xn = x + ts / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    return xn, I + ts / 6.0 * (J1x + 2 * D2x + 2 * D3x + D4x), ts / 6.0 * (J1u + 2 * D2u + 2 * D3u + D4u)


def exact_integration_odeint(x, u, ts, par, friction):
    """The reference's own accurate plant: scipy odeint over [0, ts] (session4_sol.py:37-56), one state at a time."""
    from scipy.integrate import odeint
    lr, lf, acc = par.axis_rear, par.axis_front, par.acceleration
    x = np.atleast_2d(np.asarray(x, float)); u = np.atleast_2d(np.asarray(u, float))
    fr = np.broadcast_to(np.asarray(friction, float), (x.shape[0],))
    out = np.zeros_like(x)
    for b in range(x.shape[0]):
        f_wrap = lambda xx, t: bicycle_f(xx, u[b], lr, lf, fr[b], acc)
        out[b] = odeint(f_wrap, x[b], [0, ts], rtol=1e-12, atol=1e-12)[-1]
    return out


def plant_step(x, u, ts, par, friction, method="rk4", substeps=4):
    """Plant x+ : Euler (nominal model, session4_sol.py:453) or RK4 with fixed sub-steps (stands in for
    the reference's odeint plant, session4_sol.py:37-56 -- documented deviation)."""
    return _plant(x, u, ts, par, friction, method, substeps)


# ------------------------------------------------------------------------------------------------
# RTI
# ------------------------------------------------------------------------------------------------
def rti_prepare(y, U_prev, ts, par, friction, method="euler", first=False):
    """Shift the previous plan, roll out, linearise.  y [batch,4], U_prev [N,batch,2].
    Returns Ubar [N,batch,2], A [N,batch,4,4], B [N,batch,4,2], c [N,batch,4], Xbar [N+1,batch,4]."""
    N = U_prev.shape[0]
    Ubar = U_prev.copy() if first else np.concatenate([U_prev[1:], U_prev[-1:]], axis=0)
    A, B, c, X = [], [], [], [y]
    x = y
    for k in range(N):
        xn, Ak, Bk = discretize(x, Ubar[k], ts, par, friction, method)
        A.append(Ak); B.append(Bk)
        c.append(xn - np.einsum("bij,bj->bi", Ak, x) - np.einsum("bij,bj->bi", Bk, Ubar[k]))
        X.append(xn)
        x = xn
    return Ubar, np.array(A), np.array(B), np.array(c), np.array(X)


def closed_loop(x0, n_steps, N=50, ts=0.05, par=None, friction_model=None, friction_plant=None, ocp_method="euler",
                plant_method="rk4", substeps=4, variant="sol", qp="port", max_iter=60, sqp_iters=1, sqp_tol=0.0,
                plant_par=None, keep_plans=False):
    """Batched RTI closed loop.  x0 [batch,4].  qp = "port" (numpy restatement of the GPU interior
    point method, batched) or "exact" (HiGHS active set + KKT refinement, one scenario at a time).
    sqp_iters > 1: re-linearise at the new plan and solve again (full-step SQP), per control step; a scenario whose
    plan moved by <= sqp_tol * max(1, |U|) stops early.  plant_par: parameters of the plant (default: the model's).
    Returns dict(X [steps+1,batch,4], U [steps,batch,2], status [steps,batch], cost [batch], viol [batch])."""
    par = par or VehicleParameters()
    pp = plant_par or par
    x = np.atleast_2d(np.asarray(x0, float))
    batch = x.shape[0]
    fm = par.friction if friction_model is None else friction_model
    fp = np.full(batch, pp.friction, float) if friction_plant is None else np.broadcast_to(np.asarray(friction_plant, float), (batch,))
    Q, QT, R = weights(variant)
    ulo, uhi, xlo, xhi = bounds(par)
    U_prev = np.zeros((N, batch, 2))
    Xs, Us, Ss, Ps = [x], [], [], []
    cost = np.zeros(batch); viol = np.zeros(batch)
    for t in range(n_steps):
        live = np.ones(batch, dtype=bool)
        for rnd in range(sqp_iters):
            Ubar, A, B, c, _ = rti_prepare(x, U_prev, ts, par, fm, ocp_method, first=(t == 0 or rnd > 0))
            if qp == "port":
                r = bq.ipm_riccati(list(A), list(B), Q, R, QT, N, x, ulo, uhi, xlo, xhi, c=list(c), warm_U=Ubar, max_iter=max_iter)
                U, status = r["U"], r["status"]
            else:
                U = np.zeros((N, batch, 2)); status = np.zeros(batch, dtype=np.int32)
                for b in range(batch):
                    e = bq.solve_exact(A[:, b], B[:, b], Q, R, QT, N, x[b], ulo, uhi, xlo, xhi, c=c[:, b])
                    U[:, b], status[b] = e["U"], e["status"]
            if sqp_iters > 1:
                # globalised SQP round (csrc/bicycle_core.cuh, rti_merit): backtracking on the l1 merit of the
                # nonlinear OCP along Ubar -> U_qp, beta = 1, 1/2, .., 2^-10, else 0; failed QPs keep the full step
                for b in range(batch):
                    if not live[b] or status[b] != bq.SOLVED:
                        continue
                    m0 = _merit(x[b], Ubar[:, b], ts, par, fm, ocp_method, Q, QT, R, xlo, xhi)
                    beta, h = 1.0, 0
                    while h <= 10 and not (_merit(x[b], Ubar[:, b] + beta * (U[:, b] - Ubar[:, b]), ts, par, fm, ocp_method,
                                                  Q, QT, R, xlo, xhi) < m0):
                        beta *= 0.5; h += 1
                    if h > 10:
                        beta = 0.0
                    U[:, b] = Ubar[:, b] + beta * (U[:, b] - Ubar[:, b])
            U = np.where(live[None, :, None], U, U_prev)   # scenarios that already met sqp_tol keep their plan
            if sqp_iters > 1:
                du = np.abs(U - Ubar).max(axis=(0, 2)); un = np.maximum(1.0, np.abs(U).max(axis=(0, 2)))
                live = live & ~(du <= sqp_tol * un)
            U_prev = U
            if not live.any():
                break
        u0 = U[0]
        cost += np.einsum("bi,ij,bj->b", x, Q, x) + np.einsum("bi,ij,bj->b", u0, R, u0)
        x = _plant(x, u0, ts, pp, fp, plant_method, substeps)
        viol = np.maximum(viol, np.maximum(xlo - x, x - xhi).max(axis=1).clip(min=0))
        Xs.append(x); Us.append(u0); Ss.append(status)
        if keep_plans:
            Ps.append(U.copy())
    out = {"X": np.array(Xs), "U": np.array(Us), "status": np.array(Ss), "cost": cost, "viol": viol}
    if keep_plans:
        out["plans"] = np.array(Ps)
    return out


SQP_MERIT_RHO = 100.0


def _merit(x0, U, ts, par, friction, method, Q, QT, R, xlo, xhi, x_obs=None, length=0.17, width=0.08):
    """l1 merit of the nonlinear OCP for one scenario: rollout cost + rho * summed state-box (and collision) violations."""
    x = np.asarray(x0, float); cost = 0.0; viol = 0.0
    lr, lf, acc = par.axis_rear, par.axis_front, par.acceleration
    f = lambda xx, uu: bicycle_f(xx, uu, lr, lf, friction, acc)
    step = forward_euler(f, ts) if method == "euler" else runge_kutta4(f, ts)
    if x_obs is not None:
        a, r = create_cover_circles(length, width, 3)
        r2 = (2 * r) ** 2
        ox = x_obs[0] + a * np.cos(x_obs[2]); oy = x_obs[1] + a * np.sin(x_obs[2])
    for k in range(U.shape[0]):
        cost += x @ Q @ x + U[k] @ R @ U[k]
        x = step(x, U[k])
        viol += np.maximum(0.0, np.maximum(xlo - x, x - xhi)).sum()
        if x_obs is not None:
            cx = x[0] + a * np.cos(x[2]); cy = x[1] + a * np.sin(x[2])
            g = r2 - ((cx[:, None] - ox[None]) ** 2 + (cy[:, None] - oy[None]) ** 2)
            viol += np.maximum(0.0, g).sum()
    return cost + x @ QT @ x + SQP_MERIT_RHO * viol


def _plant(x, u, ts, par, friction, method, substeps):
    lr, lf, acc = par.axis_rear, par.axis_front, par.acceleration
    f = lambda xx, uu: bicycle_f(xx, uu, lr, lf, friction, acc)
    if method == "euler":
        return x + ts * f(x, u)
    h = ts / substeps
    for _ in range(substeps):
        s1 = f(x, u); s2 = f(x + 0.5 * h * s1, u); s3 = f(x + 0.5 * h * s2, u); s4 = f(x + h * s3, u)
        x = x + h / 6.0 * (s1 + 2 * s2 + 2 * s3 + s4)
    return x


# ------------------------------------------------------------------------------------------------
# obstacle-avoidance variant (reference session_4/main.py:29-129)
# ------------------------------------------------------------------------------------------------
def create_cover_circles(l, w, n_c=3):
    """Offsets of the n_c covering-circle centres along the vehicle axis and their radius
    (main.py:191-200: d = l/(2 n_c), r = sqrt(d^2 + w^2/4), centres ((2k+1) d - l/2, 0))."""
    d = l / (2 * n_c)
    return np.array([(2 * k + 1) * d - l / 2 for k in range(n_c)]), float(np.sqrt(d * d + w * w / 4))


def obstacle_weights():
    """main.py:72-74: Q = diag(1, 6, .2, .05), Q_N = 100 Q, R = diag(1, .01)."""
    Q = np.diag([1.0, 6.0, 0.2, 0.05])
    return Q, 100.0 * Q, np.diag([1.0, 0.01])


def obstacle_rows(xbar, x_obs, length=0.17, width=0.08, n_c=3):
    """Linearisation at xbar [batch,4] of the n_c^2 collision constraints of main.py:95-104,
    g_ij(x) = |c_i(x) - o_j|^2 >= (r + r_p)^2 with c_i(x) = p + a_i (cos psi, sin psi):
    rows C x >= h with C = grad g(xbar), h = (r + r_p)^2 - g(xbar) + C xbar.  Returns C [batch, n_c^2, 4], h."""
    a, r = create_cover_circles(length, width, n_c)
    r2 = (2 * r) ** 2
    x_obs = np.asarray(x_obs, float)
    ox = x_obs[0] + a * np.cos(x_obs[2]); oy = x_obs[1] + a * np.sin(x_obs[2])
    px, py, psi = xbar[:, 0], xbar[:, 1], xbar[:, 2]
    C = np.zeros((xbar.shape[0], n_c * n_c, 4)); h = np.zeros((xbar.shape[0], n_c * n_c))
    for i in range(n_c):
        cx, cy = px + a[i] * np.cos(psi), py + a[i] * np.sin(psi)
        for j in range(n_c):
            dx, dy = cx - ox[j], cy - oy[j]
            g = dx * dx + dy * dy
            row = i * n_c + j
            C[:, row, 0] = 2 * dx
            C[:, row, 1] = 2 * dy
            C[:, row, 2] = 2 * dx * (-a[i] * np.sin(psi)) + 2 * dy * (a[i] * np.cos(psi))
            h[:, row] = r2 - g + np.einsum("bi,bi->b", C[:, row], xbar)
    return C, h


def closed_loop_obstacle(x0, x_obs, n_steps, N=30, ts=0.08, par=None, friction_plant=None, plant_method="rk4",
                         substeps=4, qp="port", max_iter=60, length=0.17, width=0.08):
    """RTI closed loop of the obstacle-avoidance controller (main.py:241-271 protocol): Euler prediction
    model, box bounds as in session4_sol, plus the linearised collision rows on every predicted state."""
    par = par or VehicleParameters()
    x = np.atleast_2d(np.asarray(x0, float))
    batch = x.shape[0]
    fp = np.full(batch, par.friction, float) if friction_plant is None else np.broadcast_to(np.asarray(friction_plant, float), (batch,))
    Q, QT, R = obstacle_weights()
    ulo, uhi, xlo, xhi = bounds(par)
    U_prev = np.zeros((N, batch, 2))
    Xs, Us, Ss = [x], [], []
    for t in range(n_steps):
        Ubar, A, B, c, Xbar = rti_prepare(x, U_prev, ts, par, par.friction, "euler", first=(t == 0))
        rows = [obstacle_rows(Xbar[k + 1], x_obs, length, width) for k in range(N)]
        Cg = np.array([r_[0] for r_ in rows]); hg = np.array([r_[1] for r_ in rows])
        if qp == "port":
            r = bq.ipm_riccati(list(A), list(B), Q, R, QT, N, x, ulo, uhi, xlo, xhi, c=list(c), warm_U=Ubar, max_iter=max_iter,
                               Cg=Cg, hg=hg)
            U, status = r["U"], r["status"]
        else:
            U = np.zeros((N, batch, 2)); status = np.zeros(batch, dtype=np.int32)
            for b in range(batch):
                e = bq.solve_exact(A[:, b], B[:, b], Q, R, QT, N, x[b], ulo, uhi, xlo, xhi, c=c[:, b], Cg=Cg[:, b], hg=hg[:, b])
                U[:, b], status[b] = e["U"], e["status"]
        u0 = U[0]
        x = _plant(x, u0, ts, par, fp, plant_method, substeps)
        U_prev = U
        Xs.append(x); Us.append(u0); Ss.append(status)
    return {"X": np.array(Xs), "U": np.array(Us), "status": np.array(Ss)}


def min_clearance(X, x_obs, length=0.17, width=0.08, n_c=3):
    """Smallest |c_i - o_j| - 2 r over a trajectory X [..., 4] (negative = overlap of covering circles)."""
    a, r = create_cover_circles(length, width, n_c)
    x_obs = np.asarray(x_obs, float)
    ox = x_obs[0] + a * np.cos(x_obs[2]); oy = x_obs[1] + a * np.sin(x_obs[2])
    cx = X[..., 0, None] + a * np.cos(X[..., 2, None]); cy = X[..., 1, None] + a * np.sin(X[..., 2, None])
    dist = np.sqrt((cx[..., :, None] - ox) ** 2 + (cy[..., :, None] - oy) ** 2)
    return float(dist.min() - 2 * r)


# ------------------------------------------------------------------------------------------------
# converged solution of the nonlinear OCP (what the reference's IPOPT call returns, session4_sol.py:126-130)
# ------------------------------------------------------------------------------------------------
def ocp_functions(x0, N, ts=0.05, par=None, friction=None, method="euler", variant="sol"):
    """The single-shooting NLP of session4_sol.MPCController.build_ocp (:132-217) for one x0 [4]:
    decision U [N,2]; cost f(U) = sum_{k<N} x_k'Qx_k + u_k'Ru_k + x_N'Q_T x_N; constraint vector g(U) = (x_1..x_N)
    (state box); variable bounds = input box.  Returns (fun, con) with fun(U) -> (f, df/dU [2N]) and
    con(U) -> (g [4N], dg/dU [4N, 2N]), derivatives by the exact chain rule through the rollout.
    tests/test_oracle_golden_s234.py pins f and g to the reference's own build_ocp evaluated numerically."""
    par = par or VehicleParameters()
    fr = par.friction if friction is None else friction
    Q, QT, R = weights(variant)
    x0 = np.asarray(x0, float).reshape(4)

    def rollout(U):
        U = np.asarray(U, float).reshape(N, 2)
        X = [x0]; As = []; Bs = []
        for k in range(N):
            xn, A, B = discretize(X[-1][None], U[k][None], ts, par, fr, method)
            X.append(xn[0]); As.append(A[0]); Bs.append(B[0])
        return U, np.array(X), As, Bs

    def sens(As, Bs):
        """S[k] = d x_{k+1} / dU  [4, 2N]"""
        S = np.zeros((N, 4, 2 * N)); cur = np.zeros((4, 2 * N))
        for k in range(N):
            cur = As[k] @ cur
            cur[:, 2 * k:2 * k + 2] += Bs[k]
            S[k] = cur
        return S

    def fun(Uf):
        U, X, As, Bs = rollout(Uf)
        f = float(sum(X[k] @ Q @ X[k] + U[k] @ R @ U[k] for k in range(N)) + X[N] @ QT @ X[N])
        S = sens(As, Bs)
        grad = (2 * U @ R).reshape(-1).copy()
        for k in range(1, N):
            grad += 2 * (Q @ X[k]) @ S[k - 1]
        grad += 2 * (QT @ X[N]) @ S[N - 1]
        return f, grad

    def con(Uf):
        U, X, As, Bs = rollout(Uf)
        return X[1:].reshape(-1), sens(As, Bs).reshape(4 * N, 2 * N)

    return fun, con


def nlp_solve(x0, N=50, ts=0.05, par=None, friction=None, method="euler", variant="sol", U_init=None, tol=1e-13,
              maxiter=500):
    """Converged local solution of the OCP by scipy SLSQP (exact derivatives), from U_init (default: zeros, the
    reference's cold start).  Returns dict(U [N,2], X [N+1,4], f, success, nit, kkt) with kkt = max-norm of the
    projected gradient of the Lagrangian."""
    from scipy.optimize import minimize
    par = par or VehicleParameters()
    fun, con = ocp_functions(x0, N, ts, par, friction, method, variant)
    ulo, uhi, xlo, xhi = bounds(par)
    lbg, ubg = np.tile(xlo, N), np.tile(xhi, N)
    cons = [{"type": "ineq", "fun": lambda U: con(U)[0] - lbg, "jac": lambda U: con(U)[1]},
            {"type": "ineq", "fun": lambda U: ubg - con(U)[0], "jac": lambda U: -con(U)[1]}]
    U0 = np.zeros(2 * N) if U_init is None else np.asarray(U_init, float).reshape(-1)
    bnds = list(zip(np.tile(ulo, N), np.tile(uhi, N)))
    res = minimize(lambda U: fun(U), np.clip(U0, np.tile(ulo, N), np.tile(uhi, N)), jac=True, bounds=bnds, constraints=cons,
                   method="SLSQP", options={"ftol": tol, "maxiter": maxiter})
    U = res.x.reshape(N, 2)
    g, _ = con(res.x)
    return {"U": U, "X": np.vstack([np.asarray(x0, float)[None], g.reshape(N, 4)]), "f": float(res.fun),
            "success": bool(res.success), "nit": int(res.nit), "message": str(res.message)}


def closed_loop_converged(x0, n_steps, N=50, ts=0.05, par=None, friction_plant=None, plant_method="euler", substeps=4,
                          warm=True):
    """Closed loop of ONE scenario with the OCP solved to convergence at every step (the reference's exercise5,
    session4_sol.py:443-465; the reference cold-starts IPOPT, here SLSQP starts from the shifted previous plan when
    ``warm`` -- the OCP is smooth and both reach the same local solution on this task)."""
    par = par or VehicleParameters()
    fp = par.friction if friction_plant is None else friction_plant
    x = np.asarray(x0, float).reshape(4)
    Xs, Us, plans = [x], [], []
    U_prev = None
    for t in range(n_steps):
        init = None if (U_prev is None or not warm) else np.vstack([U_prev[1:], U_prev[-1:]])
        r = nlp_solve(x, N, ts, par, U_init=init)
        U_prev = r["U"]
        u0 = U_prev[0]
        x = _plant(x[None], u0[None], ts, par, np.array([fp]), plant_method, substeps)[0]
        Xs.append(x); Us.append(u0); plans.append(U_prev.copy())
    return {"X": np.array(Xs), "U": np.array(Us), "plans": np.array(plans)}
