"""CPU oracle for the box-constrained linear MPC QP of sessions 2/3 (TEST INFRASTRUCTURE).

DATA PINNED, SOLVER PARITY UNPINNED BY THE REFERENCE: /root/reference ships only the problem *data*
(session_2/problem.py:4-32, session_3/problem.py:8-36) and the per-solve log schema
(session_2/log.py:8-12); the solver file of the course is not in the repository.  The data
(fields, __post_init__ matrices, properties) is pinned to tests/golden/session234.json, produced by
the reference's own problem.py (tests/golden/make_golden.py, oracle/ref_loader.py::load_problem;
tests/test_oracle_golden_s234.py).  The *solution* has no reference output to be pinned to; it is
pinned by agreement of independent exact CPU solvers on the same QP (tests/test_oracle_boxqp.py):
  * HiGHS active-set QP (scipy's bundled ``scipy.optimize._highspy``),
  * scipy SLSQP,
  * our own numpy restatement of the algorithm the GPU runs (``admm_riccati``).

The QP (x_0 given):
    min  sum_{k<N} x_k'Q x_k + u_k'R u_k  +  x_N' P_f x_N
    s.t. x_{k+1} = A_k x_k + B_k u_k + c_k,   u_lo <= u_k <= u_hi (k<N),   x_lo <= x_k <= x_hi (1<=k<=N)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SOLVED, MAX_ITER, INFEASIBLE = 1, 2, 3


# ------------------------------------------------------------------------------------------------
# problem data (restated from the reference)
# ------------------------------------------------------------------------------------------------
@dataclass
class Problem:
    """Restatement of /root/reference/session_2/problem.py:4-32 (same field names, defaults and
    __post_init__); ``default_factory`` replaces the ndarray defaults that Python >= 3.11 rejects."""
    Ts: float = 0.3
    Q: np.ndarray = field(default_factory=lambda: np.diag([10, 1]))
    R: np.ndarray = field(default_factory=lambda: np.diag([0.01]))
    p_min: float = -150
    p_max: float = 1.0
    v_min: float = -20
    v_max: float = 25.0
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


def session3_problem(**kw):
    """/root/reference/session_3/problem.py:15,17: p_min = -120, v_min = -50."""
    return Problem(p_min=-120, v_min=-50, **kw)


def problem_bounds(p):
    return (np.array([p.u_min], float), np.array([p.u_max], float),
            np.array([p.p_min, p.v_min], float), np.array([p.p_max, p.v_max], float))


# ------------------------------------------------------------------------------------------------
# condensed form
# ------------------------------------------------------------------------------------------------
def condense(A, B, Q, R, Pf, N):
    """Prediction matrices of the LTI problem: X = Phi x0 + Gamma U with X = (x_1..x_N),
    H = Gamma' Qbar Gamma + Rbar, F = Gamma' Qbar Phi (Qbar = blkdiag(Q,..,Q,Pf)), so that
    J(U) = U'HU + 2 x0'F'U + const."""
    A, B, Q, R, Pf = (np.asarray(M, float) for M in (A, B, Q, R, Pf))
    n, m = B.shape
    Phi = np.zeros((N * n, n))
    Gam = np.zeros((N * n, N * m))
    Ap = np.eye(n)
    blocks = []
    for i in range(N):
        blocks.append(Ap @ B)          # A^i B
        Ap = A @ Ap
        Phi[i * n:(i + 1) * n] = Ap    # A^(i+1)
    for i in range(N):
        for j in range(i + 1):
            Gam[i * n:(i + 1) * n, j * m:(j + 1) * m] = blocks[i - j]
    Qbar = np.kron(np.eye(N), Q)
    Qbar[-n:, -n:] = Pf
    Rbar = np.kron(np.eye(N), R)
    H = Gam.T @ Qbar @ Gam + Rbar
    F = Gam.T @ Qbar @ Phi
    return Phi, Gam, H, F


def ltv_prediction(A, B, c, x0):
    """X = Phi x0 + Gamma U + g for stage-varying A[k], B[k], c[k]; returns (Phi, Gamma, g)."""
    N, n, m = B.shape
    Phi = np.zeros((N * n, n)); Gam = np.zeros((N * n, N * m)); g = np.zeros(N * n)
    Pk = np.eye(n); gk = np.zeros(n); cols = np.zeros((n, N * m))
    for k in range(N):
        cols = A[k] @ cols
        cols[:, k * m:(k + 1) * m] = B[k]
        Pk = A[k] @ Pk
        gk = A[k] @ gk + c[k]
        Phi[k * n:(k + 1) * n] = Pk
        Gam[k * n:(k + 1) * n] = cols
        g[k * n:(k + 1) * n] = gk
    return Phi, Gam, g


# ------------------------------------------------------------------------------------------------
# exact solvers
# ------------------------------------------------------------------------------------------------
def _highs_qp(H, g, G, lo, hi, lb, ub):
    """min 1/2 z'Hz + g'z  s.t. lo <= G z <= hi, lb <= z <= ub  (HiGHS active-set QP)."""
    import scipy.sparse as sp
    from scipy.optimize._highspy import _core as hc
    nz, nc = H.shape[0], G.shape[0]
    model = hc.HighsModel()
    lp = model.lp_
    lp.num_col_, lp.num_row_ = nz, nc
    lp.col_cost_, lp.col_lower_, lp.col_upper_ = g.copy(), lb.copy(), ub.copy()
    lp.row_lower_, lp.row_upper_ = lo.copy(), hi.copy()
    Gc = sp.csc_matrix(G)
    lp.a_matrix_.format_ = hc.MatrixFormat.kColwise
    lp.a_matrix_.start_ = Gc.indptr.astype(np.int32)
    lp.a_matrix_.index_ = Gc.indices.astype(np.int32)
    lp.a_matrix_.value_ = Gc.data
    Hl = sp.csc_matrix(np.tril(H))
    model.hessian_.dim_ = nz
    model.hessian_.format_ = hc.HessianFormat.kTriangular
    model.hessian_.start_ = Hl.indptr.astype(np.int32)
    model.hessian_.index_ = Hl.indices.astype(np.int32)
    model.hessian_.value_ = Hl.data
    h = hc._Highs()
    h.setOptionValue("output_flag", False)
    h.setOptionValue("primal_feasibility_tolerance", 1e-10)
    h.setOptionValue("dual_feasibility_tolerance", 1e-10)
    h.passModel(model)
    h.run()
    status = h.getModelStatus()
    sol = h.getSolution()
    _highs_qp.last_duals = (np.array(sol.col_dual), np.array(sol.row_dual))
    return np.array(sol.col_value), status == hc.HighsModelStatus.kOptimal, status == hc.HighsModelStatus.kInfeasible


def _condensed_qp(A, B, c, Q, R, Pf, x0, u_lo, u_hi, x_lo, x_hi, Cg=None, hg=None):
    """Condensed QP in U.  Optional general stage rows Cg[k] x_{k+1} >= hg[k] (Cg [N, nc, n], hg [N, nc])
    are appended to the constraint matrix after the state-box rows."""
    N, n, m = B.shape
    Phi, Gam, g = ltv_prediction(A, B, c, x0)
    Qbar = np.kron(np.eye(N), Q); Qbar[-n:, -n:] = Pf
    Rbar = np.kron(np.eye(N), R)
    free = Phi @ x0 + g
    H = 2 * (Gam.T @ Qbar @ Gam + Rbar)
    grad = 2 * Gam.T @ Qbar @ free
    lo = np.tile(x_lo, N) - free
    hi = np.tile(x_hi, N) - free
    G = Gam
    if Cg is not None:
        nc = Cg.shape[1]
        rows = np.zeros((N * nc, Gam.shape[1])); rlo = np.zeros(N * nc)
        for k in range(N):
            rows[k * nc:(k + 1) * nc] = Cg[k] @ Gam[k * n:(k + 1) * n]
            rlo[k * nc:(k + 1) * nc] = hg[k] - Cg[k] @ free[k * n:(k + 1) * n]
        G = np.vstack([Gam, rows]); lo = np.concatenate([lo, rlo]); hi = np.concatenate([hi, np.full(N * nc, np.inf)])
    return H, grad, G, lo, hi, np.tile(u_lo, N), np.tile(u_hi, N), free


def stage_arrays(A, B, N, c=None):
    A, B = np.asarray(A, float), np.asarray(B, float)
    if A.ndim == 2:
        A = np.broadcast_to(A, (N,) + A.shape)
        B = np.broadcast_to(B, (N,) + B.shape)
    if c is None:
        c = np.zeros((N, A.shape[-1]))
    return np.ascontiguousarray(A), np.ascontiguousarray(B), np.ascontiguousarray(c, dtype=float)


def rollout(A, B, c, x0, U):
    X = [np.asarray(x0, float)]
    for k in range(B.shape[0]):
        X.append(A[k] @ X[-1] + B[k] @ U[k] + c[k])
    return np.array(X)


def cost_of(X, U, Q, R, Pf):
    return float(sum(X[k] @ Q @ X[k] + U[k] @ R @ U[k] for k in range(U.shape[0])) + X[-1] @ Pf @ X[-1])


def _kkt_refine(H, grad, Gam, lo, hi, lb, ub, U, tol=1e-6, max_rounds=50, duals=None):
    """Exact (to rounding) solution on the active set identified near U: solve the equality-
    constrained KKT system densely, then check primal feasibility and multiplier signs; constraints
    with a wrong-sign multiplier are dropped, violated ones added (a few primal active-set rounds)."""
    nz = H.shape[0]
    C_all = np.vstack([np.eye(nz), Gam])
    l_all = np.concatenate([lb, lo]); u_all = np.concatenate([ub, hi])
    val = C_all @ U
    scale = np.maximum(1.0, np.abs(val))
    side = np.where(np.abs(val - l_all) <= tol * scale, -1, np.where(np.abs(val - u_all) <= tol * scale, 1, 0))
    if duals is not None:
        # degenerate vertices: constraints that sit on their bound with a zero multiplier are only weakly active and
        # must not enter the equality set (they make it rank deficient); keep the ones the solver's duals support
        dual = np.concatenate(duals)
        strong = np.abs(dual) > 1e-7 * max(1.0, np.abs(dual).max())
        side = np.where(strong, side, 0)
    for _ in range(max_rounds):
        idx = np.nonzero(side)[0]
        C = C_all[idx]; dvec = np.where(side[idx] < 0, l_all[idx], u_all[idx])
        k = len(idx)
        KKT = np.block([[H, C.T], [C, np.zeros((k, k))]])
        sol = np.linalg.lstsq(KKT, np.concatenate([-grad, dvec]), rcond=None)[0] if k else np.linalg.solve(H, -grad)
        Un, lam = sol[:nz], sol[nz:]
        val = C_all @ Un
        scale = np.maximum(1.0, np.abs(val))
        viol_l = (val < l_all - 1e-9 * scale); viol_u = (val > u_all + 1e-9 * scale)
        # multiplier sign: for an upper-active row lam >= 0, lower-active lam <= 0 (H U + g + C'lam = 0)
        wrong = (lam * side[idx] < -1e-9 * np.maximum(1.0, np.abs(lam).max() if k else 1.0))
        if not viol_l.any() and not viol_u.any() and not wrong.any():
            return Un, True, side
        if wrong.any():
            worst = idx[np.argmin(lam * side[idx])]
            side[worst] = 0
        else:
            v = np.maximum(l_all - val, val - u_all) / scale
            j = int(np.argmax(v))
            side[j] = -1 if val[j] < l_all[j] else 1
    return Un, False, side


def solve_exact(A, B, Q, R, Pf, N, x0, u_lo, u_hi, x_lo, x_hi, c=None, method="highs", refine=True, Cg=None, hg=None):
    """Exact solution of one scenario.  HiGHS (or SLSQP) identifies the active set, a dense KKT solve
    on that set gives the solution to rounding accuracy (HiGHS alone is only ~1e-6 accurate in U when
    the cost is flat).  Returns dict(U [N,m], X [N+1,n], cost, status, sat_u, sat_x)."""
    A, B, c = stage_arrays(A, B, N, c)
    Q, R, Pf = (np.asarray(M, float) for M in (Q, R, Pf))
    x0 = np.asarray(x0, float)
    n, m = B.shape[-2], B.shape[-1]
    H, grad, Gam, lo, hi, lb, ub, free = _condensed_qp(A, B, c, Q, R, Pf, x0, u_lo, u_hi, x_lo, x_hi, Cg, hg)
    if method == "highs":
        U, ok, infeasible = _highs_qp(H, grad, Gam, lo, hi, lb, ub)
    elif method == "slsqp":
        from scipy.optimize import Bounds, LinearConstraint, minimize
        fin = np.isfinite(lo) | np.isfinite(hi)
        cons = [LinearConstraint(Gam[fin], lo[fin], hi[fin])] if fin.any() else []
        res = minimize(lambda z: 0.5 * z @ H @ z + grad @ z, np.clip(np.zeros_like(lb), lb, ub),
                       jac=lambda z: H @ z + grad, bounds=Bounds(lb, ub), constraints=cons, method="SLSQP",
                       options={"ftol": 1e-15, "maxiter": 1000})
        U, ok, infeasible = res.x, res.success, not res.success
    else:
        raise ValueError(method)
    if method == "highs" and not ok and not infeasible:
        # HiGHS' active-set QP occasionally stops without a verdict on degenerate problems: second opinion
        alt = solve_exact(A, B, Q, R, Pf, N, x0, u_lo, u_hi, x_lo, x_hi, c=c, method="slsqp", refine=refine, Cg=Cg, hg=hg)
        if alt["status"] == SOLVED:
            return alt
    status = SOLVED if ok else (INFEASIBLE if infeasible else MAX_ITER)
    side = None
    if ok and refine:
        U_raw = U
        duals = getattr(_highs_qp, "last_duals", None) if method == "highs" else None
        U, ok2, side = _kkt_refine(H, grad, Gam, lo, hi, lb, ub, U_raw, duals=duals)
        if not ok2 and duals is not None:
            U, ok2, side = _kkt_refine(H, grad, Gam, lo, hi, lb, ub, U_raw)
        if not ok2:
            status = MAX_ITER
            U = U_raw
    U = U.reshape(N, m)
    X = rollout(A, B, c, x0, U)
    out = {"U": U, "X": X, "cost": cost_of(X, U, Q, R, Pf), "status": status}
    if side is not None:
        out["sat_u"] = side[:N * m].reshape(N, m).astype(np.int8)
        out["sat_x"] = side[N * m:N * m + N * n].reshape(N, n).astype(np.int8)
        if Cg is not None:
            out["sat_c"] = side[N * m + N * n:].reshape(N, -1).astype(np.int8)
    return out


# ------------------------------------------------------------------------------------------------
# numpy restatement of the GPU algorithm (ADMM with a Riccati-structured subproblem), batched
# ------------------------------------------------------------------------------------------------
def riccati_factor(A, B, c, Q, R, Pf, rho_u, rho_x):
    """Factorisation of the ADMM subproblem  min cost + rho/2 |z - v|^2  s.t. dynamics
    (weights Q + rho_x I, R + rho_u I, terminal Pf + rho_x I).  Stage arrays may carry leading
    batch dims.  Returns K [N,...,m,n], Sinv [N,...,m,m], Pc [N,...,n] (= P_{k+1} c_k), P [N+1,...]."""
    N = B.shape[0]
    n, m = B.shape[-2], B.shape[-1]
    In, Im = np.eye(n), np.eye(m)
    P = Pf + rho_x * In
    Ks, Sinvs, Pcs, Ps = [], [], [], [P]
    for k in range(N - 1, -1, -1):
        Ak, Bk, ck = A[k], B[k], c[k]
        At, Bt = np.swapaxes(Ak, -1, -2), np.swapaxes(Bk, -1, -2)
        PA, PB = P @ Ak, P @ Bk
        S = R + rho_u * Im + Bt @ PB
        Sinv = np.linalg.inv(S)
        K = -Sinv @ (Bt @ PA)
        Pcs.append((P @ ck[..., None])[..., 0])
        Ks.append(K); Sinvs.append(Sinv)
        Qk = (Q + rho_x * In) if k > 0 else Q
        P = Qk + At @ (PA + PB @ K)
        Ps.append(P)
    return np.array(Ks[::-1]), np.array(Sinvs[::-1]), np.array(Pcs[::-1]), np.array(Ps[::-1])


def admm_riccati(A, B, Q, R, Pf, N, x0, u_lo, u_hi, x_lo, x_hi, c=None, rho_u=None, rho_x=None, alpha=1.6,
                 eps_abs=1e-9, eps_rel=1e-9, max_iter=4000, warm=None, check_infeasible=True):
    """Batched over x0 [batch, n] with a shared (possibly stage-varying) model.

    Splitting: z = (u_0, x_1, ..., u_{N-1}, x_N);  f = cost + dynamics,  g = box.  With the scaled
    dual mu and t = zhat + mu:  w = clamp(t), mu = t - w, so only t is carried between iterations:
        v      = 2 clamp(t) - t                      (= w - mu)
        ztilde = argmin f(z) + sum rho_i/2 (z_i - v_i)^2        (Riccati sweeps with FIXED gains)
        t+     = alpha ztilde + (1 - alpha) clamp(t) + (t - clamp(t))
    Output U = clamp(t_u) (exactly feasible inputs), X = rollout(U).
    """
    A, B, c = stage_arrays(A, B, N, c)
    Q, R, Pf = (np.asarray(M, float) for M in (Q, R, Pf))
    x0 = np.atleast_2d(np.asarray(x0, float))
    batch, n, m = x0.shape[0], B.shape[-2], B.shape[-1]
    if rho_u is None or rho_x is None:
        rho_u, rho_x = default_rho(Q, R)
    K, Sinv, Pc, P = riccati_factor(A, B, c, Q, R, Pf, rho_u, rho_x)
    u_lo, u_hi, x_lo, x_hi = (np.asarray(v, float) for v in (u_lo, u_hi, x_lo, x_hi))
    tu = np.zeros((N, batch, m)); tx = np.zeros((N, batch, n))  # tx[k] belongs to x_{k+1}
    if warm is not None:
        tu[:], tx[:] = warm
    else:
        # cold start: unconstrained rollout from zero inputs, clamped
        x = x0
        for k in range(N):
            tu[k] = np.clip(0.0, u_lo, u_hi)
            x = x @ A[k].T + tu[k] @ B[k].T + c[k]
            tx[k] = x
    status = np.zeros(batch, dtype=np.int32)
    iters = np.zeros(batch, dtype=np.int32)
    active = np.ones(batch, dtype=bool)
    prev_rp = np.zeros((N, batch, n + m))
    d = np.zeros((N, batch, m))
    for it in range(1, max_iter + 1):
        wu, wx = np.clip(tu, u_lo, u_hi), np.clip(tx, x_lo, x_hi)
        vu, vx = 2 * wu - tu, 2 * wx - tx
        # backward sweep: p_N = q_N;  h = P_{k+1} c_k - p_{k+1};  d_k = Sinv (r_k - B'h);
        #                 p_k = q_k - A'h + K'(r_k - B'h)
        p = rho_x * vx[N - 1]
        for k in range(N - 1, -1, -1):
            h = Pc[k] - p
            gu = rho_u * vu[k] - h @ B[k]
            d[k] = gu @ Sinv[k].T
            if k > 0:
                p = rho_x * vx[k - 1] - h @ A[k] + gu @ K[k]
        # forward sweep
        x = x0
        rp = np.zeros(batch); rd = np.zeros(batch); zn = np.zeros(batch); mun = np.zeros(batch)
        stall = np.zeros(batch)
        for k in range(N):
            u = x @ K[k].T + d[k]
            x = x @ A[k].T + u @ B[k].T + c[k]
            tun = alpha * u + (1 - alpha) * wu[k] + (tu[k] - wu[k])
            txn = alpha * x + (1 - alpha) * wx[k] + (tx[k] - wx[k])
            wun, wxn = np.clip(tun, u_lo, u_hi), np.clip(txn, x_lo, x_hi)
            r_k = np.concatenate([u - wun, x - wxn], axis=1)
            rp = np.maximum(rp, np.abs(r_k).max(axis=1))
            stall = np.maximum(stall, np.abs(r_k - prev_rp[k]).max(axis=1))
            prev_rp[k][active] = r_k[active]
            rd = np.maximum(rd, np.maximum(rho_u * np.abs(wun - wu[k]).max(axis=1), rho_x * np.abs(wxn - wx[k]).max(axis=1)))
            zn = np.maximum(zn, np.maximum(np.abs(u).max(axis=1), np.abs(x).max(axis=1)))
            zn = np.maximum(zn, np.maximum(np.abs(wun).max(axis=1), np.abs(wxn).max(axis=1)))
            mun = np.maximum(mun, np.maximum(rho_u * np.abs(tun - wun).max(axis=1), rho_x * np.abs(txn - wxn).max(axis=1)))
            tu[k][active] = tun[active]; tx[k][active] = txn[active]
        iters[active] = it
        done = active & (rp <= eps_abs + eps_rel * zn) & (rd <= eps_abs + eps_rel * mun)
        status[done] = SOLVED
        active &= ~done
        if check_infeasible and it > 10:
            # Douglas-Rachford on an infeasible problem: ztilde - w converges to the (non-zero) gap
            # vector between the dynamics subspace and the box -> residual stalls at a non-zero value
            inf = active & (stall <= 1e-7 * rp) & (rp > 1e-6 * np.maximum(1.0, zn))
            status[inf] = INFEASIBLE
            active &= ~inf
        if not active.any():
            break
    status[active] = MAX_ITER
    U = np.clip(tu, u_lo, u_hi)
    X = np.zeros((N + 1, batch, n)); X[0] = x0
    for k in range(N):
        X[k + 1] = X[k] @ A[k].T + U[k] @ B[k].T + c[k]
    cost = np.einsum("kbi,ij,kbj->b", X[:-1], Q, X[:-1]) + np.einsum("kbi,ij,kbj->b", U, R, U) \
        + np.einsum("bi,ij,bj->b", X[-1], Pf, X[-1])
    sat_u = (U <= u_lo).astype(np.int8) * -1 + (U >= u_hi).astype(np.int8)
    wx = np.clip(tx, x_lo, x_hi)
    sat_x = (tx <= x_lo).astype(np.int8) * -1 + (tx >= x_hi).astype(np.int8)
    return {"U": U, "X": X, "cost": cost, "status": status, "iters": iters, "sat_u": sat_u, "sat_x": sat_x,
            "t": (tu, tx)}


def default_rho(Q, R):
    """Penalty heuristic shared with the GPU host code: geometric mean of the weight scales."""
    q = float(np.mean(np.diag(np.asarray(Q, float))))
    r = float(np.mean(np.diag(np.asarray(R, float))))
    return max(r, 1e-6) * 10.0, max(q, 1e-6) * 1.0


# ------------------------------------------------------------------------------------------------
# numpy restatement of the algorithm the GPU runs (K4): Mehrotra predictor-corrector interior
# point method whose Newton systems are solved by a Riccati recursion over the horizon.
# The code is organised in the same passes as the CUDA kernel (csrc/boxqp_core.cuh).
# ------------------------------------------------------------------------------------------------
BIG = 1e19  # |bound| >= BIG means "no bound"


def _bmat(M, batch):
    M = np.asarray(M, float)
    return np.broadcast_to(M, (batch,) + M.shape[-2:]) if M.ndim == 2 else M


def _bvec(v, batch):
    v = np.asarray(v, float)
    return np.broadcast_to(v, (batch, v.shape[-1])) if v.ndim == 1 else v


def ipm_riccati(A, B, Q, R, Pf, N, x0, u_lo, u_hi, x_lo, x_hi, c=None, warm_U=None, max_iter=60,
                eps=1e-9, second_order=True, verbose=False, Cg=None, hg=None):
    """Batched box-constrained LQ-MPC QP.  A, B, c are lists/arrays over stages; each stage entry is
    shared ([n,n]) or per scenario ([batch,n,n]).  x0 [batch, n].

    Variables z_k = (u_k, x_{k+1}); for every finite bound a slack s > 0 and a multiplier lam > 0.
      Sigma = lam_l/s_l + lam_u/s_u
      dz    = argmin 1/2 dz'(H + Sigma) dz - rhs'dz  s.t. dx+ = A dx + B du, dx_0 = 0   (Riccati sweep)
      rhs   = -H z + (sigma mu - cc_l)/s_l - Sigma_l r_l - (sigma mu - cc_u)/s_u + Sigma_u r_u
    predictor: sigma = 0, cc = 0;  corrector: sigma = (mu_aff/mu)^3, cc = ds_aff * dlam_aff.
    One step length alpha = min(1, 0.995 * fraction to the boundary) for all variables.

    Optional general stage rows Cg[k] x_{k+1} >= hg[k] (Cg [N, nc, n] shared or [N, batch, nc, n]; hg [N, nc] or
    [N, batch, nc]): each row is treated like a lower-bounded element whose value is w = Cg x and whose direction is
    dw = Cg dx; it adds Cg' diag(lam/s) Cg to the stage Hessian and Cg' rhs_c to the stage gradient.
    """
    if np.asarray(A).ndim == 2:
        A, B, c = stage_arrays(A, B, N, c)
    Q, R, Pf = (np.asarray(M, float) for M in (Q, R, Pf))
    x0 = np.atleast_2d(np.asarray(x0, float))
    batch, n, m = x0.shape[0], np.asarray(B[0]).shape[-2], np.asarray(B[0]).shape[-1]
    d = m + n
    if c is None:
        c = np.zeros((N, n))
    Ab = [_bmat(A[k], batch) for k in range(N)]
    Bb = [_bmat(B[k], batch) for k in range(N)]
    cb = [_bvec(c[k], batch) for k in range(N)]
    lo = np.concatenate([np.asarray(u_lo, float), np.asarray(x_lo, float)])
    hi = np.concatenate([np.asarray(u_hi, float), np.asarray(x_hi, float)])
    has_l, has_u = lo > -BIG, hi < BIG
    lo_f, hi_f = np.where(has_l, lo, 0.0), np.where(has_u, hi, 0.0)
    ncons = N * int(has_l.sum() + has_u.sum())
    nc = 0
    if Cg is not None:
        Cg = np.asarray(Cg, float); hg = np.asarray(hg, float)
        nc = Cg.shape[-2]
        Cgb = [np.broadcast_to(Cg[k], (batch, nc, n)) for k in range(N)]
        hgb = [np.broadcast_to(hg[k], (batch, nc)) for k in range(N)]
        ncons += N * nc
    mv = lambda M, v: np.einsum("bij,bj->bi", M, v)
    mtv = lambda M, v: np.einsum("bji,bj->bi", M, v)
    dg = lambda v: np.einsum("bi,ij->bij", v, np.eye(v.shape[1]))

    # ---- start: inputs clamped into the box, states by rollout, slacks >= 1, lam = mu0 / s
    U0 = np.clip(np.zeros((N, batch, m)) if warm_U is None else np.asarray(warm_U, float), lo[:m], hi[:m])
    z = np.zeros((N, batch, d)); x = x0
    for k in range(N):
        x = mv(Ab[k], x) + mv(Bb[k], U0[k]) + cb[k]
        z[k, :, :m] = U0[k]; z[k, :, m:] = x
    # barrier parameter: tolerance scale from the weights; start value per scenario from the size of
    # the cost gradient at the start point (a centred start: lam ~ |H z0| / s)
    mu_scale = max(1.0, float(np.abs(Q).max()), float(np.abs(R).max()))
    g0 = np.zeros(batch)
    for k in range(N):
        g0 = np.maximum(g0, np.abs(z[k][:, :m] @ R.T).max(axis=1))
        g0 = np.maximum(g0, np.abs(z[k][:, m:] @ (Pf if k == N - 1 else Q).T).max(axis=1))
    mu0 = np.maximum(mu_scale, g0)[None, :, None]
    sl = np.where(has_l, np.maximum(z - lo_f, 1.0), 1.0)
    su = np.where(has_u, np.maximum(hi_f - z, 1.0), 1.0)
    ll = np.where(has_l, mu0 / sl, 0.0)
    lu = np.where(has_u, mu0 / su, 0.0)
    mu0 = mu0[0, :, 0]
    sc = np.ones((N, batch, nc)); lc = np.zeros((N, batch, nc)); ccc = np.zeros((N, batch, nc)); dwc = np.zeros((N, batch, nc))
    rcw = np.zeros((N, batch, nc))   # row residuals C x - h - s: carried and decayed by (1 - alpha), never recomputed
    for k in range(N):
        if nc:
            w0 = mv(Cgb[k], z[k][:, m:]) - hgb[k]
            sc[k] = np.maximum(w0, 1.0)
            lc[k] = mu0[:, None] / sc[k]
            rcw[k] = w0 - sc[k]

    status = np.zeros(batch, dtype=np.int32)
    iters = np.zeros(batch, dtype=np.int32)
    active = np.ones(batch, dtype=bool)
    K = [None] * N; Sinv = [None] * N; dff = [None] * N
    zh = np.zeros_like(z); ccl = np.zeros_like(z); ccu = np.zeros_like(z)
    rp_last = np.full(batch, np.inf)

    def backward(sig_mu, factor):
        """Pass 1 / 3: (factorisation and) feed-forward terms of the Newton step in RESIDUAL form
        (unknown dz, homogeneous dynamics): rhs = -H z + [(sig_mu - cc_l)/s_l - Sigma_l r_l]
        - [(sig_mu - cc_u)/s_u - Sigma_u r_u].  All terms stay O(lam) near convergence, so rounding
        errors of the huge barrier weights scale with |dz| and vanish at the solution."""
        Pacc = np.broadcast_to(Pf, (batch, n, n)); pacc = np.zeros((batch, n))
        for k in range(N - 1, -1, -1):
            Sl = np.where(has_l, ll[k] / sl[k], 0.0); Su = np.where(has_u, lu[k] / su[k], 0.0)
            rl, ru = z[k] - lo_f - sl[k], hi_f - z[k] - su[k]
            Hz = np.concatenate([z[k][:, :m] @ R.T, z[k][:, m:] @ (Pf if k == N - 1 else Q).T], axis=1)
            rhs = -Hz + np.where(has_l, (sig_mu[:, None] - ccl[k]) / sl[k] - Sl * rl, 0.0) - \
                np.where(has_u, (sig_mu[:, None] - ccu[k]) / su[k] - Su * ru, 0.0)
            Sig = Sl + Su
            if nc:
                Sc = lc[k] / sc[k]
                rc = rcw[k]
                rhs_c = (sig_mu[:, None] - ccc[k]) / sc[k] - Sc * rc
                rhs = rhs.copy()
                rhs[:, m:] += mtv(Cgb[k], rhs_c)
            if factor:
                P = Pacc + dg(Sig[:, m:])
                if nc:
                    P = P + np.einsum("bji,bj,bjl->bil", Cgb[k], Sc, Cgb[k])
                # Joseph form (see BoxQpIpm::backward): positive semidefinite sums only
                PB = P @ Bb[k]
                Rt = R + dg(Sig[:, :m])
                S = Rt + np.swapaxes(Bb[k], 1, 2) @ PB
                Sinv[k] = np.linalg.inv(S)
                K[k] = -Sinv[k] @ (np.swapaxes(PB, 1, 2) @ Ab[k])
                Acl = Ab[k] + Bb[k] @ K[k]
                Pacc = Q + np.swapaxes(Acl, 1, 2) @ (P @ Acl) + np.swapaxes(K[k], 1, 2) @ (Rt @ K[k])
            h = -(rhs[:, m:] + pacc)
            gu = rhs[:, :m] - mtv(Bb[k], h)
            dff[k] = mv(Sinv[k], gu)
            pacc = -mtv(Ab[k], h) + mtv(K[k], gu)

    def forward():
        dx = np.zeros((batch, n))
        for k in range(N):
            du = mv(K[k], dx) + dff[k]
            dx = mv(Ab[k], dx) + mv(Bb[k], du)
            zh[k, :, :m] = du; zh[k, :, m:] = dx
            if nc:
                dwc[k] = mv(Cgb[k], dx)

    def rows_dir(sig_mu):
        """Slack / multiplier directions of the general rows."""
        rc = rcw
        dsc = dwc + rc
        dlc = (sig_mu[None, :, None] - ccc) / sc - lc - lc / sc * dsc
        return dsc, dlc, rc

    def directions(sig_mu):
        dz = zh
        dsl, dsu = dz + (z - lo_f - sl), -dz + (hi_f - z - su)
        dll = np.where(has_l, (sig_mu[None, :, None] - ccl) / sl - ll - ll / sl * dsl, 0.0)
        dlu = np.where(has_u, (sig_mu[None, :, None] - ccu) / su - lu - lu / su * dsu, 0.0)
        return dz, np.where(has_l, dsl, 0.0), np.where(has_u, dsu, 0.0), dll, dlu

    def step_len(dsl, dsu, dll, dlu):
        # largest step keeping s, lam > 0: 1 / max_i(-ds_i/s_i, -dlam_i/lam_i)
        with np.errstate(divide="ignore", invalid="ignore"):
            q = np.maximum.reduce([np.where(has_l, -dsl / sl, 0.0), np.where(has_u, -dsu / su, 0.0),
                                   np.where(has_l, -dll / np.where(has_l, ll, 1.0), 0.0),
                                   np.where(has_u, -dlu / np.where(has_u, lu, 1.0), 0.0)])
            qmax = q.max(axis=(0, 2))
            if nc:
                qmax = np.maximum(qmax, np.maximum(-step_len.dsc / sc, -step_len.dlc / lc).max(axis=(0, 2)))
            return np.where(qmax > 0, 1.0 / qmax, 1e30)

    zero = np.zeros(batch)
    for it in range(1, max_iter + 1):
        mu = (np.where(has_l, sl * ll, 0).sum(axis=(0, 2)) + np.where(has_u, su * lu, 0).sum(axis=(0, 2)) +
              (sc * lc).sum(axis=(0, 2))) / ncons
        ccl[:] = 0; ccu[:] = 0; ccc[:] = 0
        backward(zero, True)
        forward()
        dz, dsl, dsu, dll, dlu = directions(zero)
        if nc:
            step_len.dsc, step_len.dlc, _ = rows_dir(zero)
        a_aff = np.minimum(1.0, step_len(dsl, dsu, dll, dlu))[None, :, None]
        mu_aff = (np.where(has_l, (sl + a_aff * dsl) * (ll + a_aff * dll), 0).sum(axis=(0, 2)) +
                  np.where(has_u, (su + a_aff * dsu) * (lu + a_aff * dlu), 0).sum(axis=(0, 2)))
        if nc:
            mu_aff = mu_aff + ((sc + a_aff * step_len.dsc) * (lc + a_aff * step_len.dlc)).sum(axis=(0, 2))
        mu_aff = mu_aff / ncons
        sigma = np.minimum(1.0, (mu_aff / np.maximum(mu, 1e-300)) ** 3)
        if second_order:
            ccl[:] = dsl * dll; ccu[:] = dsu * dlu
            if nc:
                ccc[:] = step_len.dsc * step_len.dlc
        sig_mu = np.maximum(sigma * mu, 1e-3 * eps * mu_scale)   # centring target, floored (see BoxQpIpm::solve)
        backward(sig_mu, False)
        forward()
        dz, dsl, dsu, dll, dlu = directions(sig_mu)
        if nc:
            step_len.dsc, step_len.dlc, _ = rows_dir(sig_mu)
        alpha = np.minimum(1.0, 0.995 * step_len(dsl, dsu, dll, dlu))
        al = alpha[None, :, None]
        act3 = active[None, :, None]          # retired scenarios are frozen (no 0 * nan leaks into them)

        def upd(arr, d_):
            arr[...] = np.where(act3, arr + al * d_, arr)
        upd(z, dz); upd(sl, dsl); upd(su, dsu); upd(ll, dll); upd(lu, dlu)
        if nc:
            upd(sc, step_len.dsc); upd(lc, step_len.dlc)
            rcw[...] = np.where(act3, (1.0 - al) * rcw, rcw)
        iters[active] = it
        mu_new = (np.where(has_l, sl * ll, 0).sum(axis=(0, 2)) + np.where(has_u, su * lu, 0).sum(axis=(0, 2)) +
                  (sc * lc).sum(axis=(0, 2))) / ncons
        rp = np.maximum(np.where(has_l, np.abs(z - lo_f - sl), 0).max(axis=(0, 2)),
                        np.where(has_u, np.abs(hi_f - z - su), 0).max(axis=(0, 2)))
        if nc:
            rp = np.maximum(rp, np.abs(rows_dir(zero)[2]).max(axis=(0, 2)))
        zn = np.maximum(1.0, np.abs(z).max(axis=(0, 2)))
        step = alpha * np.abs(dz).max(axis=(0, 2))
        if verbose:
            print(it, "mu", mu_new.max(), "rp", rp.max(), "alpha", alpha.min(), "step", step.max())
        done = active & (mu_new <= eps * mu_scale) & (rp <= eps * zn) & (step <= 1e-6 * zn)
        status[done] = SOLVED
        active &= ~done
        # stalled: the step length collapses / the barrier parameter grows 100x above its start value.  With a bound
        # residual that cannot be closed the problem is infeasible (box and dynamics do not meet).
        with np.errstate(invalid="ignore"):
            stuck = active & (~(alpha >= 1e-6) | ~(mu_new <= 100.0 * mu0))
            status[stuck & (rp <= 1e-6 * zn)] = MAX_ITER
            status[stuck & ~(rp <= 1e-6 * zn)] = INFEASIBLE
        active &= ~stuck
        if stuck.any():
            # the batched numpy code keeps evaluating retired scenarios: park diverged iterates on benign values
            # (their outputs are only the status; the GPU threads simply leave the loop)
            bad = stuck & ~np.isfinite(mu_new)
            for arr in (sl, su, sc):
                arr[:, bad] = 1.0
            for arr in (ll, lu, lc):
                arr[:, bad] = 1.0 if arr is lc else np.where(arr[:, bad] != 0, 1.0, 0.0)
            z[:, bad] = np.nan_to_num(z[:, bad], nan=0.0, posinf=0.0, neginf=0.0)
        if not active.any():
            break
    # infeasible problems keep a bound residual that cannot be closed
    rp = np.maximum(np.where(has_l, np.abs(z - lo_f - sl), 0).max(axis=(0, 2)),
                    np.where(has_u, np.abs(hi_f - z - su), 0).max(axis=(0, 2)))
    if nc:
        rp = np.maximum(rp, np.abs(rows_dir(zero)[2]).max(axis=(0, 2)))
    zn = np.maximum(1.0, np.abs(z).max(axis=(0, 2)))
    status[active & ~(rp <= 1e-6 * zn)] = INFEASIBLE   # also catches diverged (non-finite) iterates
    status[active & (rp <= 1e-6 * zn)] = MAX_ITER
    # ---- output: active set from the complementarity pairs, inputs snapped onto their bounds
    act_l = has_l & (ll > sl); act_u = has_u & (lu > su)
    zs = np.where(act_l, lo_f, np.where(act_u, hi_f, z))
    U = zs[:, :, :m]
    X = np.zeros((N + 1, batch, n)); X[0] = x0
    for k in range(N):
        X[k + 1] = mv(Ab[k], X[k]) + mv(Bb[k], U[k]) + cb[k]
    cost = np.einsum("kbi,ij,kbj->b", X[:-1], Q, X[:-1]) + np.einsum("kbi,ij,kbj->b", U, R, U) \
        + np.einsum("bi,ij,bj->b", X[-1], Pf, X[-1])
    sat = act_u.astype(np.int8) - act_l.astype(np.int8)
    return {"U": U, "X": X, "cost": cost, "status": status, "iters": iters, "sat_u": sat[:, :, :m],
            "sat_x": sat[:, :, m:], "sat_c": (lc > sc).astype(np.int8) * -1}


# ------------------------------------------------------------------------------------------------
# exact solves of many scenarios on all host cores (tests of the full-size configurations)
# ------------------------------------------------------------------------------------------------
def _exact_worker(args):
    A, B, Q, R, Pf, N, x0, u_lo, u_hi, x_lo, x_hi = args
    return solve_exact(A, B, Q, R, Pf, N, x0, u_lo, u_hi, x_lo, x_hi)


def solve_exact_many(A, B, Q, R, Pf, N, X0, u_lo, u_hi, x_lo, x_hi, processes=None):
    """solve_exact for every row of X0 [count, n] (shared LTI model), one process per host core."""
    import multiprocessing as mp
    import os
    procs = processes or max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    jobs = [(A, B, Q, R, Pf, N, np.asarray(x, float), u_lo, u_hi, x_lo, x_hi) for x in X0]
    if procs == 1 or len(jobs) < 4:
        return [_exact_worker(j) for j in jobs]
    with mp.get_context("spawn").Pool(min(procs, len(jobs))) as pool:
        return pool.map(_exact_worker, jobs, chunksize=1)
