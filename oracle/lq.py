"""fp64 numpy restatement of the reference's session-1 LQ path (oracle, test infra).

Every function names the reference lines it follows.  Shapes and quirks are
kept on purpose: gains carry the minus sign (u = +K x),
lists are returned first-stage-first, ``simulate`` yields ``steps`` states,
``prediction`` skips ``gains[0]``.

Pinned by tests/test_oracle_golden.py against tests/golden/session1.json
(outputs of the reference's own code) and against the live reference when
/root/reference is mounted.
"""
from __future__ import annotations

import numpy as np


def _as_R(R, m):
    """R may be (m,), (m,m) or scalar-like (reference FHC.py:141 passes shape (1,))."""
    R = np.asarray(R, dtype=np.float64)
    if R.ndim <= 1:
        if R.size != 1 and R.size != m:
            raise ValueError("R must be (m,), (m,m) or scalar")
        # numpy broadcasting of (m,) onto (m,m) adds R[j] to column j; only m == 1
        # is meaningful in the reference.  Restate that case exactly.
        if m != 1:
            raise ValueError("1-D R only supported for a single input (reference FHC.py:141)")
        return R.reshape(1, 1)
    return R


def ricatti_recursion(A, B, Q, R, P_f, N):
    """Backward Riccati recursion, FHC argument order.

    Follows /root/reference/session_1/FHC.py:51-61:
        K_k = -inv(R + B'PB) B'PA          (:56)
        P_k = Q + A'PA + A'PB K_k          (:57)
    returns (P reversed, K reversed)       (:61)
    Batched extension: any leading batch dims broadcast through ``@``.
    """
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    P = np.asarray(P_f, dtype=np.float64)
    m = B.shape[-1]
    R = _as_R(R, m) if np.asarray(R).ndim <= 2 else np.asarray(R, dtype=np.float64)
    At = np.swapaxes(A, -1, -2)
    Bt = np.swapaxes(B, -1, -2)
    Ps = [P]
    Ks = []
    for _ in range(int(N)):
        S = R + Bt @ P @ B
        K = -np.linalg.inv(S) @ Bt @ P @ A
        P = Q + At @ P @ A + At @ P @ B @ K
        Ks.append(K)
        Ps.append(P)
    return Ps[::-1], Ks[::-1]


def riccati_recursion(A, B, R, Q, Pf, N):
    """Same recursion, instructor-solution form and argument order (R before Q).

    Follows /root/reference/session_1/session1_sol.py:44-65:
        K_k = -solve(R + B'PB, B'PA)       (:60)
        P_k = Q + A'P(A + B K_k)           (:62)
    """
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64)
    P = np.asarray(Pf, dtype=np.float64)
    At = np.swapaxes(A, -1, -2)
    Bt = np.swapaxes(B, -1, -2)
    Ps = [P]
    Ks = []
    for _ in range(int(N)):
        K = -np.linalg.solve(R + Bt @ P @ B, Bt @ P @ A)
        Ks.append(K)
        P = Q + At @ P @ (A + B @ K)
        Ps.append(P)
    return Ps[::-1], Ks[::-1]


def step(A, B, x, u):
    """x+ = A x + B u  (/root/reference/session_1/LinearSystem.py:16-18)."""
    return A @ x + B @ u


def simulate(A, B, x0, gains, steps, mode="receding"):
    """Closed loop of LinearSystem.simulate with an AutoCruising policy.

    Follows /root/reference/session_1/LinearSystem.py:20-26 (t = 1..steps-1,
    ``steps`` states including x0, layout (n, batch, steps)) with the policies
    of FHC.py:25-29: ``receding`` -> gains[0] @ x, ``pred`` -> gains[t] @ x.
    x0 is (n, batch).
    """
    x = np.asarray(x0, dtype=np.float64)
    out = [x]
    for t in range(1, int(steps)):
        K = gains[0] if mode == "receding" else gains[t]
        x = step(A, B, x, K @ x)
        out.append(x)
    return np.stack(out, axis=2)


def prediction(A, B, xt, gains, horizon):
    """Open-loop prediction of LinearSystem.prediction + AutoCruising.pred.

    Follows /root/reference/session_1/LinearSystem.py:28-35 with FHC.py:28-29:
    the loop runs t = 1..horizon-1 and applies gains[t] (gains[0] is never used).
    """
    return simulate(A, B, xt, gains, horizon, mode="pred")


def session1_simulate(A, B, x0, gains, steps, mode="receding"):
    """Generic loop of the instructor solution.

    Follows /root/reference/session_1/session1_sol.py:68-91: t = 0..steps-1,
    returns ((steps+1, ..., n) states, instability flag ``|x_{t+1}|_2 > 100``).
    x0 is (n,) or batched (batch, n); flag is a bool or (batch,) bool array.
    mode ``receding`` -> gains[0] (:107-109), ``pred`` -> gains[t] (:121-123).
    """
    x = np.asarray(x0, dtype=np.float64)
    xs = [x]
    flag = np.zeros(x.shape[:-1], dtype=bool)
    for t in range(int(steps)):
        K = gains[0] if mode == "receding" else gains[t]
        u = x @ np.swapaxes(K, -1, -2) if x.ndim > 1 else K @ x
        x = x @ np.swapaxes(A, -1, -2) + u @ np.swapaxes(B, -1, -2) if x.ndim > 1 else A @ x + B @ u
        xs.append(x)
        flag = flag | (np.linalg.norm(x, axis=-1) > 100)
    return np.array(xs), (bool(flag) if flag.ndim == 0 else flag)


def cost_to_go(P0, x0):
    """V_N(x0) = x0' P[0] x0 (/root/reference/session_1/FHC.py:123-124). x0 is (n, batch)."""
    x0 = np.asarray(x0, dtype=np.float64)
    return np.einsum("ib,ij,jb->b", x0, np.asarray(P0, dtype=np.float64), x0)


def lq_open_loop(A, B, Q, R, P_f, x0, N):
    """Optimal open-loop plan of the finite-horizon LQ problem from x0.

    u_k = K[k] x_k for k = 0..N-1 (gains of ricatti_recursion, FHC.py:51-61), stage
    cost x'Qx + u'Ru, terminal x_N' P_f x_N.  Returns X (N+1, n), U (N, m), V.
    Single scenario; used to check the fused per-scenario solve.
    """
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    m = B.shape[1]
    Rm = _as_R(R, m)
    P, K = ricatti_recursion(A, B, Q, Rm, P_f, N)
    x = np.asarray(x0, dtype=np.float64).reshape(-1)
    X = [x]
    U = []
    V = 0.0
    for k in range(int(N)):
        u = K[k] @ x
        V += x @ Q @ x + u @ Rm @ u
        x = A @ x + B @ u
        X.append(x)
        U.append(u)
    V += x @ np.asarray(P_f, dtype=np.float64) @ x
    return np.array(X), np.array(U), float(V), P, K
