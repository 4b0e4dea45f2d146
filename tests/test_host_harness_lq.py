"""The kernels' per-scenario bodies (csrc/lq_core.cuh), run on the CPU by tests/harness, against
the oracle.  This checks algebra + indexing of the device code without a GPU; the same
comparisons run on the real kernels in test_gpu_lq.py."""
import ctypes as C

import numpy as np
import pytest

from oracle import lq

P64 = C.POINTER(C.c_double)


def p(a):
    return None if a is None else a.ctypes.data_as(P64)


def c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def models(rng, batch, n, m):
    A = np.eye(n) + 0.5 * np.diag(np.ones(n - 1), 1) + 0.05 * rng.standard_normal((batch, n, n))
    B = np.zeros((n, m)); B[-1, 0] = -0.5
    if m > 1:
        B[-2, 1] = 0.3
    B = B + 0.05 * rng.standard_normal((batch, n, m))
    Q = np.eye(n) * (1 + 0.2 * rng.random((batch, 1, 1)))
    Rm = rng.standard_normal((batch, m, m)) * 0.05
    R = 0.1 * np.eye(m) * (1 + rng.random((batch, 1, 1))) + Rm @ Rm.transpose(0, 2, 1)
    return c(A), c(B), c(Q), c(R)


@pytest.mark.parametrize("n,m", [(2, 1), (4, 1), (4, 2)])
def test_riccati_body(hh, n, m):
    rng = np.random.default_rng(1)
    batch, N = 7, 12
    A, B, Q, R = models(rng, batch, n, m)
    K = np.zeros((N, batch, m, n)); P = np.zeros((N + 1, batch, n, n))
    rc = hh.hh_riccati(p(A), C.c_int64(n * n), p(B), C.c_int64(n * m), p(Q), C.c_int64(n * n), p(R), C.c_int64(m * m),
                       p(Q), C.c_int64(n * n), p(K), p(P), 1, C.c_int64(batch), n, m, N)
    assert rc == 0
    Po, Ko = lq.ricatti_recursion(A, B, Q, R, Q, N)
    np.testing.assert_allclose(K, np.array(Ko), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(P, np.array(Po), rtol=1e-10, atol=1e-12)
    # shared model, P0 only
    K1 = np.zeros((N, 1, m, n)); P0 = np.zeros((1, n, n))
    rc = hh.hh_riccati(p(A[0]), C.c_int64(0), p(B[0]), C.c_int64(0), p(Q[0]), C.c_int64(0), p(R[0]), C.c_int64(0),
                       p(Q[0]), C.c_int64(0), p(K1), p(P0), 0, C.c_int64(1), n, m, N)
    assert rc == 0
    np.testing.assert_allclose(P0[0], Po[0][0], rtol=1e-10)
    np.testing.assert_allclose(K1[:, 0], np.array(Ko)[:, 0], rtol=1e-10, atol=1e-12)


def test_riccati_body_golden(hh, golden):
    g = golden["cfg1"]
    A, B, Q = c(g["A"]), c(g["B"]), c(g["Q"])
    R = c(g["R"]).reshape(1, 1)
    for N, rec in g["recursion"].items():
        N = int(N)
        K = np.zeros((N, 1, 1, 2)); P = np.zeros((N + 1, 1, 2, 2))
        assert hh.hh_riccati(p(A), C.c_int64(0), p(B), C.c_int64(0), p(Q), C.c_int64(0), p(R), C.c_int64(0), p(Q),
                             C.c_int64(0), p(K), p(P), 1, C.c_int64(1), 2, 1, N) == 0
        np.testing.assert_allclose(K[:, 0], np.array(rec["K"]), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(P[:, 0], np.array(rec["P"]), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("n,m", [(2, 1), (4, 1), (4, 2)])
@pytest.mark.parametrize("vec", [1, 2])
def test_rollout_shared_body(hh, n, m, vec):
    rng = np.random.default_rng(2)
    batch, N, T = 10, 9, 9
    A, B, Q, R = (M[0] for M in models(rng, 1, n, m))
    A, B, Q, R = c(A), c(B), c(Q), c(R)
    Po, Ko = lq.ricatti_recursion(A, B, Q, R, Q, N)
    K = c(np.array(Ko))
    x0 = c(rng.uniform(-10, 10, (n, batch)))
    for off, step, mode in [(0, 0, "receding"), (1, 1, "pred")]:
        X = np.zeros((T, n, batch)); U = np.zeros((T - 1, m, batch)); cost = np.zeros(batch)
        flag = np.zeros(batch, dtype=np.uint8)
        rc = hh.hh_lq_rollout(p(A), C.c_int64(0), p(B), C.c_int64(0), p(K), C.c_int64(m * n), C.c_int64(0), off, step,
                              p(x0), p(X), p(U), p(Q), p(R), p(Q), p(cost), flag.ctypes.data_as(C.POINTER(C.c_uint8)),
                              C.c_double(100.0), C.c_int64(batch), n, m, T, vec)
        assert rc == 0
        Xo = lq.simulate(A, B, x0, Ko, T, mode=mode)  # (n, batch, T)
        np.testing.assert_allclose(X.transpose(1, 2, 0), Xo, rtol=1e-10, atol=1e-10)
        # cost: sum of stage costs + terminal
        co = np.zeros(batch)
        for t in range(T - 1):
            g = Ko[0] if mode == "receding" else Ko[t + 1]
            u = g @ Xo[:, :, t]
            np.testing.assert_allclose(U[t], u, rtol=1e-9, atol=1e-9)
            co += np.einsum("ib,ij,jb->b", Xo[:, :, t], Q, Xo[:, :, t]) + np.einsum("ib,ij,jb->b", u, R, u)
        co += np.einsum("ib,ij,jb->b", Xo[:, :, -1], Q, Xo[:, :, -1])
        np.testing.assert_allclose(cost, co, rtol=1e-10)
        np.testing.assert_array_equal(flag, (np.linalg.norm(Xo[:, :, 1:], axis=0) > 100).any(axis=1))


@pytest.mark.parametrize("n,m", [(2, 1), (4, 1), (4, 2)])
def test_rollout_per_scenario_body(hh, n, m):
    rng = np.random.default_rng(3)
    batch, N = 6, 8
    A, B, Q, R = models(rng, batch, n, m)
    Po, Ko = lq.ricatti_recursion(A, B, Q, R, Q, N)
    K = c(np.array(Ko))  # [N, batch, m, n]
    x0 = c(rng.uniform(-10, 10, (n, batch)))
    T = N + 1
    X = np.zeros((T, n, batch))
    rc = hh.hh_lq_rollout(p(A), C.c_int64(n * n), p(B), C.c_int64(n * m), p(K), C.c_int64(batch * m * n), C.c_int64(m * n),
                          0, 1, p(x0), p(X), None, None, None, None, None, None, C.c_double(100.0), C.c_int64(batch),
                          n, m, T, 1)
    assert rc == 0
    for b in range(batch):
        Xb, Ub, V, _, _ = lq.lq_open_loop(A[b], B[b], Q[b], R[b], Q[b], x0[:, b], N)
        np.testing.assert_allclose(X[:, :, b], Xb, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("n,m", [(2, 1), (4, 1), (4, 2)])
def test_lq_solve_body(hh, n, m):
    rng = np.random.default_rng(4)
    batch, N = 9, 20
    A, B, Q, R = models(rng, batch, n, m)
    x0 = c(rng.uniform(-10, 10, (batch, n)))
    X = np.zeros((N + 1, batch, n)); U = np.zeros((N, batch, m)); V = np.zeros(batch)
    K = np.zeros((N, batch, m, n)); P0 = np.zeros((batch, n, n))
    rc = hh.hh_lq_solve(p(A), C.c_int64(n * n), p(B), C.c_int64(n * m), p(Q), C.c_int64(n * n), p(R), C.c_int64(m * m),
                        p(Q), C.c_int64(n * n), p(x0), p(X), p(U), p(V), p(K), p(P0), C.c_int64(batch), n, m, N)
    assert rc == 0
    for b in range(batch):
        Xb, Ub, Vb, Pb, Kb = lq.lq_open_loop(A[b], B[b], Q[b], R[b], Q[b], x0[b], N)
        np.testing.assert_allclose(X[:, b], Xb, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(U[:, b], Ub, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(V[b], Vb, rtol=1e-10)
        np.testing.assert_allclose(V[b], x0[b] @ Pb[0] @ x0[b], rtol=1e-9)  # V = x0' P0 x0 (FHC.py:123-124)
        np.testing.assert_allclose(K[:, b], np.array(Kb), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(P0[b], Pb[0], rtol=1e-10)


@pytest.mark.parametrize("n", [2, 4])
def test_lq_solve_krylov_body(hh, n):
    """Single-input solve in Krylov coordinates (lq_solve_krylov_body) against the oracle's dense
    recursion, and the conditioning guard: ill-conditioned scenarios are handed to the dense body."""
    rng = np.random.default_rng(5)
    batch, N, m = 64, 20, 1
    A, B, Q, R = models(rng, batch, n, m)
    # three scenarios the guard must reject: uncontrollable pair, b = 0, nearly dependent Krylov columns
    A[0] = np.eye(n); B[1] = 0.0
    A[2] = np.diag(1.0 + 1e-6 * np.arange(n)); B[2] = 1.0
    x0 = c(rng.uniform(-10, 10, (batch, n)))
    X = np.zeros((N + 1, batch, n)); U = np.zeros((N, batch, m)); V = np.zeros(batch)
    used = np.zeros(batch, dtype=np.uint8)
    Pf = c(2.5 * Q + 0.1 * np.eye(n))
    Pf[::2] = Q[::2]   # half of the scenarios keep P_f = Q
    rc = hh.hh_lq_solve_krylov(p(A), C.c_int64(n * n), p(B), C.c_int64(n * m), p(Q), C.c_int64(n * n), p(R),
                               C.c_int64(m * m), p(Pf), C.c_int64(n * n), p(x0), p(X), p(U), p(V), C.c_int64(batch),
                               n, N, C.c_double(1e3), used.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert not used[:3].any()
    assert used[3:].mean() > 0.8, "the well-conditioned scenarios must take the Krylov path"
    for b in range(batch):
        Xb, Ub, Vb, Pb, Kb = lq.lq_open_loop(A[b], B[b], Q[b], R[b], Pf[b], x0[b], N)
        scale = max(1.0, np.abs(Xb).max())
        np.testing.assert_allclose(X[:, b], Xb, rtol=0, atol=1e-8 * scale)
        np.testing.assert_allclose(U[:, b], Ub, rtol=0, atol=1e-8 * max(1.0, np.abs(Ub).max()))
        np.testing.assert_allclose(V[b], Vb, rtol=1e-8)


def test_lq_solve_krylov_body_against_extended_precision(hh):
    """Where the Krylov-coordinate path and the dense fp64 recursion disagree most (fast, badly conditioned models:
    spectral radius 2-3, |P| ~ 1e7), extended precision (numpy longdouble) says which one is off: the accepted
    Krylov-path solves stay within 1e-8 of the exact plan, the dense recursion -- the reference's own arithmetic --
    is the one that drifts (up to ~1e-6 on this distribution)."""
    if np.finfo(np.longdouble).eps > 1e-18:
        pytest.skip("no extended-precision long double on this platform")
    rng = np.random.default_rng(3)
    batch, n, m, N = 20000, 4, 1, 20
    A = c(np.eye(n) + 0.5 * np.diag(np.ones(n - 1), 1) + 0.5 * rng.standard_normal((batch, n, n)))
    B = np.zeros((n, m)); B[-1, 0] = -0.5
    B = c(B + 0.5 * rng.standard_normal((batch, n, m)))
    G = rng.standard_normal((batch, n, n))
    Q = c(G @ G.transpose(0, 2, 1) / n + 0.01 * np.eye(n))
    R = c(0.1 * (1 + rng.random((batch, 1, 1))))
    x0 = c(rng.uniform(-10, 10, (batch, n)))
    out = {}
    for name in ("krylov", "dense"):
        X = np.zeros((N + 1, batch, n)); U = np.zeros((N, batch, m)); V = np.zeros(batch)
        used = np.zeros(batch, dtype=np.uint8)
        common = (p(A), C.c_int64(n * n), p(B), C.c_int64(n * m), p(Q), C.c_int64(n * n), p(R), C.c_int64(1), p(Q),
                  C.c_int64(n * n), p(x0), p(X), p(U), p(V))
        if name == "krylov":
            rc = hh.hh_lq_solve_krylov(*common, C.c_int64(batch), n, N, C.c_double(1e3), used.ctypes.data_as(C.c_void_p))
        else:
            rc = hh.hh_lq_solve(*common, None, None, C.c_int64(batch), n, m, N)
        assert rc == 0
        out[name] = (U[:, :, 0], used)
    Uk, used = out["krylov"]
    Ud = out["dense"][0]
    assert used.mean() > 0.8
    scale = np.maximum(np.abs(Ud).max(0), 1.0)
    gap = np.abs(Uk - Ud).max(0) / scale
    L = np.longdouble
    worst_k = 0.0
    for b in np.argsort(gap)[-12:]:
        Ab, Bb, Qb, Rb = (M[b].astype(L) for M in (A, B, Q, R))
        P = Qb.copy(); Ks = []
        for _ in range(N):
            K = -(Bb.T @ P @ Ab) / (Rb + Bb.T @ P @ Bb)
            P = Qb + Ab.T @ P @ (Ab + Bb @ K)
            Ks.append(K)
        x = x0[b].astype(L)[:, None]; Ue = []
        for K in Ks[::-1]:
            u = K @ x
            x = Ab @ x + Bb @ u
            Ue.append(u[0, 0])
        Ue = np.array(Ue, dtype=L)
        s = max(1.0, float(np.abs(Ue).max()))
        ek = float(np.abs(Uk[:, b] - Ue).max()) / s
        ed = float(np.abs(Ud[:, b] - Ue).max()) / s
        if used[b]:
            worst_k = max(worst_k, ek)
            assert ek <= max(1e-8, ed), (b, ek, ed)
    assert worst_k <= 1e-8
