"""csrc/boxqp_core.cuh (the K4 per-scenario body) run on the CPU by tests/harness, against the
numpy restatement of the same algorithm and against the exact active-set oracle."""
import ctypes as C

import numpy as np
import pytest

from oracle import boxqp as bq

P64 = C.POINTER(C.c_double)


def p(a):
    return None if a is None else a.ctypes.data_as(P64)


def c_(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def run_harness(hh, A, B, c, ltv, Q, R, Pf, ulo, uhi, xlo, xhi, x0, N, warm=None, max_iter=60, eps=1e-9, store=1):
    """x0 [batch, n] -> dict with U [N, batch, m], X [N+1, batch, n] (oracle layout)."""
    batch, n = x0.shape
    m = len(ulo)
    x0T = c_(x0.T)
    U = np.zeros((N, m, batch)); X = np.zeros((N + 1, n, batch)); cost = np.zeros(batch)
    status = np.zeros(batch, dtype=np.int32); iters = np.zeros(batch, dtype=np.int32)
    su = np.zeros((N, m, batch), dtype=np.int8); sx = np.zeros((N, n, batch), dtype=np.int8)
    warmT = None if warm is None else c_(warm.transpose(0, 2, 1))
    rc = hh.hh_boxqp_solve(p(c_(A)), p(c_(B)), p(None if c is None else c_(c)), ltv, p(c_(Q)), p(c_(R)), p(c_(Pf)),
                           p(c_(ulo)), p(c_(uhi)), p(c_(xlo)), p(c_(xhi)), p(x0T), p(warmT), p(U), p(X), p(cost),
                           status.ctypes.data_as(C.POINTER(C.c_int32)), iters.ctypes.data_as(C.POINTER(C.c_int32)),
                           su.ctypes.data_as(C.POINTER(C.c_int8)), sx.ctypes.data_as(C.POINTER(C.c_int8)),
                           C.c_int64(batch), n, m, N, max_iter, C.c_double(eps), store)
    assert rc == 0
    return {"U": U.transpose(0, 2, 1), "X": X.transpose(0, 2, 1), "cost": cost, "status": status, "iters": iters,
            "sat_u": su.transpose(0, 2, 1), "sat_x": sx.transpose(0, 2, 1)}


def session_x0(rng, batch):
    return np.stack([rng.uniform(-100, 0, batch), rng.uniform(-10, 15, batch)], 1)


@pytest.mark.parametrize("store", [0, 1])
@pytest.mark.parametrize("make,N", [(bq.Problem, 5), (bq.Problem, 30), (bq.session3_problem, 30)])
def test_session23_problem_matches_numpy_port_and_exact(hh, make, N, store):
    prob = make(N=N)
    ulo, uhi, xlo, xhi = bq.problem_bounds(prob)
    rng = np.random.default_rng(N)
    x0 = session_x0(rng, 48)
    x0[0] = [-100.0, 0.0]  # U* = [10,10,10,10,-20] at N = 5
    x0[1] = [-1.0, 14.0]   # cannot brake in time: infeasible
    got = run_harness(hh, prob.A, prob.B, None, 0, prob.Q, prob.R, prob.Q, ulo, uhi, xlo, xhi, x0, N, store=store)
    port = bq.ipm_riccati(prob.A, prob.B, prob.Q, prob.R, prob.Q, N, x0, ulo, uhi, xlo, xhi)
    np.testing.assert_array_equal(got["status"], port["status"])
    np.testing.assert_array_equal(got["iters"], port["iters"])
    ok = got["status"] == bq.SOLVED
    np.testing.assert_allclose(got["U"][:, ok], port["U"][:, ok], rtol=1e-9, atol=1e-9)
    np.testing.assert_array_equal(got["sat_u"][:, ok], port["sat_u"][:, ok])
    np.testing.assert_array_equal(got["sat_x"][:, ok], port["sat_x"][:, ok])
    assert got["status"][1] == bq.INFEASIBLE
    if N == 5:
        np.testing.assert_array_equal(got["U"][:, 0, 0], [10, 10, 10, 10, -20])
    for b in range(x0.shape[0]):
        ex = bq.solve_exact(prob.A, prob.B, prob.Q, prob.R, prob.Q, N, x0[b], ulo, uhi, xlo, xhi)
        if ex["status"] != bq.SOLVED:
            assert got["status"][b] != bq.SOLVED or ex["status"] == bq.MAX_ITER
            continue
        assert got["status"][b] == bq.SOLVED
        su = max(1.0, np.abs(ex["U"]).max()); sx = max(1.0, np.abs(ex["X"]).max())
        assert np.abs(got["U"][:, b] - ex["U"]).max() <= 1e-6 * su
        assert np.abs(got["X"][:, b] - ex["X"]).max() <= 1e-6 * sx
        assert abs(got["cost"][b] - ex["cost"]) <= 1e-9 * abs(ex["cost"])
        np.testing.assert_array_equal(got["sat_u"][:, b], ex["sat_u"])   # bit-identical saturation pattern
        np.testing.assert_array_equal(got["sat_x"][:, b], ex["sat_x"])
        # saturated inputs are returned exactly on the bound
        assert np.all(got["U"][:, b][ex["sat_u"] > 0] == uhi[0]) and np.all(got["U"][:, b][ex["sat_u"] < 0] == ulo[0])


def random_ltv(rng, batch, N, n, m):
    A0 = np.eye(n) + 0.1 * np.diag(np.ones(n - 1), 1)
    B0 = np.zeros((n, m)); B0[-1, 0] = 0.1
    if m > 1:
        B0[-2, 1] = 0.1
    A = A0 + 0.02 * rng.standard_normal((N, batch, n, n))
    B = B0 + 0.02 * rng.standard_normal((N, batch, n, m))
    c = 0.01 * rng.standard_normal((N, batch, n))
    return A, B, c


@pytest.mark.parametrize("store", [0, 1])
@pytest.mark.parametrize("n,m", [(2, 1), (4, 1), (4, 2)])
def test_ltv_per_scenario_models(hh, n, m, store):
    rng = np.random.default_rng(n * 10 + m)
    batch, N = 12, 15
    A, B, c = random_ltv(rng, batch, N, n, m)
    Q = np.diag(rng.uniform(0.5, 2.0, n)); R = np.diag(rng.uniform(0.05, 0.2, m)); Pf = 5 * Q
    ulo, uhi = -np.ones(m), 0.5 * np.ones(m)
    xlo, xhi = -2.0 * np.ones(n), 2.0 * np.ones(n)
    xlo[0] = -1e20  # unbounded below in the first state
    x0 = rng.uniform(-1.5, 1.5, (batch, n))
    warm = rng.uniform(-2, 2, (N, batch, m))
    # harness layout: A [N][n*n][batch]
    Ah = A.reshape(N, batch, n * n).transpose(0, 2, 1); Bh = B.reshape(N, batch, n * m).transpose(0, 2, 1)
    ch = c.transpose(0, 2, 1)
    got = run_harness(hh, Ah, Bh, ch, 1, Q, R, Pf, ulo, uhi, xlo, xhi, x0, N, warm=warm, store=store)
    port = bq.ipm_riccati(list(A), list(B), Q, R, Pf, N, x0, ulo, uhi, xlo, xhi, c=list(c), warm_U=warm)
    np.testing.assert_array_equal(got["status"], port["status"])
    ok = got["status"] == bq.SOLVED
    assert ok.sum() >= batch // 2
    # store 1 keeps the corrector data in float32: its iterates leave those of the all-float64 port at the level of the
    # stopping tolerance (the bar against the exact solution below is the same 1e-6 for both)
    np.testing.assert_allclose(got["U"][:, ok], port["U"][:, ok], rtol=1e-8, atol=1e-9 if store == 0 else 1e-7)
    for b in np.nonzero(ok)[0]:
        ex = bq.solve_exact(A[:, b], B[:, b], Q, R, Pf, N, x0[b], ulo, uhi, np.where(xlo < -1e19, -np.inf, xlo), xhi,
                            c=c[:, b])
        assert ex["status"] == bq.SOLVED
        assert np.abs(got["U"][:, b] - ex["U"]).max() <= 1e-6 * max(1.0, np.abs(ex["U"]).max())
        np.testing.assert_array_equal(got["sat_u"][:, b], ex["sat_u"])
        np.testing.assert_array_equal(got["sat_x"][:, b], ex["sat_x"])


def test_unconstrained_is_the_lq_solution(hh):
    """No finite bound: the QP solution is the finite-horizon LQ plan of session 1."""
    from oracle import lq
    prob = bq.Problem(N=12)
    big = 1e20
    x0 = np.array([[-3.0, 1.0], [2.0, -0.5]])
    for store in (0, 1):
        got = run_harness(hh, prob.A, prob.B, None, 0, prob.Q, prob.R, prob.Q, [-big], [big], [-big, -big], [big, big], x0, 12,
                          store=store)
        for b in range(2):
            X, U, V, _, _ = lq.lq_open_loop(prob.A, prob.B, prob.Q.astype(float), prob.R.astype(float), prob.Q.astype(float), x0[b], 12)
            np.testing.assert_allclose(got["U"][:, b], U, rtol=1e-9, atol=1e-10)
            np.testing.assert_allclose(got["cost"][b], V, rtol=1e-10)
        # one exact Newton step (both workspaces keep the gains in float64)
        assert np.all(got["status"] == bq.SOLVED) and np.all(got["iters"] == 1)


def rows_problem(rng, batch, N, nc, n=4, m=2):
    """LTV QPs with nc general stage rows Cg x_{k+1} >= hg, built so that some rows are active."""
    A, B, c = random_ltv(rng, batch, N, n, m)
    Q = np.diag(rng.uniform(0.5, 2.0, n)); R = np.diag(rng.uniform(0.05, 0.2, m)); Pf = 5 * Q
    ulo, uhi = -np.ones(m), 0.5 * np.ones(m)
    xlo, xhi = -2.0 * np.ones(n), 2.0 * np.ones(n)
    x0 = rng.uniform(-1.0, 1.0, (batch, n))
    base = bq.ipm_riccati(list(A), list(B), Q, R, Pf, N, x0, ulo, uhi, xlo, xhi, c=list(c))
    Cg = rng.standard_normal((N, batch, nc, n))
    low = -0.03 if nc <= 3 else -0.004   # a few rows cut into the box-only optimum, the rest are slack
    hg = np.einsum("kbji,kbi->kbj", Cg, base["X"][1:]) - rng.uniform(low, 0.3, (N, batch, nc))
    return A, B, c, Q, R, Pf, ulo, uhi, xlo, xhi, x0, Cg, hg


@pytest.mark.parametrize("store", [0, 1])
@pytest.mark.parametrize("nc", [3, 9])
def test_general_stage_rows(hh, nc, store):
    """Polytopic stage constraints Cg x >= hg (the linearised collision constraints of
    session_4/main.py:95-104 have this form) against the numpy restatement and the exact oracle."""
    rng = np.random.default_rng(40 + nc)
    batch, N, n, m = 10, 10, 4, 2
    A, B, c, Q, R, Pf, ulo, uhi, xlo, xhi, x0, Cg, hg = rows_problem(rng, batch, N, nc)
    U = np.zeros((N, m, batch)); X = np.zeros((N + 1, n, batch)); cost = np.zeros(batch)
    status = np.zeros(batch, dtype=np.int32); iters = np.zeros(batch, dtype=np.int32)
    su = np.zeros((N, m, batch), dtype=np.int8); sx = np.zeros((N, n, batch), dtype=np.int8)
    scn = np.zeros((N, nc, batch), dtype=np.int8)
    I8 = C.POINTER(C.c_int8); I32 = C.POINTER(C.c_int32)
    rc = hh.hh_boxqp_solve_rows(p(c_(A.reshape(N, batch, n * n).transpose(0, 2, 1))), p(c_(B.reshape(N, batch, n * m).transpose(0, 2, 1))),
                                p(c_(c.transpose(0, 2, 1))), 1, p(c_(Q)), p(c_(R)), p(c_(Pf)), p(c_(ulo)), p(c_(uhi)), p(c_(xlo)),
                                p(c_(xhi)), p(c_(Cg.reshape(N, batch, nc * n).transpose(0, 2, 1))), p(c_(hg.transpose(0, 2, 1))), nc,
                                p(c_(x0.T)), None, p(U), p(X), p(cost), status.ctypes.data_as(I32), iters.ctypes.data_as(I32),
                                su.ctypes.data_as(I8), sx.ctypes.data_as(I8), scn.ctypes.data_as(I8), C.c_int64(batch), n, m, N,
                                60, C.c_double(1e-9), store)
    assert rc == 0
    port = bq.ipm_riccati(list(A), list(B), Q, R, Pf, N, x0, ulo, uhi, xlo, xhi, c=list(c), Cg=Cg, hg=hg)
    np.testing.assert_array_equal(status, port["status"])
    ok = status == bq.SOLVED
    assert ok.sum() >= batch // 2
    np.testing.assert_allclose(U.transpose(0, 2, 1)[:, ok], port["U"][:, ok], rtol=1e-7, atol=1e-8)
    n_active = 0
    for b in np.nonzero(ok)[0]:
        ex = bq.solve_exact(A[:, b], B[:, b], Q, R, Pf, N, x0[b], ulo, uhi, xlo, xhi, c=c[:, b], Cg=Cg[:, b], hg=hg[:, b])
        assert ex["status"] == bq.SOLVED
        assert np.abs(U[:, :, b] - ex["U"]).max() <= 1e-6 * max(1.0, np.abs(ex["U"]).max())
        np.testing.assert_array_equal(su[:, :, b], ex["sat_u"])
        np.testing.assert_array_equal(sx[:, :, b], ex["sat_x"])
        np.testing.assert_array_equal(scn[:, :, b], ex["sat_c"])
        n_active += int(np.abs(ex["sat_c"]).sum())
        assert np.all(np.einsum("kji,ki->kj", Cg[:, b], X[1:, :, b]) >= hg[:, b] - 1e-7)   # rows hold on the rollout
    assert n_active > 0
