"""GPU parity of the box-constrained linear MPC QP (K4) against the exact CPU oracle
(HiGHS active set + dense KKT refinement) and the numpy restatement of the GPU algorithm.
Tolerance 1e-6 relative on U and X (north star, fp64); saturation patterns must be identical and
saturated inputs exactly equal to their bound."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import boxqp as bq  # noqa: E402


@pytest.fixture(scope="module")
def mods():
    import torch
    from model_predictive_control_b200 import boxqp, problem, log
    assert torch.cuda.is_available()
    return boxqp, problem, log, torch


def session_x0(rng, batch):
    return np.stack([rng.uniform(-100, 0, batch), rng.uniform(-10, 15, batch)], 1)


def check_against_exact(res_U, res_X, res_cost, status, sat_u, sat_x, ex_list, ulo, uhi):
    n_checked = 0
    for b, ex in enumerate(ex_list):
        if ex["status"] == bq.MAX_ITER:  # oracle could not certify (degenerate active set)
            continue
        if ex["status"] == bq.INFEASIBLE:
            assert status[b] == bq.INFEASIBLE
            continue
        assert status[b] == bq.SOLVED
        su, sx = max(1.0, np.abs(ex["U"]).max()), max(1.0, np.abs(ex["X"]).max())
        assert np.abs(res_U[b] - ex["U"]).max() <= 1e-6 * su
        assert np.abs(res_X[b] - ex["X"]).max() <= 1e-6 * sx
        assert abs(res_cost[b] - ex["cost"]) <= 1e-9 * abs(ex["cost"])
        np.testing.assert_array_equal(sat_u[b], ex["sat_u"])
        np.testing.assert_array_equal(sat_x[b], ex["sat_x"])
        for j in range(len(ulo)):
            assert np.all(res_U[b][:, j][ex["sat_u"][:, j] > 0] == uhi[j])
            assert np.all(res_U[b][:, j][ex["sat_u"][:, j] < 0] == ulo[j])
        n_checked += 1
    return n_checked


@pytest.mark.parametrize("which,N", [("Problem", 5), ("Problem", 30), ("Problem3", 30)])
def test_session23_problem(mods, which, N):
    boxqp, problem, log, torch = mods
    prob = getattr(problem, which)(N=N)
    oprob = (bq.Problem if which == "Problem" else bq.session3_problem)(N=N)
    rng = np.random.default_rng(100 + N)
    batch = 200
    x0 = session_x0(rng, batch)
    x0[0] = [-100.0, 0.0]
    x0[1] = [-1.0, 14.0]
    mpc = problem.LinearMPC(prob)
    res = mpc.solve(x0)
    assert res.input_prediction.shape == (batch, N, 1) and res.state_prediction.shape == (batch, N + 1, 2)
    U = res.input_prediction.cpu().numpy(); X = res.state_prediction.cpu().numpy()
    status = res.status.cpu().numpy()
    ulo, uhi, xlo, xhi = bq.problem_bounds(oprob)
    ex = [bq.solve_exact(oprob.A, oprob.B, oprob.Q, oprob.R, oprob.Q, N, x0[b], ulo, uhi, xlo, xhi) for b in range(batch)]
    n = check_against_exact(U, X, res.cost.cpu().numpy(), status, res.sat_u.permute(2, 0, 1).cpu().numpy(),
                            res.sat_x.permute(2, 0, 1).cpu().numpy(), ex, ulo, uhi)
    assert n >= batch * 0.9
    assert status[1] == bq.INFEASIBLE and not bool(res.solver_success[1])
    if N == 5:
        np.testing.assert_array_equal(U[0, :, 0], [10, 10, 10, 10, -20])
    # numpy restatement of the same algorithm: same statuses and iteration counts
    port = bq.ipm_riccati(oprob.A, oprob.B, oprob.Q, oprob.R, oprob.Q, N, x0, ulo, uhi, xlo, xhi)
    np.testing.assert_array_equal(status, port["status"])
    assert np.abs(res.iters.cpu().numpy() - port["iters"]).max() <= 1


def test_policy_call_and_log(mods):
    boxqp, problem, log, torch = mods
    prob = problem.Problem(N=10)
    mpc = problem.LinearMPC(prob)
    lg = log.ControllerLog()
    u = mpc(np.array([-50.0, 5.0]), lg)
    assert u.shape == (1,) and isinstance(u, np.ndarray)
    assert len(lg.solver_success) == 1 and bool(lg.solver_success[0])
    assert lg.state_prediction[0].shape == (11, 2) and lg.input_prediction[0].shape == (10, 1)
    np.testing.assert_allclose(lg.input_prediction[0][0], u)
    X, U = problem.closed_loop(prob, np.array([[-50.0, 5.0], [-20.0, 0.0]]), 15, mpc, lg)
    assert X.shape == (2, 16, 2) and U.shape == (2, 15, 1) and len(lg.solver_success) == 16
    # closed loop obeys the plant and the bounds
    np.testing.assert_allclose(X[:, 1:], X[:, :-1] @ prob.A.T + U @ prob.B.T, rtol=1e-12, atol=1e-12)
    assert U.max() <= prob.u_max and U.min() >= prob.u_min and X[:, :, 0].max() <= prob.p_max + 1e-9


@pytest.mark.parametrize("n,m", [(2, 1), (4, 1), (4, 2)])
def test_ltv_per_scenario(mods, n, m):
    boxqp, problem, log, torch = mods
    rng = np.random.default_rng(n * 10 + m)
    batch, N = 40, 15
    A0 = np.eye(n) + 0.1 * np.diag(np.ones(n - 1), 1)
    B0 = np.zeros((n, m)); B0[-1, 0] = 0.1
    if m > 1:
        B0[-2, 1] = 0.1
    A = A0 + 0.02 * rng.standard_normal((N, batch, n, n))
    B = B0 + 0.02 * rng.standard_normal((N, batch, n, m))
    c = 0.01 * rng.standard_normal((N, batch, n))
    Q = np.diag(rng.uniform(0.5, 2.0, n)); R = np.diag(rng.uniform(0.05, 0.2, m)); Pf = 5 * Q
    ulo, uhi = -np.ones(m), 0.5 * np.ones(m)
    xlo, xhi = -2.0 * np.ones(n), 2.0 * np.ones(n)
    xlo[0] = -np.inf
    x0 = rng.uniform(-1.5, 1.5, (batch, n))
    warm = rng.uniform(-2, 2, (N, batch, m))
    dev = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
    res = boxqp.solve(dev(A.reshape(N, batch, n * n).transpose(0, 2, 1)), dev(B.reshape(N, batch, n * m).transpose(0, 2, 1)),
                      dev(Q), dev(R), dev(Pf), N, dev(x0.T), ulo, uhi, xlo, xhi, c=dev(c.transpose(0, 2, 1)),
                      warm_U=dev(warm.transpose(0, 2, 1)))
    ex = [bq.solve_exact(A[:, b], B[:, b], Q, R, Pf, N, x0[b], ulo, uhi, xlo, xhi, c=c[:, b]) for b in range(batch)]
    nchk = check_against_exact(res.input_prediction.cpu().numpy(), res.state_prediction.cpu().numpy(), res.cost.cpu().numpy(),
                               res.status.cpu().numpy(), res.sat_u.permute(2, 0, 1).cpu().numpy(),
                               res.sat_x.permute(2, 0, 1).cpu().numpy(), ex, ulo, uhi)
    assert nchk >= batch // 2


@pytest.mark.parametrize("n,m", [(4, 1), (4, 2)])
def test_lti_4x_staged_kernel_against_exact_and_unstaged(mods, n, m):
    """Shared LTI models of the (4,1) / (4,2) shapes run on the staged kernel (tile stages copied into shared memory by
    cp.async.bulk): against the exact oracle on a subsample, and bitwise against the same solve without staging
    (MPC_QP_STAGED=0), ragged batch, caller's order and a permuted one."""
    import os
    boxqp, problem, log, torch = mods
    rng = np.random.default_rng(400 + n * 10 + m)
    batch, N = 9000 + 13, 25
    A = np.eye(n) + 0.1 * np.diag(np.ones(n - 1), 1) + 0.01 * rng.standard_normal((n, n))
    B = np.zeros((n, m)); B[-1, 0] = 0.1
    if m > 1:
        B[-2, 1] = 0.1
    Q = np.diag(rng.uniform(0.5, 2.0, n)); R = np.diag(rng.uniform(0.05, 0.2, m)); Pf = 5 * Q
    ulo, uhi = -np.ones(m), 0.5 * np.ones(m)
    xlo, xhi = -2.0 * np.ones(n), 2.0 * np.ones(n)
    xlo[0] = -np.inf
    x0 = rng.uniform(-1.4, 1.4, (batch, n))
    dev = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
    args = (dev(A), dev(B), dev(Q), dev(R), dev(Pf), N, dev(x0.T), ulo, uhi, xlo, xhi)
    res = boxqp.solve(*args, order=None)
    keep = [t.clone() for t in (res.U, res.X, res.cost, res.status, res.iters, res.sat_u, res.sat_x)]
    sub = rng.choice(batch, 48, replace=False)
    ex = [bq.solve_exact(A, B, Q, R, Pf, N, x0[b], ulo, uhi, xlo, xhi) for b in sub]
    idx = torch.tensor(sub, device="cuda")
    nchk = check_against_exact(res.input_prediction[idx].cpu().numpy(), res.state_prediction[idx].cpu().numpy(),
                               res.cost[idx].cpu().numpy(), res.status[idx].cpu().numpy(),
                               res.sat_u.permute(2, 0, 1)[idx].cpu().numpy(), res.sat_x.permute(2, 0, 1)[idx].cpu().numpy(),
                               ex, ulo, uhi)
    assert nchk >= 16   # solved and certified by the oracle (the rest: infeasible starts, flagged by both)
    perm = torch.randperm(batch, device="cuda").to(torch.int32)
    old = os.environ.get("MPC_QP_STAGED")
    try:
        for staged, order in (("1", perm), ("0", None), ("0", perm)):
            os.environ["MPC_QP_STAGED"] = staged
            r2 = boxqp.solve(*args, order=order)
            for a_, b_ in zip(keep, (r2.U, r2.X, r2.cost, r2.status, r2.iters, r2.sat_u, r2.sat_x)):
                assert torch.equal(a_, b_)
    finally:
        if old is None:
            os.environ.pop("MPC_QP_STAGED", None)
        else:
            os.environ["MPC_QP_STAGED"] = old


def test_unconstrained_equals_session1_lq(mods):
    """With every bound at infinity the QP is the finite-horizon LQ problem of session 1."""
    boxqp, problem, log, torch = mods
    from model_predictive_control_b200 import lq
    prob = problem.Problem(N=12)
    rng = np.random.default_rng(3)
    x0 = rng.uniform(-5, 5, (64, 2))
    dev = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device="cuda")
    res = boxqp.solve(dev(prob.A), dev(prob.B), dev(prob.Q), dev(prob.R), dev(prob.Q), 12, dev(x0.T.copy()),
                      [-np.inf], [np.inf], [-np.inf] * 2, [np.inf] * 2)
    out = lq.lq_solve(dev(prob.A), dev(prob.B), dev(prob.Q), dev(prob.R), dev(prob.Q), dev(x0), 12)
    assert torch.allclose(res.input_prediction, out.U.permute(1, 0, 2), rtol=1e-9, atol=1e-10)
    assert torch.allclose(res.cost, out.V, rtol=1e-10)
    assert int(res.iters.max()) == 1 and bool(res.solver_success.all())


def test_full_size_properties_cfg3(mods):
    """256k scenarios, N = 30 (BASELINE config 3): size-independent properties -- predictions obey
    the dynamics and the bounds, saturated inputs sit exactly on the bound, the cost equals the
    re-evaluated objective, infeasible starts are flagged, and a subsample matches the exact oracle."""
    boxqp, problem, log, torch = mods
    prob = problem.Problem(N=30)
    batch = 1 << 18
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    x0 = torch.stack([torch.rand(batch, generator=g, device="cuda", dtype=torch.float64) * 100 - 100,
                      torch.rand(batch, generator=g, device="cuda", dtype=torch.float64) * 25 - 10], dim=1)
    res = problem.LinearMPC(prob).solve(x0)
    ok = res.solver_success
    frac_inf = float((res.status == 3).double().mean())
    assert 0.0 < frac_inf < 0.1 and int((res.status == 2).sum()) == 0
    U, X = res.input_prediction[ok], res.state_prediction[ok]
    A = torch.tensor(prob.A, device="cuda"); B = torch.tensor(prob.B, device="cuda")
    assert torch.allclose(X[:, 1:], X[:, :-1] @ A.t() + U @ B.t(), rtol=1e-12, atol=1e-10)
    assert float(U.max()) <= prob.u_max and float(U.min()) >= prob.u_min
    tol = 1e-7
    assert float(X[:, 1:, 0].max()) <= prob.p_max + tol and float(X[:, 1:, 1].max()) <= prob.v_max + tol
    assert float(X[:, 1:, 0].min()) >= prob.p_min - tol and float(X[:, 1:, 1].min()) >= prob.v_min - tol
    sat = res.sat_u.permute(2, 0, 1)[ok]
    assert bool((U[sat > 0] == prob.u_max).all()) and bool((U[sat < 0] == prob.u_min).all())
    Q = torch.tensor(np.asarray(prob.Q, float), device="cuda"); R = torch.tensor(np.asarray(prob.R, float), device="cuda")
    J = torch.einsum("bki,ij,bkj->b", X[:, :-1], Q, X[:, :-1]) + torch.einsum("bki,ij,bkj->b", U, R, U) \
        + torch.einsum("bi,ij,bj->b", X[:, -1], Q, X[:, -1])
    assert torch.allclose(res.cost[ok], J, rtol=1e-11)
    idx = torch.arange(0, batch, batch // 64)[:64]
    oprob = bq.Problem(N=30)
    ulo, uhi, xlo, xhi = bq.problem_bounds(oprob)
    x0h = x0[idx].cpu().numpy()
    ex = [bq.solve_exact(oprob.A, oprob.B, oprob.Q, oprob.R, oprob.Q, 30, x0h[i], ulo, uhi, xlo, xhi) for i in range(64)]
    check_against_exact(res.input_prediction[idx].cpu().numpy(), res.state_prediction[idx].cpu().numpy(),
                        res.cost[idx].cpu().numpy(), res.status[idx].cpu().numpy(),
                        res.sat_u.permute(2, 0, 1)[idx].cpu().numpy(), res.sat_x.permute(2, 0, 1)[idx].cpu().numpy(), ex, ulo, uhi)


def test_errors(mods):
    boxqp, problem, log, torch = mods
    prob = problem.Problem(N=5)
    dev = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device="cuda")
    x0 = dev(np.zeros((2, 4)))
    with pytest.raises(ValueError):
        boxqp.solve(dev(prob.A), dev(prob.B), dev(prob.Q), dev(prob.R), dev(prob.Q), 5, x0, [1.0], [-1.0], [-1, -1], [1, 1])
    with pytest.raises(ValueError):
        boxqp.solve(dev(prob.A), dev(prob.B), dev(prob.Q), dev(prob.R), dev(prob.Q), 5, x0.float(), [-1.0], [1.0], [-1, -1], [1, 1])
    from model_predictive_control_b200 import _lib
    with pytest.raises(_lib.MpcError):  # (n, m) without an instantiated kernel
        boxqp.solve(dev(np.eye(3)), dev(np.ones((3, 1))), dev(np.eye(3)), dev([[1.0]]), dev(np.eye(3)), 5, dev(np.zeros((3, 4))),
                    [-1.0], [1.0], [-1] * 3, [1] * 3)


@pytest.mark.parametrize("n,m,N", [(2, 1, 5), (2, 1, 30), (4, 2, 20), (12, 4, 50)])
def test_condense_matches_oracle(mods, n, m, N):
    boxqp, problem, log, torch = mods
    rng = np.random.default_rng(n + m + N)
    if n == 2:
        p = bq.Problem(N=N)
        A, B, Q, R, Pf = p.A, p.B, np.asarray(p.Q, float), np.asarray(p.R, float), 3.0 * np.asarray(p.Q, float)
    else:
        A = np.eye(n) + 0.1 * rng.standard_normal((n, n)); B = rng.standard_normal((n, m))
        Q = np.diag(rng.uniform(0.5, 2, n)); R = np.diag(rng.uniform(0.05, 0.5, m)); Pf = 2.0 * Q
    dev = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device="cuda")
    Phi, Gam, H, F = boxqp.condense(dev(A), dev(B), dev(Q), dev(R), dev(Pf), N)
    Po, Go, Ho, Fo = bq.condense(A, B, Q, R, Pf, N)
    for got, ref in ((Phi, Po), (Gam, Go), (H, Ho), (F, Fo)):
        assert got.shape == ref.shape
        assert np.abs(got.cpu().numpy() - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
    # batched models: one CTA per model
    As = np.stack([A, A + 0.01]); Bs = np.stack([B, 2 * B])
    Phi2, Gam2, H2, F2 = boxqp.condense(dev(As), dev(Bs), dev(Q), dev(R), dev(Pf), N)
    Ho2 = bq.condense(As[1], Bs[1], Q, R, Pf, N)[2]
    assert np.abs(H2[1].cpu().numpy() - Ho2).max() <= 1e-9 * max(1.0, np.abs(Ho2).max())
    assert torch.equal(H2[0], H)


def test_condensed_qp_agrees_with_solver(mods):
    """The condensed matrices and the structure-exploiting solver describe the same QP: at the solver's
    optimum the projected-gradient residual of the condensed problem vanishes on the free inputs."""
    boxqp, problem, log, torch = mods
    prob = problem.Problem(N=30)
    dev = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device="cuda")
    Phi, Gam, H, F = boxqp.condense(dev(prob.A), dev(prob.B), dev(prob.Q), dev(prob.R), dev(prob.Q), 30)
    rng = np.random.default_rng(0)
    x0 = np.stack([rng.uniform(-100, 0, 64), rng.uniform(-10, 15, 64)], 1)
    res = problem.LinearMPC(prob).solve(x0)
    ok = res.solver_success & (res.sat_x.abs().sum(dim=(0, 1)) == 0)   # no active state bound
    U = res.input_prediction[ok][:, :, 0]
    g = 2 * (U @ H + dev(x0)[ok] @ F.t())            # gradient of U'HU + 2 x0'F'U
    free = res.sat_u.permute(2, 0, 1)[ok][:, :, 0] == 0
    assert float(g[free].abs().max()) <= 1e-6 * float(g.abs().max())
    X = res.state_prediction[ok][:, 1:].reshape(U.shape[0], -1)
    assert torch.allclose(X, dev(x0)[ok] @ Phi.t() + U @ Gam.t(), rtol=1e-10, atol=1e-9)


def test_summary_kernel(mods):
    boxqp, problem, log, torch = mods
    from model_predictive_control_b200 import distributed as D
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    b = 100003
    cost = torch.rand(b, generator=g, device="cuda", dtype=torch.float64)
    viol = torch.rand(b, generator=g, device="cuda", dtype=torch.float64)
    nsat = torch.randint(0, 50, (b,), generator=g, device="cuda", dtype=torch.int32)
    status = torch.randint(1, 4, (b,), generator=g, device="cuda", dtype=torch.int32)
    iters = torch.randint(1, 60, (b,), generator=g, device="cuda", dtype=torch.int32)
    s = D.merge_summaries(D.local_summary(cost, viol, nsat, status, iters)[None].cpu())
    assert s["scenarios"] == b and abs(s["sum_cost"] - float(cost.sum())) <= 1e-9 * b
    assert s["max_violation"] == float(viol.max()) and s["sum_saturated"] == int(nsat.sum())
    assert s["n_infeasible"] == int((status == 3).sum()) and s["n_max_iter"] == int((status == 2).sum())
    assert s["n_solved"] == int((status == 1).sum()) and s["sum_iters"] == int(iters.sum())


def cfg5_model(seed=1234 + 5):
    """BASELINE config 5: four triple-integrator chains at Ts = 0.1 with weak
    seeded coupling, Q = I, R = 0.1 I, |u| <= 1, |x_i| <= 5."""
    rng = np.random.default_rng(seed)
    Ts = 0.1
    Ac = np.array([[1, Ts, Ts * Ts / 2], [0, 1, Ts], [0, 0, 1.0]])
    Bc = np.array([[Ts**3 / 6], [Ts * Ts / 2], [Ts]])
    A = np.kron(np.eye(4), Ac) + 0.01 * rng.standard_normal((12, 12))
    B = np.kron(np.eye(4), Bc)
    return A, B, np.eye(12), 0.1 * np.eye(4)


def test_cfg5_shape_n12_m4_N50(mods):
    boxqp, problem, log, torch = mods
    A, B, Q, R = cfg5_model()
    N, batch = 50, 48
    rng = np.random.default_rng(9)
    x0 = rng.uniform(-2, 2, (batch, 12))
    ulo, uhi, xlo, xhi = -np.ones(4), np.ones(4), -5 * np.ones(12), 5 * np.ones(12)
    dev = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device="cuda")
    res = boxqp.solve(dev(A), dev(B), dev(Q), dev(R), dev(Q), N, dev(x0.T.copy()), ulo, uhi, xlo, xhi)
    ex = [bq.solve_exact(A, B, Q, R, Q, N, x0[b], ulo, uhi, xlo, xhi) for b in range(12)]
    nchk = check_against_exact(res.input_prediction[:12].cpu().numpy(), res.state_prediction[:12].cpu().numpy(),
                               res.cost[:12].cpu().numpy(), res.status[:12].cpu().numpy(),
                               res.sat_u.permute(2, 0, 1)[:12].cpu().numpy(), res.sat_x.permute(2, 0, 1)[:12].cpu().numpy(),
                               ex, ulo, uhi)
    assert nchk >= 4   # the rest of the first dozen starts outside the feasible set and is flagged infeasible
    port = bq.ipm_riccati(A, B, Q, R, Q, N, x0, ulo, uhi, xlo, xhi)
    np.testing.assert_array_equal(res.status.cpu().numpy(), port["status"])
    ok = port["status"] == 1
    assert np.abs(res.input_prediction.cpu().numpy()[ok] - port["U"].transpose(1, 0, 2)[ok]).max() <= 1e-6


def test_edge_cases_boxqp(mods):
    boxqp, problem, log, torch = mods
    prob = problem.Problem(N=1)                      # horizon 1
    res = problem.LinearMPC(prob).solve(np.array([-10.0, 40.0]))   # v_1 >= 40 - 0.3*20 = 34 > v_max: infeasible
    assert int(res.status[0]) == bq.INFEASIBLE
    res = problem.LinearMPC(prob).solve(np.array([-10.0, 3.0]))
    ex = bq.solve_exact(prob.A, prob.B, prob.Q, prob.R, prob.Q, 1, np.array([-10.0, 3.0]), *bq.problem_bounds(bq.Problem(N=1)))
    np.testing.assert_allclose(res.input_prediction[0].cpu().numpy(), ex["U"], rtol=1e-6, atol=1e-9)
    # input bounds only (state bounds at infinity), one-sided state bound
    prob = problem.Problem(N=10)
    dev = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device="cuda")
    x0 = np.array([[-30.0, 8.0], [-5.0, 2.0]])
    res = boxqp.solve(dev(prob.A), dev(prob.B), dev(prob.Q), dev(prob.R), dev(prob.Q), 10, dev(x0.T.copy()),
                      [prob.u_min], [prob.u_max], [-np.inf, -np.inf], [prob.p_max, np.inf])
    for b in range(2):
        ex = bq.solve_exact(prob.A, prob.B, prob.Q.astype(float), prob.R.astype(float), prob.Q.astype(float), 10, x0[b],
                            np.array([prob.u_min]), np.array([prob.u_max]), np.array([-np.inf, -np.inf]),
                            np.array([prob.p_max, np.inf]))
        assert ex["status"] == bq.SOLVED and int(res.status[b]) == bq.SOLVED
        np.testing.assert_allclose(res.input_prediction[b].cpu().numpy(), ex["U"], rtol=1e-6, atol=1e-7)
        np.testing.assert_array_equal(res.sat_u.permute(2, 0, 1)[b].cpu().numpy(), ex["sat_u"])
    # batch of one, and a batch that is not a multiple of the CTA size
    for batch in (1, 129):
        xb = np.tile(np.array([[-50.0, 5.0]]), (batch, 1))
        r = problem.LinearMPC(problem.Problem(N=5)).solve(xb)
        assert r.input_prediction.shape == (batch, 5, 1) and bool(r.solver_success.all())
        assert torch.equal(r.U[:, :, 0], r.U[:, :, -1])


@pytest.mark.parametrize("nc", [3, 9])
def test_general_stage_rows_gpu(mods, nc):
    """K4 with polytopic stage constraints Cg x_{k+1} >= hg against the exact oracle."""
    boxqp, problem, log, torch = mods
    from test_host_harness_boxqp import rows_problem
    rng = np.random.default_rng(60 + nc)
    batch, N, n, m = 48, 12, 4, 2
    A, B, c, Q, R, Pf, ulo, uhi, xlo, xhi, x0, Cg, hg = rows_problem(rng, batch, N, nc)
    dev = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
    res = boxqp.solve(dev(A.reshape(N, batch, n * n).transpose(0, 2, 1)), dev(B.reshape(N, batch, n * m).transpose(0, 2, 1)),
                      dev(Q), dev(R), dev(Pf), N, dev(x0.T), ulo, uhi, xlo, xhi, c=dev(c.transpose(0, 2, 1)),
                      Cg=dev(Cg.reshape(N, batch, nc * n).transpose(0, 2, 1)), hg=dev(hg.transpose(0, 2, 1)))
    status = res.status.cpu().numpy()
    port = bq.ipm_riccati(list(A), list(B), Q, R, Pf, N, x0, ulo, uhi, xlo, xhi, c=list(c), Cg=Cg, hg=hg)
    np.testing.assert_array_equal(status, port["status"])
    U = res.input_prediction.cpu().numpy(); satc = res.sat_c.permute(2, 0, 1).cpu().numpy()
    nact = 0
    for b in range(0, batch, 3):
        ex = bq.solve_exact(A[:, b], B[:, b], Q, R, Pf, N, x0[b], ulo, uhi, xlo, xhi, c=c[:, b], Cg=Cg[:, b], hg=hg[:, b])
        if ex["status"] != bq.SOLVED:
            assert status[b] != bq.SOLVED or ex["status"] == bq.MAX_ITER
            continue
        assert status[b] == bq.SOLVED
        assert np.abs(U[b] - ex["U"]).max() <= 1e-6 * max(1.0, np.abs(ex["U"]).max())
        np.testing.assert_array_equal(satc[b], ex["sat_c"])
        np.testing.assert_array_equal(res.sat_u.permute(2, 0, 1)[b].cpu().numpy(), ex["sat_u"])
        nact += int(np.abs(ex["sat_c"]).sum())
    assert nact > 0


@pytest.mark.parametrize("shape", ["lti21", "ltv42", "lti124"])
def test_boxqp_writes_stay_inside_buffers(mods, shape):
    """Guard bands around every output and around the solver workspace (ragged batch): the kernels write nothing
    outside what mpc_boxqp_workspace_bytes / the output shapes promise."""
    boxqp, problem, log, torch = mods
    rng = np.random.default_rng(99)
    pad = 512
    dd = dict(dtype=torch.float64, device="cuda")
    if shape == "lti21":
        prob = problem.Problem(N=9)
        n, m, N, batch = 2, 1, 9, 193
        A, B = torch.tensor(prob.A, **dd), torch.tensor(prob.B, **dd)
        Q, R = torch.tensor(prob.Q.astype(float), **dd), torch.tensor(prob.R.astype(float), **dd)
        lo_u, hi_u, lo_x, hi_x = [prob.u_min], [prob.u_max], [prob.p_min, prob.v_min], [prob.p_max, prob.v_max]
        x0 = torch.tensor(session_x0(rng, batch).T.copy(), **dd)
        kw = {}
    elif shape == "ltv42":
        n, m, N, batch = 4, 2, 11, 161
        A0 = np.eye(n) + 0.1 * np.diag(np.ones(n - 1), 1)
        B0 = np.zeros((n, m)); B0[-1, 0] = 0.1; B0[-2, 1] = 0.1
        A = torch.tensor((A0 + 0.02 * rng.standard_normal((N, batch, n, n))).transpose(0, 2, 3, 1).reshape(N, n * n, batch).copy(), **dd)
        B = torch.tensor((B0 + 0.02 * rng.standard_normal((N, batch, n, m))).transpose(0, 2, 3, 1).reshape(N, n * m, batch).copy(), **dd)
        kw = {"c": torch.tensor(0.01 * rng.standard_normal((N, n, batch)), **dd)}
        Q, R = torch.eye(n, **dd), 0.1 * torch.eye(m, **dd)
        lo_u, hi_u, lo_x, hi_x = [-1.0] * m, [0.5] * m, [-2.0] * n, [2.0] * n
        x0 = torch.tensor(rng.uniform(-1.5, 1.5, (n, batch)), **dd)
    else:
        n, m, N, batch = 12, 4, 6, 77
        Ts = 0.1
        Ac = np.array([[1, Ts, Ts * Ts / 2], [0, 1, Ts], [0, 0, 1.0]]); Bc = np.array([[Ts**3 / 6], [Ts * Ts / 2], [Ts]])
        A = torch.tensor(np.kron(np.eye(4), Ac) + 0.01 * rng.standard_normal((12, 12)), **dd)
        B = torch.tensor(np.kron(np.eye(4), Bc), **dd)
        Q, R = torch.eye(n, **dd), 0.1 * torch.eye(m, **dd)
        lo_u, hi_u, lo_x, hi_x = [-1.0] * m, [1.0] * m, [-5.0] * n, [5.0] * n
        x0 = torch.tensor(rng.uniform(-1, 1, (n, batch)), **dd)
        kw = {}
    ws = boxqp.BoxQpWorkspace(batch, n, m, N, "cuda")
    guards = {}
    for name in ("ws", "U", "X", "cost", "status", "iters", "sat_u", "sat_x"):
        t = getattr(ws, name)
        fill = -7 if t.dtype in (torch.int8, torch.int32) else -7.25
        buf = torch.full((t.numel() + 2 * pad,), fill, dtype=t.dtype, device="cuda")
        setattr(ws, name, buf[pad:pad + t.numel()].view(t.shape))
        guards[name] = (buf, fill)
    res = boxqp.solve(A, B, Q, R, Q, N, x0, lo_u, hi_u, lo_x, hi_x, workspace=ws, **kw)
    torch.cuda.synchronize()
    for name, (buf, fill) in guards.items():
        assert bool((buf[:pad] == fill).all()) and bool((buf[-pad:] == fill).all()), name
    assert int((res.status == 0).sum()) == 0                      # every scenario got a status
    for name in ("U", "X", "cost"):
        assert not bool((getattr(res, name) == -7.25).any()), name
