"""Round-2 GPU parity tests: the reference's own call flows through `simulate` (exercise 3 / 4 open-loop policies,
sessions-2/3 `policy=controller, log=ControllerLog()`), the exact stability test, the plant-model mismatch of the fused
loop, the fused obstacle loop, the float32 products (1e-4, identical saturation patterns) and the SQP rounds."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import bicycle as bc  # noqa: E402
from oracle import boxqp as bq  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def mods():
    import torch
    from model_predictive_control_b200 import FHC, boxqp, problem, problem3, session1_sol, session4
    from model_predictive_control_b200.log import ControllerLog
    assert torch.cuda.is_available()
    return dict(torch=torch, FHC=FHC, boxqp=boxqp, problem=problem, problem3=problem3, s1=session1_sol, s4=session4,
                Log=ControllerLog)


# ---------------------------------------------------------------------------------------------------------------
# reference flows through simulate
# ---------------------------------------------------------------------------------------------------------------
def test_exercise3_open_loop_policy_flow(mods):
    """Numeric part of the reference's exercise3 / exercise4 (session4_sol.py:326-420): solve the OCP once, replay
    `controls[t]` through `simulate(x0, dynamics, n_steps=N, policy=open_loop_policy)` under the assumed (forward Euler)
    and the accurate model, nominal and with friction * 0.8."""
    s4 = mods["s4"]
    N, ts = 50, 0.05
    x0 = np.array([0.6, -0.25, 0, 0])
    controller = s4.MPCController(N=N, ts=ts, params=s4.VehicleParameters())
    solution = controller.solve(x0)
    controls = controller.reshape_input(solution)
    assert controls.shape == (N, 2)

    def open_loop_policy(t):
        return controls[t]

    bicycle = s4.KinematicBicycle(s4.VehicleParameters())
    x_model = s4.simulate(x0, s4.forward_euler(bicycle, ts), n_steps=N, policy=open_loop_policy)
    x_exact = s4.simulate(x0, s4.exact_integration(bicycle, ts), n_steps=N, policy=open_loop_policy)
    params = s4.VehicleParameters()
    params.friction *= 0.8
    x_mis = s4.simulate(x0, s4.exact_integration(s4.KinematicBicycle(params), ts), n_steps=N, policy=open_loop_policy)
    assert x_model.shape == x_exact.shape == x_mis.shape == (N + 1, 4)
    par = bc.VehicleParameters()
    xe, xr, xm = x0[None], x0[None], x0[None]
    for t in range(N):
        xe = bc.plant_step(xe, controls[t][None], ts, par, 1.0, "euler")
        xr = bc.plant_step(xr, controls[t][None], ts, par, 1.0, "rk4", 4)
        xm = bc.plant_step(xm, controls[t][None], ts, par, 0.8, "rk4", 4)
        np.testing.assert_allclose(x_model[t + 1], xe[0], rtol=0, atol=1e-12)
        np.testing.assert_allclose(x_exact[t + 1], xr[0], rtol=0, atol=1e-12)
        np.testing.assert_allclose(x_mis[t + 1], xm[0], rtol=0, atol=1e-12)
    # the replayed plan under the assumed model IS the controller's own state prediction
    np.testing.assert_allclose(x_model, solution["state_prediction"], rtol=0, atol=1e-9)
    # build_test_policy (session4_sol.py:58-62): lambda y, t
    X = s4.simulate(np.zeros(4), s4.runge_kutta4(bicycle, ts), 20, policy=s4.build_test_policy())
    x = np.zeros((1, 4))
    for t in range(20):
        x = bc.plant_step(x, np.array([[1.0, 0.1 * np.sin(t)]]), ts, par, 1.0, "rk4", 1)
    np.testing.assert_allclose(X[-1], x[0], rtol=0, atol=1e-12)


@pytest.mark.parametrize("session", [2, 3])
def test_session23_simulate_with_controller_and_log(mods, session):
    """The sessions-2/3 convention: simulate(x0, dynamics, n_steps, policy=controller, log=ControllerLog())."""
    s4, Log = mods["s4"], mods["Log"]
    pm = mods["problem"] if session == 2 else mods["problem3"]
    prob = pm.Problem(N=10)
    assert (prob.p_min, prob.v_min) == ((-150, -20) if session == 2 else (-120, -50))
    ctrl = pm.LinearMPC(prob)
    dynamics = lambda x, u: prob.A @ x + prob.B @ u
    x0 = np.array([-40.0, 5.0])
    log = Log()
    X = s4.simulate(x0, dynamics, 12, policy=ctrl, log=log)
    assert X.shape == (13, 2) and len(log.solver_success) == 12 and all(bool(v) for v in log.solver_success)
    assert log.state_prediction[0].shape == (11, 2) and log.input_prediction[0].shape == (10, 1)
    oprob = bq.Problem(N=10) if session == 2 else bq.session3_problem(N=10)
    ulo, uhi, xlo, xhi = bq.problem_bounds(oprob)
    x = x0
    for t in range(12):
        ex = bq.solve_exact(oprob.A, oprob.B, oprob.Q, oprob.R, oprob.Q, 10, x, ulo, uhi, xlo, xhi)
        assert ex["status"] == bq.SOLVED
        np.testing.assert_allclose(log.input_prediction[t], ex["U"], rtol=0, atol=1e-6 * max(1.0, np.abs(ex["U"]).max()))
        x = oprob.A @ x + oprob.B @ ex["U"][0]
        np.testing.assert_allclose(X[t + 1], x, rtol=0, atol=1e-5)
    # batched states go through the same call
    Xb = s4.simulate(np.array([[-40.0, 5.0], [-10.0, 2.0]]), lambda x, u: x @ prob.A.T + u @ prob.B.T, 5, policy=ctrl, log=Log())
    assert Xb.shape == (2, 6, 2)
    np.testing.assert_allclose(Xb[0], X[:6], rtol=0, atol=1e-9)


# ---------------------------------------------------------------------------------------------------------------
# exact stability test (golden values: SURVEY Appendix A, reference session1_sol.py:114-116)
# ---------------------------------------------------------------------------------------------------------------
def test_spectral_radius_golden_and_random(mods):
    torch, FHC, s1 = mods["torch"], mods["FHC"], mods["s1"]
    A, B = FHC.get_dynamics_discrete(0.5)
    C = np.array([[1], [-2 / 3]])
    Q = C @ C.T + 1e-3 * np.eye(2)
    gold = {4: 1.124421099332, 6: 0.622535920563, 10: 0.534904684516, 20: 0.534262585786}
    for N, rho_ref in gold.items():
        _, K = FHC.ricatti_recursion(A, B, Q, np.array([0.1]), Q, N)
        rho = FHC.closed_loop_spectral_radius(A, B, K[0])
        assert abs(rho - rho_ref) < 1e-9
        assert s1.is_stable(A, B, K) == (rho_ref < 1.0)
    rng = np.random.default_rng(5)
    for n, m in ((2, 1), (4, 1), (4, 2)):
        Ab = rng.standard_normal((300, n, n)) * 0.6
        Bb = rng.standard_normal((300, n, m))
        Kb = rng.standard_normal((300, m, n)) * 0.3
        Ab[0] = np.diag(np.ones(n - 1), 1); Bb[0] = 0; Kb[0] = 0   # nilpotent: rho = 0
        Ab[1] = np.eye(n) + np.diag(np.ones(n - 1), 1) * 0.5; Bb[1] = 0  # defective, rho = 1
        rho = FHC.closed_loop_spectral_radius(Ab, Bb, Kb)
        ref = np.abs(np.linalg.eigvals(Ab + Bb @ Kb)).max(axis=1)
        np.testing.assert_allclose(rho[2:], ref[2:], rtol=1e-9)
        assert rho[0] < 1e-6 and abs(rho[1] - 1.0) < 1e-9
    rho32 = FHC.closed_loop_spectral_radius(torch.tensor(A, dtype=torch.float32, device="cuda"),
                                            torch.tensor(B, dtype=torch.float32, device="cuda"),
                                            torch.tensor(np.asarray(K[0]), dtype=torch.float32, device="cuda"))
    assert abs(float(rho32) - gold[20]) < 1e-6


# ---------------------------------------------------------------------------------------------------------------
# fused loop: plant parameters are the PLANT's
# ---------------------------------------------------------------------------------------------------------------
def test_fused_loop_uses_the_plants_parameters(mods):
    s4 = mods["s4"]
    rng = np.random.default_rng(31)
    x0 = np.array([0.6, -0.25, 0, 0]) + rng.uniform(-0.1, 0.1, (6, 4)) * np.array([1, 1, 0.5, 0.2])
    steps, N = 20, 15
    pp = s4.VehicleParameters()
    pp.acceleration *= 0.9
    pp.axis_rear *= 1.1
    pp.friction *= 0.8
    ctrl = s4.MPCController(N=N, ts=0.05, params=s4.VehicleParameters())
    plant = s4.exact_integration(s4.KinematicBicycle(pp), 0.05)
    X = s4.simulate(x0, plant, steps, policy=ctrl)                      # fused kernel
    opp = bc.VehicleParameters(acceleration=pp.acceleration, axis_rear=pp.axis_rear, friction=pp.friction)
    ref = bc.closed_loop(x0, steps, N=N, plant_par=opp, qp="port")
    np.testing.assert_allclose(X.transpose(1, 0, 2), ref["X"], rtol=0, atol=1e-6)
    nominal = bc.closed_loop(x0, steps, N=N, friction_plant=np.full(6, 0.8), qp="port")
    assert np.abs(ref["X"] - nominal["X"]).max() > 1e-3                 # the mismatch is visible
    # step-by-step path (policy called per step) gives the same loop
    ctrl2 = s4.MPCController(N=N, ts=0.05, params=s4.VehicleParameters())
    x = x0
    for t in range(steps):
        x = plant(x, ctrl2(x))
    np.testing.assert_allclose(x, X[:, -1], rtol=0, atol=1e-6)
    with pytest.raises(ValueError):
        ctrl.closed_loop(x0, 3, plant=s4.forward_euler(lambda x, u: x, 0.05))


def test_fused_obstacle_loop_matches_stepwise(mods):
    """ObstacleMPCController.closed_loop (one launch for the whole loop) against the per-step driver and the numpy
    restatement; reference protocol session_4/main.py:241-271."""
    s4, torch = mods["s4"], mods["torch"]
    horizon, ts, steps = 30, 0.08, 25
    params = s4.VehicleParameters()
    x_obs = np.array([0.25, 0, 0.0, 0.0])
    x0 = np.array([[0.3, -0.1, 0.0, 0.0], [0.35, -0.12, 0.1, 0.0], [0.4, 0.12, 0.0, 0.0], [0.3, 0.14, 0.1, 0.0]])
    ctrl = s4.ObstacleMPCController(horizon, ts, params, s4.KinematicBicycle(params, symbolic=True), x_obs)
    plant = s4.exact_integration(s4.KinematicBicycle(params), ts)
    res = ctrl.closed_loop(x0, steps, plant=plant)
    X = res.states.cpu().numpy()
    assert int(res.n_failed.sum()) == 0
    ctrl2 = s4.ObstacleMPCController(horizon, ts, params, s4.KinematicBicycle(params, symbolic=True), x_obs)
    x = x0
    for t in range(steps):
        x = plant(x, ctrl2(x))
        np.testing.assert_allclose(X[:, t + 1], x, rtol=0, atol=2e-6)
    ref = bc.closed_loop_obstacle(x0, x_obs, steps, N=horizon, ts=ts, qp="port")
    np.testing.assert_allclose(X.transpose(1, 0, 2), ref["X"], rtol=0, atol=2e-6)
    clear = res.clearance.cpu().numpy()
    a, r = bc.create_cover_circles(0.17, 0.08, 3)
    for b in range(4):
        c2 = min((X[b, t, 0] + a[i] * np.cos(X[b, t, 2]) - (x_obs[0] + a[j])) ** 2 + (X[b, t, 1] + a[i] * np.sin(X[b, t, 2]) - x_obs[1]) ** 2
                 for t in range(1, steps + 1) for i in range(3) for j in range(3)) - (2 * r) ** 2
        assert abs(clear[b] - c2) < 1e-9


# ---------------------------------------------------------------------------------------------------------------
# float32 products: north-star tolerance 1e-4, identical saturation patterns
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("which,N", [("Problem", 5), ("Problem", 30), ("Problem3", 30)])
def test_boxqp_float32_session23(mods, which, N):
    torch, problem = mods["torch"], mods["problem"]
    prob = getattr(problem, which)(N=N)
    oprob = bq.Problem(N=N) if which == "Problem" else bq.session3_problem(N=N)
    ulo, uhi, xlo, xhi = bq.problem_bounds(oprob)
    rng = np.random.default_rng(100 + N)
    nb = 160
    x0 = np.stack([rng.uniform(-100, 0, nb), rng.uniform(-10, 15, nb)], 1)
    x0 = x0.astype(np.float32).astype(np.float64)        # the float32 product sees exactly these values
    mpc32 = problem.LinearMPC(prob, dtype=torch.float32)
    res = mpc32.solve(x0)
    assert res.U.dtype == torch.float32 and res.X.dtype == torch.float32
    mpc64 = problem.LinearMPC(prob)
    r64 = mpc64.solve(x0)
    np.testing.assert_array_equal(res.status.cpu().numpy(), r64.status.cpu().numpy())
    U = res.input_prediction.double().cpu().numpy(); X = res.state_prediction.double().cpu().numpy()
    su = res.sat_u.permute(2, 0, 1).cpu().numpy(); sx = res.sat_x.permute(2, 0, 1).cpu().numpy()
    n_ok = 0
    for b in range(nb):
        ex = bq.solve_exact(oprob.A, oprob.B, oprob.Q, oprob.R, oprob.Q, N, x0[b], ulo, uhi, xlo, xhi)
        if ex["status"] != bq.SOLVED:
            assert int(res.status[b]) != bq.SOLVED or ex["status"] == bq.MAX_ITER   # (the oracle's own failure to decide)
            continue
        assert int(res.status[b]) == bq.SOLVED
        assert np.abs(U[b] - ex["U"]).max() <= 1e-4 * max(1.0, np.abs(ex["U"]).max())
        assert np.abs(X[b] - ex["X"]).max() <= 1e-4 * max(1.0, np.abs(ex["X"]).max())
        assert abs(float(res.cost[b]) - ex["cost"]) <= 1e-4 * abs(ex["cost"])
        np.testing.assert_array_equal(su[b], ex["sat_u"])      # bit-identical saturation pattern
        np.testing.assert_array_equal(sx[b], ex["sat_x"])
        assert np.all(U[b][ex["sat_u"] > 0] == uhi[0]) and np.all(U[b][ex["sat_u"] < 0] == ulo[0])
        n_ok += 1
    assert n_ok > nb // 2


def test_rti_float32_closed_loop(mods):
    s4, torch = mods["s4"], mods["torch"]
    rng = np.random.default_rng(77)
    x0 = (np.array([0.6, -0.25, 0, 0]) + rng.uniform(-0.2, 0.2, (16, 4)) * np.array([1, 1, 0.5, 0.2])).astype(np.float32).astype(np.float64)
    fr = rng.uniform(0.7, 1.0, 16).astype(np.float32).astype(np.float64)
    steps, N = 30, 20
    c32 = s4.MPCController(N=N, ts=0.05, params=s4.VehicleParameters(), dtype=torch.float32)
    r32 = c32.closed_loop(x0, steps, friction_plant=fr)
    assert r32.X.dtype == torch.float32 and int(r32.n_failed.sum()) == 0
    ref = bc.closed_loop(x0, steps, N=N, friction_plant=fr, qp="port")
    X = r32.states.double().cpu().numpy().transpose(1, 0, 2); U = r32.inputs.double().cpu().numpy().transpose(1, 0, 2)
    np.testing.assert_allclose(X, ref["X"], rtol=0, atol=1e-4 * np.abs(ref["X"]).max())
    np.testing.assert_allclose(U, ref["U"], rtol=0, atol=1e-4)
    ulo, uhi, _, _ = bc.bounds(bc.VehicleParameters())
    u32 = lambda v: np.float64(np.float32(v))
    sat_gpu = (U == u32(uhi[0])) | (U == u32(ulo[0])) | (U == u32(uhi[1])) | (U == u32(ulo[1]))
    sat_ref = (ref["U"] == uhi) | (ref["U"] == ulo)
    # identical saturation pattern wherever the reference input is not within 1e-4 of switching
    margin = np.minimum(np.abs(ref["U"] - uhi), np.abs(ref["U"] - ulo))
    clear = sat_ref | (margin > 1e-4)
    np.testing.assert_array_equal(sat_gpu[clear], sat_ref[clear])


# ---------------------------------------------------------------------------------------------------------------
# SQP rounds towards the converged OCP solution (what the reference's IPOPT call returns)
# ---------------------------------------------------------------------------------------------------------------
def test_sqp_rounds_match_restatement_and_converged_fixture(mods):
    s4 = mods["s4"]
    with open(os.path.join(HERE, "golden", "session4_nlp.json")) as fh:
        g = json.load(fh)
    x0 = np.asarray(g["x0"]); Ustar = np.asarray(g["U_star"])
    N, ts = g["N"], g["ts"]
    errs = {}
    for k in (1, 2, 40):
        ctrl = s4.MPCController(N=N, ts=ts, params=s4.VehicleParameters(), sqp_iters=k)
        U = ctrl.reshape_input(ctrl.solve(x0))
        errs[k] = float(np.abs(U - Ustar).max())
        if k == 2:
            ref = bc.closed_loop(x0[None], 1, N=N, ts=ts, qp="port", sqp_iters=k, plant_method="euler", keep_plans=True)
            np.testing.assert_allclose(U, ref["plans"][0][:, 0], rtol=0, atol=1e-6)
    assert errs[1] > 1e-2                      # one RTI QP from a cold start is far from the converged plan
    assert errs[40] < 2e-5                     # 40 globalised rounds reach it (fixture: 1.3e-6, SLSQP's own accuracy)
    # fused closed loop: RTI and 3-round SQP against the converged closed loop of the fixture
    Xs = np.asarray(g["closed_loop"]["X"]); steps = Xs.shape[0] - 1
    dev = {}
    for k, tol in ((1, 0.0), (3, 0.0), (60, 1e-8)):
        ctrl = s4.MPCController(N=N, ts=ts, params=s4.VehicleParameters(), sqp_iters=k, sqp_tol=tol)
        res = ctrl.closed_loop(x0, steps, plant=s4.forward_euler(s4.KinematicBicycle(s4.VehicleParameters()), ts))
        dev[k] = float(np.abs(res.states.cpu().numpy()[0] - Xs).max())
    table = {r["sqp_iters"]: r["max_dx"] for r in g["closed_loop_vs_converged"]}
    assert dev[1] < 0.3 and dev[3] < dev[1]          # fixture (numpy restatement): 0.165 and 0.008
    assert abs(dev[1] - table[1]) < 1e-3
    assert dev[60] < 3e-4                            # rounds to tolerance (fixture: 9.6e-5; Gauss-Newton SQP converges linearly)


# ---------------------------------------------------------------------------------------------------------------
# full-size parity of the north-star configurations
# ---------------------------------------------------------------------------------------------------------------
def test_cfg2b_krylov_kernel_vs_vectorised_oracle_16k(mods):
    """The bench kernel (lq_solve_krylov_kernel) against the oracle restatement of FHC.ricatti_recursion + rollout,
    vectorised over 16 384 scenarios of the bench distribution (bench.cfg2b_inputs_numpy)."""
    import sys
    torch = mods["torch"]
    sys.path.insert(0, os.path.dirname(HERE))
    import bench
    from model_predictive_control_b200 import lq
    from oracle import lq as olq
    batch, N = 16384, 20
    A, B, Q, R, Pf, x0 = bench.cfg2b_inputs_numpy(batch, 4321)
    dev = lambda a: torch.tensor(a, dtype=torch.float64, device="cuda")
    out = lq.lq_solve(dev(A), dev(B), dev(Q), dev(R), dev(Pf), dev(x0), N)
    assert lq.lq_solve_kernel_name(4, 1, torch.float64) == "lq_solve_krylov_kernel"
    P, K = olq.ricatti_recursion(A, B, Q, R, Pf, N)           # batched through numpy's @ (FHC.py:56-57 verbatim)
    X = [x0]; U = []
    for k in range(N):
        u = np.einsum("bij,bj->bi", K[k], X[-1]); U.append(u)
        X.append(np.einsum("bij,bj->bi", A, X[-1]) + np.einsum("bij,bj->bi", B, u))
    X, U = np.array(X), np.array(U)
    V = np.einsum("bi,bij,bj->b", x0, P[0], x0)
    scale = np.maximum(1.0, np.abs(X).max(axis=(0, 2)))
    assert (np.abs(out.U.cpu().numpy() - U).max(axis=(0, 2)) / scale).max() <= 1e-8
    assert (np.abs(out.X.cpu().numpy() - X).max(axis=(0, 2)) / scale).max() <= 1e-8
    np.testing.assert_allclose(out.V.cpu().numpy(), V, rtol=1e-8)


def test_cfg5_full_size_against_exact_oracle(mods):
    """BASELINE configs[4] at its per-GPU size (2^20 scenarios, nx=12, nu=4, N=50): size-independent properties on the
    whole batch, and 64 solved + 64 infeasible scenarios against the exact oracle (HiGHS active set + KKT refinement;
    HiGHS' own model status for infeasibility), solved in parallel on the host cores."""
    import sys
    torch, boxqp = mods["torch"], mods["boxqp"]
    sys.path.insert(0, HERE)
    from test_gpu_boxqp import cfg5_model
    A, B, Q, R = cfg5_model()
    N, batch = 50, 1 << 20
    g = torch.Generator(device="cuda"); g.manual_seed(1234 + 5)
    x0T = torch.rand(12, batch, generator=g, device="cuda", dtype=torch.float64) * 4 - 2
    dev = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device="cuda")
    Ad, Bd = dev(A), dev(B)
    ws = boxqp.BoxQpWorkspace(batch, 12, 4, N, "cuda", sat=True)
    res = boxqp.solve(Ad, Bd, dev(Q), dev(R), dev(Q), N, x0T, -1.0, 1.0, -5.0, 5.0, workspace=ws)
    torch.cuda.synchronize()
    status = res.status
    solved, infeas = status == 1, status == 3
    assert int(solved.sum()) > batch // 4 and int(infeas.sum()) > batch // 4 and int((status == 2).sum()) < 64
    # (i) the returned states are the rollout of the returned inputs, (ii) inputs inside their box, exactly,
    # (iii) solved scenarios respect the state box, (iv) the cost is the cost of (X, U)
    Xn = torch.einsum("ij,kjb->kib", Ad, res.X[:-1]) + torch.einsum("ij,kjb->kib", Bd, res.U)
    assert float((Xn - res.X[1:]).abs().max()) <= 1e-9
    assert float(res.U.abs().max()) <= 1.0
    assert float(res.X[1:, :, solved].abs().max()) <= 5.0 * (1 + 1e-6)   # the 1e-6 parity bar on the state box
    cost = (res.X[:-1] ** 2).sum(dim=(0, 1)) + 0.1 * (res.U ** 2).sum(dim=(0, 1)) + (res.X[-1] ** 2).sum(dim=0)
    assert float(((cost - res.cost).abs() / cost.clamp(min=1.0))[solved].max()) <= 1e-12
    # saturation flags are consistent with the returned inputs
    assert bool(((res.sat_u != 0) == (res.U.abs() == 1.0))[:, :, solved].all())
    idx_s = torch.nonzero(solved).flatten()[:: max(1, int(solved.sum()) // 64)][:64]
    idx_i = torch.nonzero(infeas).flatten()[:: max(1, int(infeas.sum()) // 64)][:64]
    idx = torch.cat([idx_s, idx_i])
    X0 = x0T[:, idx].t().cpu().numpy()
    ulo, uhi, xlo, xhi = -np.ones(4), np.ones(4), -5 * np.ones(12), 5 * np.ones(12)
    ex = bq.solve_exact_many(A, B, Q, R, Q, N, X0, ulo, uhi, xlo, xhi)
    U = res.input_prediction[idx].cpu().numpy(); su = res.sat_u.permute(2, 0, 1)[idx].cpu().numpy()
    sx = res.sat_x.permute(2, 0, 1)[idx].cpu().numpy()
    n_s = n_i = 0
    for j, e in enumerate(ex):
        if j < len(idx_s):
            assert e["status"] == bq.SOLVED, j
            assert np.abs(U[j] - e["U"]).max() <= 1e-6 * max(1.0, np.abs(e["U"]).max())
            np.testing.assert_array_equal(su[j], e["sat_u"])
            np.testing.assert_array_equal(sx[j], e["sat_x"])
            n_s += 1
        else:
            assert e["status"] == bq.INFEASIBLE, (j, e["status"])     # HiGHS: kInfeasible
            n_i += 1
    assert n_s >= 64 and n_i >= 64


def test_state_ordered_batches_are_bitwise_identical(mods):
    """Difficulty-sorted batches (mpc_state_order_keys + mpc_boxqp_solve_ordered): lane b solves scenario order[b];
    every scenario's arithmetic is independent of its lane, so all outputs are bitwise those of the unordered launch."""
    torch, boxqp, problem = mods["torch"], mods["boxqp"], mods["problem"]
    prob = problem.Problem(N=30)
    # several full waves and a ragged last warp: the races this guards against (a staging buffer refilled while loads
    # from it are still queued) only show when the whole machine is busy with scattered accesses
    batch = (1 << 18) - 37
    g = torch.Generator(device="cuda"); g.manual_seed(9)
    x0T = torch.stack([torch.rand(batch, generator=g, device="cuda", dtype=torch.float64) * 100 - 100,
                       torch.rand(batch, generator=g, device="cuda", dtype=torch.float64) * 25 - 10], 0).contiguous()
    dev = lambda M: torch.tensor(np.asarray(M, dtype=np.float64), device="cuda")
    A, B, Q, R = dev(prob.A), dev(prob.B), dev(prob.Q), dev(prob.R)
    bounds = ([prob.u_min], [prob.u_max], [prob.p_min, prob.v_min], [prob.p_max, prob.v_max])
    order = boxqp.state_order(x0T)
    assert order.dtype == torch.int32 and torch.equal(torch.sort(order.long()).values, torch.arange(batch, device="cuda"))
    # neighbours in the order are neighbours in state space: mean jump far below that of a random order
    xs = x0T[:, order.long()]
    jump = (xs[:, 1:] - xs[:, :-1]).abs().mean(dim=1)
    rand = (x0T[:, 1:] - x0T[:, :-1]).abs().mean(dim=1)
    assert bool((jump < 0.2 * rand).all())
    import os
    old = os.environ.get("MPC_QP_STAGED")
    os.environ["MPC_QP_STAGED"] = "0"      # the kernel without shared-memory staging is the reference of the staged one
    try:
        plain = boxqp.solve(A, B, Q, R, Q, 30, x0T, *bounds, order=None)
        keep = [t.clone() for t in (plain.U, plain.X, plain.cost, plain.status, plain.iters, plain.sat_u, plain.sat_x)]
    finally:
        if old is None:
            del os.environ["MPC_QP_STAGED"]
        else:
            os.environ["MPC_QP_STAGED"] = old
    perm = torch.randperm(batch, device="cuda", generator=g).to(torch.int32)
    for how in (None, "auto", order, perm, perm):
        res = boxqp.solve(A, B, Q, R, Q, 30, x0T, *bounds, order=how)
        for a_, b_ in zip(keep, (res.U, res.X, res.cost, res.status, res.iters, res.sat_u, res.sat_x)):
            assert torch.equal(a_, b_)
    # warps of the ordered launch are more homogeneous than those of the caller's order
    it = keep[4].double()
    wmax = lambda v: v[: batch // 32 * 32].reshape(-1, 32).max(dim=1).values.mean()
    assert float(wmax(it[order.long()])) < 0.9 * float(wmax(it))
    with pytest.raises(ValueError):
        boxqp.solve(A, B, Q, R, Q, 30, x0T, *bounds, order=order.long())
