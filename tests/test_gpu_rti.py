"""GPU parity of the session-4 path (K5: RTI preparation, plant step, fused closed loop) against
oracle/bicycle.py -- the CPU restatement of the same RTI algorithm, with the QP solved by the exact
active-set oracle or by the numpy restatement of the GPU interior-point method.  PARITY UNPINNED by
the reference (CasADi / IPOPT / rcracers are not available); tolerance 1e-6 relative, saturation
patterns identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import bicycle as bc  # noqa: E402


@pytest.fixture(scope="module")
def mods():
    import torch
    from model_predictive_control_b200 import session4
    assert torch.cuda.is_available()
    return session4, torch


def scenarios(rng, batch):
    x0 = np.array([0.6, -0.25, 0, 0]) + rng.uniform(-0.2, 0.2, (batch, 4)) * np.array([1, 1, 0.5, 0.2])
    return x0, rng.uniform(0.7, 1.0, batch)


@pytest.mark.parametrize("integrator", ["euler", "rk4"])
def test_fused_closed_loop_vs_restatement(mods, integrator):
    s4, torch = mods
    rng = np.random.default_rng(11)
    x0, fr = scenarios(rng, 24)
    steps, N = 30, 20
    ctrl = s4.MPCController(N=N, ts=0.05, params=s4.VehicleParameters(), integrator=integrator)
    res = ctrl.closed_loop(x0, steps, friction_plant=fr)
    ref = bc.closed_loop(x0, steps, N=N, friction_plant=fr, ocp_method=integrator, qp="port")
    X = res.states.cpu().numpy().transpose(1, 0, 2); U = res.inputs.cpu().numpy().transpose(1, 0, 2)
    assert int(res.n_failed.sum()) == 0 and np.all(ref["status"] == 1)
    np.testing.assert_allclose(X, ref["X"], rtol=0, atol=1e-6 * np.abs(ref["X"]).max())
    np.testing.assert_allclose(U, ref["U"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(res.cost.cpu().numpy(), ref["cost"], rtol=1e-6)
    np.testing.assert_allclose(res.violation.cpu().numpy(), ref["viol"], atol=1e-7)
    ulo, uhi, _, _ = bc.bounds(bc.VehicleParameters())
    np.testing.assert_array_equal(U == uhi, ref["U"] == uhi)
    np.testing.assert_array_equal(U == ulo, ref["U"] == ulo)
    np.testing.assert_array_equal(res.n_saturated.cpu().numpy(), ((ref["U"] == uhi) | (ref["U"] == ulo)).sum(axis=(0, 2)))


def test_closed_loop_vs_exact_qp(mods):
    s4, torch = mods
    rng = np.random.default_rng(12)
    x0, fr = scenarios(rng, 3)
    steps, N = 15, 15
    ctrl = s4.MPCController(N=N, ts=0.05, params=s4.VehicleParameters())
    res = ctrl.closed_loop(x0, steps, friction_plant=fr)
    ref = bc.closed_loop(x0, steps, N=N, friction_plant=fr, qp="exact")
    X = res.states.cpu().numpy().transpose(1, 0, 2); U = res.inputs.cpu().numpy().transpose(1, 0, 2)
    np.testing.assert_allclose(X, ref["X"], rtol=0, atol=1e-6 * np.abs(ref["X"]).max())
    np.testing.assert_allclose(U, ref["U"], rtol=0, atol=1e-6)
    ulo, uhi, _, _ = bc.bounds(bc.VehicleParameters())
    np.testing.assert_array_equal(U == uhi, np.abs(ref["U"] - uhi) < 1e-12)
    np.testing.assert_array_equal(U == ulo, np.abs(ref["U"] - ulo) < 1e-12)


def test_reference_call_surface_exercise5(mods):
    """The reference's exercise5 (session4_sol.py:443-465) written against our module: N = 50,
    ts = 0.05, x0 = [.6, -.25, 0, 0], 100 closed-loop steps, nominal and mismatched plant."""
    s4, torch = mods
    N, ts, nstep = 50, 0.05, 100
    x0 = np.array([0.6, -0.25, 0, 0])
    bicycle = s4.KinematicBicycle(s4.VehicleParameters())
    dynamics_assumed = s4.forward_euler(bicycle, ts)
    controller = s4.MPCController(N=N, ts=ts, params=s4.VehicleParameters())
    assert controller.N == N and controller.ts == ts and set(controller.bounds) == {"lbx", "ubx", "lbg", "ubg"}
    x_model = s4.simulate(x0, dynamics_assumed, n_steps=nstep, policy=controller)
    assert isinstance(x_model, np.ndarray) and x_model.shape == (nstep + 1, 4)
    params = s4.VehicleParameters()
    params.friction *= 0.8
    dynamics_accurate = s4.exact_integration(s4.KinematicBicycle(params), ts)
    x_exact = s4.simulate(x0, dynamics_accurate, n_steps=nstep, policy=controller)
    assert x_exact.shape == (nstep + 1, 4)
    # the car is parked at the origin in both cases, and the mismatch changes the trajectory
    assert np.abs(x_model[-1]).max() < 0.05 and np.abs(x_exact[-1]).max() < 0.05
    assert np.abs(x_model - x_exact).max() > 1e-3
    ref = bc.closed_loop(x0[None], nstep, N=N, ts=ts, friction_plant=[1.0], plant_method="euler", qp="port")
    np.testing.assert_allclose(x_model, ref["X"][:, 0], rtol=0, atol=1e-6)
    ref2 = bc.closed_loop(x0[None], nstep, N=N, ts=ts, friction_plant=[0.8], plant_method="rk4", substeps=4, qp="port")
    np.testing.assert_allclose(x_exact, ref2["X"][:, 0], rtol=0, atol=1e-6)


def test_step_by_step_controller_equals_fused_loop(mods):
    """MPCController.__call__ (one prepare + QP launch per step, warm-started) driven by the generic
    simulate loop gives the same closed loop as the fused kernel."""
    s4, torch = mods
    rng = np.random.default_rng(13)
    x0, _ = scenarios(rng, 8)
    N, ts, steps = 20, 0.05, 12
    par = s4.VehicleParameters()
    ctrl = s4.MPCController(N=N, ts=ts, params=par)
    fused = ctrl.closed_loop(x0, steps, plant=s4.forward_euler(s4.KinematicBicycle(par), ts)).states.cpu().numpy()
    ctrl.reset()
    dyn = s4.forward_euler(s4.KinematicBicycle(par), ts)
    xs = [x0]
    for t in range(steps):
        u = ctrl(xs[-1])
        assert u.shape == (8, 2)
        xs.append(dyn(xs[-1], u))
    np.testing.assert_allclose(np.array(xs).transpose(1, 0, 2), fused, rtol=0, atol=1e-9)
    sol = ctrl.solve(x0[0])
    assert sol["x"].shape == (2 * N,) and ctrl.reshape_input(sol).shape == (N, 2)
    u0 = ctrl(x0[0])
    assert u0.shape == (2,)


def test_integrators_and_plant(mods):
    s4, torch = mods
    par = s4.VehicleParameters()
    rng = np.random.default_rng(14)
    x, fr = scenarios(rng, 16)
    u = rng.uniform(-0.3, 0.3, (16, 2))
    for make, meth, sub in ((s4.forward_euler, "euler", 1), (s4.runge_kutta4, "rk4", 1), (s4.exact_integration, "rk4", 4)):
        xn = make(s4.KinematicBicycle(par), 0.05)(x, u)
        np.testing.assert_allclose(xn, bc.plant_step(x, u, 0.05, bc.VehicleParameters(), 1.0, meth, sub), rtol=1e-12, atol=1e-14)
    # generic callables keep working on the host path of the factories (reference semantics)
    f = s4.runge_kutta4(lambda x_, u_: -x_ + u_, 0.1)
    np.testing.assert_allclose(f(np.ones(2), np.zeros(2)), np.exp(-0.1) * np.ones(2), rtol=1e-6)


def test_full_size_properties_cfg4(mods):
    """64k scenarios x 200 steps (BASELINE config 4): every QP solves, inputs stay in their box, the
    closed loop obeys the plant, parks the car, and a subsample matches the CPU restatement."""
    s4, torch = mods
    batch, steps, N = 1 << 16, 200, 50
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    scale = torch.tensor([1, 1, 0.5, 0.2], device="cuda", dtype=torch.float64)
    x0 = torch.tensor([0.6, -0.25, 0, 0], device="cuda", dtype=torch.float64) + \
        (torch.rand(batch, 4, generator=g, device="cuda", dtype=torch.float64) * 0.4 - 0.2) * scale
    fr = torch.rand(batch, generator=g, device="cuda", dtype=torch.float64) * 0.3 + 0.7
    par = s4.VehicleParameters()
    ctrl = s4.MPCController(N=N, ts=0.05, params=par)
    res = ctrl.closed_loop(x0, steps, friction_plant=fr)
    assert int(res.n_failed.sum()) == 0
    U, X = res.inputs, res.states
    assert float(U[..., 0].max()) <= par.max_drive and float(U[..., 0].min()) >= par.min_drive
    assert float(U[..., 1].abs().max()) <= par.max_steer
    assert float(X[:, -1].abs().max()) < 0.2  # parked: every scenario ends near the origin
    assert float(res.violation.max()) < 0.05
    idx = torch.arange(0, batch, batch // 4)[:4]
    ref = bc.closed_loop(x0[idx].cpu().numpy(), 60, N=N, friction_plant=fr[idx].cpu().numpy(), qp="port")
    np.testing.assert_allclose(X[idx, :61].cpu().numpy().transpose(1, 0, 2), ref["X"], rtol=0, atol=1e-6)
    # one plant step re-evaluated by the standalone kernel
    xt = X[:, 100].t().contiguous(); ut = U[:, 100].t().contiguous()
    xn = s4.plant_step(par, 0.05, xt, ut, friction=fr, substeps=4)
    assert torch.allclose(xn.t(), X[:, 101], rtol=1e-12, atol=1e-13)


def test_adaptive_plant_and_open_loop_comparison(mods):
    """exact_integration(adaptive=True) against scipy odeint (what the reference's exact_integration
    calls), and the numbers of the reference's compare_open_loop (session4_sol.py:65-104)."""
    s4, torch = mods
    par = s4.VehicleParameters()
    rng = np.random.default_rng(31)
    x, fr = scenarios(rng, 12)
    u = np.stack([rng.uniform(-1, 1, 12), rng.uniform(-0.384, 0.384, 12)], 1)
    gt = s4.exact_integration(s4.KinematicBicycle(par), 0.25, adaptive=True)
    np.testing.assert_allclose(gt(x, u), bc.exact_integration_odeint(x, u, 0.25, bc.VehicleParameters(), 1.0), rtol=0, atol=1e-9)
    results, errors = s4.compare_open_loop(0.05, np.array([0.6, -0.25, 0.0, 0.0]), 40)
    assert results["Ground truth"].shape == (41, 4)
    assert errors["RK 4"].max() < 1e-5 < errors["Forward Euler"].max()   # RK4 is far more accurate than Euler
    # ground truth by odeint, step by step with the same test policy
    pol = s4.build_test_policy()
    xs = [np.array([0.6, -0.25, 0.0, 0.0])]
    for t in range(40):
        xs.append(bc.exact_integration_odeint(xs[-1], pol(xs[-1], t), 0.05, bc.VehicleParameters(), 1.0)[0])
    np.testing.assert_allclose(results["Ground truth"], np.array(xs), rtol=0, atol=1e-8)


def test_prediction_bundles(mods):
    s4, torch = mods
    rng = np.random.default_rng(32)
    x0, fr = scenarios(rng, 5)
    N, steps = 15, 6
    ctrl = s4.MPCController(N=N, ts=0.05, params=s4.VehicleParameters())
    res = ctrl.closed_loop(x0, steps, friction_plant=fr, keep_predictions=True)
    assert res.X_bundle.shape == (steps, N + 1, 4, 5) and res.U_bundle.shape == (steps, N, 2, 5)
    assert res.bundle(2).shape == (steps, N + 1, 4)
    # every prediction starts at the measured closed-loop state and its first input is the applied one
    assert torch.equal(res.X_bundle[:, 0], res.X[:-1])
    assert torch.equal(res.U_bundle[:, 0], res.U)
    # the bundles are those of the step-by-step controller
    ctrl2 = s4.MPCController(N=N, ts=0.05, params=s4.VehicleParameters())
    sol = ctrl2.solve(x0)
    np.testing.assert_allclose(res.X_bundle[0].permute(2, 0, 1).cpu().numpy(), sol["state_prediction"], rtol=0, atol=1e-12)
    # a mismatched adaptive plant inside the fused loop
    params = s4.VehicleParameters(); params.friction *= 0.8
    plant = s4.exact_integration(s4.KinematicBicycle(params), 0.05, adaptive=True)
    r2 = ctrl.closed_loop(x0, steps, plant=plant)
    ref = bc.closed_loop(x0, steps, N=N, friction_plant=np.full(5, 0.8), plant_method="rk4", substeps=16, qp="port")
    np.testing.assert_allclose(r2.states.cpu().numpy().transpose(1, 0, 2), ref["X"], rtol=0, atol=1e-7)


def test_obstacle_avoidance_controller(mods):
    """the obstacle-avoidance controller of the reference's session_4/main.py
    (:29-129, protocol :241-271) as RTI with linearised collision rows, against the CPU restatement."""
    s4, torch = mods
    horizon, ts, steps = 30, 0.08, 40
    params = s4.VehicleParameters()
    x_obs = np.array([0.25, 0, 0.0, 0.0])
    x0 = np.array([[0.3, -0.1, 0.0, 0.0], [0.35, -0.12, 0.1, 0.0], [0.4, 0.12, 0.0, 0.0]])
    controller = s4.ObstacleMPCController(horizon, ts, params, s4.KinematicBicycle(params, symbolic=True), x_obs)
    assert controller.bounds["lbg"].shape == (horizon * 13,)
    sol = controller.solve(x0[0])
    assert controller.reshape_input(sol).shape == (horizon, 2) and bool(sol["success"])
    dynamics_accurate = s4.exact_integration(s4.KinematicBicycle(params), ts)
    X = s4.simulate(x0, dynamics_accurate, n_steps=steps, policy=controller)
    assert X.shape == (3, steps + 1, 4)
    ref = bc.closed_loop_obstacle(x0, x_obs, steps, N=horizon, ts=ts, qp="port")
    assert np.all(ref["status"] == 1)
    np.testing.assert_allclose(X.transpose(1, 0, 2), ref["X"], rtol=0, atol=2e-6)
    assert bc.min_clearance(X, x_obs) > -1e-3          # the covering circles do not overlap (up to linearisation error)
    assert np.abs(X[:, -1, :2]).max() < 0.08            # and the car reaches the parking spot
    # exact-QP restatement on the first steps
    ref2 = bc.closed_loop_obstacle(x0[:1], x_obs, 6, N=horizon, ts=ts, qp="exact")
    np.testing.assert_allclose(X[0, :7], ref2["X"][:, 0], rtol=0, atol=2e-6)


def test_device_kernels_against_reference_fixture(mods):
    """tests/golden/session234.json = outputs of the reference's own session-4 code (integrators of
    session4_sol.py:22-56; OCP of main.py:41-113 evaluated numerically).  The device plant steps reproduce the
    reference's integrators, and the linearisation kernels reproduce -- at the linearisation point -- the reference's
    predicted states and collision-constraint values."""
    import ctypes
    import json
    import os
    s4, torch = mods
    from model_predictive_control_b200 import _lib
    with open(os.path.join(os.path.dirname(__file__), "golden", "session234.json")) as fh:
        g = json.load(fh)
    dev = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device="cuda")
    for c in g["integrators"]:
        p = s4.VehicleParameters(); p.friction = c["friction"]
        bike = s4.KinematicBicycle(p)
        x, u = np.asarray(c["x"]), np.asarray(c["u"])
        np.testing.assert_allclose(s4.forward_euler(bike, c["ts"])(x, u), c["forward_euler"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(s4.runge_kutta4(bike, c["ts"])(x, u), c["runge_kutta4"], rtol=1e-13, atol=1e-15)
        # the reference's ground truth is odeint at its default tolerances (~1.5e-8)
        np.testing.assert_allclose(s4.exact_integration(bike, c["ts"], adaptive=True)(x, u), c["exact_integration"], rtol=0, atol=2e-8)
        np.testing.assert_allclose(s4.exact_integration(bike, c["ts"], substeps=16)(x, u), c["exact_integration"], rtol=0, atol=2e-8)
    o = g["ocp_main"]
    N, ts, x_obs = o["N"], o["ts"], np.asarray(o["x_obs"])
    p = s4.VehicleParameters()
    a_c, r_c = bc.create_cover_circles(p.length, p.width, 3)
    r2 = (2 * r_c) ** 2
    for c in o["cases"]:
        y = dev(c["x0"]).reshape(4, 1).contiguous()
        plan = dev(c["U"]).reshape(N, 2, 1).contiguous()
        warm, A, B, cc, Cg, hg = (torch.empty((N, k, 1), dtype=torch.float64, device="cuda") for k in (2, 16, 8, 4, 36, 9))
        xo = (ctypes.c_double * 4)(*[float(v) for v in x_obs])
        _lib.check(_lib.lib().mpc_bicycle_rti_prepare_obstacle(
            p.axis_rear, p.axis_front, float(p.acceleration), float(p.friction), ts, 0, float(p.length), float(p.width), xo,
            _lib.ptr(y), _lib.ptr(plan), 1, _lib.ptr(warm), _lib.ptr(A), _lib.ptr(B), _lib.ptr(cc), _lib.ptr(Cg), _lib.ptr(hg),
            1, N, _lib.MPC_F64, _lib.stream(y.device)))
        torch.cuda.synchronize()
        gref = np.asarray(c["g"]).reshape(N, 13)
        U = np.asarray(c["U"])
        A_, B_, c_ = A[:, :, 0].cpu().numpy().reshape(N, 4, 4), B[:, :, 0].cpu().numpy().reshape(N, 4, 2), cc[:, :, 0].cpu().numpy()
        x = np.asarray(c["x0"])
        for k in range(N):
            # affine model of stage k at the linearisation point reproduces the reference's Euler rollout (its g rows 0..3)
            xn = A_[k] @ x + B_[k] @ U[k] + c_[k]
            np.testing.assert_allclose(xn, gref[k, :4], rtol=1e-12, atol=1e-14)
            rows = Cg[k, :, 0].cpu().numpy().reshape(9, 4)
            vals = r2 - hg[k, :, 0].cpu().numpy() + rows @ gref[k, :4]
            np.testing.assert_allclose(vals, gref[k, 4:], rtol=1e-10, atol=1e-13)   # squared centre distances, main.py:95-104
            x = gref[k, :4]
        np.testing.assert_array_equal(warm[:, :, 0].cpu().numpy(), U)
