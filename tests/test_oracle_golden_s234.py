"""Pin the restatements of sessions 2-4 (oracle/boxqp.py, oracle/bicycle.py) and the product's data classes to what
the reference itself defines there.  tests/golden/session234.json was produced by running the reference's own,
unmodified code (tests/golden/make_golden.py through oracle/ref_loader.py): the Problem data of
session_{2,3}/problem.py, session_4/parameters.py, the integrators of session4_sol.py:22-56 and the OCPs of
session4_sol.py:131-217 and main.py:41-113 evaluated numerically (cost, constraint vector, bounds).
Not pinned, because the reference does not contain it: a QP/NLP *solver* (sessions 2/3 ship none, session 4 calls
IPOPT) and the bicycle ODE of the absent rcracers package (ours, plugged into the reference's code when the
fixture was generated)."""
import json
import os

import numpy as np
import pytest

from oracle import bicycle as obc
from oracle import boxqp as obq
from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ["Ts", "Q", "R", "p_min", "p_max", "v_min", "v_max", "u_min", "u_max", "N", "A", "B", "n_state", "n_input"]


@pytest.fixture(scope="module")
def g234():
    with open(os.path.join(ROOT, "tests", "golden", "session234.json")) as fh:
        return json.load(fh)


def arr(x):
    return np.asarray(x, dtype=np.float64)


def check_problem(p, ref):
    for k in FIELDS:
        np.testing.assert_array_equal(arr(getattr(p, k)), arr(ref[k]), err_msg=k)


@pytest.mark.parametrize("session", [2, 3])
def test_problem_data(g234, session):
    ref = g234["problem"][str(session)]
    from model_predictive_control_b200 import problem as prod
    for make in ((obq.Problem if session == 2 else obq.session3_problem), (prod.Problem if session == 2 else prod.Problem3)):
        check_problem(make(), ref["default"])
        check_problem(make(N=30), ref["N30"])
        check_problem(make(Ts=0.1, N=7), ref["Ts01_N7"])
    # the bounds the restatement hands to its solvers are exactly these fields
    p = (obq.Problem if session == 2 else obq.session3_problem)()
    ulo, uhi, xlo, xhi = obq.problem_bounds(p)
    d = ref["default"]
    np.testing.assert_array_equal(ulo, [d["u_min"]]); np.testing.assert_array_equal(uhi, [d["u_max"]])
    np.testing.assert_array_equal(xlo, [d["p_min"], d["v_min"]]); np.testing.assert_array_equal(xhi, [d["p_max"], d["v_max"]])


def test_controller_log_schema(g234):
    """session_{2,3}/log.py:8-12: three list fields, empty at construction."""
    import dataclasses
    from model_predictive_control_b200.log import ControllerLog
    ours = {f.name: type(getattr(ControllerLog(), f.name)).__name__ for f in dataclasses.fields(ControllerLog)}
    assert ours == g234["log_fields"]["2"] == g234["log_fields"]["3"]
    assert all(getattr(ControllerLog(), k) == [] for k in ours)


def test_vehicle_parameters(g234):
    ref = g234["parameters"]
    from model_predictive_control_b200.parameters import VehicleParameters
    p = VehicleParameters()
    assert set(ref) == set(p.__dataclass_fields__)
    for k, v in ref.items():
        assert float(getattr(p, k)) == v, k
    o = obc.VehicleParameters()
    for k in o.__dataclass_fields__:
        assert float(getattr(o, k)) == ref[k], k


def test_integrators(g234):
    """oracle/bicycle.py's Euler / RK4 / accurate plant against session4_sol.forward_euler, runge_kutta4,
    exact_integration (run on the same ODE)."""
    par = obc.VehicleParameters()
    for c in g234["integrators"]:
        x, u, ts, fr = arr(c["x"]), arr(c["u"]), c["ts"], c["friction"]
        f = lambda x_, u_: obc.bicycle_f(x_, u_, par.axis_rear, par.axis_front, fr, par.acceleration)
        np.testing.assert_allclose(f(x, u), c["f"], rtol=1e-14, atol=1e-16)
        np.testing.assert_allclose(obc.forward_euler(f, ts)(x, u), c["forward_euler"], rtol=1e-14, atol=1e-16)
        np.testing.assert_allclose(obc.runge_kutta4(f, ts)(x, u), c["runge_kutta4"], rtol=1e-14, atol=1e-16)
        for method, key in (("euler", "forward_euler"), ("rk4", "runge_kutta4")):
            xn, A, B = obc.discretize(x[None], u[None], ts, par, fr, method=method)
            np.testing.assert_allclose(xn[0], c[key], rtol=1e-13, atol=1e-15)
            eps = 1e-6   # the Jacobians of the restatement are the derivatives of the reference's integrator
            for j in range(4):
                d = np.zeros(4); d[j] = eps
                fd = (obc.discretize((x + d)[None], u[None], ts, par, fr, method)[0][0]
                      - obc.discretize((x - d)[None], u[None], ts, par, fr, method)[0][0]) / (2 * eps)
                np.testing.assert_allclose(A[0][:, j], fd, rtol=0, atol=1e-8)
        np.testing.assert_allclose(obc.plant_step(x[None], u[None], ts, par, fr, method="rk4", substeps=1)[0],
                                   c["runge_kutta4"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(obc.plant_step(x[None], u[None], ts, par, fr, method="rk4", substeps=16)[0],
                                   c["exact_integration"], rtol=0, atol=2e-8)   # odeint's own tolerance is ~1.5e-8
        np.testing.assert_allclose(np.ravel(obc.exact_integration_odeint(x, u, ts, par, fr)), c["exact_integration"], rtol=0, atol=2e-8)   # restatement: odeint at 1e-12; the reference: odeint's defaults


def test_open_loop_protocol(g234):
    """compare_open_loop's protocol (session4_sol.py:65-104): simulate(x0, dynamics, steps, policy(y, t))."""
    o = g234["open_loop"]
    par = obc.VehicleParameters()
    f = lambda x_, u_: obc.bicycle_f(x_, u_, par.axis_rear, par.axis_front, par.friction, par.acceleration)
    for key, step in (("forward_euler", obc.forward_euler(f, o["ts"])), ("runge_kutta4", obc.runge_kutta4(f, o["ts"]))):
        x = arr(o["x0"]); X = [x]
        for t in range(o["steps"]):
            x = step(x, np.array([1.0, 0.1 * np.sin(t)]))   # build_test_policy, session4_sol.py:58-62
            X.append(x)
        np.testing.assert_allclose(np.array(X), o[key], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(g234["test_policy"], [[1.0, 0.0], [1.0, 0.1 * np.sin(1)], [1.0, 0.1 * np.sin(2.5)]], rtol=1e-15)


def rollout_cost(x0, U, ts, Q, QT, R, par):
    f = lambda x_, u_: obc.bicycle_f(x_, u_, par.axis_rear, par.axis_front, par.friction, par.acceleration)
    step = obc.forward_euler(f, ts)
    x = x0; cost = 0.0; X = []
    for u in U:
        cost += x @ Q @ x + u @ R @ u
        x = step(x, u)
        X.append(x)
    return cost + x @ QT @ x, np.array(X)


def test_session4_ocp(g234):
    """Weights, shooting order, integrator and bounds of session4_sol.MPCController.build_ocp (:131-217)."""
    o = g234["ocp_sol"]
    par = obc.VehicleParameters()
    Q, QT, R = obc.weights("sol")
    N = o["N"]
    for c in o["cases"]:
        x0, U = arr(c["x0"]), arr(c["U"])
        np.testing.assert_array_equal(c["x"], U.ravel())      # decision vector = [u_0; u_1; ...], u = (a, delta)
        np.testing.assert_array_equal(c["p"], x0)
        cost, X = rollout_cost(x0, U, o["ts"], Q, QT, R, par)
        np.testing.assert_allclose(cost, c["f"], rtol=1e-13)
        np.testing.assert_allclose(X.ravel(), c["g"], rtol=1e-13, atol=1e-15)   # g = [x_1; ...; x_N]
    ulo, uhi, xlo, xhi = obc.bounds(par)
    for key, vec, reps in (("lbx", ulo, N), ("ubx", uhi, N), ("lbg", xlo, N), ("ubg", xhi, N)):
        np.testing.assert_array_equal(o["bounds"][key], np.tile(vec, reps))
    from model_predictive_control_b200 import session4
    ctrl = session4.MPCController(N=50, ts=0.05, params=session4.VehicleParameters())
    for key in ("lbx", "ubx", "lbg", "ubg"):
        np.testing.assert_array_equal(np.ravel(ctrl.bounds[key]), o["N50_bounds"][key])
    np.testing.assert_array_equal(ctrl.Q, Q); np.testing.assert_array_equal(ctrl.QT, QT); np.testing.assert_array_equal(ctrl.R, R)


def collision_values(X, x_obs, par_len=0.17, par_w=0.08):
    a, r = obc.create_cover_circles(par_len, par_w, 3)
    ox = x_obs[0] + a * np.cos(x_obs[2]); oy = x_obs[1] + a * np.sin(x_obs[2])
    out = []
    for x in X:
        cx = x[0] + a * np.cos(x[2]); cy = x[1] + a * np.sin(x[2])
        out.append([(cx[i] - ox[j]) ** 2 + (cy[i] - oy[j]) ** 2 for i in range(3) for j in range(3)])
    return np.array(out), (2 * r) ** 2


def test_obstacle_ocp(g234):
    """main.MPCController.build_ocp (:41-113): weights, constraint layout [x_k; 9 squared centre distances] per stage,
    bounds, covering circles -- and the linearised rows of the RTI restatement reproduce the reference's constraint
    values at the linearisation point."""
    o = g234["ocp_main"]
    par = obc.VehicleParameters()
    Q, QT, R = obc.obstacle_weights()
    N, x_obs = o["N"], arr(o["x_obs"])
    a, r = obc.create_cover_circles(0.17, 0.08, 3)
    np.testing.assert_allclose(a, [c[0] for c in o["cover_circles"]["centers"]], rtol=1e-15)
    np.testing.assert_allclose(r, o["cover_circles"]["r"], rtol=1e-15)
    for c in o["cases"]:
        x0, U = arr(c["x0"]), arr(c["U"])
        cost, X = rollout_cost(x0, U, o["ts"], Q, QT, R, par)
        np.testing.assert_allclose(cost, c["f"], rtol=1e-13)
        g = arr(c["g"]).reshape(N, 13)
        np.testing.assert_allclose(X, g[:, :4], rtol=1e-13, atol=1e-15)
        vals, r2 = collision_values(X, x_obs)
        np.testing.assert_allclose(vals, g[:, 4:], rtol=1e-12, atol=1e-15)
        Cg, hg = obc.obstacle_rows(X, x_obs)              # rows C x >= h, h = r2 - g(xbar) + C xbar
        np.testing.assert_allclose(r2 - hg + np.einsum("kri,ki->kr", Cg, X), g[:, 4:], rtol=1e-11, atol=1e-14)
    ulo, uhi, xlo, xhi = obc.bounds(par)
    lbg = np.tile(np.concatenate([xlo, np.full(9, r2)]), N); ubg = np.tile(np.concatenate([xhi, np.full(9, np.inf)]), N)
    np.testing.assert_allclose(o["bounds"]["lbg"], lbg, rtol=1e-15)
    np.testing.assert_array_equal(o["bounds"]["ubg"], ubg)
    np.testing.assert_array_equal(o["bounds"]["lbx"], np.tile(ulo, N)); np.testing.assert_array_equal(o["bounds"]["ubx"], np.tile(uhi, N))
    from model_predictive_control_b200 import session4
    p = session4.VehicleParameters()
    ctrl = session4.ObstacleMPCController(N, o["ts"], p, session4.KinematicBicycle(p, symbolic=True), x_obs)
    np.testing.assert_allclose(ctrl.bounds["lbg"], o["bounds"]["lbg"], rtol=1e-15)
    np.testing.assert_array_equal(ctrl.bounds["ubg"], o["bounds"]["ubg"])
    np.testing.assert_array_equal(ctrl.Q, Q); np.testing.assert_array_equal(ctrl.QT, QT); np.testing.assert_array_equal(ctrl.R, R)
    np.testing.assert_allclose(session4.x2T(np.array([0.3, -0.1, 0.7, 0.0])), o["x2T"], rtol=1e-15)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is not mounted here")
def test_live_reference_matches_fixture(g234):
    """Where the reference is mounted the fixture is re-derived from it (guards against a stale file)."""
    for s in (2, 3):
        P = ref_loader.load_problem(s)
        check_problem(P(), g234["problem"][str(s)]["default"])
    VP = ref_loader.load_parameters()
    assert {k: float(getattr(VP(), k)) for k in VP.__dataclass_fields__} == g234["parameters"]
    sol, cs = ref_loader.load_session4("session4_sol")
    o = g234["ocp_sol"]; c = o["cases"][0]
    cs.values = {"x0": arr(c["x0"]), **{f"u_{t}": arr(c["U"])[t] for t in range(o["N"])}}
    ctrl = sol.MPCController(o["N"], o["ts"], params=VP())
    np.testing.assert_allclose(float(np.squeeze(ctrl.ipopt_solver.nlp["f"])), c["f"], rtol=1e-15)
    np.testing.assert_allclose(np.ravel(ctrl.ipopt_solver.nlp["g"]), c["g"], rtol=1e-15)
