"""csrc/bicycle_core.cuh (K5 bodies: RTI preparation, plant step, fused closed loop) run on the CPU
by tests/harness, against oracle/bicycle.py."""
import ctypes as C

import numpy as np
import pytest

from oracle import bicycle as bc

P64 = C.POINTER(C.c_double)
P32 = C.POINTER(C.c_int32)


def p(a):
    return None if a is None else a.ctypes.data_as(P64)


def c_(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def scenarios(rng, batch):
    x0 = np.array([0.6, -0.25, 0, 0]) + rng.uniform(-0.2, 0.2, (batch, 4)) * np.array([1, 1, 0.5, 0.2])
    return x0, rng.uniform(0.7, 1.0, batch)


def test_jacobians_match_finite_differences():
    par = bc.VehicleParameters()
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.5, 0.5, (6, 4)); u = rng.uniform(-0.3, 0.3, (6, 2))
    for meth in ("euler", "rk4"):
        _, A, B = bc.discretize(x, u, 0.05, par, 1.0, meth)
        eps = 1e-6
        for j in range(4):
            d = np.zeros(4); d[j] = eps
            fd = (bc.discretize(x + d, u, 0.05, par, 1.0, meth)[0] - bc.discretize(x - d, u, 0.05, par, 1.0, meth)[0]) / (2 * eps)
            np.testing.assert_allclose(A[:, :, j], fd, atol=1e-8)
        for j in range(2):
            d = np.zeros(2); d[j] = eps
            fd = (bc.discretize(x, u + d, 0.05, par, 1.0, meth)[0] - bc.discretize(x, u - d, 0.05, par, 1.0, meth)[0]) / (2 * eps)
            np.testing.assert_allclose(B[:, :, j], fd, atol=1e-8)


@pytest.mark.parametrize("rk4", [0, 1])
@pytest.mark.parametrize("first", [0, 1])
def test_rti_prepare_body(hh, rk4, first):
    par = bc.VehicleParameters()
    rng = np.random.default_rng(1)
    batch, N, ts = 5, 12, 0.05
    y, _ = scenarios(rng, batch)
    Uprev = rng.uniform(-0.3, 0.3, (N, batch, 2))
    warm = np.zeros((N, 2, batch)); A = np.zeros((N, 16, batch)); B = np.zeros((N, 8, batch)); c = np.zeros((N, 4, batch))
    assert hh.hh_rti_prepare(C.c_double(par.axis_rear), C.c_double(par.axis_front), C.c_double(par.acceleration),
                             C.c_double(0.9), C.c_double(ts), rk4, p(c_(y.T)), p(c_(Uprev.transpose(0, 2, 1))), first,
                             p(warm), p(A), p(B), p(c), C.c_int64(batch), N) == 0
    Ub, Ao, Bo, co, _ = bc.rti_prepare(y, Uprev, ts, par, 0.9, "rk4" if rk4 else "euler", first=bool(first))
    np.testing.assert_allclose(warm.transpose(0, 2, 1), Ub, rtol=0, atol=0)
    np.testing.assert_allclose(A.transpose(0, 2, 1).reshape(N, batch, 4, 4), Ao, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(B.transpose(0, 2, 1).reshape(N, batch, 4, 2), Bo, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(c.transpose(0, 2, 1), co, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("substeps", [0, 1, 4])
def test_plant_step_body(hh, substeps):
    par = bc.VehicleParameters()
    rng = np.random.default_rng(2)
    batch = 7
    x, fr = scenarios(rng, batch)
    u = rng.uniform(-0.3, 0.3, (batch, 2))
    xn = np.zeros((4, batch))
    assert hh.hh_plant_step(C.c_double(par.axis_rear), C.c_double(par.axis_front), C.c_double(par.acceleration),
                            C.c_double(0.05), p(c_(fr)), substeps, p(c_(x.T)), p(c_(u.T)), p(xn), C.c_int64(batch)) == 0
    ref = bc.plant_step(x, u, 0.05, par, fr, "euler" if substeps == 0 else "rk4", max(substeps, 1))
    np.testing.assert_allclose(xn.T, ref, rtol=1e-12, atol=1e-14)


def run_loop(hh, x0, fr, steps, N, rk4=0, substeps=4, ts=0.05, variant="sol", store=1, sqp_iters=1, sqp_tol=0.0,
             plant_par=None, x_obs=None):
    par = bc.VehicleParameters()
    pp = plant_par or par
    batch = x0.shape[0]
    Q, QT, R = bc.weights(variant)
    ulo, uhi, xlo, xhi = bc.bounds(par)
    U_plan = np.zeros((N, 2, batch)); X_pred = np.zeros((N + 1, 4, batch))
    X_cl = np.zeros((steps + 1, 4, batch)); U_cl = np.zeros((steps, 2, batch))
    cost = np.zeros(batch); viol = np.zeros(batch); clear = np.zeros(batch)
    ints = [np.zeros(batch, dtype=np.int32) for _ in range(4)]
    nc = 0 if x_obs is None else 9
    xo = None if x_obs is None else c_(x_obs)
    rc = hh.hh_rti_closed_loop(C.c_double(par.axis_rear), C.c_double(par.axis_front), C.c_double(par.acceleration),
                               C.c_double(par.friction), C.c_double(ts), rk4, C.c_double(pp.axis_rear),
                               C.c_double(pp.axis_front), C.c_double(pp.acceleration), p(c_(fr)), substeps, steps,
                               sqp_iters, C.c_double(sqp_tol), p(c_(Q)), p(c_(R)),
                               p(c_(QT)), p(c_(ulo)), p(c_(uhi)), p(c_(xlo)), p(c_(xhi)), nc, C.c_double(0.17),
                               C.c_double(0.08), p(xo), p(c_(x0.T)), p(U_plan), p(X_pred),
                               p(X_cl), p(U_cl), p(cost), p(viol), p(clear), *[a.ctypes.data_as(P32) for a in ints],
                               C.c_int64(batch), N, 60, C.c_double(1e-9), store)
    assert rc == 0
    return {"X": X_cl.transpose(0, 2, 1), "U": U_cl.transpose(0, 2, 1), "cost": cost, "viol": viol, "n_sat": ints[0],
            "n_fail": ints[1], "iters": ints[2], "last_status": ints[3], "U_plan": U_plan.transpose(0, 2, 1),
            "X_pred": X_pred.transpose(0, 2, 1), "clear": clear}


@pytest.mark.parametrize("store", [0, 1])
@pytest.mark.parametrize("rk4", [0, 1])
def test_closed_loop_matches_numpy_restatement(hh, rk4, store):
    rng = np.random.default_rng(3)
    x0, fr = scenarios(rng, 6)
    steps, N = 25, 20
    got = run_loop(hh, x0, fr, steps, N, rk4=rk4, store=store)
    ref = bc.closed_loop(x0, steps, N=N, friction_plant=fr, ocp_method="rk4" if rk4 else "euler", qp="port")
    assert np.all(got["n_fail"] == 0) and np.all(ref["status"] == 1)
    np.testing.assert_allclose(got["X"], ref["X"], rtol=0, atol=1e-6 * np.abs(ref["X"]).max())
    np.testing.assert_allclose(got["U"], ref["U"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(got["cost"], ref["cost"], rtol=1e-6)
    np.testing.assert_allclose(got["viol"], ref["viol"], atol=1e-7)
    par = bc.VehicleParameters()
    ulo, uhi, _, _ = bc.bounds(par)
    np.testing.assert_array_equal(got["U"] == uhi, ref["U"] == uhi)   # identical saturation pattern
    np.testing.assert_array_equal(got["U"] == ulo, ref["U"] == ulo)
    assert np.array_equal(got["n_sat"], ((ref["U"] == uhi) | (ref["U"] == ulo)).sum(axis=(0, 2)))


def test_closed_loop_matches_exact_qp_oracle(hh):
    """Same RTI loop with every QP solved by the exact active-set oracle."""
    rng = np.random.default_rng(4)
    x0, fr = scenarios(rng, 2)
    steps, N = 12, 15
    got = run_loop(hh, x0, fr, steps, N)
    ref = bc.closed_loop(x0, steps, N=N, friction_plant=fr, qp="exact")
    assert np.all(ref["status"] == 1)
    np.testing.assert_allclose(got["X"], ref["X"], rtol=0, atol=1e-6 * np.abs(ref["X"]).max())
    np.testing.assert_allclose(got["U"], ref["U"], rtol=0, atol=1e-6)
    par = bc.VehicleParameters()
    ulo, uhi, _, _ = bc.bounds(par)
    np.testing.assert_array_equal(got["U"] == uhi, np.abs(ref["U"] - uhi) < 1e-12)
    np.testing.assert_array_equal(got["U"] == ulo, np.abs(ref["U"] - ulo) < 1e-12)


def test_adaptive_plant_matches_scipy_odeint(hh):
    """substeps < 0: Dormand-Prince 5(4).  Checked against scipy odeint, the integrator behind the
    reference's exact_integration (session4_sol.py:37-56)."""
    par = bc.VehicleParameters()
    rng = np.random.default_rng(21)
    batch = 9
    x, fr = scenarios(rng, batch)
    x[:, 3] = rng.uniform(-0.5, 0.5, batch)
    u = np.stack([rng.uniform(-1, 1, batch), rng.uniform(-0.384, 0.384, batch)], 1)
    for ts in (0.05, 0.5):
        xn = np.zeros((4, batch))
        assert hh.hh_plant_step(C.c_double(par.axis_rear), C.c_double(par.axis_front), C.c_double(par.acceleration),
                                C.c_double(ts), p(c_(fr)), -10, p(c_(x.T)), p(c_(u.T)), p(xn), C.c_int64(batch)) == 0
        ref = bc.exact_integration_odeint(x, u, ts, par, fr)
        np.testing.assert_allclose(xn.T, ref, rtol=0, atol=1e-9)
        # and the fixed RK4 x4 stand-in is close to it at the controller's sampling time
        if ts == 0.05:
            np.testing.assert_allclose(bc.plant_step(x, u, ts, par, fr, "rk4", 4), ref, rtol=0, atol=1e-8)
