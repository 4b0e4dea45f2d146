"""session4.simulate hands arguments to the policy by PARAMETER NAME, as the course simulator the reference imports
from rcracers does (call sites /root/reference/session_4/session4_sol.py:353-361 `open_loop_policy(t)`, :58-62
`lambda y, t`, :222 `__call__(self, y)`; sessions 2/3 `simulate(..., policy=controller, log=ControllerLog())`,
/root/reference/session_2/log.py:8-12).  Plain-Python dynamics and policies: the dispatch runs without a GPU."""
import numpy as np
import pytest

from model_predictive_control_b200 import session4
from model_predictive_control_b200.log import ControllerLog


def dyn(x, u):
    return 0.9 * x + np.array([0.1, 0.0, 0.0, 0.0]) * u[0] + np.array([0.0, 0.2, 0.0, 0.0]) * u[1]


def test_open_loop_policy_receives_the_step_index():
    """The reference's exercise 3 / 4 pattern: `def open_loop_policy(t): return controls[t]`."""
    controls = np.arange(12.0).reshape(6, 2)

    def open_loop_policy(t):
        return controls[t]

    x0 = np.array([0.6, -0.25, 0.0, 0.0])
    X = session4.simulate(x0, dyn, n_steps=6, policy=open_loop_policy)
    assert X.shape == (7, 4)
    x = x0
    for t in range(6):
        x = dyn(x, controls[t])
        np.testing.assert_allclose(X[t + 1], x, rtol=0, atol=0)


def test_state_and_time_policy():
    """build_test_policy of the reference (session4_sol.py:58-62): lambda y, t."""
    pol = session4.build_test_policy()
    x0 = np.zeros(4)
    X = session4.simulate(x0, dyn, 5, policy=pol)
    x = x0
    for t in range(5):
        x = dyn(x, np.array([1.0, 0.1 * np.sin(t)]))
    np.testing.assert_allclose(X[-1], x)


def test_policy_named_y_only_and_unknown_names():
    seen = []

    def by_state(y):
        seen.append(("y", y.copy()))
        return np.array([y[0], 0.0])

    session4.simulate(np.ones(4), dyn, 2, policy=by_state)
    assert len(seen) == 2 and seen[0][0] == "y" and seen[0][1].shape == (4,)
    calls = []
    session4.simulate(np.ones(4), dyn, 3, policy=lambda x, k: (calls.append(k), np.zeros(2))[1])
    assert calls == [0, 1, 2]                      # unknown names: (state, step index) in order
    calls = []
    session4.simulate(np.ones(4), dyn, 2, policy=lambda state: (calls.append(state.shape), np.zeros(2))[1])
    assert calls == [(4,), (4,)]                   # a single unknown name is the measurement


class LoggingController:
    """Sessions-2/3 convention: controller(y, log) appends to the ControllerLog and returns u_0."""

    def __call__(self, y, log):
        log.solver_success.append(True)
        log.state_prediction.append(np.tile(y, (3, 1)))
        log.input_prediction.append(np.zeros((2, 2)))
        return np.zeros(2)


def test_log_is_passed_by_name():
    log = ControllerLog()
    X = session4.simulate(np.ones(4), dyn, 4, policy=LoggingController(), log=log)
    assert X.shape == (5, 4)
    assert len(log.solver_success) == 4 and log.state_prediction[0].shape == (3, 4)

    def t_and_log(t, log):
        log.solver_success.append(t)
        return np.zeros(2)

    log2 = ControllerLog()
    session4.simulate(np.ones(4), dyn, 3, policy=t_and_log, log=log2)
    assert log2.solver_success == [0, 1, 2]


def test_missing_policy_raises():
    with pytest.raises(ValueError):
        session4.simulate(np.ones(4), dyn, 3)


def test_swapped_state_bounds_flag():
    """template.py:132-133 lists the state bounds as [x, y, vel, heading] against the state order
    [x, y, heading, vel]; the flag reproduces that, the default is session4_sol.py:176-177."""
    par = session4.VehicleParameters()
    good = session4.MPCController(N=3, ts=0.05, params=par)
    swapped = session4.MPCController(N=3, ts=0.05, params=par, swap_state_bounds=True)
    np.testing.assert_array_equal(good.bounds["lbg"][:4], [par.min_pos_x, par.min_pos_y, par.min_heading, par.min_vel])
    np.testing.assert_array_equal(swapped.bounds["lbg"][:4], [par.min_pos_x, par.min_pos_y, par.min_vel, par.min_heading])
    np.testing.assert_array_equal(swapped.bounds["ubg"][:4], [par.max_pos_x, par.max_pos_y, par.max_vel, par.max_heading])
    np.testing.assert_array_equal(swapped.bounds["lbx"], good.bounds["lbx"])
