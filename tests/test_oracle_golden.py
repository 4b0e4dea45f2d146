"""Pin the numpy oracle (oracle/lq.py) to the reference's own outputs.

tests/golden/session1.json was produced by running /root/reference/session_1 code
(tests/golden/make_golden.py).  Where /root/reference is mounted the oracle is also
compared with the live reference on fresh random inputs.
"""
import numpy as np
import pytest

from oracle import lq, ref_loader

RTOL = 1e-12


def arr(x):
    return np.asarray(x, dtype=np.float64)


def close(a, b, rtol=RTOL, atol=1e-13):
    np.testing.assert_allclose(arr(a), arr(b), rtol=rtol, atol=atol)


def test_survey_appendix_a_values(golden):
    """Known answers quoted in SURVEY.md Appendix A (generated from the reference)."""
    g = golden["cfg1"]
    close(g["V_N_1to9"][0], 44.705384340181936)
    close(g["V_N_1to9"][8], 1004.7592308787273)
    close(g["V_inf"], 1006.403292761319)
    close(g["recursion"]["20"]["K"][0], [[1.28645066125583, 2.312564925545413]])
    close(g["recursion"]["4"]["K"][0], [[-0.3468200818212525, 1.1448928016005042]])


def test_recursion_matches_reference_outputs(golden):
    g = golden["cfg1"]
    A, B, Q, R = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"])
    for N, rec in g["recursion"].items():
        P, K = lq.ricatti_recursion(A, B, Q, R, Q, int(N))
        assert len(P) == int(N) + 1 and len(K) == int(N)
        for p, pr in zip(P, rec["P"]):
            close(p, pr)
        for k, kr in zip(K, rec["K"]):
            close(k, kr)
        Ps, Ks = lq.riccati_recursion(A, B, R.reshape(1, 1), Q, Q, int(N))
        for p, pr in zip(Ps, rec["P_sol"]):
            close(p, pr)
        for k, kr in zip(Ks, rec["K_sol"]):
            close(k, kr)


def test_cost_to_go_matches_compare_term_cost(golden):
    g = golden["cfg1"]
    A, B, Q, R, x0 = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]), arr(g["x0"])
    for N in range(1, 10):
        P, _ = lq.ricatti_recursion(A, B, Q, R, Q, N)
        close(lq.cost_to_go(P[0], x0)[0], g["V_N_1to9"][N - 1])


def test_closed_loop_and_prediction(golden):
    g = golden["cfg1"]
    A, B, Q, R, x0 = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]), arr(g["x0"])
    for N, cl in g["closed_loop"].items():
        N = int(N)
        _, K = lq.ricatti_recursion(A, B, Q, R, Q, N)
        X = lq.simulate(A, B, x0, K, 30)
        assert X.shape == (2, 1, 30)
        # N=4 is unstable (|x| ~ 4e2): relative comparison
        close(X, cl["X"], rtol=1e-10)
        for t, pr in zip((0, 1, 7), cl["pred_t0_t1_t7"]):
            close(lq.prediction(A, B, X[:, :, t], K, N), pr, rtol=1e-10)
        xs, flag = lq.session1_simulate(A, B, 10 * np.ones(2), K, 30)
        assert xs.shape == (31, 2)
        close(xs, cl["sol_X"], rtol=1e-10)
        assert flag == cl["sol_flag"]
        xp, _ = lq.session1_simulate(A, B, 10 * np.ones(2), K, N, mode="pred")
        close(xp, cl["sol_pred"], rtol=1e-10)
    assert g["closed_loop"]["4"]["sol_flag"] is True
    assert g["closed_loop"]["20"]["sol_flag"] is False


def test_prediction_skips_gain0(golden):
    """Quirk 4 of SURVEY Appendix B: first predicted step uses gains[1]."""
    g = golden["cfg1"]
    A, B, Q, R, x0 = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]), arr(g["x0"])
    _, K = lq.ricatti_recursion(A, B, Q, R, Q, 10)
    xp = lq.prediction(A, B, x0, K, 10)
    close(xp[:, 0, 1], (A @ x0 + B @ (K[1] @ x0))[:, 0])
    close(xp[:, 0, 1], [15.0, -7.941782569134741])


def test_batched_shared_model(golden):
    g = golden["cfg2a"]
    A, B, Q, R, x0 = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]), arr(g["x0"])
    P, K = lq.ricatti_recursion(A, B, Q, R, Q, g["N"])
    for k, kr in zip(K, g["K"]):
        close(k, kr)
    close(lq.simulate(A, B, x0, K, 30), g["simulate_30"], rtol=1e-10)
    close(lq.prediction(A, B, x0, K, 20), g["prediction_20"], rtol=1e-10)
    close(lq.cost_to_go(P[0], x0), g["V"])


def test_batched_per_scenario_models(golden):
    g = golden["cfg2b"]
    A, B, Q, R, x0 = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]), arr(g["x0"])
    # batched oracle call: leading batch dim on every matrix
    P, K = lq.ricatti_recursion(A, B, Q, R.reshape(-1, 1, 1), Q, g["N"])
    for b in range(A.shape[0]):
        for k in range(g["N"]):
            close(K[k][b], g["K"][b][k], rtol=1e-10)
            close(P[k][b], g["P"][b][k], rtol=1e-10)
        Kb = [K[k][b] for k in range(g["N"])]
        close(lq.simulate(A[b], B[b], x0[b].reshape(4, 1), Kb, 21), g["simulate_21"][b], rtol=1e-9)


def test_wide_shape(golden):
    g = golden["n12m4"]
    A, B, Q, R = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"])
    P, K = lq.riccati_recursion(A, B, R, Q, Q, g["N"])
    close(P[0], g["P0"], rtol=1e-9)
    for k, kr in zip(K, g["K"]):
        close(k, kr, rtol=1e-9, atol=1e-12)
    P2, K2 = lq.ricatti_recursion(A, B, Q, R, Q, g["N"])
    close(P2[0], g["P0_fhc"], rtol=1e-9)
    close(K2[0], g["K0_fhc"], rtol=1e-9, atol=1e-12)


def test_open_loop_plan_cost_equals_cost_to_go(golden):
    """V accumulated along the optimal plan equals x0'P0x0 (FHC.py:123-124)."""
    g = golden["cfg2a"]
    A, B, Q, R, x0 = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]), arr(g["x0"])
    X, U, V, P, K = lq.lq_open_loop(A, B, Q, R, Q, x0[:, 3], g["N"])
    close(V, g["V"][3], rtol=1e-10)
    assert X.shape == (21, 4) and U.shape == (20, 1)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_live_reference_random_models():
    FHC, LinearSystem, sol = ref_loader.load_session1()
    rng = np.random.default_rng(7)
    for n, m, N in ((2, 1, 9), (4, 1, 20), (4, 2, 15), (12, 4, 50)):
        A = np.eye(n) + 0.2 * rng.standard_normal((n, n))
        B = rng.standard_normal((n, m))
        Q = np.eye(n) * rng.uniform(0.5, 2)
        R = np.eye(m) * rng.uniform(0.05, 1)
        P_ref, K_ref = FHC.ricatti_recursion(A, B, Q, R, Q, N)
        P, K = lq.ricatti_recursion(A, B, Q, R, Q, N)
        for a, b in zip(P + K, P_ref + K_ref):
            close(a, b, rtol=1e-11)
        Ps_ref, Ks_ref = sol.riccati_recursion(A, B, R, Q, Q, N)
        Ps, Ks = lq.riccati_recursion(A, B, R, Q, Q, N)
        for a, b in zip(Ps + Ks, Ps_ref + Ks_ref):
            close(a, b, rtol=1e-11)
        x0 = rng.uniform(-10, 10, size=(n, 5))
        s = FHC.AutoCruising(A, B)
        s.set_opti_gain(K_ref)
        s.simulate(x0, s.control_law, 12)
        close(lq.simulate(A, B, x0, K, 12), s.x, rtol=1e-10)
        close(lq.prediction(A, B, x0, K, N), s.prediction(x0, s.pred, N), rtol=1e-10)
