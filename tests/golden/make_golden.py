"""Generate tests/golden/session1.json and tests/golden/session234.json by RUNNING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference modules are imported unmodified through oracle.ref_loader; every
entry records which reference function produced it.  The fixture travels to the
GPU box; /root/reference does not.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import scipy
from scipy import linalg

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def tolist(a):
    return np.asarray(a, dtype=np.float64).tolist()


def cfg2_model():
    """n=4, m=1 chain of integrators: the reference's double integrator
    (FHC.py:33-48, ts=0.5) extended to four states (cfg 2)."""
    n = 4
    A = np.eye(n) + 0.5 * np.diag(np.ones(n - 1), 1)
    B = np.zeros((n, 1))
    B[-1, 0] = -0.5
    C = np.array([[1.0], [-2.0 / 3.0], [0.0], [0.0]])
    Q = C @ C.T + 1e-3 * np.eye(n)
    R = np.array([0.1])
    return A, B, Q, R


def main():
    FHC, LinearSystem, sol = ref_loader.load_session1()
    out = {"_meta": {"numpy": np.__version__, "scipy": scipy.__version__,
                     "generated_by": "tests/golden/make_golden.py",
                     "source": "outputs of /root/reference/session_1/{FHC,LinearSystem,session1_sol}.py"}}

    # ---------------- cfg 1: FHC.main() data (FHC.py:134-147)
    A, B = FHC.get_dynamics_discrete(0.5)
    C = np.array([[1], [-2 / 3]])
    Q = np.matmul(C, C.T) + 1e-3 * np.eye(2, 2)
    R = np.array([0.1])
    P_f = Q
    x0 = np.array([[10.0], [10.0]])
    g = {"A": tolist(A), "B": tolist(B), "Q": tolist(Q), "R": tolist(R), "x0": tolist(x0)}

    # compare_term_cost numeric part (FHC.py:117-127)
    V_N = []
    for N in range(1, 10):
        P_n, _ = FHC.ricatti_recursion(A, B, Q, R, P_f, N)
        V_N.append(float(np.squeeze(x0.T @ P_n[0] @ x0)))
    P_inf = linalg.solve_discrete_are(A, B, Q, R)
    g["V_N_1to9"] = V_N
    g["P_inf"] = tolist(P_inf)
    g["V_inf"] = float(np.squeeze(x0.T @ P_inf @ x0))
    g["K_inf"] = tolist(-np.linalg.inv(R + B.T @ P_inf @ B) @ B.T @ P_inf @ A)

    # full recursion outputs for several horizons
    g["recursion"] = {}
    for N in (1, 4, 6, 10, 20, 50):
        P, K = FHC.ricatti_recursion(A, B, Q, R, P_f, N)
        Ps, Ks = sol.riccati_recursion(A, B, R.reshape(1, 1), Q, P_f, N)
        g["recursion"][str(N)] = {"P": [tolist(p) for p in P], "K": [tolist(k) for k in K],
                                  "P_sol": [tolist(p) for p in Ps], "K_sol": [tolist(k) for k in Ks]}

    # run_and_plot_traj numeric part (FHC.py:64-91): simulate 30 steps + predictions
    g["closed_loop"] = {}
    for N in (4, 6, 10, 20):
        _, gains = FHC.ricatti_recursion(A, B, Q, R, P_f, N)
        sys_ = FHC.AutoCruising(A, B)
        sys_.set_opti_gain(gains)
        sys_.simulate(x0, sys_.control_law, 30)
        X = sys_.x
        preds = [sys_.prediction(X[:, :, t], sys_.pred, N) for t in (0, 1, 7)]
        # instructor-solution loop (session1_sol.py:68-91,:155-170)
        f = lambda x, u: A @ x + B @ u
        xs, flag = sol.simulate(10 * np.ones(2), f, lambda x, t: gains[0] @ x, 30)
        xp, _ = sol.simulate(10 * np.ones(2), f, lambda x, t: gains[t] @ x, N)
        g["closed_loop"][str(N)] = {"X": tolist(X), "pred_t0_t1_t7": [tolist(p) for p in preds],
                                    "sol_X": tolist(xs), "sol_flag": bool(flag), "sol_pred": tolist(xp)}
    out["cfg1"] = g

    # ---------------- cfg 2a: shared n=4 model, column-batched x0 (reference code runs batched unchanged)
    A4, B4, Q4, R4 = cfg2_model()
    rng = np.random.default_rng(1235)
    X0 = rng.uniform(-10, 10, size=(4, 64))
    P, K = FHC.ricatti_recursion(A4, B4, Q4, R4, Q4, 20)
    sys_ = FHC.AutoCruising(A4, B4)
    sys_.set_opti_gain(K)
    sys_.simulate(X0, sys_.control_law, 30)
    pred = sys_.prediction(X0, sys_.pred, 20)
    out["cfg2a"] = {"A": tolist(A4), "B": tolist(B4), "Q": tolist(Q4), "R": tolist(R4), "N": 20,
                    "x0": tolist(X0), "P": [tolist(p) for p in P], "K": [tolist(k) for k in K],
                    "simulate_30": tolist(sys_.x), "prediction_20": tolist(pred),
                    "V": tolist(np.einsum("ib,ij,jb->b", X0, P[0], X0))}

    # ---------------- cfg 2b: per-scenario models, reference recursion looped over scenarios
    nb = 24
    As = A4 + 0.05 * rng.standard_normal((nb, 4, 4))
    Bs = B4 + 0.05 * rng.standard_normal((nb, 4, 1))
    Qs = Q4 * (1 + 0.2 * rng.uniform(size=(nb, 1, 1)))
    Rs = 0.1 * (1 + rng.uniform(size=(nb, 1)))
    x0s = rng.uniform(-10, 10, size=(nb, 4))
    Pl, Kl, Xl = [], [], []
    for b in range(nb):
        P, K = FHC.ricatti_recursion(As[b], Bs[b], Qs[b], Rs[b], Qs[b], 20)
        Pl.append([tolist(p) for p in P])
        Kl.append([tolist(k) for k in K])
        sys_ = FHC.AutoCruising(As[b], Bs[b])
        sys_.set_opti_gain(K)
        sys_.simulate(x0s[b].reshape(4, 1), sys_.control_law, 21)
        Xl.append(tolist(sys_.x))
    out["cfg2b"] = {"A": tolist(As), "B": tolist(Bs), "Q": tolist(Qs), "R": tolist(Rs), "N": 20,
                    "x0": tolist(x0s), "P": Pl, "K": Kl, "simulate_21": Xl}

    # ---------------- a wider shape (n=12, m=4) through the instructor-solution recursion (2-D R)
    n, m = 12, 4
    A12 = np.eye(n) + 0.1 * rng.standard_normal((n, n))
    B12 = rng.standard_normal((n, m))
    Q12 = np.eye(n)
    R12 = 0.1 * np.eye(m)
    P, K = sol.riccati_recursion(A12, B12, R12, Q12, Q12, 50)
    Pf_, Kf_ = FHC.ricatti_recursion(A12, B12, Q12, R12, Q12, 50)
    out["n12m4"] = {"A": tolist(A12), "B": tolist(B12), "Q": tolist(Q12), "R": tolist(R12), "N": 50,
                    "P0": tolist(P[0]), "K": [tolist(k) for k in K],
                    "P0_fhc": tolist(Pf_[0]), "K0_fhc": tolist(Kf_[0])}

    path = os.path.join(ROOT, "tests", "golden", "session1.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print("wrote", path, os.path.getsize(path), "bytes")



def _fields(obj, names):
    out = {}
    for k in names:
        v = getattr(obj, k)
        out[k] = tolist(v) if isinstance(v, np.ndarray) else (float(v) if isinstance(v, (int, float, np.floating, np.integer)) else v)
    return out


def main_sessions234():
    """Sessions 2-4: what the reference DOES define there, produced by its own (unmodified) code:
    the Problem data of session_2/3 (problem.py), VehicleParameters (parameters.py), the integrators of
    session4_sol.py:22-56, and the OCPs of session4_sol.py:131-217 / main.py:41-113 evaluated numerically
    (cost f(U; x0), constraint vector g(U; x0), bound vectors) through oracle.ref_loader.NumericCasadi.
    The only ingredient that is NOT the reference's is the bicycle ODE behind rcracers' KinematicBicycle
    (package absent and unpinned): oracle/bicycle.py's definition is plugged in, as documented in DESIGN.md."""
    out = {"_meta": {"numpy": np.__version__, "scipy": scipy.__version__, "generated_by": "tests/golden/make_golden.py",
                     "source": "/root/reference/session_{2,3}/problem.py, session_4/{parameters,session4_sol,main}.py run "
                               "through oracle/ref_loader.py (numeric casadi stand-in; bicycle ODE = oracle/bicycle.py)"}}
    names = ["Ts", "Q", "R", "p_min", "p_max", "v_min", "v_max", "u_min", "u_max", "N", "A", "B", "n_state", "n_input"]
    out["problem"] = {}
    for s in (2, 3):
        Problem = ref_loader.load_problem(s)
        out["problem"][str(s)] = {"default": _fields(Problem(), names), "N30": _fields(Problem(N=30), names),
                                  "Ts01_N7": _fields(Problem(Ts=0.1, N=7), names)}
    out["log_fields"] = {}
    for s in (2, 3):
        Log = ref_loader.load_log(s)
        lg = Log()
        out["log_fields"][str(s)] = {f.name: type(getattr(lg, f.name)).__name__ for f in __import__("dataclasses").fields(Log)}
    VP = ref_loader.load_parameters()
    vp = VP()
    out["parameters"] = {k: float(getattr(vp, k)) for k in VP.__dataclass_fields__}

    rng = np.random.default_rng(4321)
    sol, cs = ref_loader.load_session4("session4_sol")
    bike = sol.KinematicBicycle(vp)
    # ---- integrators (session4_sol.py:22-56) on the bicycle, nominal and mismatched friction (:461-463)
    cases = []
    for i in range(6):
        x = np.array([0.6, -0.25, 0.0, 0.0]) + rng.uniform(-0.3, 0.3, 4)
        u = np.array([rng.uniform(-1, 1), rng.uniform(-0.384, 0.384)])
        ts = [0.05, 0.08, 0.2][i % 3]
        vpm = VP(); vpm.friction = vp.friction * (0.8 if i % 2 else 1.0)
        bk = sol.KinematicBicycle(vpm)
        cases.append({"x": tolist(x), "u": tolist(u), "ts": ts, "friction": float(vpm.friction),
                      "f": tolist(bk(x, u)),
                      "forward_euler": tolist(sol.forward_euler(bk, ts)(x, u)),
                      "runge_kutta4": tolist(sol.runge_kutta4(bk, ts)(x, u)),
                      "exact_integration": tolist(sol.exact_integration(bk, ts)(x, u))})
    out["integrators"] = cases
    pol = sol.build_test_policy()
    out["test_policy"] = [tolist(pol(None, t)) for t in (0, 1, 2.5)]
    # open-loop protocol of compare_open_loop (:65-104): simulate(x0, dynamics, steps, policy) with the three integrators
    x0 = np.zeros(4)
    out["open_loop"] = {"x0": tolist(x0), "ts": 0.05, "steps": 40,
                        "forward_euler": tolist(sol.simulate(x0, sol.forward_euler(bike, 0.05), 40, policy=pol)),
                        "runge_kutta4": tolist(sol.simulate(x0, sol.runge_kutta4(bike, 0.05), 40, policy=pol)),
                        "exact_integration": tolist(sol.simulate(x0, sol.exact_integration(bike, 0.05), 40, policy=pol))}

    # ---- session4_sol.MPCController.build_ocp, numerically
    def ocp_cases(make, N, nrep, spread):
        res = []
        for _ in range(nrep):
            xv = np.array([0.6, -0.25, 0.0, 0.0]) + rng.uniform(-1, 1, 4) * spread
            U = np.stack([rng.uniform(-1, 1, N), rng.uniform(-0.384, 0.384, N)], 1)
            cs_, ctrl = make()
            cs_.values = {"x0": xv, **{f"u_{t}": U[t] for t in range(N)}}
            c = ctrl()
            nlp = c.ipopt_solver.nlp
            res.append({"x0": tolist(xv), "U": tolist(U), "f": float(np.squeeze(nlp["f"])), "g": tolist(np.ravel(nlp["g"])),
                        "x": tolist(np.ravel(nlp["x"])), "p": tolist(np.ravel(nlp["p"]))})
        bounds = {k: tolist(np.ravel(np.asarray(v, dtype=float))) for k, v in c.bounds.items()}
        return res, bounds

    N = 5
    res, bounds = ocp_cases(lambda: (cs, lambda: sol.MPCController(N, 0.05, params=vp)), N, 4, np.array([0.2, 0.2, 0.5, 0.2]))
    out["ocp_sol"] = {"N": N, "ts": 0.05, "cases": res, "bounds": bounds}
    cs.values = {"x0": np.zeros(4), **{f"u_{t}": np.zeros(2) for t in range(50)}}
    c50 = sol.MPCController(50, 0.05, params=vp)
    out["ocp_sol"]["N50_bounds"] = {k: tolist(np.ravel(np.asarray(v, dtype=float))) for k, v in c50.bounds.items()}

    # ---- main.MPCController.build_ocp (obstacle avoidance), numerically
    mn, cs2 = ref_loader.load_session4("main")
    x_obs = np.array([0.25, 0.0, 0.0, 0.0])
    N = 4
    res, bounds = ocp_cases(lambda: (cs2, lambda: mn.MPCController(N, 0.08, vp, mn.KinematicBicycle(vp, symbolic=True), x_obs)),
                            N, 4, np.array([0.3, 0.2, 0.6, 0.2]))
    centers, r = mn.create_cover_circles(vp.length, vp.width, 3)
    out["ocp_main"] = {"N": N, "ts": 0.08, "x_obs": tolist(x_obs), "cases": res, "bounds": bounds,
                       "cover_circles": {"centers": [tolist(c_) for c_ in centers], "r": float(r)},
                       "x2T": tolist(mn.x2T(np.array([0.3, -0.1, 0.7, 0.0]), False))}
    path = os.path.join(ROOT, "tests", "golden", "session234.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if "--only-234" not in sys.argv:
        main()
    main_sessions234()
