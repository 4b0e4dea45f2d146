"""Generate tests/golden/session1.json by RUNNING THE REFERENCE'S OWN CODE.

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


if __name__ == "__main__":
    main()
