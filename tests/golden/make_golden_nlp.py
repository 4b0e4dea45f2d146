"""Generates tests/golden/session4_nlp.json: the CONVERGED solution of the reference's session-4 OCP
(/root/reference/session_4/session4_sol.py:132-217, N = 50, ts = 0.05, x0 = [0.6, -0.25, 0, 0], :445-447) and the closed
loop of exercise5's first simulation (:458, nominal forward-Euler plant) with the OCP solved to convergence at every
step -- what the reference's IPOPT call returns (:126-130), computed here with scipy SLSQP on oracle.bicycle's
restatement of the NLP (exact derivatives).

Pinning: at the solution, the reference's OWN build_ocp is evaluated numerically (oracle/ref_loader.NumericCasadi, source
unmodified): its cost f and constraint vector g must equal the restatement's, and U* must be a KKT point of the
reference's f (projected finite-difference gradient of the reference's f on the free inputs with no active state bound).
Needs /root/reference (this container); the GPU box only reads the JSON.

    python tests/golden/make_golden_nlp.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bicycle as bc, ref_loader  # noqa: E402


def reference_fg(sol, cs, vp, x0, U, N):
    cs.values = {"x0": np.asarray(x0, float), **{f"u_{t}": U[t] for t in range(N)}}
    c = sol.MPCController(N, 0.05, params=vp)
    nlp = c.ipopt_solver.nlp
    return float(np.squeeze(nlp["f"])), np.ravel(nlp["g"]).astype(float)


def main():
    N, ts = 50, 0.05
    x0 = np.array([0.6, -0.25, 0.0, 0.0])
    t0 = time.time()
    r = bc.nlp_solve(x0, N, ts)
    r = bc.nlp_solve(x0, N, ts, U_init=r["U"], tol=1e-15)        # polish
    assert r["success"], r["message"]
    fun, con = bc.ocp_functions(x0, N, ts)
    out = {"about": "converged solution of the reference's session-4 OCP (session4_sol.py:132-217) by scipy SLSQP on the "
                    "restated NLP; pinned to the reference's own build_ocp evaluated numerically at the solution",
           "generator": "tests/golden/make_golden_nlp.py", "N": N, "ts": ts, "x0": x0.tolist(), "U_star": r["U"].tolist(),
           "X_star": r["X"].tolist(), "f_star": r["f"]}
    # ---- the reference's own f, g at U*
    sol, cs = ref_loader.load_session4("session4_sol")
    vp = ref_loader.load_parameters()()
    f_ref, g_ref = reference_fg(sol, cs, vp, x0, r["U"], N)
    g_or = con(r["U"].reshape(-1))[0]
    out["pin"] = {"f_reference": f_ref, "f_restatement": r["f"], "max_abs_g_diff": float(np.abs(g_ref - g_or).max())}
    assert abs(f_ref - r["f"]) < 1e-10 and np.abs(g_ref - g_or).max() < 1e-12
    # KKT of the reference's f: finite-difference gradient on the inputs that are free and do not move an active state bound
    ulo, uhi, xlo, xhi = bc.bounds(bc.VehicleParameters())
    Xs = r["X"][1:]
    x_active = (np.abs(Xs - xlo) < 1e-7) | (np.abs(Xs - xhi) < 1e-7)
    u_free = (np.abs(r["U"] - ulo) > 1e-6) & (np.abs(r["U"] - uhi) > 1e-6)
    out["pin"]["active_state_bounds"] = int(x_active.sum())
    gr = fun(r["U"].reshape(-1))[1].reshape(N, 2)
    eps = 1e-6
    fd_max = 0.0
    probe = [(k, j) for k in range(N) for j in range(2) if u_free[k, j]][:12]
    for k, j in probe:
        Up, Um = r["U"].copy(), r["U"].copy()
        Up[k, j] += eps; Um[k, j] -= eps
        fd = (reference_fg(sol, cs, vp, x0, Up, N)[0] - reference_fg(sol, cs, vp, x0, Um, N)[0]) / (2 * eps)
        fd_max = max(fd_max, abs(fd - gr[k, j]))
    out["pin"]["max_fd_vs_analytic_gradient_of_reference_f"] = fd_max
    if not x_active.any():
        out["pin"]["max_abs_gradient_on_free_inputs"] = float(np.abs(gr[u_free]).max())
    # ---- SQP rounds (the GPU algorithm, numpy restatement) against U*
    table = []
    for k in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 40):
        ref = bc.closed_loop(x0[None], 1, N=N, ts=ts, qp="port", sqp_iters=k, plant_method="euler", keep_plans=True)
        table.append({"sqp_iters": k, "max_abs_U_minus_Ustar": float(np.abs(ref["plans"][0][:, 0] - r["U"]).max())})
    out["cold_start_convergence"] = table
    # ---- closed loop with converged solves (exercise5's nominal simulation, forward-Euler plant)
    steps = 60
    cl = bc.closed_loop_converged(x0, steps, N, ts, plant_method="euler")
    out["closed_loop"] = {"steps": steps, "X": cl["X"].tolist(), "U": cl["U"].tolist()}
    rows = []
    for k, tol in ((1, 0.0), (2, 0.0), (3, 0.0), (5, 0.0), (8, 0.0), (60, 1e-8)):
        ref = bc.closed_loop(x0[None], steps, N=N, ts=ts, qp="port", sqp_iters=k, sqp_tol=tol, plant_method="euler")
        dx = np.abs(ref["X"][:, 0] - cl["X"]).max(axis=1); du = np.abs(ref["U"][:, 0] - cl["U"]).max(axis=1)
        rows.append({"sqp_iters": k, "sqp_tol": tol, "max_dx": float(dx.max()), "max_du": float(du.max()), "dx_at_step": [float(dx[i]) for i in (1, 5, 10, 20, 40, steps)],
                     "first_step_with_du_below_1e-6": int(next((i for i in range(steps) if np.all(du[i:] < 1e-6)), -1))})
    out["closed_loop_vs_converged"] = rows
    out["seconds"] = time.time() - t0
    with open(os.path.join(ROOT, "tests", "golden", "session4_nlp.json"), "w") as fh:
        json.dump(out, fh)
    print(json.dumps({k: out[k] for k in ("pin", "cold_start_convergence", "closed_loop_vs_converged", "seconds")}, indent=1))


if __name__ == "__main__":
    main()
