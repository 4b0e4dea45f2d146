"""GPU parity of the LQ path (K1, K2, fused solve) through the reference-shaped API and the
device-level operators, against the numpy oracle and the golden vectors generated from the
reference's own session-1 code.  Tolerances: 1e-6 relative fp64 / 1e-4 fp32 (north star); the
fp64 kernels are in fact held to 1e-9."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import lq as olq  # noqa: E402

RTOL64, RTOL32 = 1e-9, 1e-4


def arr(x):
    return np.asarray(x, dtype=np.float64)


@pytest.fixture(scope="module")
def mods():
    import torch
    from model_predictive_control_b200 import FHC, LinearSystem, lq, session1_sol
    assert torch.cuda.is_available()
    return FHC, LinearSystem, session1_sol, lq, torch


def models(rng, batch, n, m):
    A = np.eye(n) + 0.5 * np.diag(np.ones(n - 1), 1) + 0.05 * rng.standard_normal((batch, n, n))
    B = np.zeros((n, m)); B[-1, 0] = -0.5
    if m > 1:
        B[-2, 1] = 0.3
    B = B + 0.05 * rng.standard_normal((batch, n, m))
    Q = np.eye(n) * (1 + 0.2 * rng.random((batch, 1, 1)))
    R = 0.1 * np.eye(m) * (1 + rng.random((batch, 1, 1)))
    return A, B, Q, R


def test_cfg1_golden_recursion(mods, golden):
    FHC, *_ = mods
    g = golden["cfg1"]
    A, B, Q, R = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"])  # R is 1-D (FHC.py:141)
    for N, rec in g["recursion"].items():
        P, K = FHC.ricatti_recursion(A, B, Q, R, Q, int(N))
        assert isinstance(P, list) and len(P) == int(N) + 1 and len(K) == int(N)
        assert P[0].shape == (2, 2) and K[0].shape == (1, 2) and isinstance(K[0], np.ndarray)
        np.testing.assert_allclose(np.array(K), arr(rec["K"]), rtol=RTOL64, atol=1e-12)
        np.testing.assert_allclose(np.array(P), arr(rec["P"]), rtol=RTOL64, atol=1e-12)
        np.testing.assert_array_equal(P[-1], Q)


def test_cfg1_cost_sweep(mods, golden):
    FHC, *_ = mods
    g = golden["cfg1"]
    V, V_inf = FHC.terminal_cost_sweep(arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]), arr(g["Q"]), arr(g["x0"]))
    np.testing.assert_allclose(V, g["V_N_1to9"], rtol=RTOL64)
    np.testing.assert_allclose(V_inf, g["V_inf"], rtol=1e-9)


def test_cfg1_closed_loop_and_prediction(mods, golden):
    FHC, LinearSystem, sol, _, _ = mods
    g = golden["cfg1"]
    A, B, Q, R, x0 = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]), arr(g["x0"])
    for N, cl in g["closed_loop"].items():
        N = int(N)
        _, gains = FHC.ricatti_recursion(A, B, Q, R, Q, N)
        sys = FHC.AutoCruising(A, B)
        sys.set_opti_gain(gains)
        assert sys.simulate(x0, sys.control_law, 30) is None
        assert sys.x.shape == (2, 1, 30)
        np.testing.assert_allclose(sys.x, arr(cl["X"]), rtol=1e-8, atol=1e-12)
        for t, pr in zip((0, 1, 7), cl["pred_t0_t1_t7"]):
            xp = sys.prediction(sys.x[:, :, t], sys.pred, N)
            assert xp.shape == (2, 1, N)
            np.testing.assert_allclose(xp, arr(pr), rtol=1e-8, atol=1e-12)
        # instructor-solution loop (session1_sol.py:68-91)
        f = sol.linear_dynamics(A, B)
        xs, flag = sol.simulate(10 * np.ones(2), f, sol.feedback_policy(gains, receding=True), 30)
        assert xs.shape == (31, 2)
        np.testing.assert_allclose(xs, arr(cl["sol_X"]), rtol=1e-8, atol=1e-12)
        assert flag == cl["sol_flag"]
        xp, _ = sol.simulate(10 * np.ones(2), f, sol.feedback_policy(gains, receding=False), N)
        np.testing.assert_allclose(xp, arr(cl["sol_pred"]), rtol=1e-8, atol=1e-12)


def test_prediction_skips_gain0(mods, golden):
    """Reference quirk: LinearSystem.prediction never applies gains[0] (LinearSystem.py:30-31)."""
    FHC, *_ = mods
    g = golden["cfg1"]
    A, B, Q, R, x0 = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]), arr(g["x0"])
    _, K = FHC.ricatti_recursion(A, B, Q, R, Q, 10)
    sys = FHC.AutoCruising(A, B)
    sys.set_opti_gain(K)
    xp = sys.prediction(x0, sys.pred, 10)
    np.testing.assert_allclose(xp[:, 0, 1], [15.0, -7.941782569134741], rtol=1e-9)


def test_sol_argument_order(mods, golden):
    _, _, sol, _, _ = mods
    g = golden["cfg1"]
    A, B, Q, R = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]).reshape(1, 1)
    rec = g["recursion"]["20"]
    P, K = sol.riccati_recursion(A, B, R, Q, Q, 20)
    np.testing.assert_allclose(np.array(K), arr(rec["K_sol"]), rtol=RTOL64, atol=1e-12)
    np.testing.assert_allclose(np.array(P), arr(rec["P_sol"]), rtol=RTOL64, atol=1e-12)
    A2, B2, Q2, R2 = sol.setup()
    np.testing.assert_allclose(R2, [[0.1]])


def test_generic_callables_take_the_step_path(mods, golden):
    """Arbitrary Python policies still work (one device step per call), as in the reference."""
    FHC, LinearSystem, *_ = mods
    g = golden["cfg1"]
    A, B, Q, R, x0 = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]), arr(g["x0"])
    _, K = FHC.ricatti_recursion(A, B, Q, R, Q, 6)
    sys = LinearSystem.LinearSystem(A, B)
    sys.simulate(x0, lambda x, t: K[0] @ x, 30)
    np.testing.assert_allclose(sys.x, arr(g["closed_loop"]["6"]["X"]), rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(sys.f(x0, K[0] @ x0), A @ x0 + B @ (K[0] @ x0), rtol=1e-12)


def test_cfg2a_golden_batched_shared_model(mods, golden):
    FHC, *_ = mods
    g = golden["cfg2a"]
    A, B, Q, R, x0 = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]), arr(g["x0"])
    P, K = FHC.ricatti_recursion(A, B, Q, R, Q, g["N"])
    np.testing.assert_allclose(np.array(K), arr(g["K"]), rtol=RTOL64, atol=1e-12)
    np.testing.assert_allclose(np.array(P), arr(g["P"]), rtol=RTOL64, atol=1e-12)
    sys = FHC.AutoCruising(A, B)
    sys.set_opti_gain(K)
    sys.simulate(x0, sys.control_law, 30)
    np.testing.assert_allclose(sys.x, arr(g["simulate_30"]), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(sys.prediction(x0, sys.pred, 20), arr(g["prediction_20"]), rtol=1e-8, atol=1e-10)


def test_cfg2b_golden_per_scenario_models(mods, golden):
    FHC, _, _, lq, torch = mods
    g = golden["cfg2b"]
    A, B, Q, R, x0 = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]), arr(g["x0"])
    nb = A.shape[0]
    P, K = FHC.ricatti_recursion(A, B, Q, R.reshape(-1, 1, 1), Q, g["N"])  # batched leading dim
    assert K[0].shape == (nb, 1, 4)
    Kg = arr(g["K"]).transpose(1, 0, 2, 3)  # golden is [scenario][stage]
    Pg = arr(g["P"]).transpose(1, 0, 2, 3)
    np.testing.assert_allclose(np.array(K), Kg, rtol=RTOL64, atol=1e-12)
    np.testing.assert_allclose(np.array(P), Pg, rtol=RTOL64, atol=1e-12)
    # per-scenario closed loop with gains[0] (reference: one AutoCruising.simulate per scenario)
    dev = lambda a: torch.tensor(a, dtype=torch.float64, device="cuda")
    res = lq.lq_rollout(dev(A), dev(B), dev(np.array(K)), dev(x0.T.copy()), 21, gain_offset=0, gain_step=0)
    Xg = arr(g["simulate_21"])[:, :, 0, :]  # [scenario][n][T]
    np.testing.assert_allclose(res["X"].permute(2, 1, 0).cpu().numpy(), Xg, rtol=1e-8, atol=1e-10)


def test_n12_m4_generic_kernel(mods, golden):
    FHC, *_ = mods
    g = golden["n12m4"]
    A, B, Q, R = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"])
    P, K = FHC.ricatti_recursion(A, B, Q, R, Q, g["N"])
    np.testing.assert_allclose(P[0], arr(g["P0_fhc"]), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(K[0], arr(g["K0_fhc"]), rtol=1e-8, atol=1e-10)
    Po, Ko = olq.ricatti_recursion(A, B, Q, R, Q, g["N"])
    np.testing.assert_allclose(np.array(K), np.array(Ko), rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("n,m", [(2, 1), (4, 1), (4, 2), (3, 2), (12, 4)])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_riccati_random_batched(mods, n, m, dtype):
    _, _, _, lq, torch = mods
    rng = np.random.default_rng(10 + n + m)
    batch, N = (1000, 20) if n <= 4 else (37, 10)
    A, B, Q, R = models(rng, batch, n, m)
    dt = torch.float64 if dtype == "f64" else torch.float32
    dev = lambda a: torch.tensor(a, dtype=dt, device="cuda")
    K, P = lq.riccati(dev(A), dev(B), dev(Q), dev(R), dev(Q), N)
    Po, Ko = olq.ricatti_recursion(A, B, Q, R, Q, N)
    rtol = RTOL64 if dtype == "f64" else RTOL32
    scaleK, scaleP = np.abs(np.array(Ko)).max(), np.abs(np.array(Po)).max()
    assert np.abs(K.cpu().numpy() - np.array(Ko)).max() <= rtol * scaleK * 10
    assert np.abs(P.cpu().numpy() - np.array(Po)).max() <= rtol * scaleP * 10
    K0, P0 = lq.riccati(dev(A), dev(B), dev(Q), dev(R), dev(Q), N, all_P=False)
    assert torch.equal(K0, K) and torch.equal(P0, P[0])
    # caller-owned result buffers
    out = (torch.empty_like(K), torch.empty_like(P))
    K1, P1 = lq.riccati(dev(A), dev(B), dev(Q), dev(R), dev(Q), N, out=out)
    assert K1 is out[0] and P1 is out[1] and torch.equal(K1, K) and torch.equal(P1, P)
    with pytest.raises(ValueError):
        lq.riccati(dev(A), dev(B), dev(Q), dev(R), dev(Q), N, out=(torch.empty_like(K), torch.empty_like(P)[1:]))


@pytest.mark.parametrize("batch", [1, 2, 31, 33, 255, 1000, 4097])
def test_rollout_ragged_batches(mods, batch):
    """Odd batch sizes take the scalar path, even ones the vectorised path; both match the oracle."""
    _, _, _, lq, torch = mods
    rng = np.random.default_rng(batch)
    A, B, Q, R = (M[0] for M in models(rng, 1, 4, 1))
    N, T = 20, 21
    Po, Ko = olq.ricatti_recursion(A, B, Q, R, Q, N)
    x0 = rng.uniform(-10, 10, (4, batch))
    dev = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64, device="cuda")
    res = lq.lq_rollout(dev(A), dev(B), dev(np.array(Ko)), dev(x0), T, gain_offset=0, gain_step=1,
                        Q=dev(Q), R=dev(R), Pf=dev(Q), want_U=True, want_cost=True, want_unstable=True)
    Xo = olq.simulate(A, B, x0, [None] + list(Ko), T, mode="pred")  # gains[t-1] at step t
    np.testing.assert_allclose(res["X"].permute(1, 2, 0).cpu().numpy(), Xo, rtol=1e-9, atol=1e-10)
    V = olq.cost_to_go(Po[0], x0)
    np.testing.assert_allclose(res["cost"].cpu().numpy(), V, rtol=1e-9)  # optimal plan: cost = x0'P0 x0
    assert res["U"].shape == (T - 1, 1, batch)


@pytest.mark.parametrize("n,m", [(2, 1), (4, 1), (4, 2)])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("shared", [False, True])
def test_lq_solve_fused(mods, n, m, dtype, shared):
    _, _, _, lq, torch = mods
    rng = np.random.default_rng(100 + n + m)
    batch, N = 777, 20
    A, B, Q, R = models(rng, 1 if shared else batch, n, m)
    x0 = rng.uniform(-10, 10, (batch, n))
    dt = torch.float64 if dtype == "f64" else torch.float32
    dev = lambda a: torch.tensor(a, dtype=dt, device="cuda")
    sq = (lambda a: a[0]) if shared else (lambda a: a)
    out = lq.lq_solve(dev(sq(A)), dev(sq(B)), dev(sq(Q)), dev(sq(R)), dev(sq(Q)), dev(x0), N, want_K=True, want_P0=True)
    rtol = 1e-8 if dtype == "f64" else RTOL32
    X, U, V = out.X.cpu().numpy(), out.U.cpu().numpy(), out.V.cpu().numpy()
    for b in range(0, batch, 97):
        i = 0 if shared else b
        Xb, Ub, Vb, Pb, Kb = olq.lq_open_loop(A[i], B[i], Q[i], R[i], Q[i], x0[b], N)
        sx, su = np.abs(Xb).max(), max(np.abs(Ub).max(), 1e-30)
        assert np.abs(X[:, b] - Xb).max() <= rtol * sx * 10
        assert np.abs(U[:, b] - Ub).max() <= rtol * su * 10
        assert abs(V[b] - Vb) <= rtol * abs(Vb) * 10
        assert np.abs(out.K[:, b].cpu().numpy() - np.array(Kb)).max() <= rtol * np.abs(np.array(Kb)).max() * 10
        assert np.abs(out.P0[b].cpu().numpy() - Pb[0]).max() <= rtol * np.abs(Pb[0]).max() * 10


@pytest.mark.parametrize("n", [2, 4])
@pytest.mark.parametrize("shared", [False, True])
def test_lq_solve_krylov_kernel(mods, n, shared, monkeypatch):
    """Single-input fp64 solve without K/P0 outputs = lq_solve_krylov_kernel: against the oracle's dense
    recursion (FHC.py:51-61), bit-identical to the dense kernel where the conditioning guard rejects a
    scenario (uncontrollable pair, b = 0, nearly dependent Krylov columns), and within 1e-8 of it elsewhere."""
    _, _, _, lq, torch = mods
    rng = np.random.default_rng(300 + n)
    batch, N, m = 1500, 20, 1
    A, B, Q, R = models(rng, 1 if shared else batch, n, m)
    if not shared:
        A[0] = np.eye(n); B[1] = 0.0
        A[2] = np.diag(1.0 + 1e-6 * np.arange(n)); B[2] = 1.0
    x0 = rng.uniform(-10, 10, (batch, n))
    dev = lambda a: torch.tensor(a, dtype=torch.float64, device="cuda")
    sq = (lambda a: a[0]) if shared else (lambda a: a)
    args = (dev(sq(A)), dev(sq(B)), dev(sq(Q)), dev(sq(R)), dev(sq(Q)), dev(x0), N)
    assert lq.lq_solve_kernel_name(n, m, torch.float64) == "lq_solve_krylov_kernel"
    out = lq.lq_solve(*args)
    monkeypatch.setenv("MPC_LQ_KRYLOV_COND", "0")
    assert lq.lq_solve_kernel_name(n, m, torch.float64) == "lq_solve_kernel"
    dense = lq.lq_solve(*args)
    monkeypatch.delenv("MPC_LQ_KRYLOV_COND")
    if not shared:
        for t_k, t_d in ((out.X, dense.X), (out.U, dense.U)):
            assert torch.equal(t_k[:, :3], t_d[:, :3]), "guarded scenarios must take the dense body"
        assert torch.equal(out.V[:3], dense.V[:3])
        differ = (out.U != dense.U).any(dim=0).any(dim=-1)
        assert differ[3:].double().mean().item() > 0.5, "the Krylov path does not seem to be taken"
    sx = dense.X.abs().amax(dim=(0, 2)).clamp_min(1.0)[None, :, None]
    su = dense.U.abs().amax(dim=(0, 2)).clamp_min(1.0)[None, :, None]
    assert ((out.X - dense.X).abs() / sx).max().item() < 1e-8
    assert ((out.U - dense.U).abs() / su).max().item() < 1e-8
    assert torch.allclose(out.V, dense.V, rtol=1e-8, atol=1e-12)
    X, U, V = out.X.cpu().numpy(), out.U.cpu().numpy(), out.V.cpu().numpy()
    for b in list(range(0, 8)) + list(range(8, batch, 149)):
        i = 0 if shared else b
        Xb, Ub, Vb, _, _ = olq.lq_open_loop(A[i], B[i], Q[i], R[i], Q[i], x0[b], N)
        assert np.abs(X[:, b] - Xb).max() <= 1e-8 * max(1.0, np.abs(Xb).max())
        assert np.abs(U[:, b] - Ub).max() <= 1e-8 * max(1.0, np.abs(Ub).max())
        assert abs(V[b] - Vb) <= 1e-8 * abs(Vb)
    # terminal weight different from Q
    Pf = 2.5 * Q + 0.1 * np.eye(n)
    Pf[::2] = Q[::2]                          # every other scenario keeps P_f = Q
    out2 = lq.lq_solve(*args[:4], dev(sq(Pf)), *args[5:])
    U2 = out2.U.cpu().numpy()
    for b in list(range(0, 8)) + list(range(8, batch, 149)):
        i = 0 if shared else b
        _, Ub, Vb, _, _ = olq.lq_open_loop(A[i], B[i], Q[i], R[i], Pf[i], x0[b], N)
        assert np.abs(U2[:, b] - Ub).max() <= 1e-8 * max(1.0, np.abs(Ub).max())
        assert abs(out2.V[b].item() - Vb) <= 1e-8 * abs(Vb)


@pytest.mark.parametrize("N,batch", [(1, 1), (2, 63), (7, 65), (33, 1000), (120, 257)])
@pytest.mark.parametrize("misaligned", [False, True])
def test_lq_solve_krylov_edge_shapes(mods, N, batch, misaligned, monkeypatch):
    """Horizons 1 / odd / long (on-chip gains up to 120 stages), ragged batches, and inputs that are only 8-byte
    aligned (views at an odd element offset: the kernel's non-vectorised path): Krylov path == dense kernel == oracle."""
    _, _, _, lq, torch = mods
    rng = np.random.default_rng(1000 + N + batch)
    n, m = 4, 1
    A, B, Q, R = models(rng, batch, n, m)
    x0 = rng.uniform(-10, 10, (batch, n))

    def dev(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if not misaligned:
            return torch.tensor(a, device="cuda")
        buf = torch.zeros(a.size + 1, dtype=torch.float64, device="cuda")
        view = buf[1:].view(a.shape)          # 8-byte aligned, not 16/32
        view.copy_(torch.tensor(a, device="cuda"))
        assert view.data_ptr() % 16 == 8
        return view
    args = (dev(A), dev(B), dev(Q), dev(R), dev(Q), dev(x0), N)
    out = lq.lq_solve(*args)
    monkeypatch.setenv("MPC_LQ_KRYLOV_COND", "0")
    dense = lq.lq_solve(*args)
    monkeypatch.delenv("MPC_LQ_KRYLOV_COND")
    assert out.X.shape == (N + 1, batch, n) and out.U.shape == (N, batch, m) and out.V.shape == (batch,)
    sx = dense.X.abs().amax(dim=(0, 2)).clamp_min(1.0)[None, :, None]
    su = dense.U.abs().amax(dim=(0, 2)).clamp_min(1.0)[None, :, None]
    assert ((out.X - dense.X).abs() / sx).max().item() < 1e-8
    assert ((out.U - dense.U).abs() / su).max().item() < 1e-8
    assert torch.allclose(out.V, dense.V, rtol=1e-8, atol=1e-12)
    assert torch.equal(out.X[0], args[5])
    for b in sorted({0, batch // 2, batch - 1}):
        Xb, Ub, Vb, _, _ = olq.lq_open_loop(A[b], B[b], Q[b], R[b], Q[b], x0[b], N)
        assert np.abs(out.X[:, b].cpu().numpy() - Xb).max() <= 1e-8 * max(1.0, np.abs(Xb).max())
        assert np.abs(out.U[:, b].cpu().numpy() - Ub).max() <= 1e-8 * max(1.0, np.abs(Ub).max())
        assert abs(out.V[b].item() - Vb) <= 1e-8 * abs(Vb)


@pytest.mark.parametrize("n,m,shared", [(3, 1, False), (6, 2, False), (12, 4, True), (5, 3, True)])
def test_lq_solve_other_shapes(mods, n, m, shared):
    """Shapes without a fused kernel: lq_solve composes K1 + K2 on the device; same contract, same oracle."""
    _, _, _, lq, torch = mods
    rng = np.random.default_rng(500 + n + m)
    batch, N = 33, 15
    A, B, Q, R = models(rng, 1 if shared else batch, n, m)
    for j in range(2, m):
        B[:, j % n, j] += 0.2
    x0 = rng.uniform(-5, 5, (batch, n))
    dev = lambda a: torch.tensor(a, dtype=torch.float64, device="cuda")
    sq = (lambda a: a[0]) if shared else (lambda a: a)
    out = lq.lq_solve(dev(sq(A)), dev(sq(B)), dev(sq(Q)), dev(sq(R)), dev(sq(Q)), dev(x0), N, want_K=True, want_P0=True)
    assert out.X.shape == (N + 1, batch, n) and out.U.shape == (N, batch, m) and out.K.shape == (N, batch, m, n)
    for b in range(0, batch, 8):
        i = 0 if shared else b
        Xb, Ub, Vb, Pb, Kb = olq.lq_open_loop(A[i], B[i], Q[i], R[i], Q[i], x0[b], N)
        np.testing.assert_allclose(out.X[:, b].cpu().numpy(), Xb, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(Xb).max()))
        np.testing.assert_allclose(out.U[:, b].cpu().numpy(), Ub, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(Ub).max()))
        np.testing.assert_allclose(out.V[b].item(), Vb, rtol=1e-9)
        np.testing.assert_allclose(out.K[:, b].cpu().numpy(), np.array(Kb), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(out.P0[b].cpu().numpy(), Pb[0], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("want_K", [False, True])
def test_lq_solve_writes_stay_inside_the_outputs(mods, want_K):
    """Guard bands around every output buffer (ragged batch, odd horizon): both fused kernels write exactly their
    outputs and nothing else (compute-sanitizer is not available on the GPU pool)."""
    _, _, _, lq, torch = mods
    rng = np.random.default_rng(4242)
    batch, n, m, N, pad = 193, 4, 1, 7, 256
    A, B, Q, R = models(rng, batch, n, m)
    x0 = rng.uniform(-10, 10, (batch, n))
    dev = lambda a: torch.tensor(a, dtype=torch.float64, device="cuda")

    def guarded(shape):
        numel = int(np.prod(shape))
        buf = torch.full((numel + 2 * pad,), -7.25, dtype=torch.float64, device="cuda")
        return buf, buf[pad:pad + numel].view(shape)
    out = lq.LqSolveBuffers(1, n, m, N, torch.float64, "cuda")
    bufs = {}
    for name, shape in (("X", (N + 1, batch, n)), ("U", (N, batch, m)), ("V", (batch,)), ("K", (N, batch, m, n)), ("P0", (batch, n, n))):
        if name in ("K", "P0") and not want_K:
            setattr(out, name, None)
            continue
        bufs[name], view = guarded(shape)
        setattr(out, name, view)
    res = lq.lq_solve(dev(A), dev(B), dev(Q), dev(R), dev(Q), dev(x0), N, out=out)
    torch.cuda.synchronize()
    assert lq.lq_solve_kernel_name(n, m, torch.float64, want_K, want_K) == ("lq_solve_kernel" if want_K else "lq_solve_krylov_kernel")
    for name, buf in bufs.items():
        assert bool((buf[:pad] == -7.25).all()) and bool((buf[-pad:] == -7.25).all()), name
        assert not bool((getattr(res, name) == -7.25).any()), name      # and every output element was written
    Xb, Ub, Vb, _, _ = olq.lq_open_loop(A[batch - 1], B[batch - 1], Q[batch - 1], R[batch - 1], Q[batch - 1], x0[batch - 1], N)
    np.testing.assert_allclose(res.U[:, batch - 1].cpu().numpy(), Ub, rtol=0, atol=1e-8 * max(1.0, np.abs(Ub).max()))


def test_full_size_properties_cfg2(mods):
    """1M scenarios (BASELINE config 2): size-independent properties of the fused solve:
    V == x0' P0 x0, X satisfies the dynamics, U = K X, linearity in x0, and agreement with the
    unfused K1 -> K2 pipeline."""
    _, _, _, lq, torch = mods
    torch.manual_seed(0)
    B_, n, m, N = 1 << 20, 4, 1, 20
    dd = dict(dtype=torch.float64, device="cuda")
    A0 = torch.eye(n, **dd) + 0.5 * torch.diag(torch.ones(n - 1, **dd), 1)
    B0 = torch.zeros(n, m, **dd); B0[-1, 0] = -0.5
    A = A0 + 0.05 * torch.randn(B_, n, n, **dd)
    Bm = B0 + 0.05 * torch.randn(B_, n, m, **dd)
    Q = torch.eye(n, **dd) * (1 + 0.2 * torch.rand(B_, 1, 1, **dd))
    R = 0.1 * torch.eye(m, **dd) * (1 + torch.rand(B_, 1, 1, **dd))
    x0 = torch.rand(B_, n, **dd) * 20 - 10
    out = lq.lq_solve(A, Bm, Q, R, Q, x0, N, want_K=True, want_P0=True)
    V_quad = torch.einsum("bi,bij,bj->b", x0, out.P0, x0)
    assert torch.allclose(out.V, V_quad, rtol=1e-9, atol=1e-9)
    Xn = torch.einsum("bij,kbj->kbi", A, out.X[:-1]) + torch.einsum("bij,kbj->kbi", Bm, out.U)
    assert torch.allclose(out.X[1:], Xn, rtol=1e-10, atol=1e-9)
    assert torch.allclose(out.U, torch.einsum("kbij,kbj->kbi", out.K, out.X[:-1]), rtol=1e-10, atol=1e-9)
    # without K/P0 outputs a single-input fp64 solve runs in Krylov coordinates (lq_solve_krylov_kernel):
    # same plan within 1e-8 of each scenario's scale (bar: 1e-6), exactly linear in x0, and its X obeys
    # the dynamics of the ORIGINAL model
    out1 = lq.lq_solve(A, Bm, Q, R, Q, x0, N)
    sx = out.X.abs().amax(dim=(0, 2)).clamp_min(1.0)[None, :, None]
    su = out.U.abs().amax(dim=(0, 2)).clamp_min(1.0)[None, :, None]
    assert ((out1.X - out.X).abs() / sx).max().item() < 1e-8
    assert ((out1.U - out.U).abs() / su).max().item() < 1e-8
    assert torch.allclose(out1.V, out.V, rtol=1e-8)
    Xn1 = torch.einsum("bij,kbj->kbi", A, out1.X[:-1]) + torch.einsum("bij,kbj->kbi", Bm, out1.U)
    assert ((out1.X[1:] - Xn1).abs() / sx).max().item() < 1e-8
    out2 = lq.lq_solve(A, Bm, Q, R, Q, 2.0 * x0, N)
    assert torch.allclose(out2.U, 2.0 * out1.U, rtol=1e-12, atol=1e-12)
    assert torch.allclose(out2.V, 4.0 * out1.V, rtol=1e-12)
    K, P = lq.riccati(A, Bm, Q, R, Q, N, all_P=False)
    assert torch.allclose(K, out.K, rtol=1e-12, atol=1e-14) and torch.allclose(P, out.P0, rtol=1e-12)
    res = lq.lq_rollout(A, Bm, K, x0.t().contiguous(), N + 1, gain_offset=0, gain_step=1)
    assert torch.allclose(res["X"].permute(0, 2, 1), out.X, rtol=1e-12, atol=1e-12)


def test_torch_in_torch_out_and_errors(mods, golden):
    FHC, _, _, lq, torch = mods
    g = golden["cfg1"]
    dev = lambda a: torch.tensor(arr(a), device="cuda")
    P, K = FHC.ricatti_recursion(dev(g["A"]), dev(g["B"]), dev(g["Q"]), dev(g["R"]), dev(g["Q"]), 5)
    assert isinstance(K[0], torch.Tensor) and K[0].is_cuda and K[0].shape == (1, 2)
    with pytest.raises(ValueError):
        lq.riccati(dev(g["A"]), dev(g["B"]), dev(np.eye(3)), dev([[0.1]]), dev(g["Q"]), 5)
    with pytest.raises(ValueError):
        lq.riccati(dev(g["A"]).cpu(), dev(g["B"]), dev(g["Q"]), dev([[0.1]]), dev(g["Q"]), 5)
    with pytest.raises(IndexError):  # prediction beyond the available gains, as gains[t] would raise
        lq.lq_rollout(dev(g["A"]), dev(g["B"]), torch.stack(K), dev(g["x0"]), 10, gain_offset=1, gain_step=1)
    # N = 0: P = [P_f], K = []
    P0, K0 = FHC.ricatti_recursion(arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]), arr(g["Q"]), 0)
    assert len(P0) == 1 and len(K0) == 0


def test_host_pipeline_matches_direct_call(mods):
    _, _, _, lq, torch = mods
    rng = np.random.default_rng(77)
    batch, n, m, N = 5000, 4, 1, 20
    A, B, Q, R = models(rng, batch, n, m)
    pin = lambda a: torch.tensor(a, dtype=torch.float64).pin_memory()
    pipe = lq.LqHostPipeline(batch, n, m, N)
    outs = []
    for i in range(5):   # more submissions than buffers: exercises the reuse hand-shakes
        x0 = rng.uniform(-10, 10, (batch, n)) * (i + 1)
        hX, hU, hV = pipe.submit(pin(A), pin(B), pin(Q), pin(R), pin(Q), pin(x0))
        pipe.wait()
        dev = lambda a: torch.tensor(a, dtype=torch.float64, device="cuda")
        ref = lq.lq_solve(dev(A), dev(B), dev(Q), dev(R), dev(Q), dev(x0), N)
        assert torch.equal(hX, ref.X.cpu()) and torch.equal(hU, ref.U.cpu()) and torch.equal(hV, ref.V.cpu())


def test_infinite_horizon_matches_dare_golden(mods, golden):
    """P_inf / K_inf of FHC.py:97-98,126 (scipy DARE in the reference) as the fixed point of K1."""
    FHC, *_ = mods
    g = golden["cfg1"]
    P_inf, K_inf = FHC.infinite_horizon(arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]))
    np.testing.assert_allclose(P_inf, arr(g["P_inf"]).reshape(2, 2), rtol=1e-9)
    np.testing.assert_allclose(K_inf, arr(g["K_inf"]).reshape(1, 2), rtol=1e-9)


def test_run_and_plot_traj_numbers(mods, golden):
    """Closed loops + per-step prediction bundles of FHC.run_and_plot_traj (FHC.py:64-91)."""
    FHC, *_ = mods
    g = golden["cfg1"]
    A, B, Q, R, x0 = arr(g["A"]), arr(g["B"]), arr(g["Q"]), arr(g["R"]), arr(g["x0"])
    out, (x_inf, b_inf) = FHC.run_and_plot_traj(A, B, Q, R, Q, x0)
    for N in (4, 6, 10):
        x, bundle = out[N]
        cl = g["closed_loop"][str(N)]
        np.testing.assert_allclose(x, arr(cl["X"]), rtol=1e-8, atol=1e-12)
        assert bundle.shape == (30, 2, 1, N)
        for t, pr in zip((0, 1, 7), cl["pred_t0_t1_t7"]):
            np.testing.assert_allclose(bundle[t], arr(pr), rtol=1e-8, atol=1e-12)
    assert x_inf.shape == (2, 1, 30) and b_inf.shape == (30, 2, 1, 10)
    assert np.abs(x_inf[:, 0, -1]).max() < 1e-5   # the infinite-horizon loop is stable


def test_edge_cases_lq(mods):
    FHC, LinearSystem, sol, lq, torch = mods
    A, B = FHC.get_dynamics_discrete(0.5)
    Q = np.eye(2); R = np.array([[0.1]])
    P, K = FHC.ricatti_recursion(A, B, Q, R, Q, 1)
    Po, Ko = olq.ricatti_recursion(A, B, Q, R, Q, 1)
    np.testing.assert_allclose(K[0], Ko[0], rtol=1e-12)
    sys = FHC.AutoCruising(A, B); sys.set_opti_gain(K)
    sys.simulate(np.array([[1.0], [2.0]]), sys.control_law, 1)    # steps = 1: only x0
    assert sys.x.shape == (2, 1, 1)
    xp = sys.prediction(np.array([[1.0], [2.0]]), sys.pred, 1)    # horizon 1: no gain needed
    assert xp.shape == (2, 1, 1)
    with pytest.raises(IndexError):                                # reference: gains[1] does not exist
        sys.prediction(np.array([[1.0], [2.0]]), sys.pred, 3)
    with pytest.raises(ValueError):                                # x0 must be a column, as in the reference
        sys.simulate(np.array([1.0, 2.0]), sys.control_law, 5)
    # float32 tensors through the reference-shaped API stay float32 and meet 1e-4
    t32 = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    P32, K32 = FHC.ricatti_recursion(t32(A), t32(B), t32(Q), t32(R), t32(Q), 20)
    Po, Ko = olq.ricatti_recursion(A, B, Q, R, Q, 20)
    assert K32[0].dtype == torch.float32
    assert np.abs(torch.stack(K32).cpu().numpy() - np.array(Ko)).max() <= 1e-4 * np.abs(np.array(Ko)).max()
    # empty batch
    out = lq.lq_solve(t32(A), t32(B), t32(Q), t32(R), t32(Q), torch.empty((0, 2), dtype=torch.float32, device="cuda"), 5)
    assert out.X.shape == (6, 0, 2)
