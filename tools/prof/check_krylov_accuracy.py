"""Krylov-coordinate LQ kernel against the dense-recursion kernel on millions of random scenarios (GPU):
largest per-scenario relative difference in X, U, V and the share of scenarios the conditioning guard hands to the
dense body, for the cfg-2b distribution and for wider model spreads.
Usage: python tools/prof/check_krylov_accuracy.py [scenarios_per_case]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import bench
from model_predictive_control_b200 import lq


def case(batch, noise, seed, N=20):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(seed)
    dd = dict(dtype=torch.float64, device=dev)
    n = 4
    A = torch.eye(n, **dd) + 0.5 * torch.diag(torch.ones(n - 1, **dd), 1) + noise * torch.randn(batch, n, n, generator=g, **dd)
    B = torch.zeros(n, 1, **dd); B[-1, 0] = -0.5
    B = B + noise * torch.randn(batch, n, 1, generator=g, **dd)
    C = torch.tensor([[1.0], [-2.0 / 3.0], [0.0], [0.0]], **dd)
    Q = (C @ C.t() + 1e-3 * torch.eye(n, **dd)) * (1 + 0.2 * torch.rand(batch, 1, 1, generator=g, **dd))
    R = 0.1 * (1 + torch.rand(batch, 1, 1, generator=g, **dd))
    x0 = torch.rand(batch, n, generator=g, **dd) * 20 - 10
    os.environ.pop("MPC_LQ_KRYLOV_COND", None)
    k = lq.lq_solve(A, B, Q, R, Q, x0, N)
    os.environ["MPC_LQ_KRYLOV_COND"] = "0"
    d = lq.lq_solve(A, B, Q, R, Q, x0, N)
    os.environ.pop("MPC_LQ_KRYLOV_COND", None)
    fin = torch.isfinite(d.X).all(dim=0).all(dim=-1) & torch.isfinite(d.U).all(dim=0).all(dim=-1)
    sx = d.X.abs().amax(dim=(0, 2)).clamp_min(1.0); su = d.U.abs().amax(dim=(0, 2)).clamp_min(1.0)
    ex = (k.X - d.X).abs().amax(dim=(0, 2)) / sx
    eu = (k.U - d.U).abs().amax(dim=(0, 2)) / su
    ev = (k.V - d.V).abs() / d.V.abs().clamp_min(1e-300)
    e = torch.maximum(torch.maximum(ex, eu), ev)[fin]
    same = ((k.U == d.U).all(dim=0).all(dim=-1)).double().mean().item()   # guarded scenarios are bit-identical
    frac = bench.krylov_path_fraction(A, B)
    q = torch.quantile(e[:: max(1, e.numel() // 1000000)], torch.tensor([0.5, 0.9999], **dd))
    # the scenarios where the two kernels differ most, against extended precision (numpy longdouble, host)
    import numpy as np
    L = np.longdouble
    worst = torch.argsort(torch.where(fin, torch.maximum(torch.maximum(ex, eu), ev), torch.zeros_like(ex)))[-3:].tolist()
    for b in worst:
        Ab, Bb, Qb, Rb = (t[b].cpu().numpy().astype(L) for t in (A, B, Q, R))
        P = Qb.copy(); Ks = []
        for _ in range(N):
            K = -(Bb.T @ P @ Ab) / (Rb + Bb.T @ P @ Bb)
            P = Qb + Ab.T @ P @ (Ab + Bb @ K)
            Ks.append(K)
        x = x0[b].cpu().numpy().astype(L)[:, None]; Ue = []
        for K in Ks[::-1]:
            u = K @ x
            x = Ab @ x + Bb @ u
            Ue.append(u[0, 0])
        Ue = np.array(Ue, dtype=L); s_ = max(1.0, float(np.abs(Ue).max()))
        ek = float(np.abs(k.U[:, b, 0].cpu().numpy() - Ue).max()) / s_
        ed = float(np.abs(d.U[:, b, 0].cpu().numpy() - Ue).max()) / s_
        print(f"    scenario {b}: |U - exact| / scale  krylov-path kernel {ek:.1e}   dense kernel {ed:.1e}   max|P0| {float(np.abs(P).max()):.1e}")
    print(f"noise {noise:4.2f}  scenarios {batch}  krylov path {frac:.4f} (bit-identical to dense: {same:.4f})  "
          f"rel diff median {q[0].item():.1e}  99.99% {q[1].item():.1e}  max {e.max().item():.1e}", flush=True)


if __name__ == "__main__":
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
    for noise, seed in ((0.05, 1), (0.05, 2), (0.1, 3), (0.2, 4), (0.5, 5)):
        case(nb, noise, seed)
