"""Debug: staged vs unstaged box-QP kernel, bitwise comparison on a cfg-3 batch."""
import os, sys, torch
sys.path.insert(0, ".")
from model_predictive_control_b200 import boxqp, problem
batch, N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18, 30
prob = problem.Problem(N=N)
dd = dict(dtype=torch.float64, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(7)
x0T = torch.stack([torch.rand(batch, generator=g, **dd) * 100 - 100, torch.rand(batch, generator=g, **dd) * 25 - 10], 0).contiguous()
A, B = torch.tensor(prob.A, **dd), torch.tensor(prob.B, **dd)
Q, R = torch.tensor(prob.Q.astype(float), **dd), torch.tensor(prob.R.astype(float), **dd)
mpc = problem.LinearMPC(prob)
u_lo, u_hi, x_lo, x_hi = mpc.bounds()
ws = boxqp.BoxQpWorkspace(batch, 2, 1, N, "cuda", dtype=torch.float64)
def run(staged, order):
    os.environ["MPC_QP_STAGED"] = "1" if staged else "0"
    r = boxqp.solve(A, B, Q, R, Q, N, x0T, u_lo, u_hi, x_lo, x_hi, workspace=ws, order=order)
    torch.cuda.synchronize()
    return r.U.clone(), r.status.clone(), r.iters.clone()
for order in (None, "auto"):
    U0, s0, i0 = run(False, order)
    for rep in range(3):
        U1, s1, i1 = run(True, order)
        bad = ((U0 != U1).flatten(0, -2).any(0) | (s0 != s1) | (i0 != i1)).nonzero().flatten()
        print("order", order, "rep", rep, "differing scenarios:", bad.numel(), "of", batch)
        if bad.numel():
            b = bad[:16].tolist()
            print("  idx", b)
            print("  lane", [x % 32 for x in b], "warp", [x // 32 for x in b[:16]])
            print("  status0", s0[bad[:16]].tolist(), "status1", s1[bad[:16]].tolist())
            print("  iters0", i0[bad[:16]].tolist(), "iters1", i1[bad[:16]].tolist())
            w = torch.unique(bad // 32)
            print("  warps affected", w.numel(), "first", w[:10].tolist())
            if order is None:
                for ww in w[:4].tolist():
                    print("   warp", ww, "iters0", i0[ww*32:(ww+1)*32].tolist(), "status0", s0[ww*32:(ww+1)*32].tolist())
                    print("   warp", ww, "iters1", i1[ww*32:(ww+1)*32].tolist())
