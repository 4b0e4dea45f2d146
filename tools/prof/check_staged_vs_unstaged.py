"""Staged vs unstaged box-QP kernel on a full-machine cfg-3 batch: bitwise comparison in the caller's order, with an
identity / random / Morton permutation and sorted by iteration count (the check that found the early-release race of the
staging pipeline; every line must print `differ: 0`)."""
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
def cmp(tag, a, b):
    bad = ((a[0] != b[0]).flatten(0, -2).any(0) | (a[1] != b[1]) | (a[2] != b[2]))
    ok = (a[1] == 1)
    d = (a[0] - b[0]).abs().flatten(0, -2).amax(0)
    print(tag, "differ:", int(bad.sum()), " max|dU| (solved):", float(d[ok].max()), " n(|dU|>1e-6):", int((d[ok] > 1e-6).sum()))
ref = run(False, None)
order_sorted = boxqp.state_order(x0T)
ident = torch.arange(batch, device="cuda", dtype=torch.int32)
perm = torch.randperm(batch, device="cuda", generator=g).to(torch.int32)
cmp("unstaged sorted vs unordered", run(False, order_sorted), ref)
cmp("unstaged perm   vs unordered", run(False, perm), ref)
for rep in range(2):
    cmp("staged unordered", run(True, None), ref)
    cmp("staged identity ", run(True, ident), ref)
    cmp("staged perm     ", run(True, perm), ref)
    cmp("staged sorted   ", run(True, order_sorted), ref)
# sorted by iteration count (most homogeneous warps possible)
byit = torch.argsort(ref[2].to(torch.int64) * 4 + ref[1].to(torch.int64)).to(torch.int32)
cmp("staged by-iters ", run(True, byit), ref)
cmp("unstaged by-iters", run(False, byit), ref)
