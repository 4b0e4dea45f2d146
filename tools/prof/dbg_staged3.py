import os, sys, torch
sys.path.insert(0, ".")
from model_predictive_control_b200 import boxqp, problem
batch, N = 1 << 18, 30
prob = problem.Problem(N=N)
dd = dict(dtype=torch.float64, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(7)
x0T = torch.stack([torch.rand(batch, generator=g, **dd) * 100 - 100, torch.rand(batch, generator=g, **dd) * 25 - 10], 0).contiguous()
A, B = torch.tensor(prob.A, **dd), torch.tensor(prob.B, **dd)
Q, R = torch.tensor(prob.Q.astype(float), **dd), torch.tensor(prob.R.astype(float), **dd)
mpc = problem.LinearMPC(prob)
u_lo, u_hi, x_lo, x_hi = mpc.bounds()
ws = boxqp.BoxQpWorkspace(batch, 2, 1, N, "cuda", dtype=torch.float64)
def run(staged, order, v=None):
    os.environ["MPC_QP_STAGED"] = "1" if staged else "0"
    if v is None: os.environ.pop("MPC_QP_PREFETCH", None)
    else: os.environ["MPC_QP_PREFETCH"] = str(v)
    r = boxqp.solve(A, B, Q, R, Q, N, x0T, u_lo, u_hi, x_lo, x_hi, workspace=ws, order=order)
    torch.cuda.synchronize()
    return r.U.clone(), r.status.clone(), r.iters.clone()
def cmp(tag, a, b):
    bad = ((a[0] != b[0]).flatten(0, -2).any(0) | (a[1] != b[1]) | (a[2] != b[2]))
    print(tag, "differ:", int(bad.sum()))
ref = run(False, None)
perm = torch.randperm(batch, device="cuda", generator=g).to(torch.int32)
for v in (1,):
    for rep in range(3):
        cmp(f"variant {v} perm", run(True, perm, v), ref)
