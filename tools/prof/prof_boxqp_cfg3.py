"""Launches of the thread-per-scenario box-QP kernel at the cfg-3 shape for ncu: `prof_boxqp_cfg3.py [batch] [f64|f32]`
(default 151 552 scenarios = two full waves of 592 resident CTAs x 128 threads)."""
import sys, torch
sys.path.insert(0, ".")
from model_predictive_control_b200 import boxqp, problem
batch, N = (int(sys.argv[1]) if len(sys.argv) > 1 else 151552), 30
dt = torch.float32 if (len(sys.argv) > 2 and sys.argv[2] == 'f32') else torch.float64
prob = problem.Problem(N=N)
dd = dict(dtype=dt, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(3)
x0T = torch.stack([torch.rand(batch, generator=g, **dd) * 100 - 100, torch.rand(batch, generator=g, **dd) * 25 - 10], 0).contiguous()
A, B = torch.tensor(prob.A, **dd), torch.tensor(prob.B, **dd)
Q, R = torch.tensor(prob.Q.astype(float), **dd), torch.tensor(prob.R.astype(float), **dd)
mpc = problem.LinearMPC(prob)
u_lo, u_hi, x_lo, x_hi = mpc.bounds()
ws = boxqp.BoxQpWorkspace(batch, 2, 1, N, "cuda", dtype=dt)
for _ in range(2):
    res = boxqp.solve(A, B, Q, R, Q, N, x0T, u_lo, u_hi, x_lo, x_hi, workspace=ws)
torch.cuda.synchronize()
print("iters", float(res.iters.double().mean()))
