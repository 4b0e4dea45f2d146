import os, sys, time, torch, numpy as np
sys.path.insert(0, ".")  # run from the repository root
from model_predictive_control_b200 import boxqp, problem, session4
dev = torch.device("cuda")
prob = problem.Problem(N=30)
A, B = (torch.tensor(M, dtype=torch.float64, device=dev) for M in (prob.A, prob.B))
Q, R = (torch.tensor(M.astype(float), device=dev) for M in (prob.Q, prob.R))
mpc = problem.LinearMPC(prob); u_lo, u_hi, x_lo, x_hi = mpc.bounds()
batch = 262144
g = torch.Generator(device=dev); g.manual_seed(7)
x0 = torch.stack([torch.rand(batch, generator=g, device=dev, dtype=torch.float64) * 100 - 100,
                  torch.rand(batch, generator=g, device=dev, dtype=torch.float64) * 25 - 10], dim=0).contiguous()
ws = boxqp.BoxQpWorkspace(batch, 2, 1, 30, dev)
def timeit(f, reps=3):
    f(); torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
for minb in (2, 3, 4, 6):
    os.environ["MPC_QP_MINB"] = str(minb)
    ms = timeit(lambda: boxqp.solve(A, B, Q, R, Q, 30, x0, u_lo, u_hi, x_lo, x_hi, workspace=ws))
    print("cfg3 minb", minb, f"{ms:.2f} ms  {batch/ms*1e3:.3e} solves/s")
# ltv (4,2) via the step-wise controller path: one prepare + one QP on 65536 scenarios
b4 = 65536
par = session4.VehicleParameters(); ctrl = session4.MPCController(N=50, ts=0.05, params=par)
x4 = (torch.tensor([0.6, -0.25, 0, 0], device=dev, dtype=torch.float64) + (torch.rand(b4, 4, device=dev, dtype=torch.float64) * 0.4 - 0.2)).t().contiguous()
for minb in (2, 3, 4, 6):
    os.environ["MPC_QP_MINB"] = str(minb)
    ctrl.reset()
    ms = timeit(lambda: (ctrl.reset(), ctrl._solve_dev(x4)))
    print("cfg4-step(cold) minb", minb, f"{ms:.2f} ms  {b4/ms*1e3:.3e} solves/s")
