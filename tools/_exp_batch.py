import sys, time, torch, numpy as np
sys.path.insert(0, ".")
from model_predictive_control_b200 import boxqp, problem
prob = problem.Problem(N=30)
dev = torch.device("cuda")
A, B = (torch.tensor(M, dtype=torch.float64, device=dev) for M in (prob.A, prob.B))
Q, R = (torch.tensor(M.astype(float), device=dev) for M in (prob.Q, prob.R))
mpc = problem.LinearMPC(prob); u_lo, u_hi, x_lo, x_hi = mpc.bounds()
for batch in (2048, 4096, 8192, 16384, 32768, 65536, 262144):
    g = torch.Generator(device=dev); g.manual_seed(7)
    x0 = torch.stack([torch.rand(batch, generator=g, device=dev, dtype=torch.float64) * 100 - 100,
                      torch.rand(batch, generator=g, device=dev, dtype=torch.float64) * 25 - 10], dim=0).contiguous()
    ws = boxqp.BoxQpWorkspace(batch, 2, 1, 30, dev)
    for _ in range(2):
        boxqp.solve(A, B, Q, R, Q, 30, x0, u_lo, u_hi, x_lo, x_hi, workspace=ws)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 5; e0.record()
    for _ in range(reps):
        r = boxqp.solve(A, B, Q, R, Q, 30, x0, u_lo, u_hi, x_lo, x_hi, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(batch, f"{ms:.3f} ms", f"{batch/ms*1e3:.3e} solves/s", "ws MB", ws.nbytes / 1e6)
