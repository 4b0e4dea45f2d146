"""A few launches of mpc_lq_solve at cfg 2b for ncu (run from the repository root)."""
import sys, torch
sys.path.insert(0, ".")
import bench
from model_predictive_control_b200 import lq
batch, N = 1 << 20, 20
A, B, Q, R, Pf, x0 = bench.cfg2b_inputs_torch(batch, 1237, torch.device("cuda:0"), torch.float64)
out = lq.LqSolveBuffers(batch, 4, 1, N, torch.float64, torch.device("cuda:0"))
for _ in range(4):
    lq.lq_solve(A, B, Q, R, Pf, x0, N, out=out)
torch.cuda.synchronize()
print(lq.lq_solve_kernel_name(4, 1, torch.float64), float(out.V.sum()))
