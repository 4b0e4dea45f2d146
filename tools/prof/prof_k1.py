"""One launch of riccati_reg_kernel with every K_k and P_k written (a1's literal contract) for ncu."""
import sys, torch
sys.path.insert(0, ".")
import bench
from model_predictive_control_b200 import lq
batch = 1 << 20
A, B, Q, R, Pf, _ = bench.cfg2b_inputs_torch(batch, 7, torch.device("cuda"), torch.float64)
for _ in range(2):
    K, P = lq.riccati(A, B, Q, R, Pf, 20, all_P=True)
torch.cuda.synchronize()
print(float(K.abs().max()))
