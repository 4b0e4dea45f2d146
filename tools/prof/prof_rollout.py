import sys, torch
sys.path.insert(0, ".")  # run from the repository root
from model_predictive_control_b200 import lq
n, m, N, batch = 4, 1, 20, 1 << 20
dd = dict(dtype=torch.float64, device="cuda")
A = torch.eye(n, **dd) + 0.5 * torch.diag(torch.ones(n - 1, **dd), 1)
B = torch.zeros(n, m, **dd); B[-1, 0] = -0.5
Q = torch.eye(n, **dd); R = torch.tensor([[0.1]], **dd)
x0 = torch.rand(n, batch, **dd) * 20 - 10
K, P = lq.riccati(A, B, Q, R, Q, N, all_P=False)
for _ in range(4):
    res = lq.lq_rollout(A, B, K[:, 0], x0, N + 1, gain_offset=0, gain_step=1, Q=Q, R=R, Pf=Q, want_U=True, want_cost=True)
torch.cuda.synchronize()
print(float(res["cost"].sum()))
