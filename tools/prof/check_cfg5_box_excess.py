import sys, torch, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_gpu_boxqp import cfg5_model
from model_predictive_control_b200 import boxqp
A, B, Q, R = cfg5_model()
N, batch = 50, 1 << 20
g = torch.Generator(device="cuda"); g.manual_seed(1234 + 5)
x0T = torch.rand(12, batch, generator=g, device="cuda", dtype=torch.float64) * 4 - 2
dev = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device="cuda")
ws = boxqp.BoxQpWorkspace(batch, 12, 4, N, "cuda", sat=True)
res = boxqp.solve(dev(A), dev(B), dev(Q), dev(R), dev(Q), N, x0T, -1.0, 1.0, -5.0, 5.0, workspace=ws)
solved = res.status == 1
v = res.X[1:, :, solved].abs().amax(dim=(0, 1)) - 5.0
print("solved", int(solved.sum()), "max excess rel", float(v.max()) / 5.0, "n > 5e-7 rel", int((v > 2.5e-6).sum()), "n > 1e-7 rel", int((v > 5e-7).sum()))
