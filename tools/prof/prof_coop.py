import sys, torch, numpy as np
sys.path.insert(0, ".")  # run from the repository root; sys.path.insert(0, "tests")
from test_gpu_boxqp import cfg5_model
from model_predictive_control_b200 import boxqp
A, B, Q, R = cfg5_model()
dev = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device="cuda")
batch = 4736 * 2
g = torch.Generator(device="cuda"); g.manual_seed(3)
x0 = torch.rand(12, batch, generator=g, device="cuda", dtype=torch.float64) * 2 - 1
ws = boxqp.BoxQpWorkspace(batch, 12, 4, 50, x0.device)
for _ in range(2):
    r = boxqp.solve(dev(A), dev(B), dev(Q), dev(R), dev(Q), 50, x0, -np.ones(4), np.ones(4), -5 * np.ones(12), 5 * np.ones(12), workspace=ws)
torch.cuda.synchronize()
print(torch.bincount(r.status, minlength=4).tolist())
