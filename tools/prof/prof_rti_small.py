"""RTI closed loop, 18 944 scenarios x 20 control steps (ncu comparisons of workspace types: MPC_QP_STORE)."""
import sys, torch
sys.path.insert(0, ".")
from model_predictive_control_b200 import session4
batch, steps = 18944, 20
g = torch.Generator(device="cuda"); g.manual_seed(5)
x0 = torch.tensor([0.6, -0.25, 0, 0], device="cuda", dtype=torch.float64) + (torch.rand(batch, 4, generator=g, device="cuda", dtype=torch.float64) * 0.4 - 0.2) * torch.tensor([1, 1, 0.5, 0.2], device="cuda", dtype=torch.float64)
fr = torch.rand(batch, generator=g, device="cuda", dtype=torch.float64) * 0.3 + 0.7
ctrl = session4.MPCController(N=50, ts=0.05, params=session4.VehicleParameters())
for _ in range(2):
    res = ctrl.closed_loop(x0, steps, friction_plant=fr)
torch.cuda.synchronize()
print("iters/QP", float(res.iters.double().mean()) / steps, "failed", int(res.n_failed.sum()))
