import os, sys, time, torch, numpy as np
sys.path.insert(0, ".")  # run from the repository root; sys.path.insert(0, "tests")
from test_gpu_boxqp import cfg5_model
from model_predictive_control_b200 import boxqp
A, B, Q, R = cfg5_model()
dev = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device="cuda")
for batch in (4736, 4736 * 8, 4736 * 32):
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    x0 = torch.rand(12, batch, generator=g, device="cuda", dtype=torch.float64) * 4 - 2
    ws = boxqp.BoxQpWorkspace(batch, 12, 4, 50, x0.device)
    f = lambda: boxqp.solve(dev(A), dev(B), dev(Q), dev(R), dev(Q), 50, x0, -np.ones(4), np.ones(4), -5 * np.ones(12), 5 * np.ones(12), workspace=ws)
    r = f(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); r = f(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(batch, f"{ms:.1f} ms {batch/ms*1e3:.3e} solves/s ws {ws.nbytes/1e6:.0f} MB", torch.bincount(r.status, minlength=4).tolist(),
          "mean iters", float(r.iters.double().mean()))
batch = 4736 * 8
x0 = torch.rand(12, batch, device="cuda", dtype=torch.float64) * 2 - 1
ws = boxqp.BoxQpWorkspace(batch, 12, 4, 50, x0.device)
for minb in (2, 3, 4):
    os.environ["MPC_COOP_MINB"] = str(minb)
    f = lambda: boxqp.solve(dev(A), dev(B), dev(Q), dev(R), dev(Q), 50, x0, -np.ones(4), np.ones(4), -5 * np.ones(12), 5 * np.ones(12), workspace=ws)
    r = f(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); r = f(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("x0 in [-1,1] minb", minb, f"{ms:.1f} ms {batch/ms*1e3:.3e} solves/s", torch.bincount(r.status, minlength=4).tolist(), "mean iters", float(r.iters.double().mean()))
