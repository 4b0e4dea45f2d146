"""Times mpc_lq_solve at cfg 2b (2^20 scenarios, n=4, m=1, N=20, fp64) under environment-selected
variants (MPC_LQ_KRYLOV_COND, MPC_LQ_THREADS, ...) and checks each against the dense kernel.
Usage: python tools/prof/exp_lq_variants.py "VAR=val,VAR2=val" "VAR=val" ...   ("-" = defaults)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import bench
from model_predictive_control_b200 import lq


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    batch, N = 1 << 20, 20
    A, B, Q, R, Pf, x0 = bench.cfg2b_inputs_torch(batch, 1237, dev, torch.float64)
    os.environ["MPC_LQ_KRYLOV_COND"] = "0"
    ref = lq.lq_solve(A, B, Q, R, Pf, x0, N)
    torch.cuda.synchronize()
    Xr, Ur, Vr = ref.X.clone(), ref.U.clone(), ref.V.clone()
    out = lq.LqSolveBuffers(batch, 4, 1, N, torch.float64, dev)
    for spec in sys.argv[1:] or ["-"]:
        for k in [k for k in os.environ if k.startswith("MPC_LQ_")]:
            del os.environ[k]
        if spec != "-":
            for kv in spec.split(","):
                k, v = kv.split("=")
                os.environ[k] = v
        ms = timed(lambda: lq.lq_solve(A, B, Q, R, Pf, x0, N, out=out))
        ex = ((out.X - Xr).abs().amax() / Xr.abs().amax()).item()
        eu = ((out.U - Ur).abs().amax() / Ur.abs().amax()).item()
        ev = ((out.V - Vr).abs() / Vr.abs()).amax().item()
        rel_row = ((out.X - Xr).abs().amax(dim=(0, 2)) / Xr.abs().amax(dim=(0, 2))).amax().item()
        print(f"{spec:40s} {ms:8.4f} ms  {batch / ms * 1e-6:7.3f} Gsolves/s  {1296 * batch / ms * 1e-6:8.1f} GB/s  "
              f"errX {ex:.2e} (per-scenario {rel_row:.2e}) errU {eu:.2e} errV {ev:.2e}", flush=True)


if __name__ == "__main__":
    main()
