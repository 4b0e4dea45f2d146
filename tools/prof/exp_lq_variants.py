"""Times mpc_lq_solve at cfg 2b (2^20 scenarios, n=4, m=1, N=20, fp64) under environment-selected
variants (MPC_LQ_KRYLOV_COND, MPC_LQ_THREADS, ...) and checks each against the dense kernel.
Usage: python tools/prof/exp_lq_variants.py "VAR=val,VAR2=val" "VAR=val" ...   ("-" = defaults)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import bench
from model_predictive_control_b200 import lq


def timed(fn, iters=60, warm=10):
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
    specs = sys.argv[1:] or ["-"]
    rounds = int(os.environ.get("EXP_ROUNDS", "5"))
    times = {s: [] for s in specs}
    errs = {}
    for r in range(rounds):   # round-robin: clocks drift with the power/thermal state, so interleave the variants
        for spec in specs:
            for k in [k for k in os.environ if k.startswith("MPC_LQ_")]:
                del os.environ[k]
            if spec != "-":
                for kv in spec.split(","):
                    k, v = kv.split("=")
                    os.environ[k] = v
            times[spec].append(timed(lambda: lq.lq_solve(A, B, Q, R, Pf, x0, N, out=out), iters=30, warm=3))
            if r == 0:
                ex = ((out.X - Xr).abs().amax(dim=(0, 2)) / Xr.abs().amax(dim=(0, 2))).amax().item()
                eu = ((out.U - Ur).abs().amax() / Ur.abs().amax()).item()
                ev = ((out.V - Vr).abs() / Vr.abs()).amax().item()
                errs[spec] = (ex, eu, ev)
    for spec in specs:
        t = sorted(times[spec])
        ms = t[len(t) // 2]
        ex, eu, ev = errs[spec]
        print(f"{spec:44s} median {ms:7.4f} ms (min {t[0]:.4f} max {t[-1]:.4f})  {batch / ms * 1e-6:6.3f} Gsolves/s  "
              f"{1296 * batch / ms * 1e-6:7.1f} GB/s  errX {ex:.1e} errU {eu:.1e} errV {ev:.1e}", flush=True)


if __name__ == "__main__":
    main()
