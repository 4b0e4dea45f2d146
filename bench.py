#!/usr/bin/env python
"""Benchmark of the batched finite-horizon MPC solve on B200 (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2b] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic scenarios.  Default workload
= BASELINE.json configs[1]: finite-horizon LQ (Riccati) solves, nx=4, nu=1, N=20, 2^20 scenarios
per GPU, each scenario with its own model and initial state ("cfg2b"), fp64.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "closed-loop MPC solves/sec (batched scenarios, horizon N)"
UNIT = "solves/s"


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def cfg2_shapes():
    return dict(n=4, m=1, N=20, batch=1 << 20)


def cfg2b_inputs_numpy(batch, seed):
    """Synthetic per-scenario models of SURVEY.md section 8d cfg 2b (numpy, for the CPU arms)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    n, m = 4, 1
    A0 = np.eye(n) + 0.5 * np.diag(np.ones(n - 1), 1)
    B0 = np.zeros((n, m)); B0[-1, 0] = -0.5
    C = np.array([[1.0], [-2.0 / 3.0], [0.0], [0.0]])
    Q0 = C @ C.T + 1e-3 * np.eye(n)
    A = A0 + 0.05 * rng.standard_normal((batch, n, n))
    B = B0 + 0.05 * rng.standard_normal((batch, n, m))
    Q = Q0 * (1 + 0.2 * rng.random((batch, 1, 1)))
    R = 0.1 * (1 + rng.random((batch, 1, 1)))
    x0 = rng.uniform(-10, 10, (batch, n))
    return A, B, Q, R, Q.copy(), x0


def cfg2b_inputs_torch(batch, seed, device, dtype):
    import torch
    g = torch.Generator(device=device); g.manual_seed(seed)
    n, m = 4, 1
    dd = dict(dtype=torch.float64, device=device)
    A0 = torch.eye(n, **dd) + 0.5 * torch.diag(torch.ones(n - 1, **dd), 1)
    B0 = torch.zeros(n, m, **dd); B0[-1, 0] = -0.5
    C = torch.tensor([[1.0], [-2.0 / 3.0], [0.0], [0.0]], **dd)
    Q0 = C @ C.t() + 1e-3 * torch.eye(n, **dd)
    A = A0 + 0.05 * torch.randn(batch, n, n, generator=g, **dd)
    B = B0 + 0.05 * torch.randn(batch, n, m, generator=g, **dd)
    Q = Q0 * (1 + 0.2 * torch.rand(batch, 1, 1, generator=g, **dd))
    R = 0.1 * (1 + torch.rand(batch, 1, 1, generator=g, **dd))
    x0 = torch.rand(batch, n, generator=g, **dd) * 20 - 10
    return [t.to(dtype).contiguous() for t in (A, B, Q, R, Q.clone(), x0)]


def cfg2b_bytes_per_solve(w, n=4, m=1, N=20):
    """Algorithmic HBM bytes per solve: read A, Q, Pf (3 n^2), B (n m), R (m^2), x0 (n);
    write X ((N+1) n), U (N m), V (1).  """
    return w * (3 * n * n + n * m + m * m + n) + w * ((N + 1) * n + N * m + 1)


def cfg2b_flops_per_solve(n=4, m=1, N=20):
    """Flops the fused kernel actually performs per solve (mul+add = 2):
    per stage  W=PA 2n^3, G=B'W 2n^2 m, PB 2n^2 m, S=R+B'PB 2nm^2+m^2, solve ~2m^3 + mn,
    W+=PB K 2n^2 m, upper triangle of P=Q+A'W n^2(n+1);  rollout u=Kx 2nm, x+=Ax+Bu 2n^2+2nm;
    V = x0'P0 x0 2n^2+2n."""
    f_ric = 2 * n**3 + 6 * n * n * m + 2 * n * m * m + m * m + 2 * m**3 + m * n + n * n * (n + 1)
    f_roll = 2 * n * m + 2 * n * n + 2 * n * m
    return N * (f_ric + f_roll) + 2 * n * n + 2 * n


def cfg2b_flops_per_solve_krylov(n=4, N=20):
    """Flops lq_solve_krylov_kernel performs per single-input solve (lq_core.cuh, mul+add = 2):
    set-up: n Krylov mat-vecs 2n^3, adjugate + determinant (n=4: 127), the two Frobenius norms 4n^2,
    -a = C^-1 A^n b and z0 = C^-1 x0 4n^2+2n, the two congruences C'QC, C'PfC 2(2n^3+n^2(n+1));
    backward stage: w = -Pa 2n^2, S and its reciprocal 2, K n, upper triangle of P+ 3n(n-1)/2 + 2n + (n-1) + 2n;
    forward stage: u = Kz 2n, z+ 2n, x = Cz 2n^2;  V = z0'P0 z0 2n^2+2n."""
    adj = 127 if n == 4 else 7
    setup = 2 * n**3 + adj + 4 * n * n + 4 + 4 * n * n + 2 * n + 2 * (2 * n**3 + n * n * (n + 1))
    back = 2 * n * n + 2 + n + 3 * n * (n - 1) // 2 + 2 * n + (n - 1) + 2 * n
    fwd = 2 * n + 2 * n + 2 * n * n
    return setup + N * (back + fwd) + 2 * n * n + 2 * n


def krylov_path_fraction(A, B, cond_max=1e3):
    """Share of the scenarios lq_solve_krylov_kernel accepts (cond_F of [b, Ab, A^2b, A^3b] <= cond_max);
    the others take the dense recursion inside the same launch.  torch on the device, outside any timed region."""
    import torch
    cols = [B[:, :, 0]]
    for _ in range(A.shape[-1] - 1):
        cols.append(torch.einsum("bij,bj->bi", A, cols[-1]))
    C = torch.stack(cols, dim=2)
    cond = torch.linalg.matrix_norm(C) * torch.linalg.matrix_norm(torch.linalg.inv(C))
    return float((cond <= cond_max).double().mean())


# ------------------------------------------------------------------------------------------------
# CPU arms (oracle = restated reference; the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One worker of the CPU arm: `count` cfg-2b scenarios, one finite-horizon solve each, in a Python loop.
    kind == "reference": the reference's OWN functions, unmodified (FHC.ricatti_recursion, FHC.py:51-61; the open-loop
    plan by AutoCruising.pred through LinearSystem.prediction exactly as FHC.py:87-88 calls it; V = x0'P_0 x0, :123-124),
    imported from /root/reference or its staged copy oracle/_ref (oracle/ref_loader.py).
    kind == "port": the oracle restatement of the same loop (when the reference files are not available)."""
    seed, count, kind = args
    A, B, Q, R, Pf, x0 = cfg2b_inputs_numpy(count, seed)
    acc = 0.0
    if kind == "reference":
        from oracle import ref_loader
        FHC, _, _ = ref_loader.load_session1()
        t0 = time.perf_counter()
        for b in range(count):
            P, K = FHC.ricatti_recursion(A[b], B[b], Q[b], R[b], Pf[b], 20)
            sys_ = FHC.AutoCruising(A[b], B[b])
            sys_.set_opti_gain(K)
            xb = x0[b][:, None]
            sys_.prediction(xb, sys_.pred, 20)
            acc += float(xb.T @ P[0] @ xb)
        return time.perf_counter() - t0, count, acc
    from oracle import lq as olq
    t0 = time.perf_counter()
    for b in range(count):  # the reference as shipped: one ricatti_recursion + rollout per scenario
        X, U, V, _, _ = olq.lq_open_loop(A[b], B[b], Q[b], R[b], Pf[b], x0[b], 20)
        acc += V
    return time.perf_counter() - t0, count, acc


def cpu_arm_kind():
    """ "reference" when the reference's session-1 files are mounted or staged (oracle/_ref, built by
    __graft_entry__.build()), else "port"."""
    try:
        from oracle import ref_loader
        return "reference" if ref_loader.session1_root() is not None else "port"
    except Exception:
        return "port"


CPU_SAMPLE = {"reference": "python loop of the reference's own FHC.ricatti_recursion + AutoCruising.pred / LinearSystem.prediction "
                           "+ x0'P_0x0 per scenario (unmodified session_1 files), one process per core",
              "port": "python loop of the oracle port of FHC.ricatti_recursion + rollout, one process per core"}


def cpu_reference_rate(total_scenarios, cores, kind=None):
    """Solves/s of the CPU arm, one Python loop per core."""
    import multiprocessing as mp
    kind = kind or cpu_arm_kind()
    per = max(1, total_scenarios // cores)
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(9000 + i, per, kind) for i in range(cores)])
    wall = time.perf_counter() - t0
    inner = max(r[0] for r in res)
    done = sum(r[1] for r in res)
    return done / inner, done, wall


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    kind = cpu_arm_kind()
    per_step = (700 if kind == "reference" else 1500) * cores  # ~1-2 s of CPU work per step on every core
    rates = []
    for i in range(args.warmup + args.steps):
        rate, done, wall = cpu_reference_rate(per_step, cores, kind)
        if i >= args.warmup:
            rates.append((rate, done))
    value = sum(r for r, _ in rates) / len(rates)
    sample = f"{rates[0][1]} scenarios/step of cfg2b (nx=4,nu=1,N=20), {CPU_SAMPLE[kind]}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * rates[0][1] / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2b: FHC Riccati LQ, nx=4 nu=1 N=20, per-scenario model + x0 (bounded CPU sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,power.limit")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw, pl = [], [], set(), [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
            try:
                pw.append(float(parts[6])); pl.append(float(parts[7]))
            except (ValueError, IndexError):
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort(); pw.sort()
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        if pw:
            out["power_w"] = pw[len(pw) // 2]
            out["power_limit_w"] = max(pl) if pl else None
        return out


# ------------------------------------------------------------------------------------------------
# process context: one process per GPU (torchrun), NUMA placement, host-link probe
# ------------------------------------------------------------------------------------------------
def gpu_numa_cpus(index):
    """(numa_node, cpu set) of the PCIe root the GPU hangs off (sysfs), or (None, None)."""
    try:
        bus = subprocess.run(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if not bus:
            return None, None
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:]}:{rest}"
        with open(path + "/numa_node") as fh:
            node = int(fh.read().strip())
        with open(path + "/local_cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                if "-" in part:
                    lo, hi = part.split("-")
                    cpus.update(range(int(lo), int(hi) + 1))
                elif part:
                    cpus.add(int(part))
        return node, cpus
    except Exception:
        return None, None


class Ctx:
    """One rank of the bench: device, process group, barrier and the max-over-ranks reduction of a timing."""

    def __init__(self, bind_numa=True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
        self.affinity0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
        self.numa = None
        if bind_numa and self.affinity0 is not None:
            # pinned host buffers are allocated on the node of the thread that allocates them: run next to the GPU
            node, cpus = gpu_numa_cpus(self.local)
            if cpus:
                use = cpus & self.affinity0
                if use:
                    os.sched_setaffinity(0, use)
                    self.numa = {"node": node, "cpus": len(use)}
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local])
        self.torch.cuda.synchronize()

    def allmax(self, value):
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, value):
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def restore_affinity(self):
        if self.affinity0 is not None:
            os.sched_setaffinity(0, self.affinity0)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def host_link_probe(ctx, mbytes=256, reps=6):
    """Ceiling of the host link under the e2e protocol: plain pinned cudaMemcpyAsync, one call per buffer, all ranks
    at the same time.  Returns GB/s per GPU for H2D alone, D2H alone and both directions concurrently (the e2e
    pipeline runs full duplex), each the MIN over ranks (the slowest link bounds the max-over-ranks timing)."""
    torch = ctx.torch
    n = mbytes * (1 << 20) // 8
    h_in = torch.empty(n, dtype=torch.float64, pin_memory=True).fill_(1.0)
    h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.float64, device=ctx.dev)
    d_out = torch.ones(n, dtype=torch.float64, device=ctx.dev)
    s1, s2 = torch.cuda.Stream(ctx.dev), torch.cuda.Stream(ctx.dev)

    def timed(do_h2d, do_d2h):
        ctx.barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(s1); e[2].record(s2)
        for _ in range(reps):
            if do_h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if do_d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        e[1].record(s1); e[3].record(s2)
        s1.synchronize(); s2.synchronize()
        ms = max(e[0].elapsed_time(e[1]) if do_h2d else 0.0, e[2].elapsed_time(e[3]) if do_d2h else 0.0)
        return ctx.allmax(ms)

    timed(True, True)
    nbytes = reps * n * 8
    out = {"h2d_gbs": nbytes / (timed(True, False) * 1e-3) / 1e9, "d2h_gbs": nbytes / (timed(False, True) * 1e-3) / 1e9}
    out["duplex_gbs_each"] = nbytes / (timed(True, True) * 1e-3) / 1e9
    out["note"] = (f"pinned cudaMemcpyAsync of {mbytes} MiB x {reps}, one call per buffer, all {ctx.world} rank(s) concurrently, "
                   "slowest rank; duplex = H2D and D2H at the same time on two streams (GB/s per direction per GPU)")
    del h_in, h_out, d_in, d_out
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(key, solves_per_launch=None):
    """DRAM bytes per launch of kernel `key` from the committed ncu capture, scaled to this launch's size."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        ent = json.load(fh).get(key)
    if not ent:
        return None
    if solves_per_launch is None:
        return ent["bytes"]
    return ent["bytes"] / ent["solves"] * solves_per_launch


def time_e2e(ctx, pipe, host_in, steps):
    """solves/s through LqHostPipeline.submit with pinned host buffers (H2D of every input and D2H of the selected
    outputs inside the timed region), max over ranks."""
    torch = ctx.torch
    for _ in range(3):
        host_res = pipe.submit(*host_in)
    pipe.wait()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(pipe.s_in)
    for _ in range(steps):
        host_res = pipe.submit(*host_in)
    e1.record(pipe.s_out)
    pipe.wait()
    ctx.barrier()
    ms = ctx.allmax(e0.elapsed_time(e1))
    return ctx.world * pipe.batch * steps / (ms * 1e-3), host_res


def k1_riccati_block(ctx, args, peak):
    """a1's literal contract (FHC.ricatti_recursion returns EVERY P_k and K_k, FHC.py:61): mpc_riccati on per-scenario
    models with all gains and all cost-to-go matrices written.  HBM-bound: 3n^2+nm+m^2 values read, N m n + (N+1) n^2 written."""
    torch = ctx.torch
    from model_predictive_control_b200 import lq
    n, m, N = 4, 1, 20
    batch = args.batch or (1 << 20)
    A, B, Q, R, Pf, _ = cfg2b_inputs_torch(batch, 1234 + 2 + 1000 * ctx.rank, ctx.dev, torch.float64)

    # caller-owned result buffers: 3.4 GB of fresh torch allocations per call inside the timed loop made this block's
    # number depend on the state of the caching allocator (0.85 - 1.86 ms per step between runs)
    out = (torch.empty((N, batch, m, n), dtype=torch.float64, device=ctx.dev),
           torch.empty((N + 1, batch, n, n), dtype=torch.float64, device=ctx.dev))

    def step():
        return lq.riccati(A, B, Q, R, Pf, N, all_P=True, out=out)

    for _ in range(3):
        step()
    ctx.barrier()
    steps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        K, P = step()
    e1.record()
    ctx.barrier()
    ms = ctx.allmax(e0.elapsed_time(e1)) / steps
    bytes_solve = 8 * (3 * n * n + n * m + m * m) + 8 * (N * m * n + (N + 1) * n * n)
    f_ric = 4 * n**3 + 6 * n * n * m + 4 * n * m * m + 2 * m**3 + 2 * n * n + m * m
    achieved = bytes_solve * batch / (ms * 1e-3) / 1e9
    fp_peak = lq.fma_peak(torch.float64)
    del K, P, out
    return {"value": ctx.world * batch / (ms * 1e-3), "unit": "recursions/s", "ms_per_step": ms, "steps": steps,
            "config": {"workload": f"k1: FHC.ricatti_recursion (FHC.py:51-61) with every K_k and P_k returned, nx=4 nu=1 N=20, "
                                   f"{batch} per-scenario models per GPU", "batch_per_gpu": batch,
                       "l2": f"{bytes_solve * batch / 1e6:.0f} MB written/read per step > 126 MB L2"},
            "roofline": {"bound": "hbm", "kernel": "riccati_reg_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "bytes_per_solve": bytes_solve, "traffic": load_traffic("riccati_reg_kernel_f64", batch),
                         "fp_pipe": {"flops_per_solve": N * f_ric, "frac": N * f_ric * batch / (ms * 1e-3) / fp_peak}},
            "gpu_launches": steps}


def run_ours(args):
    ctx = Ctx()
    torch, dist = ctx.torch, ctx.dist
    from model_predictive_control_b200 import lq

    world, rank, local, dev = ctx.world, ctx.rank, ctx.local, ctx.dev
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    w = 8 if args.dtype == "f64" else 4
    shp = cfg2_shapes()
    n, m, N, batch = shp["n"], shp["m"], shp["N"], args.batch or shp["batch"]

    A, B, Q, R, Pf, x0 = cfg2b_inputs_torch(batch, 1234 + 2 + 1000 * rank, dev, dtype)
    out = lq.LqSolveBuffers(batch, n, m, N, dtype, dev)

    def step():
        lq.lq_solve(A, B, Q, R, Pf, x0, N, out=out)

    barrier = ctx.barrier
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # keep the clock sampler fed for >= ~0.3 s: untimed extra passes before the timed ones
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.3:
        step()
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ctx.allmax(total_ms)
    ms_per_step = total_ms / args.steps
    value = world * batch * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the public API with HOST buffers (pinned): every step copies all of its
    # inputs host->device and its results device->host inside the timed region.  The public
    # call is lq.LqHostPipeline.submit, which overlaps the copies of consecutive steps (PCIe full duplex).
    # Two output sets: the full one (X, U, V: 840 B per solve) is the headline; (U, V) -- the plan and its cost, the
    # states being a function of inputs and plan -- is 168 B per solve.
    host_in = [torch.empty(t_.shape, dtype=t_.dtype, pin_memory=True).copy_(t_) for t_ in (A, B, Q, R, Pf, x0)]
    pipe = lq.LqHostPipeline(batch, n, m, N, dtype, dev)
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    e2e_steps = max(4, min(args.steps, 8))
    e2e_value, host_res = time_e2e(ctx, pipe, host_in, e2e_steps)
    e2e_check = bool(torch.equal(host_res[2], out.V.cpu()))   # the host result is the device result
    del pipe
    pipe_uv = lq.LqHostPipeline(batch, n, m, N, dtype, dev, outputs=("U", "V"))
    d2h_uv = pipe_uv.d2h_bytes
    e2e_uv, host_uv = time_e2e(ctx, pipe_uv, host_in, e2e_steps)
    e2e_uv_check = bool(torch.equal(host_uv[1], out.V.cpu()))
    del pipe_uv
    link = host_link_probe(ctx)

    # ---- the other reading of configs[1] (cfg 2a: ONE shared model, random initial states): Riccati recursion once per
    # step (a single CTA) + the K2 rollout kernel for every scenario.  Reported beside the headline, not instead of it.
    A0 = torch.eye(n, dtype=dtype, device=dev) + 0.5 * torch.diag(torch.ones(n - 1, dtype=dtype, device=dev), 1)
    B0 = torch.zeros(n, m, dtype=dtype, device=dev); B0[-1, 0] = -0.5
    C0 = torch.tensor([[1.0], [-2.0 / 3.0], [0.0], [0.0]], dtype=dtype, device=dev)
    Q0 = C0 @ C0.t() + 1e-3 * torch.eye(n, dtype=dtype, device=dev); R0 = torch.tensor([[0.1]], dtype=dtype, device=dev)
    x0T = x0.t().contiguous()

    def step2a():
        K2a, _ = lq.riccati(A0, B0, Q0, R0, Q0, N, all_P=False)
        return lq.lq_rollout(A0, B0, K2a[:, 0], x0T, N + 1, gain_offset=0, gain_step=1, Q=Q0, R=R0, Pf=Q0, want_U=True, want_cost=True)

    for _ in range(5):
        step2a()
    barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for _ in range(args.steps):
        step2a()
    eb.record()
    barrier()
    ms2a = ctx.allmax(ea.elapsed_time(eb)) / args.steps

    # ---- the Krylov guard under divergence: the same launch on models spread 4x wider (sigma = 0.2), where ~15 % of
    # the scenarios take the dense recursion inside the launch and most warps execute both bodies
    guard = None
    if dtype == torch.float64 and not args.quick:
        gsd = torch.Generator(device=dev); gsd.manual_seed(77 + rank)
        Aw = A + 0.15 * torch.randn(A.shape, generator=gsd, device=dev, dtype=dtype)
        Bw = B + 0.15 * torch.randn(B.shape, generator=gsd, device=dev, dtype=dtype)
        for _ in range(3):
            lq.lq_solve(Aw, Bw, Q, R, Pf, x0, N, out=out)
        barrier()
        eg0, eg1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eg0.record()
        for _ in range(5):
            lq.lq_solve(Aw, Bw, Q, R, Pf, x0, N, out=out)
        eg1.record()
        barrier()
        msg = ctx.allmax(eg0.elapsed_time(eg1)) / 5
        guard = {"value": world * batch / (msg * 1e-3), "unit": UNIT, "ms_per_step": msg,
                 "krylov_path_fraction": krylov_path_fraction(Aw, Bw, float(os.environ.get("MPC_LQ_KRYLOV_COND", "1e3"))),
                 "note": "same kernel and shapes, per-scenario models perturbed with sigma ~ 0.2 instead of 0.05: the "
                         "conditioning guard sends part of every warp through the dense recursion (divergent warps)"}
        del Aw, Bw
        step()  # restore the headline outputs for the summary below

    # ---- device copy bandwidth under the SAME protocol as the timed loop (0.3 s of back-to-back launches first, then
    # 10 timed ones): what a pure streaming kernel sustains on this board once the power cap has set the clocks.
    # Context only: roofline.peak stays the burst figure of MEASURED_PEAKS.json.
    cp_src = torch.empty(1 << 27, dtype=torch.float64, device=dev)      # 1 GiB
    cp_dst = torch.empty_like(cp_src)
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.3:
        cp_dst.copy_(cp_src)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(10):
        cp_dst.copy_(cp_src)
    c1.record()
    torch.cuda.synchronize()
    copy_sustained_gbs = 10 * 2 * cp_src.numel() * 8 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del cp_src, cp_dst

    # ---- the same launch timed alone after a cool-down (burst clocks), like MEASURED_PEAKS.json's copy bandwidth
    # (best of 10): the timed loop above runs power-capped (sw_power_cap), which the burst figure separates from the kernel
    time.sleep(1.0)
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    burst = []
    for _ in range(10):
        b0.record(); step(); b1.record()
        torch.cuda.synchronize()
        burst.append(b0.elapsed_time(b1))
    barrier()

    # ---- final gather of summaries over NCCL (outside the solve, outside the timed steps)
    summ = torch.stack([out.V.sum(), out.U[0].abs().max(), torch.tensor(float(batch), device=dev, dtype=dtype)]).double()
    if world > 1:
        gathered = [torch.empty_like(summ) for _ in range(world)]
        dist.all_gather(gathered, summ)
        summ_all = torch.stack(gathered)
    else:
        summ_all = summ[None]
    peak, peak_src = load_peaks()
    kernel = lq.lq_solve_kernel_name(n, m, dtype)
    kfrac = (krylov_path_fraction(A, B, float(os.environ.get("MPC_LQ_KRYLOV_COND", "1e3")))
             if kernel == "lq_solve_krylov_kernel" else 0.0)
    del A, B, Q, R, Pf, host_in, out
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs and a1's literal outputs, each a short block of its own (every rank runs them;
    # cfg 5 at 2^20 scenarios per GPU is the named 1/2/4/8-GPU sweep)
    secondary = {}
    if not args.quick:
        secondary["k1_riccati"] = k1_riccati_block(ctx, args, peak)
        for wl, st in (("cfg3", 5), ("cfg4", 2), ("cfg5", 2)):
            line_w = secondary_line(ctx, wl, steps=st, warmup=3 if wl == "cfg3" else 1, cpu=False)
            if rank == 0:
                secondary[wl] = {k: line_w[k] for k in ("value", "unit", "ms_per_step", "steps", "config", "roofline", "e2e",
                                                        "gpu_launches", "summary", "solved_only") if k in line_w}

    if rank == 0:
        bytes_solve = cfg2b_bytes_per_solve(w, n, m, N)
        if kernel == "lq_solve_krylov_kernel":
            flops_solve = kfrac * cfg2b_flops_per_solve_krylov(n, N) + (1 - kfrac) * cfg2b_flops_per_solve(n, m, N)
        else:
            flops_solve = cfg2b_flops_per_solve(n, m, N)
        kern_ms = sum(per_launch) / len(per_launch)
        achieved = bytes_solve * batch / (kern_ms * 1e-3) / 1e9
        fp_peak = lq.fma_peak(dtype)
        per_gpu = lambda v: v / world / batch
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "cfg2b: FHC Riccati LQ solve (backward recursion + optimal plan + cost), nx=4 nu=1 N=20, "
                                   f"{batch} scenarios per GPU, per-scenario model and initial state",
                       "batch_per_gpu": batch, "nx": n, "nu": m, "horizon": N, "parallelism": f"scenario-shard x{world}",
                       "numa": ctx.numa,
                       "l2": (f"inputs+outputs {bytes_solve * batch / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)"
                              if bytes_solve * batch > 126e6 else
                              f"inputs+outputs {bytes_solve * batch / 1e6:.0f} MB per step FIT in the 126 MB L2: not a valid "
                              "timing configuration (use the default batch)")},
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": load_traffic(kernel + "_" + args.dtype, batch),
                         "krylov_path_fraction": kfrac,
                         "sustained_copy": {"gbs": copy_sustained_gbs, "frac_of_it": achieved / copy_sustained_gbs,
                                            "note": "torch copy_ of 1 GiB timed under the same protocol as the loop above "
                                                    "(0.3 s of back-to-back launches, then 10 timed): the streaming rate this "
                                                    "board sustains once the power cap has set the clocks; context for frac"},
                         "burst": {"best_ms": min(burst), "median_ms": sorted(burst)[len(burst) // 2],
                                   "achieved": bytes_solve * batch / (min(burst) * 1e-3) / 1e9,
                                   "frac": bytes_solve * batch / (min(burst) * 1e-3) / 1e9 / peak,
                                   "note": "same launch timed alone after a 1 s cool-down, best of 10 (how the peak itself was "
                                           "measured); achieved/frac above are the average over the timed loop, which runs power-capped"},
                         "peak_source": peak_src, "bytes_per_solve": bytes_solve, "kernel_ms": kern_ms,
                         "fp_pipe": {"flops_per_solve": flops_solve, "achieved_tflops": flops_solve * batch / (kern_ms * 1e-3) / 1e12,
                                     "measured_fma_peak_tflops": fp_peak / 1e12,
                                     "frac": flops_solve * batch / (kern_ms * 1e-3) / fp_peak},
                         "guard_divergence": guard},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "d2h_gbs_per_gpu": d2h * per_gpu(e2e_value) / 1e9, "h2d_gbs_per_gpu": h2d * per_gpu(e2e_value) / 1e9,
                    "host_link": link,
                    "frac_of_host_link": d2h * per_gpu(e2e_value) / 1e9 / link["d2h_gbs"],
                    "frac_of_host_link_duplex": d2h * per_gpu(e2e_value) / 1e9 / link["duplex_gbs_each"],
                    "plan_only": {"value": e2e_uv, "unit": UNIT, "outputs": ["U", "V"], "h2d_bytes_per_step": h2d,
                                  "d2h_bytes_per_step": d2h_uv, "h2d_gbs_per_gpu": h2d * per_gpu(e2e_uv) / 1e9,
                                  "frac_of_host_link": h2d * per_gpu(e2e_uv) / 1e9 / link["h2d_gbs"],
                                  "result_matches_device": e2e_uv_check,
                                  "note": "same call with outputs=('U','V'): the plan and its cost travel back, the predicted "
                                          "states (a function of inputs and plan) stay on the device; now the H2D stream of the "
                                          "per-scenario models (456 B per solve) is the bound"},
                    "note": "lq.LqHostPipeline: pinned host buffers (allocated on the GPU's NUMA node), all model/x0 inputs H2D and "
                            "X/U/V D2H every step, copies of consecutive steps overlapped on separate streams (full duplex); "
                            "the D2H stream (840 B per solve) runs at the host-link rate, which bounds this figure: "
                            "frac_of_host_link = D2H GB/s of this run / the D2H ceiling measured by host_link (all ranks "
                            "concurrently); _duplex = against the ceiling with both directions fully loaded",
                    "result_matches_device": e2e_check},
            "gpu_launches": args.steps, "clocks": clocks,
            "cfg2a_shared_model": {"value": world * batch / (ms2a * 1e-3), "unit": UNIT, "ms_per_step": ms2a,
                                   "kernels": "riccati_reg_kernel (1 CTA) + rollout_shared_kernel",
                                   "bytes_per_solve": w * (n + N * m + (N + 1) * n + 1),
                                   "hbm_frac": w * (n + N * m + (N + 1) * n + 1) * batch / (ms2a * 1e-3) / 1e9 / peak,
                                   "note": "same shapes with ONE shared model: Riccati recursion once per step, optimal plan X, U, cost per scenario"},
            "summary": {"sum_cost": float(summ_all[:, 0].sum()), "max_abs_u0": float(summ_all[:, 1].max()),
                        "scenarios": int(summ_all[:, 2].sum())},
            "secondary": secondary,
        }
        if world == 1 and not args.no_cpu:
            ctx.restore_affinity()
            cores = host_cores()
            kind = cpu_arm_kind()
            rate, done, wall = cpu_reference_rate((8000 if kind == "reference" else 25000) * cores, cores, kind)  # ~10-20 s per core
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"{done} scenarios of cfg2b, {CPU_SAMPLE[kind]} ({wall:.1f} s wall)"}
        print(json.dumps(line), flush=True)
    ctx.close()


def run_cfg2a(args):
    """cfg 2a: shared model, 2^20 random initial states; K1 once (outside the timed steps, it is one
    CTA), then per step ONE K2 launch: N-step rollout with the time-varying gains, X, U and cost written."""
    import torch
    import torch.distributed as dist
    from model_predictive_control_b200 import lq
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    w = 8 if args.dtype == "f64" else 4
    n, m, N, batch = 4, 1, 20, args.batch or (1 << 20)
    dd = dict(dtype=dtype, device=dev)
    A = torch.eye(n, **dd) + 0.5 * torch.diag(torch.ones(n - 1, **dd), 1)
    B = torch.zeros(n, m, **dd); B[-1, 0] = -0.5
    C = torch.tensor([[1.0], [-2.0 / 3.0], [0.0], [0.0]], **dd)
    Q = C @ C.t() + 1e-3 * torch.eye(n, **dd); R = torch.tensor([[0.1]], **dd)
    g = torch.Generator(device=dev); g.manual_seed(1234 + 2 + 1000 * rank)
    x0 = (torch.rand(n, batch, generator=g, device=dev, dtype=torch.float64) * 20 - 10).to(dtype)
    def step():
        # the whole solve of the batch: one Riccati recursion for the shared model (a single CTA), then the rollouts
        K, P = lq.riccati(A, B, Q, R, Q, N, all_P=False)
        return lq.lq_rollout(A, B, K[:, 0], x0, N + 1, gain_offset=0, gain_step=1, Q=Q, R=R, Pf=Q, want_U=True, want_cost=True)

    for _ in range(max(args.warmup, 3)):
        res = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.3:
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local])
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    if rank == 0:
        peak, peak_src = load_peaks()
        bytes_solve = w * (n + N * m + (N + 1) * n + 1)
        achieved = bytes_solve * batch / (ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": world * batch / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": f"cfg2a: shared-model finite-horizon LQ, nx=4 nu=1 N=20, {batch} random initial states per GPU; "
                                       "gains from one Riccati recursion, per-scenario optimal plan X, U and cost by K2",
                           "batch_per_gpu": batch, "l2": f"{bytes_solve * batch / 1e6:.0f} MB per step > 126 MB L2"},
                "roofline": {"bound": "hbm", "kernel": "rollout_shared_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": load_traffic("rollout_shared_kernel_" + args.dtype, batch),
                             "peak_source": peak_src, "bytes_per_solve": bytes_solve, "kernel_ms": ms},
                "gpu_launches": 2 * args.steps, "clocks": clocks,
                "summary": {"sum_cost": float(res["cost"].sum())}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ipm_flops_per_iter(n, m, N):
    """Flops of one interior-point iteration as performed (csrc/boxqp_core.cuh): one factorising
    backward sweep, one feed-forward-only backward sweep, two forward sweeps, one update pass."""
    d = n + m
    fac = 2 * n**3 + 2 * n * n * m + 2 * n * m * m + 2 * n * n * m + 2 * m * m * n + 2 * n * n * m + n * n * (n + 1) + 2 * m**3
    rhs = 2 * (m * m + n * n) + 14 * d
    psweep = 2 * n * m + 2 * m * m + 2 * n * n + 2 * n * m
    fwd = 2 * n * m + 2 * n * n + 2 * n * m + 24 * d
    upd = 16 * d
    return N * (fac + 2 * (rhs + psweep) + 2 * fwd + upd)


def ipm_workspace_bytes_per_iter(n, m, N, model_doubles_per_stage):
    """HBM bytes one interior-point iteration moves BY DESIGN in the thread-per-scenario kernels (csrc/boxqp_core.cuh):
    the solver state lives in a [stage][element][batch] workspace that is streamed once per pass.  Per stage and
    iteration, in doubles: the iterate (z, s_l, s_u, lam_l, lam_u = 5(n+m)) is read by the five passes and written by the
    update (x6); the gains K (mn) are written once and read three times; S^-1 (m^2) written + read; the feed-forward
    (m) written twice + read twice; dz_aff (n+m) written + read three times; dz (n+m) written + read; the per-scenario
    stage model (LTV only) is read by all five passes."""
    d = n + m
    per_stage = 6 * 5 * d + 4 * m * n + 2 * m * m + 4 * m + 4 * d + 2 * d + 5 * model_doubles_per_stage
    return 8 * N * per_stage


def _cpu_qp_worker(args):
    """Exact CPU solves (HiGHS active set + KKT refinement, oracle/boxqp.py) of a slice of cfg3 / cfg5 scenarios."""
    import numpy as np
    from oracle import boxqp as obq
    workload, seed, count = args
    rng = np.random.default_rng(seed)
    if workload == "cfg3":
        p = obq.Problem(N=30)
        A, B, Q, R, N = p.A, p.B, np.asarray(p.Q, float), np.asarray(p.R, float), 30
        ulo, uhi, xlo, xhi = obq.problem_bounds(p)
        x0 = np.stack([rng.uniform(-100, 0, count), rng.uniform(-10, 15, count)], 1)
    else:
        mrng = np.random.default_rng(1234 + 5)
        Ts = 0.1
        Ac = np.array([[1, Ts, Ts * Ts / 2], [0, 1, Ts], [0, 0, 1.0]]); Bc = np.array([[Ts**3 / 6], [Ts * Ts / 2], [Ts]])
        A = np.kron(np.eye(4), Ac) + 0.01 * mrng.standard_normal((12, 12)); B = np.kron(np.eye(4), Bc)
        Q, R, N = np.eye(12), 0.1 * np.eye(4), 50
        ulo, uhi, xlo, xhi = -np.ones(4), np.ones(4), -5 * np.ones(12), 5 * np.ones(12)
        x0 = rng.uniform(-2, 2, (count, 12))
    t0 = time.perf_counter()
    for b in range(count):
        obq.solve_exact(A, B, Q, R, Q, N, x0[b], ulo, uhi, xlo, xhi)
    return time.perf_counter() - t0, count


def cpu_baseline_secondary(workload, cores):
    """Bounded CPU sample of the secondary workloads (restated oracle, not reference code: the reference has no
    solver for cfg 3/5 and its cfg-4 solver, CasADi + IPOPT, is not installable here)."""
    import multiprocessing as mp
    if workload in ("cfg3", "cfg5"):
        per = 1500 if workload == "cfg3" else 8   # ~10-20 s of CPU work
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_cpu_qp_worker, [(workload, 7000 + i, per) for i in range(cores)])
        rate = sum(r[1] for r in res) / max(r[0] for r in res)
        return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{per * cores} scenarios of {workload}, exact active-set QP solves (HiGHS + KKT refinement, oracle/boxqp.py), one process per core; restatement, not reference code"}
    import numpy as np
    from oracle import bicycle as obc
    rng = np.random.default_rng(11)
    nb, nsteps = 64, 48   # ~15 s of CPU work
    x0 = np.array([0.6, -0.25, 0, 0]) + rng.uniform(-0.2, 0.2, (nb, 4)) * np.array([1, 1, 0.5, 0.2])
    t0 = time.perf_counter()
    obc.closed_loop(x0, nsteps, N=50, friction_plant=rng.uniform(0.7, 1.0, nb), qp="port")
    dt = time.perf_counter() - t0
    return {"value": nb * nsteps / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{nb} scenarios x {nsteps} control steps of cfg4, numpy restatement of the RTI loop and of the interior-point QP solver (oracle/bicycle.py, batched numpy, one process); restatement, not reference code"}


def ipm_workspace_bytes_v2(n, m, N, model_elems, narrow=True, elem=8):
    """HBM bytes one interior-point iteration of the round-2 kernels moves BY DESIGN (csrc/boxqp_core.cuh, four sweeps;
    `narrow` = the float64 product's mixed workspace: slacks, multipliers, dz_aff and the corrector data e, g in float32):
      A reads z, s, lam (4 d), dz, dz_aff (+ model) and writes z, s, lam, K, S^-1, d;  B reads z, s, lam, K, d (+ model) and
      writes dz_aff, e, g;  C reads e, g, K, S^-1, d (+ model) and writes d;  D reads z, s, lam, dz_aff, K, d (+ model), writes dz."""
    d = n + m
    z, sl, da = elem * d, (4 if narrow else elem) * 4 * d, (4 if narrow else elem) * d
    dz, eg = elem * d, 2 * (4 if narrow else elem) * d
    K, S, ff, mod = elem * m * n, elem * m * m, elem * m, elem * model_elems
    a = (z + sl + dz + da + mod) + (z + sl + K + S + ff)
    b = (z + sl + K + ff + mod) + (da + eg)
    c = (eg + K + S + ff + mod) + ff
    dd = (z + sl + da + K + ff + mod) + dz
    return N * (a + b + c + dd)


def secondary_line(ctx, workload, steps, warmup, batch=0, horizon=0, cpu=False, dtype="f64"):
    """One bench line (same JSON schema as the headline) for cfg3 / cfg4 / cfg5.  Collective: every rank calls it."""
    torch, dist = ctx.torch, ctx.dist
    from model_predictive_control_b200 import boxqp, distributed as D, lq, problem, session4

    world, rank, local, dev = ctx.world, ctx.rank, ctx.local, ctx.dev
    tdt = torch.float64 if dtype == "f64" else torch.float32
    es = 8 if dtype == "f64" else 4
    g = torch.Generator(device=dev); g.manual_seed(1234 + int(workload[-1]) + 1000 * rank)
    rnd = lambda *shape: torch.rand(*shape, generator=g, device=dev, dtype=torch.float64)
    make_step_on = None
    if workload == "cfg3":
        batch = batch or (1 << 18)
        N = horizon or 30
        prob = problem.Problem(N=N)
        n, m = 2, 1
        x0 = torch.stack([rnd(batch) * 100 - 100, rnd(batch) * 25 - 10], dim=1)
        mpc = problem.LinearMPC(prob)
        x0T = x0.t().contiguous().to(tdt)
        ws = boxqp.BoxQpWorkspace(batch, n, m, N, dev, dtype=tdt)
        A, B = (torch.tensor(M, dtype=tdt, device=dev) for M in (prob.A, prob.B))
        Q, R = (torch.tensor(M.astype(float), dtype=tdt, device=dev) for M in (prob.Q, prob.R))
        u_lo, u_hi, x_lo, x_hi = mpc.bounds()

        def step():
            return boxqp.solve(A, B, Q, R, Q, N, x0T, u_lo, u_hi, x_lo, x_hi, workspace=ws)

        solves_per_step = batch
        name = f"cfg3: session-2 Problem box-QP (input + state bounds), nx=2 nu=1 N={N}, {batch} scenarios per GPU"
        io_bytes = es * (n + N * m + (N + 1) * n + 1) + 8 + N * (n + m)
        host_in, host_out = [x0T], lambda r: [r.U, r.X, r.cost, r.status]
        # shared LTI model, box constraints: the staged kernel (tile stages copied into shared memory by cp.async.bulk)
        kname, model_elems = ("boxqp_ipm_staged_kernel" if os.environ.get("MPC_QP_STAGED") != "0" else "boxqp_ipm_kernel"), 0
    elif workload == "cfg5":
        import numpy as np
        batch = batch or (1 << 20)
        n, m, N = 12, 4, 50
        rng = np.random.default_rng(1234 + 5)      # the shared model is the same on every rank
        Ts = 0.1
        Ac = np.array([[1, Ts, Ts * Ts / 2], [0, 1, Ts], [0, 0, 1.0]]); Bc = np.array([[Ts**3 / 6], [Ts * Ts / 2], [Ts]])
        A = torch.tensor(np.kron(np.eye(4), Ac) + 0.01 * rng.standard_normal((12, 12)), device=dev)
        B = torch.tensor(np.kron(np.eye(4), Bc), device=dev)
        Q = torch.eye(12, dtype=torch.float64, device=dev); R = 0.1 * torch.eye(4, dtype=torch.float64, device=dev)
        x0T = rnd(12, batch) * 4 - 2
        ws = boxqp.BoxQpWorkspace(batch, n, m, N, dev, sat=False)

        def step():
            return boxqp.solve(A, B, Q, R, Q, N, x0T, -1.0, 1.0, -5.0, 5.0, workspace=ws)

        def make_step_on(x_sub):
            ws_sub = boxqp.BoxQpWorkspace(x_sub.shape[1], n, m, N, dev, sat=False)
            return lambda: boxqp.solve(A, B, Q, R, Q, N, x_sub, -1.0, 1.0, -5.0, 5.0, workspace=ws_sub)

        solves_per_step = batch
        name = (f"cfg5: box-QP nx=12 nu=4 N=50 (four coupled triple integrators, |u|<=1, |x|<=5, x0~U[-2,2]^12), "
                f"{batch} scenarios per GPU")
        io_bytes = 8 * (n + N * m + (N + 1) * n + 1) + 8
        host_in, host_out = [x0T], lambda r: [r.U, r.X, r.cost, r.status]
        kname, model_elems = "boxqp_ipm_coop_kernel", 0
    elif workload == "cfg4":
        batch = batch or (1 << 16)
        n, m, N, steps_cl = 4, 2, 50, 200
        par = session4.VehicleParameters()
        scale = torch.tensor([1, 1, 0.5, 0.2], device=dev, dtype=torch.float64)
        x0 = (torch.tensor([0.6, -0.25, 0, 0], device=dev, dtype=torch.float64) + (rnd(batch, 4) * 0.4 - 0.2) * scale).to(tdt)
        fr = (rnd(batch) * 0.3 + 0.7).to(tdt)
        ctrl = session4.MPCController(N=N, ts=0.05, params=par, dtype=tdt)

        def step():
            return ctrl.closed_loop(x0, steps_cl, friction_plant=fr)

        solves_per_step = batch * steps_cl
        name = f"cfg4: session-4 bicycle RTI closed loop, nx=4 nu=2 N=50, {batch} scenarios x {steps_cl} control steps per GPU"
        io_bytes = es * (4 + 1) / steps_cl + es * 6  # per solve: x0 + friction amortised, X_cl/U_cl rows written
        host_in, host_out = [x0, fr], lambda r: [r.X, r.U, r.cost, r.violation]
        kname, model_elems = "rti_closed_loop_kernel", 14
    else:
        raise SystemExit(f"unknown workload {workload}")

    barrier = ctx.barrier
    for _ in range(max(warmup, 1)):
        res = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        res = step()
        evs[i + 1].record()
    barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ctx.allmax(total_ms)
    value = world * solves_per_step * steps / (total_ms * 1e-3)
    kern_ms = total_ms / steps
    # e2e: host x0 in, predictions / closed-loop trajectories out (pinned buffers)
    pin_in = [torch.empty(t_.shape, dtype=t_.dtype, pin_memory=True).copy_(t_) for t_ in host_in]
    outs = host_out(res)
    pin_out = [torch.empty(t_.shape, dtype=t_.dtype, pin_memory=True) for t_ in outs]

    def e2e_step():
        for d_, h_ in zip(host_in, pin_in):
            d_.copy_(h_, non_blocking=True)
        r = step()
        for h_, d_ in zip(pin_out, host_out(r)):
            h_.copy_(d_, non_blocking=True)

    e2e_n = 2 if kern_ms > 500 else 3
    e2e_step(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_n):
        e2e_step()
    e1.record(); barrier()
    e2e_value = world * solves_per_step * e2e_n / (ctx.allmax(e0.elapsed_time(e1)) * 1e-3)
    if workload in ("cfg3", "cfg5"):
        nsat = (res.sat_u != 0).sum(dim=(0, 1)).to(torch.int32).contiguous() if res.sat_u is not None else None
        summ = D.local_summary(cost=res.cost.double(), status=res.status, iters=res.iters, n_saturated=nsat)
        solved = res.status == 1
        it_s = float(res.iters[solved].double().mean()) if bool(solved.any()) else 0.0
        it_f = float(res.iters[~solved].double().mean()) if bool((~solved).any()) else 0.0
    else:
        summ = D.local_summary(cost=res.cost.double(), violation=res.violation.double(), n_saturated=res.n_saturated, iters=res.iters)
        solved, it_s, it_f = None, None, None
    merged = D.gather_summaries(summ)
    # solved-only rate: the same launch on the feasible scenarios alone (infeasible ones leave early through the stall
    # test and would flatter the rate) -- measured, not derived
    solved_only = None
    if make_step_on is not None:
        idx = torch.nonzero(solved).flatten()
        x_sub = x0T[:, idx].contiguous()
        step_sub = make_step_on(x_sub)
        step_sub(); barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        r_sub = step_sub()
        s1.record(); barrier()
        ms_sub = ctx.allmax(s0.elapsed_time(s1))
        n_sub = ctx.allsum(float(idx.numel()))
        solved_only = {"value": n_sub / (ms_sub * 1e-3), "unit": UNIT, "scenarios": n_sub, "ms": ms_sub,
                       "all_solved": bool((r_sub.status == 1).all()), "mean_iters": float(r_sub.iters.double().mean()),
                       "note": "the feasible scenarios of the batch alone, one more launch"}
    line = None
    if rank == 0:
        peak, peak_src = load_peaks()
        fp_peak = lq.fma_peak(torch.float64)
        iters_total = merged["sum_iters"] / world  # per rank (ranks run the same distribution)
        flops = iters_total * ipm_flops_per_iter(n, m, N)
        achieved = io_bytes * solves_per_step / (kern_ms * 1e-3) / 1e9
        traffic = load_traffic(kname + ("_" + workload), solves_per_step)
        wsb = None
        if workload != "cfg5":
            narrow = dtype == "f64" and not (workload == "cfg4" and os.environ.get("MPC_QP_STORE") != "mix") \
                and os.environ.get("MPC_QP_STORE") != "f64"
            wb = ipm_workspace_bytes_v2(n, m, N, model_elems, narrow=narrow, elem=es)
            wsb = {"bytes_per_iter": wb, "achieved": wb * iters_total / (kern_ms * 1e-3) / 1e9,
                   "frac": wb * iters_total / (kern_ms * 1e-3) / 1e9 / peak,
                   "dram_bytes_per_solve_ncu": (traffic / solves_per_step) if traffic else None,
                   "note": "HBM bytes the kernel moves by design (solver workspace streamed once per sweep, "
                           "ipm_workspace_bytes_v2) x iterations performed; self-inflicted traffic, reported for "
                           "transparency -- no roofline credit is claimed for it"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(warmup, 1), "ms_per_step": kern_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": name, "batch_per_gpu": batch, "nx": n, "nu": m, "horizon": N,
                       "parallelism": f"scenario-shard x{world}",
                       "l2": "solver state is a per-lane workspace streamed through L2/HBM every iteration (> 126 MB)"},
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "note": "algorithmic I/O only; the kernel is bound by its FP64 instruction stream and its workspace "
                                 "traffic, see fp_pipe / workspace_stream",
                         "kernel_ms": kern_ms, "workspace_stream": wsb,
                         "fp_pipe": {"mean_iters_per_solve": iters_total / solves_per_step,
                                     "mean_iters_solved": it_s, "mean_iters_not_solved": it_f,
                                     "flops_per_iter": ipm_flops_per_iter(n, m, N),
                                     "achieved_tflops": flops / (kern_ms * 1e-3) / 1e12,
                                     "measured_fma_peak_tflops": fp_peak / 1e12, "frac": flops / (kern_ms * 1e-3) / fp_peak}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": sum(t_.numel() * t_.element_size() for t_ in pin_in),
                    "d2h_bytes_per_step": sum(t_.numel() * t_.element_size() for t_ in pin_out)},
            # cfg 3: mpc_state_order_keys + the solve per step (the device sort between them is torch's)
            "gpu_launches": steps * (2 if workload == "cfg3" else 1), "clocks": clocks, "summary": merged,
        }
        if solved_only is not None:
            line["solved_only"] = solved_only
        if world == 1 and cpu:
            ctx.restore_affinity()
            line["cpu_baseline"] = cpu_baseline_secondary(workload, host_cores())
    res = ws = ctrl = None
    torch.cuda.empty_cache()
    return line


def run_secondary(args):
    """--workload cfg3 | cfg4 | cfg5: one bench line of that config (same JSON schema; the default line is cfg2b with
    these as `secondary` blocks)."""
    ctx = Ctx()
    line = secondary_line(ctx, args.workload, args.steps, max(args.warmup, 3), batch=args.batch, horizon=args.horizon,
                          cpu=not args.no_cpu, dtype=args.dtype)
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()


# ------------------------------------------------------------------------------------------------
# SURVEY 8(f) rows ("next"): the callers either side of the solve, each with its own bench line
# ------------------------------------------------------------------------------------------------
def run_next(args):
    """--workload obstacle | bundles | dare | plant: one JSON line (same schema) per SURVEY 8(f) item.
    obstacle: session_4/main.py controller as RTI with linearised collision rows, step-wise closed loop;
    bundles : prediction bundles of FHC.run_and_plot_traj (closed loop + one prediction per closed-loop state);
    dare    : infinite-horizon cost-to-go / gain as the Riccati fixed point, one per scenario model;
    plant   : the accurate plant step (adaptive Dormand-Prince 5(4)) of session4_sol.exact_integration."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from model_predictive_control_b200 import FHC, lq, session4

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    seed = 1234 + 6 + 1000 * rank
    g = torch.Generator(device=dev); g.manual_seed(seed)
    rnd = lambda *shape: torch.rand(*shape, generator=g, device=dev, dtype=torch.float64)
    w = args.workload
    launches = 1
    cpu = None
    if w == "obstacle":
        batch = args.batch or (1 << 16)
        N, ts, steps_cl = 30, 0.08, 100     # reference protocol: 100 closed-loop steps (session_4/main.py:241-271)
        par = session4.VehicleParameters()
        x_obs = np.array([0.25, 0.0, 0.0, 0.0])
        x0 = torch.tensor([0.3, -0.1, 0.0, 0.0], device=dev, dtype=torch.float64) + (rnd(batch, 4) * 0.1 - 0.05) * \
            torch.tensor([1, 1, 1, 0.0], device=dev, dtype=torch.float64)
        ctrl = session4.ObstacleMPCController(N, ts, par, session4.KinematicBicycle(par, symbolic=True), x_obs)
        plant = session4.exact_integration(session4.KinematicBicycle(par), ts)

        def step():
            return session4.simulate(x0, plant, n_steps=steps_cl, policy=ctrl)

        units, n, m = batch * steps_cl, 4, 2
        launches = 1              # the whole closed loop is ONE rti_closed_loop_kernel launch (round 1: three per control step)
        name = (f"8(f).1 obstacle-avoidance RTI closed loop (session_4/main.py controller, 9 linearised collision rows per stage), "
                f"nx=4 nu=2 N={N}, {batch} scenarios x {steps_cl} control steps per GPU, fused closed loop")
        kname, io_bytes = "rti_closed_loop_kernel", 8 * 6
        host_in, host_out = [x0], lambda r: [r]

        def cpu_fn():
            from oracle import bicycle as obc
            nb, ns = 8, 4
            t0 = time.perf_counter()
            obc.closed_loop_obstacle(x0[:nb].cpu().numpy(), x_obs, ns, N=N, ts=ts, qp="port")
            dt = time.perf_counter() - t0
            return {"value": nb * ns / dt, "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": f"{nb} scenarios x {ns} control steps, numpy restatement of the obstacle RTI loop (oracle/bicycle.py); restatement, not reference code"}
        cpu = cpu_fn
    elif w == "bundles":
        batch = args.batch or (1 << 18)
        n, m, N, n_steps = 2, 1, 10, 30
        A, B = FHC.get_dynamics_discrete(0.5)
        C = np.array([[1.0], [-2.0 / 3.0]])
        Q = C @ C.T + 1e-3 * np.eye(2); R = np.array([0.1])
        dv = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device=dev)
        Ad, Bd, Qd, Rd = dv(A), dv(B), dv(Q), dv(R)
        x0 = rnd(2, batch) * 20 - 10
        _, gains = FHC.ricatti_recursion(Ad, Bd, Qd, Rd, Qd, N)

        def step():
            return FHC.closed_loop_with_predictions(Ad, Bd, Qd, Rd, Qd, x0, N, n_steps=n_steps, gains=gains)[2]

        units = batch * n_steps              # one open-loop prediction per closed-loop state
        launches = 2                         # closed-loop rollout + the batched prediction rollout (rollout_shared_kernel)
        name = (f"8(f).2 prediction bundles of FHC.run_and_plot_traj (FHC.py:87-91): {batch} scenarios x {n_steps} closed-loop "
                f"states x horizon-{N} predictions, nx=2 nu=1")
        kname, io_bytes = "rollout_shared_kernel", 8 * n * (N + 1)   # per prediction: read the start state, write N states
        host_in, host_out = [x0], lambda r: [r]

        def cpu_fn():
            from oracle import lq as olq
            nb = 20000
            xs = x0[:, :nb].cpu().numpy()
            Po, Ko = olq.ricatti_recursion(A, B, Q, R, Q, N)
            t0 = time.perf_counter()
            X = olq.simulate(A, B, xs, Ko, n_steps, mode="receding")
            for t in range(n_steps):
                olq.prediction(A, B, X[:, :, t], Ko, N)
            dt = time.perf_counter() - t0
            return {"value": nb * n_steps / dt, "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": f"{nb} scenarios, column-batched numpy port of LinearSystem.simulate/prediction (the reference's own code runs column-batched unchanged)"}
        cpu = cpu_fn
    elif w == "dare":
        batch = args.batch or (1 << 16)
        n, m, N = 4, 1, 0
        A, B, Q, R, _, _ = cfg2b_inputs_torch(batch, seed, dev, torch.float64)

        def step():
            return FHC.infinite_horizon(A, B, Q, R)

        units = batch
        launches = 5                         # riccati_reg_kernel per doubling of the horizon (64, 128, ...), data dependent
        name = (f"8(f).3 infinite-horizon cost-to-go and gain (FHC.py:97-98,126) as the Riccati fixed point, one per scenario model, "
                f"nx=4 nu=1, {batch} models per GPU")
        kname, io_bytes = "riccati_reg_kernel", 8 * (3 * n * n + n * m + m * m) + 8 * (n * n + m * n)
        host_in, host_out = [A, B, Q, R], lambda r: [r[0], r[1]]

        def cpu_fn():
            from scipy import linalg
            nb = 2000
            An, Bn, Qn, Rn = (t_[:nb].cpu().numpy() for t_ in (A, B, Q, R))
            t0 = time.perf_counter()
            for i in range(nb):
                P = linalg.solve_discrete_are(An[i], Bn[i], Qn[i], Rn[i])
                -np.linalg.inv(Rn[i] + Bn[i].T @ P @ Bn[i]) @ Bn[i].T @ P @ An[i]
            dt = time.perf_counter() - t0
            return {"value": nb / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                    "sample": f"{nb} models, scipy.linalg.solve_discrete_are + K_inf exactly as the reference calls them (FHC.py:97-98,126), python loop"}
        cpu = cpu_fn
    else:  # plant
        batch = args.batch or (1 << 20)
        n, m, N = 4, 2, 0
        par = session4.VehicleParameters()
        ts = 0.05
        x = torch.stack([rnd(batch) * 2 - 1, rnd(batch) * 2 - 1, rnd(batch) * 2 - 1, rnd(batch) - 0.5], 0)
        u = torch.stack([rnd(batch) * 2 - 1, rnd(batch) * 0.768 - 0.384], 0)

        def step():
            return session4.plant_step(par, ts, x, u, substeps=-10)   # adaptive Dormand-Prince, rtol = atol = 1e-10

        units = batch
        name = (f"8(f).4 accurate plant step (session4_sol.exact_integration, :37-56) as adaptive Dormand-Prince 5(4), "
                f"rtol=atol=1e-10, ts={ts}, {batch} states per GPU")
        kname, io_bytes = "bicycle_plant_kernel", 8 * (4 + 2 + 4)
        host_in, host_out = [x, u], lambda r: [r]

        def cpu_fn():
            from oracle import bicycle as obc
            nb = 2000
            xs, us = x[:, :nb].t().cpu().numpy(), u[:, :nb].t().cpu().numpy()
            op = obc.VehicleParameters()
            t0 = time.perf_counter()
            for i in range(nb):
                obc.exact_integration_odeint(xs[i], us[i], ts, op, op.friction)
            dt = time.perf_counter() - t0
            return {"value": nb / dt, "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": f"{nb} states, scipy odeint over [0, ts] per state as the reference's exact_integration does, python loop"}
        cpu = cpu_fn

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        res = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms = total_ms / args.steps
    value = world * units * args.steps / (total_ms * 1e-3)
    pin_in = [torch.empty(t_.shape, dtype=t_.dtype, pin_memory=True).copy_(t_) for t_ in host_in]
    pin_out = [torch.empty(t_.shape, dtype=t_.dtype, pin_memory=True) for t_ in host_out(res)]

    def e2e_step():
        for d_, h_ in zip(host_in, pin_in):
            d_.copy_(h_, non_blocking=True)
        r = step()
        for h_, d_ in zip(pin_out, host_out(r)):
            h_.copy_(d_, non_blocking=True)

    e2e_step(); barrier()
    e0.record()
    for _ in range(3):
        e2e_step()
    e1.record(); barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * units * 3 / (float(t.item()) * 1e-3)
    if rank == 0:
        peak, peak_src = load_peaks()
        achieved = io_bytes * units / (ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "batch_per_gpu": batch, "nx": n, "nu": m, "horizon": N,
                       "parallelism": f"scenario-shard x{world}",
                       "l2": "per-step footprint (inputs, outputs, solver workspace) exceeds the 126 MB L2; no flush needed"},
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "bytes_per_unit": io_bytes,
                         "note": "algorithmic I/O of the dominant kernel per unit of work; step time includes the other launches of the step"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": sum(t_.numel() * t_.element_size() for t_ in pin_in),
                    "d2h_bytes_per_step": sum(t_.numel() * t_.element_size() for t_ in pin_out)},
            "gpu_launches": launches * args.steps, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--batch", type=int, default=0, help="scenarios per GPU (default: the named config's)")
    ap.add_argument("--workload", default="cfg2b", choices=["cfg2b", "cfg2a", "cfg3", "cfg4", "cfg5", "obstacle", "bundles", "dare", "plant"],
                    help="cfg2b (default, BASELINE configs[1]); cfg2a = same shapes with ONE shared model: the Riccati "
                         "recursion runs once, the per-scenario work is the K2 rollout (HBM-bound); cfg3 = session-2 box-QP N=30, 256k scenarios; "
                         "cfg4 = session-4 RTI closed loop, 64k scenarios x 200 steps; "
                         "cfg5 = nx=12 nu=4 N=50 box-QP, 2^20 scenarios per GPU (8M over 8 GPUs); "
                         "obstacle | bundles | dare | plant = the SURVEY 8(f) rows (see run_next)")
    ap.add_argument("--horizon", type=int, default=0, help="cfg3 only: horizon N (default 30, the BASELINE config; the north star's "
                                                          "throughput target is quoted at N = 20)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="cfg2b headline only: skip the secondary blocks (cfg3/4/5, k1) "
                                                         "and the guard-divergence leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "cfg2b":
        run_ours(args)
    elif args.workload == "cfg2a":
        run_cfg2a(args)
    elif args.workload in ("obstacle", "bundles", "dare", "plant"):
        run_next(args)
    else:
        run_secondary(args)


if __name__ == "__main__":
    main()
