"""Device-level LQ operators on torch CUDA tensors (thin wrappers over the C ABI).

These are the building blocks under the reference-shaped API in ``FHC.py``,
``LinearSystem.py`` and ``session1_sol.py``.  Layouts are the library's (see
include/mpc_b200.h): models ``[batch, n, n]`` (or shared ``[n, n]``), rollouts batch-contiguous
``[T, n, batch]``, fused solves stage-major ``[N+1, batch, n]``.
"""
from __future__ import annotations

import torch

from ctypes import c_int, c_int64, c_void_p

from . import _lib


def _model(t, rows, cols, name):
    """Return (contiguous tensor, batch or None, stride in elements)."""
    if t.dim() == 2:
        if tuple(t.shape) != (rows, cols):
            raise ValueError(f"{name} must be ({rows},{cols}), got {tuple(t.shape)}")
        return t.contiguous(), None, 0
    if t.dim() == 3 and tuple(t.shape[1:]) == (rows, cols):
        return t.contiguous(), t.shape[0], rows * cols
    raise ValueError(f"{name} must be ({rows},{cols}) or (batch,{rows},{cols}), got {tuple(t.shape)}")


def _same_kind(ref, **named):
    """Every operand shares dtype and device with ``ref`` -- the kernels reinterpret raw pointers, so a float32 matrix
    next to a float64 state would be read as garbage without this check."""
    for name, t in named.items():
        if t is not None and (t.dtype != ref.dtype or t.device != ref.device):
            raise ValueError(f"{name} must share dtype and device with the other operands "
                             f"({ref.dtype}, {ref.device}); got {t.dtype}, {t.device}")


def _common_batch(bs, extra=None):
    vals = {b for b in bs if b is not None}
    if extra is not None:
        vals.add(extra)
    if len(vals) > 1:
        raise ValueError(f"inconsistent batch sizes {sorted(vals)}")
    return vals.pop() if vals else None


def riccati(A, B, Q, R, Pf, N, all_P=True, out=None):
    """Backward Riccati recursion (K1).  Returns K [N, batch, m, n], P [N+1, batch, n, n]
    (or [batch, n, n] = P_0 when ``all_P`` is false).  batch = 1 when every matrix is shared.
    ``out=(K, P)``: caller-owned result buffers of exactly those shapes (repeated solves without allocation).
    Replaces reference session_1/FHC.py:51-61."""
    _lib.require_cuda(A, B, Q, R, Pf)
    _same_kind(A, B=B, Q=Q, R=R, P_f=Pf)
    n, m = A.shape[-1], B.shape[-1]
    A, bA, sA = _model(A, n, n, "A")
    B, bB, sB = _model(B, n, m, "B")
    Q, bQ, sQ = _model(Q, n, n, "Q")
    R, bR, sR = _model(R, m, m, "R")
    Pf, bP, sP = _model(Pf, n, n, "P_f")
    batch = _common_batch([bA, bB, bQ, bR, bP]) or 1
    N = int(N)
    if N < 0:
        raise ValueError("horizon must be non-negative")
    k_shape, p_shape = (N, batch, m, n), ((N + 1, batch, n, n) if all_P else (batch, n, n))
    if out is None:
        K = torch.empty(k_shape, dtype=A.dtype, device=A.device)
        P = torch.empty(p_shape, dtype=A.dtype, device=A.device)
    else:
        K, P = out
        for name, t, shp in (("K", K, k_shape), ("P", P, p_shape)):
            if tuple(t.shape) != shp or t.dtype != A.dtype or t.device != A.device or not t.is_contiguous():
                raise ValueError(f"out {name} must be a contiguous {A.dtype} tensor of shape {shp} on {A.device}")
    with torch.cuda.device(A.device):
        _lib.check(_lib.lib().mpc_riccati(
            _lib.ptr(A), sA, _lib.ptr(B), sB, _lib.ptr(Q), sQ, _lib.ptr(R), sR, _lib.ptr(Pf), sP,
            _lib.ptr(K), _lib.ptr(P), 1 if all_P else 0, batch, n, m, N, _lib.dtype_enum(A),
            _lib.stream(A.device)))
    return K, P


def lq_rollout(A, B, K, x0, T, gain_offset=0, gain_step=0, Q=None, R=None, Pf=None,
               want_U=False, want_cost=False, want_unstable=False, norm_limit=100.0):
    """Batched linear rollout under state feedback (K2).

    x0 [n, batch]; K [ng, m, n] (shared) or [ng, batch, m, n]; A, B shared or [batch, ., .].
    Returns dict with X [T, n, batch] and optionally U [T-1, m, batch], cost [batch],
    unstable [batch] (uint8).  Transition i uses gain ``gain_offset + gain_step * i``.
    Replaces reference session_1/LinearSystem.py:20-35 + FHC.py:25-29."""
    _lib.require_cuda(A, B, K, x0)
    _same_kind(x0, A=A, B=B, K=K, Q=Q if want_cost else None, R=R if want_cost else None, Pf=Pf if want_cost else None)
    n, m = A.shape[-1], B.shape[-1]
    if x0.dim() != 2 or x0.shape[0] != n:
        raise ValueError(f"x0 must be (n={n}, batch), got {tuple(x0.shape)}")
    batch = x0.shape[1]
    x0 = x0.contiguous()
    A, bA, sA = _model(A, n, n, "A")
    B, bB, sB = _model(B, n, m, "B")
    T = int(T)
    if T < 1:
        raise ValueError("need at least one state")
    ng_needed = gain_offset + gain_step * (T - 2) + 1 if T >= 2 else 0
    if K.dim() == 3:
        K = K.contiguous()
        sKs, sK, bK = m * n, 0, None
    elif K.dim() == 4:
        K = K.contiguous()
        bK = K.shape[1]
        sKs, sK = bK * m * n, m * n
    else:
        raise ValueError("gains must be [ng, m, n] or [ng, batch, m, n]")
    if tuple(K.shape[-2:]) != (m, n):
        raise ValueError(f"gain blocks must be ({m},{n})")
    if K.shape[0] < ng_needed:
        raise IndexError(f"policy needs gains[{ng_needed - 1}] but only {K.shape[0]} gains were set")
    _common_batch([bA, bB, bK], batch)
    X = torch.empty((T, n, batch), dtype=x0.dtype, device=x0.device)
    U = torch.empty((max(T - 1, 0), m, batch), dtype=x0.dtype, device=x0.device) if want_U else None
    cost = torch.empty((batch,), dtype=x0.dtype, device=x0.device) if want_cost else None
    unstable = torch.empty((batch,), dtype=torch.uint8, device=x0.device) if want_unstable else None
    if want_cost:
        if Q is None or R is None or Pf is None:
            raise ValueError("cost needs Q, R and Pf")
        Q, R, Pf = Q.contiguous(), R.contiguous(), Pf.contiguous()
    with torch.cuda.device(x0.device):
        _lib.check(_lib.lib().mpc_lq_rollout(
            _lib.ptr(A), sA, _lib.ptr(B), sB, _lib.ptr(K), sKs, sK, int(gain_offset), int(gain_step),
            _lib.ptr(x0), _lib.ptr(X), _lib.ptr(U), _lib.ptr(Q) if want_cost else None,
            _lib.ptr(R) if want_cost else None, _lib.ptr(Pf) if want_cost else None, _lib.ptr(cost),
            _lib.ptr(unstable), float(norm_limit), batch, n, m, T, _lib.dtype_enum(x0),
            _lib.stream(x0.device)))
    return {"X": X, "U": U, "cost": cost, "unstable": unstable}


def linear_step(A, B, x, u):
    """x+ = A x + B u for x [n, batch], u [m, batch] (reference LinearSystem.py:16-18)."""
    _lib.require_cuda(A, B, x, u)
    _same_kind(x, A=A, B=B, u=u)
    n, m = A.shape[-1], B.shape[-1]
    x, u = x.contiguous(), u.contiguous()
    if x.shape[0] != n or u.shape[0] != m or x.shape[1] != u.shape[1]:
        raise ValueError(f"shape mismatch: x {tuple(x.shape)}, u {tuple(u.shape)} for n={n}, m={m}")
    xn = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().mpc_linear_step(_lib.ptr(A.contiguous()), _lib.ptr(B.contiguous()), _lib.ptr(x),
                                              _lib.ptr(u), _lib.ptr(xn), x.shape[1], n, m,
                                              _lib.dtype_enum(x), _lib.stream(x.device)))
    return xn


class LqSolveBuffers:
    """Pre-allocated outputs of :func:`lq_solve` so that a timed loop allocates nothing."""

    def __init__(self, batch, n, m, N, dtype, device, want_K=False, want_P0=False):
        self.X = torch.empty((N + 1, batch, n), dtype=dtype, device=device)
        self.U = torch.empty((N, batch, m), dtype=dtype, device=device)
        self.V = torch.empty((batch,), dtype=dtype, device=device)
        self.K = torch.empty((N, batch, m, n), dtype=dtype, device=device) if want_K else None
        self.P0 = torch.empty((batch, n, n), dtype=dtype, device=device) if want_P0 else None

    def check(self, batch, n, m, N, dtype, device, want_K, want_P0):
        """A caller-supplied buffer set must match the call it is used for (the kernel writes through raw pointers)."""
        want = {"X": (N + 1, batch, n), "U": (N, batch, m), "V": (batch,), "K": (N, batch, m, n), "P0": (batch, n, n)}
        for name, shape in want.items():
            t = getattr(self, name)
            if t is None:
                if name in ("X", "U", "V") or (name == "K" and want_K) or (name == "P0" and want_P0):
                    raise ValueError(f"out.{name} is missing (build the buffers with want_K / want_P0 as the call needs)")
                continue
            if tuple(t.shape) != shape or t.dtype != dtype or t.device != torch.device(device) or not t.is_contiguous():
                raise ValueError(f"out.{name} must be a contiguous {shape} {dtype} tensor on {device}; "
                                 f"got {tuple(t.shape)}, {t.dtype}, {t.device}")


def lq_solve(A, B, Q, R, Pf, x0, N, want_K=False, want_P0=False, out=None):
    """Fused per-scenario finite-horizon LQ solve (K1+K2): x0 [batch, n] ->
    X [N+1, batch, n], U [N, batch, m], V [batch] (+ K [N, batch, m, n], P0 [batch, n, n]).
    Models shared ([n,n]) or per scenario ([batch,n,n]).  Q and P_f must be symmetric (their upper
    triangles are used).  Single-input float64 solves without K/P0 run in Krylov coordinates
    (``lq_solve_kernel_name``; env ``MPC_LQ_KRYLOV_COND`` = conditioning bound of the guard, 0 = off):
    same plan as the dense recursion within ~1e-10 of each scenario's scale on well-conditioned
    models, with a per-scenario fallback to the dense recursion otherwise."""
    _lib.require_cuda(A, B, Q, R, Pf, x0)
    _same_kind(x0, A=A, B=B, Q=Q, R=R, P_f=Pf)
    n, m = A.shape[-1], B.shape[-1]
    if x0.dim() != 2 or x0.shape[1] != n:
        raise ValueError(f"x0 must be (batch, n={n}), got {tuple(x0.shape)}")
    batch = x0.shape[0]
    x0 = x0.contiguous()
    A, bA, sA = _model(A, n, n, "A")
    B, bB, sB = _model(B, n, m, "B")
    Q, bQ, sQ = _model(Q, n, n, "Q")
    R, bR, sR = _model(R, m, m, "R")
    Pf, bP, sP = _model(Pf, n, n, "P_f")
    _common_batch([bA, bB, bQ, bR, bP], batch)
    N = int(N)
    if out is None:
        out = LqSolveBuffers(batch, n, m, N, x0.dtype, x0.device, want_K, want_P0)
    else:
        out.check(batch, n, m, N, x0.dtype, x0.device, want_K, want_P0)
    # the fused kernels keep the gains on chip (N m n elements per thread of shared memory): longer horizons and the
    # shapes without a register-resident kernel take the two-launch path
    if (n, m) not in FUSED_SHAPES or N * m * n * x0.element_size() > FUSED_GAIN_BYTES:
        return _lq_solve_composed(A, B, Q, R, Pf, x0, N, out)
    with torch.cuda.device(x0.device):
        _lib.check(_lib.lib().mpc_lq_solve(
            _lib.ptr(A), sA, _lib.ptr(B), sB, _lib.ptr(Q), sQ, _lib.ptr(R), sR, _lib.ptr(Pf), sP,
            _lib.ptr(x0), _lib.ptr(out.X), _lib.ptr(out.U), _lib.ptr(out.V), _lib.ptr(out.K),
            _lib.ptr(out.P0), batch, n, m, N, _lib.dtype_enum(x0), _lib.stream(x0.device)))
    return out


FUSED_SHAPES = {(2, 1), (4, 1), (4, 2)}   # (n, m) with a register-resident fused kernel behind mpc_lq_solve
FUSED_GAIN_BYTES = 220 * 1024 // 32        # on-chip gains per thread: a 32-thread CTA's slice must fit 220 KB (csrc/lq.cu)


def _lq_solve_composed(A, B, Q, R, Pf, x0, N, out):
    """Same result for shapes without a fused kernel (n <= 32, m <= 16): the recursion (mpc_riccati, K1) followed by the
    gain rollout (mpc_lq_rollout, K2) -- two launches, the gains take a round trip through HBM -- and
    V = x0' P_0 x0 (FHC.py:123-124).  Still entirely on the device."""
    batch, n = x0.shape
    K, P0 = riccati(A, B, Q, R, Pf, N, all_P=False)            # K [N, 1 or batch, m, n], P0 [1 or batch, n, n]
    res = lq_rollout(A, B, K if K.shape[1] > 1 else K[:, 0], x0.t().contiguous(), N + 1, gain_offset=0, gain_step=1,
                     want_U=True)
    out.X.copy_(res["X"].permute(0, 2, 1))
    out.U.copy_(res["U"].permute(0, 2, 1))
    out.V.copy_(torch.einsum("bi,bij,bj->b", x0, P0.expand(batch, n, n), x0))
    if out.K is not None:
        out.K.copy_(K.expand(N, batch, K.shape[2], n))
    if out.P0 is not None:
        out.P0.copy_(P0.expand(batch, n, n))
    return out


def lq_solve_kernel_name(n, m, dtype, want_K=False, want_P0=False):
    """Name of the kernel :func:`lq_solve` launches for these arguments (``mpc_lq_solve_variant``):
    single-input fp64 solves without K/P0 outputs run in Krylov coordinates."""
    enum = 0 if dtype == torch.float64 else 1
    v = _lib.lib().mpc_lq_solve_variant(int(n), int(m), enum, int(bool(want_K or want_P0)))
    return "lq_solve_krylov_kernel" if v == 1 else "lq_solve_kernel"


class LqHostPipeline:
    """End-to-end fused LQ solves for callers whose data lives in HOST memory.

    ``submit(A, B, Q, R, Pf, x0)`` takes pinned host tensors of one batch and returns pinned host
    tensors, one per entry of ``outputs`` (default ``(X, U, V)``).  Consecutive submissions are software-pipelined over three CUDA streams
    with double-buffered device storage: the H2D copy of batch i+1 overlaps the kernel and the D2H
    copy of batch i (PCIe is full duplex), so the steady-state cost per batch is
    max(H2D, D2H, kernel) rather than their sum.  ``wait()`` blocks until everything submitted has
    landed in the returned host buffers.
    """

    def __init__(self, batch, n, m, N, dtype=torch.float64, device=None, per_scenario_model=True,
                 outputs=("X", "U", "V")):
        """``outputs``: which results travel back to the host, in this order -- any of "X" (predicted states,
        (N+1) n values per solve), "U" (the optimal plan, N m) and "V" (the optimal cost).  The states are a function of
        the inputs and the plan (x+ = A x + B u), so a caller that only applies / logs the plan can leave X on the
        device: ("U", "V") is 168 B instead of 840 B of D2H per solve at cfg 2b."""
        if not outputs or any(o not in ("X", "U", "V") for o in outputs):
            raise ValueError('outputs must be a non-empty selection of "X", "U", "V"')
        self.outputs = tuple(outputs)
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.batch, self.n, self.m, self.N, self.dtype = batch, n, m, N, dtype
        mb = (batch,) if per_scenario_model else ()
        shapes = [mb + (n, n), mb + (n, m), mb + (n, n), mb + (m, m), mb + (n, n), (batch, n)]
        self.dev_in = [[torch.empty(sh, dtype=dtype, device=self.dev) for sh in shapes] for _ in range(2)]
        self.dev_out = [LqSolveBuffers(batch, n, m, N, dtype, self.dev) for _ in range(2)]
        self.host_out = [[torch.empty(getattr(o, name).shape, dtype=dtype, pin_memory=True) for name in self.outputs]
                         for o in self.dev_out]
        self.s_in, self.s_c, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_c = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        self.count = 0
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.dev_in[0])
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.host_out[0])

    def submit(self, A, B, Q, R, Pf, x0):
        s = self.count % 2
        first_use = self.count < 2
        with torch.cuda.stream(self.s_in):
            if not first_use:
                self.s_in.wait_event(self.ev_c[s])        # the kernel that last read these inputs is done
            for d, h in zip(self.dev_in[s], (A, B, Q, R, Pf, x0)):
                d.copy_(h, non_blocking=True)
            self.ev_in[s].record(self.s_in)
        with torch.cuda.stream(self.s_c):
            self.s_c.wait_event(self.ev_in[s])
            if not first_use:
                self.s_c.wait_event(self.ev_out[s])       # the D2H that last read these outputs is done
            lq_solve(*self.dev_in[s], self.N, out=self.dev_out[s])
            self.ev_c[s].record(self.s_c)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_c[s])
            o = self.dev_out[s]
            for h, name in zip(self.host_out[s], self.outputs):
                h.copy_(getattr(o, name), non_blocking=True)
            self.ev_out[s].record(self.s_out)
        self.count += 1
        return self.host_out[s]

    def wait(self):
        for st in (self.s_in, self.s_c, self.s_out):
            st.synchronize()


_lib.register("mpc_spectral_radius", c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                             c_int, c_int, c_int, c_void_p])


def spectral_radius(A, B, K):
    """rho(A + B K) per scenario: the exact stability test of the linear closed loop under u = K x that the
    reference leaves as an exercise (session_1/session1_sol.py:114-116).  A, B, K shared ([n,n], [n,m], [m,n]) or
    batched with a leading axis.  Returns a tensor [batch] (batch = 1 when everything is shared)."""
    _lib.require_cuda(A, B, K)
    _same_kind(A, B=B, K=K)
    n, m = A.shape[-1], B.shape[-1]
    A, bA, sA = _model(A, n, n, "A")
    B, bB, sB = _model(B, n, m, "B")
    K, bK, sK = _model(K, m, n, "K")
    batch = _common_batch([bA, bB, bK]) or 1
    rho = torch.empty((batch,), dtype=A.dtype, device=A.device)
    with torch.cuda.device(A.device):
        _lib.check(_lib.lib().mpc_spectral_radius(_lib.ptr(A), sA, _lib.ptr(B), sB, _lib.ptr(K), sK, _lib.ptr(rho),
                                                  batch, n, m, _lib.dtype_enum(A), _lib.stream(A.device)))
    return rho


def fma_peak(dtype=torch.float64):
    """Measured FMA-pipe FLOP/s of the current device (roofline denominator for fp kernels)."""
    import ctypes
    out = ctypes.c_double(0.0)
    _lib.check(_lib.lib().mpc_fma_peak_probe(_lib.MPC_F64 if dtype == torch.float64 else _lib.MPC_F32,
                                             ctypes.byref(out)))
    return out.value
