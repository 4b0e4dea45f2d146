"""Device-level box-constrained linear MPC QP solver (K4) on torch CUDA tensors.

    min  sum_{k<N} x_k'Q x_k + u_k'R u_k + x_N'Pf x_N
    s.t. x_{k+1} = A_k x_k + B_k u_k + c_k,  u_lo <= u_k <= u_hi,  x_lo <= x_k <= x_hi (k >= 1)

The QP is the one posed by the reference's ``Problem`` data (session_2/problem.py:8-24,
session_3/problem.py:12-28); results carry the fields of the reference's per-solve log
(session_2/log.py:8-12).  Library layout is batch-contiguous: x0 [n, batch], U [N, m, batch],
X [N+1, n, batch].
"""
from __future__ import annotations

from ctypes import POINTER, c_double, c_int, c_int8, c_int32, c_int64, c_void_p
from dataclasses import dataclass

import torch

from . import _lib

_lib.register("mpc_boxqp_workspace_bytes", c_int64, [c_int64, c_int, c_int, c_int, c_int])
_lib.register("mpc_boxqp_solve", c_int,
              [c_void_p] * 3 + [c_int] + [c_void_p] * 16 + [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int,
                                                           c_double, c_int, c_void_p])

_lib.register("mpc_state_order_keys", c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p])
_lib.register("mpc_boxqp_solve_ordered", c_int,
              [c_void_p] * 3 + [c_int] + [c_void_p] * 17 + [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int,
                                                           c_double, c_int, c_void_p])
_lib.register("mpc_boxqp_rows_workspace_bytes", c_int64, [c_int64, c_int, c_int, c_int, c_int, c_int])
_lib.register("mpc_boxqp_solve_rows", c_int,
              [c_void_p] * 3 + [c_int] + [c_void_p] * 9 + [c_int] + [c_void_p] * 11 + [c_int64, c_int64, c_int, c_int, c_int, c_int,
                                                                                  c_double, c_int, c_void_p])

BIG = 1e20  # "no bound"


@dataclass
class BoxQpResult:
    """Batched counterpart of the reference's ControllerLog entry (session_2/log.py:10-12)."""
    U: torch.Tensor        # input_prediction  [N, m, batch]
    X: torch.Tensor        # state_prediction  [N+1, n, batch]
    cost: torch.Tensor     # [batch]
    status: torch.Tensor   # int32 [batch]: 1 solved, 2 max_iter, 3 infeasible
    iters: torch.Tensor    # int32 [batch]
    sat_u: torch.Tensor    # int8 [N, m, batch]: -1 lower, +1 upper, 0 free
    sat_x: torch.Tensor    # int8 [N, n, batch]
    sat_c: torch.Tensor = None   # int8 [N, nc, batch]: -1 where a general row C x >= h is active

    @property
    def solver_success(self):
        return self.status == _lib.MPC_SOLVED

    @property
    def input_prediction(self):   # [batch, N, m] view
        return self.U.permute(2, 0, 1)

    @property
    def state_prediction(self):   # [batch, N+1, n] view
        return self.X.permute(2, 0, 1)


class BoxQpWorkspace:
    """Caller-owned scratch + outputs, reusable across solves of the same shape."""

    def __init__(self, batch, n, m, N, device, sat=True, nc=0, dtype=torch.float64):
        if dtype not in (torch.float64, torch.float32):
            raise ValueError("the box-QP solver takes float64 or float32 arrays")
        dd = dict(dtype=dtype, device=device)
        en = _lib.MPC_F64 if dtype == torch.float64 else _lib.MPC_F32
        nbytes = (_lib.lib().mpc_boxqp_rows_workspace_bytes(batch, n, m, N, nc, en) if nc else
                  _lib.lib().mpc_boxqp_workspace_bytes(batch, n, m, N, en))
        self.sat_c = torch.empty((N, nc, batch), dtype=torch.int8, device=device) if (sat and nc) else None
        self.nc = nc
        self.dtype = dtype
        self.ws = torch.empty(max(nbytes // 8, 1) + 2, dtype=torch.float64, device=device)   # 16-byte aligned scratch
        self.nbytes = nbytes
        self.U = torch.empty((N, m, batch), **dd)
        self.X = torch.empty((N + 1, n, batch), **dd)
        self.cost = torch.empty(batch, **dd)
        self.status = torch.empty(batch, dtype=torch.int32, device=device)
        self.iters = torch.empty(batch, dtype=torch.int32, device=device)
        self.sat_u = torch.empty((N, m, batch), dtype=torch.int8, device=device) if sat else None
        self.sat_x = torch.empty((N, n, batch), dtype=torch.int8, device=device) if sat else None
        self.shape = (batch, n, m, N)


def _vec(v, k, device, name, dtype=torch.float64):
    t = torch.as_tensor(v, dtype=torch.float64, device=device).reshape(-1)
    if t.numel() == 1 and k > 1:
        t = t.expand(k)
    if t.numel() != k:
        raise ValueError(f"{name} must have {k} entries, got {t.numel()}")
    if bool(torch.isnan(t).any()):
        raise ValueError(f"{name} contains NaN (use +-inf for 'no bound')")
    t = torch.clamp(t, min=-BIG, max=BIG)   # +-inf -> "no bound"
    return t.to(dtype).contiguous()


ORDER_MIN_BATCH = 8192   # below this a launch is one partial wave: ordering the scenarios buys nothing


def state_order(x0):
    """Permutation that sorts the scenarios along the Morton (Z-order) curve of their initial states x0 [n, batch]:
    neighbouring initial states -- similar active sets, similar iteration counts -- land in the same warp
    (``mpc_state_order_keys`` + a device sort).  int32 [batch]."""
    n, batch = x0.shape
    lo, hi = torch.aminmax(x0, dim=1)
    lohi = torch.cat([lo, hi]).contiguous()
    keys = torch.empty(batch, dtype=torch.int32, device=x0.device)
    with torch.cuda.device(x0.device):
        _lib.check(_lib.lib().mpc_state_order_keys(_lib.ptr(x0), _lib.ptr(lohi), _lib.ptr(keys), batch, n,
                                                   _lib.dtype_enum(x0), _lib.stream(x0.device)))
    return torch.argsort(keys).to(torch.int32)


def solve(A, B, Q, R, Pf, N, x0, u_lo, u_hi, x_lo, x_hi, c=None, warm_U=None, max_iter=60, eps=1e-9,
          workspace=None, Cg=None, hg=None, order="auto"):
    """Solve ``batch`` QPs.  x0 [n, batch] (CUDA; float64 = the reference's arithmetic, or float32: float32 arrays
    and solver workspace, float64 arithmetic inside the kernel -- the north star's 1e-4 tolerance class).

    LTI: A [n,n], B [n,m] shared (c must be None).
    LTV: A [N, n*n, batch], B [N, n*m, batch], c [N, n, batch] per scenario and stage.
    Bounds are per coordinate, shared by all stages and scenarios; +-inf = unbounded.
    Optional general stage rows  Cg_k x_{k+1} >= hg_k:  Cg [N, nc*n, batch] (row-major rows), hg [N, nc, batch].
    ``order``: which workspace lane solves which scenario.  "auto" (default): shared-model problems of at least
    ORDER_MIN_BATCH scenarios on the thread-per-scenario kernels are solved in the order of :func:`state_order`
    (difficulty-sorted warps; results are bitwise independent of the order and land at each scenario's own index);
    None: lane b solves scenario b; or an int32 permutation of range(batch).
    """
    _lib.require_cuda(A, B, Q, R, Pf, x0)
    dt = x0.dtype
    if dt not in (torch.float64, torch.float32):
        raise ValueError("x0 must be float64 (the reference's arithmetic) or float32")
    for name, t in (("A", A), ("B", B), ("c", c), ("warm_U", warm_U), ("Cg", Cg), ("hg", hg)):
        if t is not None and (t.dtype != dt or t.device != x0.device):
            raise ValueError(f"{name} must share dtype and device with x0 ({dt}, {x0.device}); got {t.dtype}, {t.device}")
    en = _lib.MPC_F64 if dt == torch.float64 else _lib.MPC_F32
    dev = x0.device
    N = int(N)
    ltv = A.dim() == 3
    if ltv:
        n = x0.shape[0]
        m = B.shape[1] // n
        batch = x0.shape[1]
        if tuple(A.shape) != (N, n * n, batch) or tuple(B.shape) != (N, n * m, batch) or c is None or \
                tuple(c.shape) != (N, n, batch):
            raise ValueError("LTV model must be A [N,n*n,batch], B [N,n*m,batch], c [N,n,batch]")
        c = c.contiguous()
    else:
        n, m = B.shape
        if c is not None:
            raise ValueError("affine term c is only supported with per-scenario stage matrices")
        if x0.dim() != 2 or x0.shape[0] != n:
            raise ValueError(f"x0 must be (n={n}, batch), got {tuple(x0.shape)}")
        batch = x0.shape[1]
    A, B, x0 = A.contiguous(), B.contiguous(), x0.contiguous()
    Q, R, Pf = (torch.as_tensor(M, device=dev).to(dt).contiguous() for M in (Q, R, Pf))
    if tuple(Q.shape) != (n, n) or tuple(R.shape) != (m, m) or tuple(Pf.shape) != (n, n):
        raise ValueError("Q, Pf must be (n,n) and R (m,m)")
    ulo, uhi = _vec(u_lo, m, dev, "u_lo", dt), _vec(u_hi, m, dev, "u_hi", dt)
    xlo, xhi = _vec(x_lo, n, dev, "x_lo", dt), _vec(x_hi, n, dev, "x_hi", dt)
    if bool((ulo > uhi).any()) or bool((xlo > xhi).any()):
        raise ValueError("lower bound above upper bound")
    if warm_U is not None:
        if tuple(warm_U.shape) != (N, m, batch):
            raise ValueError(f"warm_U must be (N, m, batch) = {(N, m, batch)}")
        warm_U = warm_U.contiguous()
    nc = 0
    if Cg is not None:
        if hg is None or Cg.dim() != 3 or Cg.shape[0] != N or Cg.shape[2] != batch or Cg.shape[1] % n:
            raise ValueError("rows must be Cg [N, nc*n, batch] with hg [N, nc, batch]")
        nc = Cg.shape[1] // n
        if tuple(hg.shape) != (N, nc, batch):
            raise ValueError(f"hg must be {(N, nc, batch)}")
        Cg, hg = Cg.contiguous(), hg.contiguous()
    w = workspace if workspace is not None else BoxQpWorkspace(batch, n, m, N, dev, nc=nc, dtype=dt)
    if w.shape != (batch, n, m, N) or getattr(w, "nc", 0) != nc or w.dtype != dt or w.U.device != dev:
        raise ValueError(f"workspace was built for {w.shape} (nc={getattr(w, 'nc', 0)}, {w.dtype}, {w.U.device}), "
                         f"need {(batch, n, m, N)} (nc={nc}, {dt}, {dev})")
    if isinstance(order, str):
        if order != "auto":
            raise ValueError('order must be "auto", None or an int32 permutation')
        order = state_order(x0) if (not ltv and not nc and n <= 4 and batch >= ORDER_MIN_BATCH) else None
    if order is not None:
        if nc or order.dtype != torch.int32 or tuple(order.shape) != (batch,) or order.device != dev:
            raise ValueError("order must be an int32 permutation [batch] on the device of x0 (not available with rows)")
        order = order.contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mpc_boxqp_solve_ordered(
                _lib.ptr(A), _lib.ptr(B), _lib.ptr(c), 1 if ltv else 0, _lib.ptr(Q), _lib.ptr(R), _lib.ptr(Pf),
                _lib.ptr(ulo), _lib.ptr(uhi), _lib.ptr(xlo), _lib.ptr(xhi), _lib.ptr(x0), _lib.ptr(warm_U),
                _lib.ptr(w.U), _lib.ptr(w.X), _lib.ptr(w.cost), _lib.ptr(w.status), _lib.ptr(w.iters),
                _lib.ptr(w.sat_u), _lib.ptr(w.sat_x), _lib.ptr(order), _lib.ptr(w.ws), w.nbytes, batch, n, m, N,
                int(max_iter), float(eps), en, _lib.stream(dev)))
        return BoxQpResult(w.U, w.X, w.cost, w.status, w.iters, w.sat_u, w.sat_x)
    if nc:
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mpc_boxqp_solve_rows(
                _lib.ptr(A), _lib.ptr(B), _lib.ptr(c), 1 if ltv else 0, _lib.ptr(Q), _lib.ptr(R), _lib.ptr(Pf),
                _lib.ptr(ulo), _lib.ptr(uhi), _lib.ptr(xlo), _lib.ptr(xhi), _lib.ptr(Cg), _lib.ptr(hg), nc, _lib.ptr(x0),
                _lib.ptr(warm_U), _lib.ptr(w.U), _lib.ptr(w.X), _lib.ptr(w.cost), _lib.ptr(w.status), _lib.ptr(w.iters),
                _lib.ptr(w.sat_u), _lib.ptr(w.sat_x), _lib.ptr(w.sat_c), _lib.ptr(w.ws), w.nbytes, batch, n, m, N,
                int(max_iter), float(eps), en, _lib.stream(dev)))
        return BoxQpResult(w.U, w.X, w.cost, w.status, w.iters, w.sat_u, w.sat_x, w.sat_c)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().mpc_boxqp_solve(
            _lib.ptr(A), _lib.ptr(B), _lib.ptr(c), 1 if ltv else 0, _lib.ptr(Q), _lib.ptr(R), _lib.ptr(Pf),
            _lib.ptr(ulo), _lib.ptr(uhi), _lib.ptr(xlo), _lib.ptr(xhi), _lib.ptr(x0), _lib.ptr(warm_U),
            _lib.ptr(w.U), _lib.ptr(w.X), _lib.ptr(w.cost), _lib.ptr(w.status), _lib.ptr(w.iters),
            _lib.ptr(w.sat_u), _lib.ptr(w.sat_x), _lib.ptr(w.ws), w.nbytes, batch, n, m, N, int(max_iter),
            float(eps), en, _lib.stream(dev)))
    return BoxQpResult(w.U, w.X, w.cost, w.status, w.iters, w.sat_u, w.sat_x)


_lib.register("mpc_condense", c_int, [c_void_p, c_int64] * 5 + [c_void_p] * 4 + [c_int64, c_int, c_int, c_int, c_int, c_void_p])


def condense(A, B, Q, R, Pf, N):
    """Condensed prediction matrices (K3).  Models shared ([n,n], ...) or batched ([batch,n,n], ...).
    Returns Phi [.., N n, n], Gamma [.., N n, N m], H [.., N m, N m], F [.., N m, n] with
    H = Gamma' Qbar Gamma + Rbar, F = Gamma' Qbar Phi, so J(U) = U'HU + 2 x0'F'U + const.  Q and Pf must be symmetric
    (as every weight of the reference is); a non-symmetric weight raises."""
    from .lq import _common_batch, _model, _same_kind
    _lib.require_cuda(A, B, Q, R, Pf)
    _same_kind(A, B=B, Q=Q, R=R, Pf=Pf)
    # the kernel contracts (Q Gamma) with Phi, i.e. it forms Gamma' Qbar' Phi: F = Gamma' Qbar Phi needs symmetric weights
    for name, M in (("Q", Q), ("Pf", Pf)):
        if float((M - M.transpose(-1, -2)).abs().max()) > 1e-12 * max(1.0, float(M.abs().max())):
            raise ValueError(f"{name} must be symmetric (F = Gamma' Qbar Phi is formed from Q Gamma)")
    n, m = A.shape[-1], B.shape[-1]
    A, bA, sA = _model(A, n, n, "A")
    B, bB, sB = _model(B, n, m, "B")
    Q, bQ, sQ = _model(Q, n, n, "Q")
    R, bR, sR = _model(R, m, m, "R")
    Pf, bP, sP = _model(Pf, n, n, "Pf")
    b = _common_batch([bA, bB, bQ, bR, bP])
    batch = b or 1
    N = int(N)
    dd = dict(dtype=A.dtype, device=A.device)
    Phi = torch.empty((batch, N * n, n), **dd)
    Gam = torch.empty((batch, N * n, N * m), **dd)
    H = torch.empty((batch, N * m, N * m), **dd)
    F = torch.empty((batch, N * m, n), **dd)
    with torch.cuda.device(A.device):
        _lib.check(_lib.lib().mpc_condense(_lib.ptr(A), sA, _lib.ptr(B), sB, _lib.ptr(Q), sQ, _lib.ptr(R), sR,
                                           _lib.ptr(Pf), sP, _lib.ptr(Phi), _lib.ptr(Gam), _lib.ptr(H), _lib.ptr(F),
                                           batch, n, m, N, _lib.dtype_enum(A), _lib.stream(A.device)))
    if b is None:
        return Phi[0], Gam[0], H[0], F[0]
    return Phi, Gam, H, F
