"""Batched N-field minimum finder / phase follower on the device, and the two-field
parameter scan built on it (BASELINE.json configs[4] on the real examples/testModel1.py model).

Reference code this stands in for (one scipy.optimize.fmin -- Nelder-Mead -- call per point and temperature):
  generic_potential.findMinimum                   generic_potential.py:477-484
  transitionFinder.traceMinimum / traceMultiMin   transitionFinder.py:34-128, 371-553
  the polish of both minima before a tunnelling   transitionFinder.py:658-664
All numerical work is one kernel launch per call (ctb_trace_minima: one warp per instance, one finite-difference
stencil point per lane); torch is the container."""
import numpy as np

from . import _lib, _tables, batch

FLAG_NAMES = {_lib.MIN_OK: "ok", _lib.MIN_NOT_CONVERGED: "not converged", _lib.MIN_NOT_A_MINIMUM: "not a minimum",
              _lib.MIN_JUMPED: "jumped to another phase", _lib.MIN_NOT_TRACED: "not traced"}


def trace_minima(model8, X0, T, t_begin=None, t_dir=None, x_eps=1e-3, xtol=1e-9, jump_tol=0.1, max_iter=60, model=None):
    """Follow the minimum of Vtot(., T) that starts at X0[i] along a temperature grid.

    model = None: the built-in model1 (examples/testModel1.py); model8: [8] (one model for all instances) or [n, 8] rows
    (potentials.model1_model8); X0: [n, 2].  model = "thermal1f": the general one-field model, rows of 64
    (potentials.thermal1f_params; the T entry is ignored), X0 [n, 1].  model = a jit.CompiledModel: model8 holds its
    parameter rows ([nparam] or [n, nparam]) and X0 is [n, model.ndim].
    T: [nT] shared grid or [n, nT] one row per instance; t_begin / t_dir: optional [n] start index and +1 / -1
    direction of the walk.  Returns device tensors dict(X [n, nT, ndim], V [n, nT], flag [n, nT] (MIN_*), iters [n]).
    An instance's walk stops at the first temperature whose flag is not MIN_OK; that entry holds the point the
    iteration ended on (for MIN_JUMPED: a minimum of the neighbouring phase)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    builtin = {None: (_lib.MODEL_MODEL1, 2, 8), "model1": (_lib.MODEL_MODEL1, 2, 8),
               "thermal1f": (_lib.MODEL_THERMAL1F, 1, _lib.NPARAM[_lib.POT_THERMAL1F])}
    if model is None or isinstance(model, str):
        kind_id, ndim, nparam = builtin[model]
    else:
        kind_id, ndim, nparam = None, model.ndim, max(1, model.nparam)
    dev = X0.device if isinstance(X0, torch.Tensor) and X0.is_cuda else torch.device("cuda", torch.cuda.current_device())
    X0d = batch._as_device(torch, X0, dev).reshape(-1, ndim).contiguous()
    n = X0d.shape[0]
    m8 = batch._as_device(torch, model8, dev).reshape(-1, nparam).contiguous()
    if m8.shape[0] not in (1, n):
        raise ValueError("the model parameters must be one row or one row per instance")
    Td = batch._as_device(torch, T, dev).contiguous()
    if Td.dim() == 1:
        nT, T_stride = Td.numel(), 0
    elif Td.dim() == 2 and Td.shape[0] == n:
        nT, T_stride = Td.shape[1], Td.shape[1]
    else:
        raise ValueError("T must be [nT] or [n, nT]")
    i32 = lambda v: None if v is None else torch.as_tensor(v, device=dev).to(torch.int32).reshape(n).contiguous()
    tb, td = i32(t_begin), i32(t_dir)
    if tb is not None and n and nT and (int(tb.min()) < 0 or int(tb.max()) >= nT):
        raise ValueError("t_begin out of range")
    X = torch.empty((n, nT, ndim), dtype=torch.float64, device=dev)
    V = torch.empty((n, nT), dtype=torch.float64, device=dev)
    flag = torch.empty((n, nT), dtype=torch.int32, device=dev)
    iters = torch.empty((n,), dtype=torch.int32, device=dev)
    opts = _lib.MinOpts(float(x_eps), float(xtol), float(jump_tol), int(max_iter), 0)
    stride = nparam if (m8.shape[0] == n and n > 1) else 0
    if kind_id is None:
        model.trace_minima_raw(m8, stride, X0d, Td, T_stride, tb, td, n, nT, opts, X, V, flag, iters, dev)
        return dict(X=X, V=V, flag=flag, iters=iters)
    tabs = _tables.device_tables(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ctb_trace_minima(kind_id, m8.data_ptr(), stride, tabs.jb,
                                        tabs.jf if kind_id == _lib.MODEL_THERMAL1F else None, X0d.data_ptr(), Td.data_ptr(), T_stride, tb.data_ptr() if tb is not None else None,
                                        td.data_ptr() if td is not None else None, n, nT, opts, X.data_ptr(), V.data_ptr(),
                                        flag.data_ptr(), iters.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                   "ctb_trace_minima")
    return dict(X=X, V=V, flag=flag, iters=iters)


def find_minimum(model8, X, T, **knobs):
    """generic_potential.findMinimum for n (X_i, T_i) pairs at once: X [n, ndim], T [n] -> dict(X [n, ndim], V [n],
    flag [n]); model= as in trace_minima."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    Td = batch._as_device(torch, T, dev).reshape(-1, 1)
    r = trace_minima(model8, X, Td, **knobs)
    return dict(X=r["X"][:, 0, :], V=r["V"][:, 0], flag=r["flag"][:, 0], iters=r["iters"])


def model1_phase_scan(model8, T, alpha=2, target=140.0, bounce_mask=None, **knobs):
    """P parameter points x nT temperatures of the two-field model1: both phases of its first-order transition are
    followed on the device, then one bounce per (point, T) is solved along the straight line between the two minima.

    model8 [P, 8]; T [nT] ascending.  Steps (each one launch over all points):
      1. phase A: from approxZeroTMin()[0] = (v, v) at T[0] upwards until it ends (spinodal: the Newton iterate falls
         into the neighbouring basin -> MIN_JUMPED, and the point it lands on is a minimum of phase B);
      2. phase B: from that landing point downwards in T (what traceMultiMin does after a phase ends,
         transitionFinder.py:466-500);
      3. wherever both exist and V_B > V_A: SingleFieldInstanton along X(s) = X_A + s nhat, s in [0, |X_B - X_A|], from
         the metastable B to A (CTB_POT_MODEL1_LINE), the straight-line stand-in for pathDeformation.fullTunneling.
    Returns device tensors: XA, XB [P, nT, 2], VA, VB, flagA, flagB, S, S_over_T, status [P, nT] (NaN / -1 where no
    bounce was solved), Tcrit [P] (V_A = V_B), Tnuc [P] (highest-T crossing of S/T = target, transitionFinder.py:773)."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    m8 = batch._as_device(torch, model8, dev).reshape(-1, 8).contiguous()
    P = m8.shape[0]
    Td = batch._as_device(torch, T, dev).reshape(-1).contiguous()
    nT = Td.numel()
    v = torch.sqrt(m8[:, 7])
    A = trace_minima(m8, torch.stack([v, v], dim=-1), Td)
    okA = A["flag"] == _lib.MIN_OK
    ar = torch.arange(nT, device=dev)
    # index of the entry that ended phase A (nT: it never ended on this grid)
    eA = torch.where(okA.all(dim=1), torch.full((P,), nT, device=dev), (~okA).to(torch.int64).argmax(dim=1))
    e_idx = eA.clamp(max=nT - 1)
    jumped = (eA < nT) & (A["flag"].gather(1, e_idx[:, None])[:, 0] == _lib.MIN_JUMPED)
    seedB = A["X"].gather(1, e_idx[:, None, None].expand(P, 1, 2))[:, 0, :]
    seedB = torch.where(jumped[:, None], seedB, torch.stack([v, -v], dim=-1))
    tbB = torch.where(jumped, e_idx, torch.zeros_like(e_idx))
    dirB = torch.where(jumped, -torch.ones_like(e_idx), torch.ones_like(e_idx))
    B = trace_minima(m8, seedB, Td, t_begin=tbB, t_dir=dirB)
    okB = B["flag"] == _lib.MIN_OK
    both = okA & okB
    dX = B["X"] - A["X"]
    L = torch.linalg.norm(dX, dim=-1)
    distinct = both & (L > 1e-3 * (torch.linalg.norm(A["X"], dim=-1) + 1.0))
    want = distinct & (B["V"] > A["V"])
    if bounce_mask is not None:
        want = want & batch._as_device(torch, bounce_mask, dev).to(torch.bool)
    S = torch.full((P, nT), float("nan"), dtype=torch.float64, device=dev)
    status = torch.full((P, nT), -1, dtype=torch.int32, device=dev)
    idx = want.nonzero()
    if idx.shape[0]:
        p_i, t_i = idx[:, 0], idx[:, 1]
        xa, Li = A["X"][p_i, t_i], L[p_i, t_i]
        rows = torch.cat([m8[p_i], Td[t_i][:, None], xa, dX[p_i, t_i] / Li[:, None]], dim=1).contiguous()
        r = batch.find_bounce_batch(_lib.POT_MODEL1_LINE, rows, torch.zeros_like(Li), Li, alpha=alpha, **knobs)
        S[p_i, t_i] = r["S"]
        status[p_i, t_i] = r["status"]
    soT = torch.where((status == 0) | (status == 1), S / Td[None, :], torch.full_like(S, float("nan")))
    # critical temperature: V_A - V_B changes sign (linear interpolation on the grid)
    d = torch.where(distinct, A["V"] - B["V"], torch.full_like(S, float("nan")))
    sc = (d[:, :-1] < 0) & (d[:, 1:] >= 0)
    has_c = sc.any(dim=1)
    ic = (sc.to(torch.int64) * ar[1:]).amax(dim=1).clamp(min=1)
    d0, d1 = d.gather(1, (ic - 1)[:, None])[:, 0], d.gather(1, ic[:, None])[:, 0]
    Tcrit = torch.where(has_c, Td[ic - 1] + (0.0 - d0) * (Td[ic] - Td[ic - 1]) / (d1 - d0), torch.full_like(d0, float("nan")))
    # nucleation: S/T grows with T below Tcrit; the highest-T crossing of the target from below
    cr = (soT[:, :-1] <= target) & (soT[:, 1:] > target)
    has_n = cr.any(dim=1)
    i_n = (cr.to(torch.int64) * ar[1:]).amax(dim=1).clamp(min=1)
    y0, y1 = soT.gather(1, (i_n - 1)[:, None])[:, 0], soT.gather(1, i_n[:, None])[:, 0]
    Tnuc = torch.where(has_n, Td[i_n - 1] + (target - y0) * (Td[i_n] - Td[i_n - 1]) / (y1 - y0),
                       torch.full_like(y0, float("nan")))
    return dict(T=Td, XA=A["X"], XB=B["X"], VA=A["V"], VB=B["V"], flagA=A["flag"], flagB=B["flag"], S=S, status=status,
                S_over_T=soT, Tcrit=Tcrit, Tnuc=Tnuc, n_bounces=int(idx.shape[0]), endA=eA,
                newton_iters=int(A["iters"].sum() + B["iters"].sum()))


def line_bounces(model, params, T, X_abs, X_meta, alpha=2, method="exact", pieces=255, pad=0.06, chunk=32768, **knobs):
    """One bounce per instance along the straight line from X_meta (metastable) to X_abs for a jit.CompiledModel.

    method = "exact" (default): the bounce kernels instantiated for the model (CompiledModel.line_functor(): built on
    first use, ~25 s of nvcc, cached) -- Vtot along the line with the reference's finite-difference dV / d2V, trajectory-
    faithful like the built-in CTB_POT_MODEL1_LINE.
    method = "sampled": no second build -- the line potential V(s) = Vtot(X_abs + s nhat, T), s in [0, L], is
    SAMPLED by the model's own kernel on `pieces` + 5 equidistant points covering [-pad L, (1 + pad) L], turned into the
    piecewise cubic Hermite interpolant on the device (derivatives at the knots by the 5-point stencil over the samples)
    and solved by the tabulated functor (CTB_POT_SPLINE).  Spline-accurate, like potentials.SampledPotential: against
    the exact functor on model1, S differs by 6e-6 (median), 1e-3 (99th percentile); T_nuc by 2e-4 K (median).
    params [n, nparam] (or one row), T [n], X_abs / X_meta [n, ndim].  Returns dict(S, status, L) device tensors [n]."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    if method not in ("exact", "sampled"):
        raise ValueError("method must be 'exact' or 'sampled'")
    if not (8 <= pieces <= _lib.SPLINE_MAX_PIECES):
        raise ValueError("pieces must be in [8, %d]" % _lib.SPLINE_MAX_PIECES)
    Xa = batch._as_device(torch, X_abs, dev).reshape(-1, model.ndim)
    Xm = batch._as_device(torch, X_meta, dev).reshape(-1, model.ndim)
    n = Xa.shape[0]
    Td = batch._as_device(torch, T, dev, (n,))
    P = batch._as_device(torch, params, dev).reshape(-1, max(1, model.nparam))
    if P.shape[0] == 1:
        P = P.expand(n, -1)
    d = Xm - Xa
    L = torch.linalg.norm(d, dim=-1)
    nhat = d / L[:, None]
    if method == "exact":
        f = model.line_functor()
        rows = f.rows(torch, P[:, :model.nparam] if model.nparam else P[:, :0], Td, Xa, nhat)
        r = batch.find_bounce_batch(f, rows, torch.zeros(n, dtype=torch.float64, device=dev), L, alpha=alpha, **knobs)
        return dict(S=r["S"], status=r["status"], L=L)
    K = pieces
    # knots k = 0..K at s_k = -pad L + k h; two more samples on either side for the derivative stencil
    h = (1.0 + 2.0 * pad) * L / K
    j = torch.arange(-2, K + 3, dtype=torch.float64, device=dev)
    S_out = torch.empty(n, dtype=torch.float64, device=dev)
    st_out = torch.empty(n, dtype=torch.int32, device=dev)
    off = 1 + _lib.SPLINE_MAX_PIECES + 1
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        m = hi - lo
        s = -pad * L[lo:hi, None] + j[None, :] * h[lo:hi, None]                       # [m, K + 5]
        Xs = (Xa[lo:hi, None, :] + s[:, :, None] * nhat[lo:hi, None, :]).reshape(-1, model.ndim)
        Ts = Td[lo:hi, None].expand(m, K + 5).reshape(-1)
        Ps = P[lo:hi, None, :].expand(m, K + 5, P.shape[1]).reshape(-1, P.shape[1])
        v = model.vtot(Ps, Xs, Ts).reshape(m, K + 5)
        v = v - v[:, 2:3]                                                              # (offset: keeps the cubic's c0 small)
        dv = (v[:, :-4] - 8.0 * v[:, 1:-3] + 8.0 * v[:, 3:-1] - v[:, 4:]) / (12.0 * h[lo:hi, None])    # at the K + 1 knots
        vk = v[:, 2:-2]
        hh = h[lo:hi, None]
        v0, v1, d0, d1 = vk[:, :-1], vk[:, 1:], dv[:, :-1], dv[:, 1:]
        c2 = (3.0 * (v1 - v0) / hh - 2.0 * d0 - d1) / hh
        c3 = (2.0 * (v0 - v1) / hh + d0 + d1) / (hh * hh)
        rows = torch.zeros((m, _lib.NPARAM[_lib.POT_SPLINE]), dtype=torch.float64, device=dev)
        rows[:, 0] = K
        rows[:, 1:K + 2] = s[:, 2:-2]
        rows[:, off:off + 4 * K] = torch.stack([v0, d0, c2, c3], dim=-1).reshape(m, 4 * K)
        r = batch.find_bounce_batch(_lib.POT_SPLINE, rows, torch.zeros(m, dtype=torch.float64, device=dev), L[lo:hi],
                                    alpha=alpha, **knobs)
        S_out[lo:hi] = r["S"]
        st_out[lo:hi] = r["status"]
    return dict(S=S_out, status=st_out, L=L)


def phase_scan(model, params, T, seed_A, seed_B=None, alpha=2, target=140.0, method="exact", pieces=255, **knobs):
    """model1_phase_scan for any compiled model: phase A from seed_A [P, ndim] at T[0] upwards until it ends, phase B from
    A's landing point downwards (or, where A never jumps, from seed_B at T[0] upwards), then line_bounces wherever both
    exist and B is metastable.  Returns the same dict as model1_phase_scan."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    P8 = batch._as_device(torch, params, dev).reshape(-1, max(1, model.nparam)).contiguous()
    P = P8.shape[0]
    Td = batch._as_device(torch, T, dev).reshape(-1).contiguous()
    nT = Td.numel()
    sA = batch._as_device(torch, seed_A, dev).reshape(-1, model.ndim)
    sA = sA.expand(P, -1) if sA.shape[0] == 1 else sA
    A = trace_minima(P8, sA, Td, model=model)
    okA = A["flag"] == _lib.MIN_OK
    ar = torch.arange(nT, device=dev)
    eA = torch.where(okA.all(dim=1), torch.full((P,), nT, device=dev), (~okA).to(torch.int64).argmax(dim=1))
    e_idx = eA.clamp(max=nT - 1)
    jumped = (eA < nT) & (A["flag"].gather(1, e_idx[:, None])[:, 0] == _lib.MIN_JUMPED)
    seedB = A["X"].gather(1, e_idx[:, None, None].expand(P, 1, model.ndim))[:, 0, :]
    if seed_B is not None:
        sB = batch._as_device(torch, seed_B, dev).reshape(-1, model.ndim)
        seedB = torch.where(jumped[:, None], seedB, sB.expand(P, -1) if sB.shape[0] == 1 else sB)
    tbB = torch.where(jumped, e_idx, torch.zeros_like(e_idx))
    dirB = torch.where(jumped, -torch.ones_like(e_idx), torch.ones_like(e_idx))
    B = trace_minima(P8, seedB, Td, t_begin=tbB, t_dir=dirB, model=model)
    okB = B["flag"] == _lib.MIN_OK
    if seed_B is None:
        okB = okB & jumped[:, None]
    dX = B["X"] - A["X"]
    L = torch.linalg.norm(dX, dim=-1)
    distinct = okA & okB & (L > 1e-3 * (torch.linalg.norm(A["X"], dim=-1) + 1.0))
    want = distinct & (B["V"] > A["V"])
    S = torch.full((P, nT), float("nan"), dtype=torch.float64, device=dev)
    status = torch.full((P, nT), -1, dtype=torch.int32, device=dev)
    idx = want.nonzero()
    if idx.shape[0]:
        p_i, t_i = idx[:, 0], idx[:, 1]
        r = line_bounces(model, P8[p_i], Td[t_i], A["X"][p_i, t_i], B["X"][p_i, t_i], alpha=alpha, method=method,
                         pieces=pieces, **knobs)
        S[p_i, t_i] = r["S"]
        status[p_i, t_i] = r["status"]
    soT = torch.where((status == 0) | (status == 1), S / Td[None, :], torch.full_like(S, float("nan")))
    d = torch.where(distinct, A["V"] - B["V"], torch.full_like(S, float("nan")))
    sc = (d[:, :-1] < 0) & (d[:, 1:] >= 0)
    has_c = sc.any(dim=1)
    ic = (sc.to(torch.int64) * ar[1:]).amax(dim=1).clamp(min=1)
    d0, d1 = d.gather(1, (ic - 1)[:, None])[:, 0], d.gather(1, ic[:, None])[:, 0]
    Tcrit = torch.where(has_c, Td[ic - 1] + (0.0 - d0) * (Td[ic] - Td[ic - 1]) / (d1 - d0), torch.full_like(d0, float("nan")))
    cr = (soT[:, :-1] <= target) & (soT[:, 1:] > target)
    has_n = cr.any(dim=1)
    i_n = (cr.to(torch.int64) * ar[1:]).amax(dim=1).clamp(min=1)
    y0, y1 = soT.gather(1, (i_n - 1)[:, None])[:, 0], soT.gather(1, i_n[:, None])[:, 0]
    Tnuc = torch.where(has_n, Td[i_n - 1] + (target - y0) * (Td[i_n] - Td[i_n - 1]) / (y1 - y0),
                       torch.full_like(y0, float("nan")))
    return dict(T=Td, XA=A["X"], XB=B["X"], VA=A["V"], VB=B["V"], flagA=A["flag"], flagB=B["flag"], S=S, status=status,
                S_over_T=soT, Tcrit=Tcrit, Tnuc=Tnuc, n_bounces=int(idx.shape[0]), endA=eA)
