"""Batched helpers around the bounce solver for finite-temperature parameter scans (BASELINE.json
configs[4]): per-instance potential evaluation, bounded 1-D
minimisation (phi_absMin(T)), the temperature window (T0, Tc) of the one-field thermal model, and the
S3(T)/T sweep with the nucleation criterion S3/T = 140 (transitionFinder.py:773).

Everything numerical runs on the device in batches; torch is the container and does the bookkeeping
(bisection brackets, interpolation of the crossing)."""
import numpy as np

from . import _lib, _tables, batch


def _prep(torch, kind, params, *vecs):
    dev = params.device if isinstance(params, torch.Tensor) and params.is_cuda else torch.device(
        "cuda", torch.cuda.current_device())
    p = batch._as_device(torch, params, dev).reshape(-1, _lib.NPARAM[kind])
    n = p.shape[0]
    return dev, p, n, [batch._as_device(torch, v, dev, (n,)) for v in vecs]


def potential_batch(kind, params, phi):
    """V_i(phi_i) for N parameter rows (device tensor [N])."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev, p, n, (ph,) = _prep(torch, kind, params, phi)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    tabs = _tables.device_tables(dev) if kind in _lib.THERMAL_KINDS else None
    with torch.cuda.device(dev):
        _lib.check(lib.ctb_potential_batch(kind, p.data_ptr(), tabs.jb if tabs else None, tabs.jf if tabs else None,
                                           ph.data_ptr(), out.data_ptr(), n,
                                           torch.cuda.current_stream(dev).cuda_stream), "ctb_potential_batch")
    return out


def minimize_1d(kind, params, lo, hi, xtol_rel=1e-10):
    """argmin of V_i on [lo_i, hi_i] (scipy.optimize.fminbound's algorithm, one thread per row).
    Returns (x, V(x)) as device tensors."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev, p, n, (a, b) = _prep(torch, kind, params, lo, hi)
    x = torch.empty(n, dtype=torch.float64, device=dev)
    v = torch.empty(n, dtype=torch.float64, device=dev)
    tabs = _tables.device_tables(dev) if kind in _lib.THERMAL_KINDS else None
    with torch.cuda.device(dev):
        _lib.check(lib.ctb_minimize_1d(kind, p.data_ptr(), tabs.jb if tabs else None, tabs.jf if tabs else None,
                                       a.data_ptr(), b.data_ptr(), float(xtol_rel), n, x.data_ptr(), v.data_ptr(),
                                       torch.cuda.current_stream(dev).cuda_stream), "ctb_minimize_1d")
    return x, v


# ---- one-field thermal model (CTB_POT_MODEL1F): params = [lam, v2, Y, n, renormScaleSq, T] -------------

def model1f_params(lam, Y, n, T, v2=246. ** 2):
    torch = _lib.require_cuda()
    lam, Y, n, T = (torch.as_tensor(x, dtype=torch.float64).cuda() for x in (lam, Y, n, T))
    lam, Y, n, T = torch.broadcast_tensors(lam, Y, n, T)
    v2t = torch.full_like(lam, v2)
    return torch.stack([lam, v2t, Y, n, v2t, T], dim=-1).contiguous()


def model1f_temperature_window(lam, Y, n, v2=246. ** 2, T_lo=20.0, T_hi=300.0, phi_box=(50.0, 500.0), iters=48):
    """(T0, Tc) per parameter point, by batched bisection in T:
    T0 -- phi = 0 turns into a local minimum (V(h) - V(0) changes sign at h = 0.5);
    Tc -- the broken minimum becomes degenerate with phi = 0.  Device tensors [P]."""
    torch = _lib.require_cuda()
    lam, Y, n = (torch.as_tensor(x, dtype=torch.float64).cuda().reshape(-1) for x in (lam, Y, n))
    P = lam.numel()
    kind = _lib.POT_MODEL1F
    zero = torch.zeros(P, dtype=torch.float64, device=lam.device)
    half = torch.full_like(zero, 0.5)

    def curv0(T):
        p = model1f_params(lam, Y, n, T, v2)
        return potential_batch(kind, p, half) - potential_batch(kind, p, zero)

    def gap(T):
        p = model1f_params(lam, Y, n, T, v2)
        _, vmin = minimize_1d(kind, p, torch.full_like(zero, phi_box[0]), torch.full_like(zero, phi_box[1]))
        return vmin - potential_batch(kind, p, zero)

    def bisect(f, a, b):
        fa = f(a)
        for _ in range(iters):
            m = 0.5 * (a + b)
            fm = f(m)
            left = (fm > 0) == (fa > 0)
            a = torch.where(left, m, a)
            fa = torch.where(left, fm, fa)
            b = torch.where(left, b, m)
        return 0.5 * (a + b)

    T0 = bisect(curv0, torch.full_like(zero, T_lo), torch.full_like(zero, T_hi))
    Tc = bisect(gap, T0 + 0.5, torch.full_like(zero, T_hi))
    return T0, Tc


def model1f_bounce_scan(lam, Y, n, nT=256, v2=246. ** 2, window=None, phi_box=(50.0, 500.0), alpha=2, **knobs):
    """P parameter points x nT temperatures uniform in (T0, Tc): phi_absMin(T) by minimize_1d, then one
    bounce per (point, T).  Returns a dict of device tensors shaped [P, nT]: T, phi_abs, S, status, plus
    T0, Tc [P] and Tnuc [P] (first T, from above, where S/T crosses 140; NaN if it never does)."""
    torch = _lib.require_cuda()
    lam, Y, n = (torch.as_tensor(x, dtype=torch.float64).cuda().reshape(-1) for x in (lam, Y, n))
    P = lam.numel()
    T0, Tc = window if window is not None else model1f_temperature_window(lam, Y, n, v2, phi_box=phi_box)
    frac = (torch.arange(nT, dtype=torch.float64, device=lam.device) + 0.5) / nT
    T = T0[:, None] + frac[None, :] * (Tc - T0)[:, None]
    params = model1f_params(lam[:, None], Y[:, None], n[:, None], T, v2).reshape(-1, 6)
    lo = torch.full((P * nT,), phi_box[0], dtype=torch.float64, device=lam.device)
    hi = torch.full((P * nT,), phi_box[1], dtype=torch.float64, device=lam.device)
    phi_abs, _ = minimize_1d(_lib.POT_MODEL1F, params, lo, hi)
    res = batch.find_bounce_batch(_lib.POT_MODEL1F, params, phi_abs, torch.zeros_like(phi_abs), alpha=alpha, **knobs)
    S = res["S"].reshape(P, nT)
    st = res["status"].reshape(P, nT)
    soT = torch.where(st <= 1, S / T, torch.full_like(S, float("nan")))
    # nucleation: S/T decreases as T falls below Tc; find the highest-T crossing of 140
    above = soT[:, 1:] > 140.0
    below = soT[:, :-1] <= 140.0
    cross = above & below
    idx = torch.where(cross.any(dim=1), (cross.to(torch.int64) * torch.arange(1, nT, device=S.device)).amax(dim=1),
                      torch.zeros(P, dtype=torch.int64, device=S.device))
    i0 = (idx - 1).clamp(min=0)
    y0, y1 = soT.gather(1, i0[:, None])[:, 0], soT.gather(1, idx[:, None])[:, 0]
    t0, t1 = T.gather(1, i0[:, None])[:, 0], T.gather(1, idx[:, None])[:, 0]
    Tnuc = torch.where(cross.any(dim=1), t0 + (140.0 - y0) * (t1 - t0) / (y1 - y0), torch.full_like(t0, float("nan")))
    return dict(T=T, phi_abs=phi_abs.reshape(P, nT), S=S, status=st, S_over_T=soT, T0=T0, Tc=Tc, Tnuc=Tnuc,
                n_rkck=res["n_rkck"].reshape(P, nT))


def model1f_tnuc(lam, Y, n, v2=246. ** 2, target=140.0, nT_coarse=24, rounds=10, T_tol=1e-7, phi_box=(50.0, 500.0),
                 window=None, beta_step=0.0, **knobs):
    """Nucleation temperature per parameter point: the root of S3(T)/T - 140 (the reference's default
    nuclCriterion, transitionFinder.py:773, found there by a sequential brentq over T, :836-837).
    Batched form: a coarse sweep brackets the crossing, then every round solves ONE bounce per point at
    an Illinois (regula-falsi) abscissa, all points in one launch.  Returns device tensors
    dict(Tnuc, S, T0, Tc, bracket_lo, bracket_hi, rounds) and, with beta_step > 0 (e.g. 2e-3),
    beta_over_H = T d(S3/T)/dT at T_nuc by a central difference over T_nuc (1 +- beta_step)."""
    torch = _lib.require_cuda()
    lam, Y, n = (torch.as_tensor(x, dtype=torch.float64).cuda().reshape(-1) for x in (lam, Y, n))
    P = lam.numel()
    sweep = model1f_bounce_scan(lam, Y, n, nT=nT_coarse, v2=v2, window=window, phi_box=phi_box, **knobs)
    T, soT = sweep["T"], sweep["S_over_T"]
    f = soT - target
    cross = (f[:, :-1] <= 0) & (f[:, 1:] > 0)
    has = cross.any(dim=1)
    idx = (cross.to(torch.int64) * torch.arange(1, nT_coarse, device=T.device)).amax(dim=1)
    i0 = (idx - 1).clamp(min=0)
    a, b = T.gather(1, i0[:, None])[:, 0], T.gather(1, idx[:, None])[:, 0]
    fa, fb = f.gather(1, i0[:, None])[:, 0], f.gather(1, idx[:, None])[:, 0]
    lo = torch.full((P,), phi_box[0], dtype=torch.float64, device=T.device)
    hi = torch.full((P,), phi_box[1], dtype=torch.float64, device=T.device)
    side = torch.zeros(P, dtype=torch.int64, device=T.device)
    S_last = torch.full((P,), float("nan"), dtype=torch.float64, device=T.device)
    done_rounds = 0
    for _ in range(rounds):
        if not bool(((b - a) > T_tol).any()):
            break
        done_rounds += 1
        Tm = (a * fb - b * fa) / (fb - fa)
        bad = ~torch.isfinite(Tm) | (Tm <= a) | (Tm >= b)
        Tm = torch.where(bad, 0.5 * (a + b), Tm)
        params = model1f_params(lam, Y, n, Tm, v2)
        phi_abs, _ = minimize_1d(_lib.POT_MODEL1F, params, lo, hi)
        r = batch.find_bounce_batch(_lib.POT_MODEL1F, params, phi_abs, torch.zeros_like(phi_abs), **knobs)
        fm = torch.where(r["status"] <= 1, r["S"] / Tm - target, torch.full_like(Tm, float("inf")))
        S_last = r["S"]
        left = fm <= 0          # root is to the right of Tm
        # Illinois: halve the retained end's function value when the same side is kept twice in a row
        fb = torch.where(left & (side == 1), 0.5 * fb, fb)
        fa = torch.where(~left & (side == 2), 0.5 * fa, fa)
        side = torch.where(left, torch.ones_like(side), torch.full_like(side, 2))
        a, fa = torch.where(left, Tm, a), torch.where(left, fm, fa)
        b, fb = torch.where(left, b, Tm), torch.where(left, fb, fm)
    Tn = torch.where(has, torch.where(fa.abs() < fb.abs(), a, b), torch.full_like(a, float("nan")))
    out = dict(Tnuc=Tn, S=S_last, T0=sweep["T0"], Tc=sweep["Tc"], bracket_lo=a, bracket_hi=b, rounds=done_rounds)
    if beta_step:
        # beta / H = T d(S3/T)/dT at T_nuc -- the quantity generic_potential.py:608-609 lists as "NOT IMPLEMENTED YET":
        # central difference over T_nuc (1 +- beta_step), two more bounces per point in one launch
        Tq = torch.where(has, Tn, 0.5 * (sweep["T0"] + sweep["Tc"]))
        Tpm = torch.cat([Tq * (1.0 - beta_step), Tq * (1.0 + beta_step)])
        rep = lambda v: torch.cat([v, v])
        params = model1f_params(rep(lam), rep(Y), rep(n), Tpm, v2)
        phi_abs, _ = minimize_1d(_lib.POT_MODEL1F, params, rep(lo), rep(hi))
        # S(T) at the default tolerances carries ~1e-4 of bisection noise, which a derivative amplifies: these two
        # bounces are solved two orders of magnitude tighter
        tight = dict(knobs)
        tight["xtol"], tight["phitol"] = min(knobs.get("xtol", 1e-4), 1e-6), min(knobs.get("phitol", 1e-4), 1e-6)
        r = batch.find_bounce_batch(_lib.POT_MODEL1F, params, phi_abs, torch.zeros_like(phi_abs), **tight)
        soT = torch.where(r["status"] <= 1, r["S"] / Tpm, torch.full_like(Tpm, float("nan")))
        out["beta_over_H"] = torch.where(has, (soT[P:] - soT[:P]) / (2.0 * beta_step), torch.full_like(Tn, float("nan")))
    return out


# ---- model1 along a fixed straight line (BASELINE.json configs[3]) -------------------

def model1_line_bounce_scan(model8, x_low, x_high, T, alpha=2, **knobs):
    """Per-temperature 1-D bounce along the fixed line X(s) = x_low + s nhat, s in [-0.2 L, 1.2 L]:
    for every T the two minima along the line are located on the device (minimize_1d on the two halves
    of the line), then one bounce per T is solved from the deeper to the shallower one.
    Returns device tensors [len(T)]: T, s_abs, s_meta, S, S_over_T, status and Tnuc (scalar tensor: the
    S/T = 140 crossing, transitionFinder.py:773; NaN if there is none)."""
    torch = _lib.require_cuda()
    kind = _lib.POT_MODEL1_LINE
    x_low, x_high = np.asarray(x_low, float), np.asarray(x_high, float)
    L = float(np.linalg.norm(x_high - x_low))
    nhat = (x_high - x_low) / L
    T = torch.as_tensor(T, dtype=torch.float64).cuda().reshape(-1)
    n = T.numel()
    base = torch.tensor(list(model8) + [0.0] + list(x_low) + list(nhat), dtype=torch.float64, device=T.device)
    params = base[None, :].repeat(n, 1)
    params[:, 8] = T
    full = lambda v: torch.full((n,), v, dtype=torch.float64, device=T.device)
    sa, va = minimize_1d(kind, params, full(-0.2 * L), full(0.45 * L))
    sm, vm = minimize_1d(kind, params, full(0.55 * L), full(1.2 * L))
    # tunnel from the higher minimum (metastable) to the lower one
    swap = va > vm
    s_abs, s_meta = torch.where(swap, sm, sa), torch.where(swap, sa, sm)
    res = batch.find_bounce_batch(kind, params, s_abs, s_meta, alpha=alpha, **knobs)
    S, st = res["S"], res["status"]
    soT = torch.where(st <= 1, S / T, torch.full_like(S, float("nan")))
    Tn = torch.tensor(float("nan"), dtype=torch.float64, device=T.device)
    order = torch.argsort(T)
    y, t = soT[order], T[order]
    good = torch.isfinite(y)
    y, t = y[good], t[good]
    if y.numel() > 1:
        c = ((y[:-1] <= 140.0) & (y[1:] > 140.0)).nonzero()
        if c.numel():
            i = int(c[-1])
            Tn = t[i] + (140.0 - y[i]) * (t[i + 1] - t[i]) / (y[i + 1] - y[i])
    return dict(T=T, s_abs=s_abs, s_meta=s_meta, S=S, S_over_T=soT, status=st, Tnuc=Tn, L=L)
