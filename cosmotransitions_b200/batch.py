"""Batched entry point: SingleFieldInstanton(...).findProfile() + .findAction() for N independent
instances in one kernel launch (tunneling1D.py:128-791; behaviour spec in SURVEY.md Appendix A)."""
import math

import numpy as np

from . import _lib, _tables

FINDPROFILE_DEFAULTS = dict(xguess=None, xtol=1e-4, phitol=1e-4, thinCutoff=.01, npoints=500, rmin=1e-4, rmax=1e4,
                            max_interior_pts=None)

_F64_OUT = ["S", "phi0", "r0", "rf", "x", "phi_bar", "rscale"]
_I32_OUT = ["status", "n_shots", "n_rkck", "n_points"]


def make_opts(alpha=2, phi_eps=1e-3, block_threads=0, **knobs):
    kw = dict(FINDPROFILE_DEFAULTS)
    for k, v in knobs.items():
        if k not in kw:
            raise TypeError("unexpected findProfile argument %r" % k)
        kw[k] = v
    if alpha not in (2, 3):
        raise NotImplementedError("alpha must be 2 or 3 (nu = 1/2 or 1 Bessel start)")
    o = _lib.Opts()
    o.xtol, o.phitol, o.thinCutoff, o.rmin, o.rmax = kw["xtol"], kw["phitol"], kw["thinCutoff"], kw["rmin"], kw["rmax"]
    o.phi_eps = phi_eps
    o.xguess = math.nan if kw["xguess"] is None else float(kw["xguess"])
    o.npoints = int(kw["npoints"])
    o.max_interior_pts = -1 if kw["max_interior_pts"] is None else int(kw["max_interior_pts"])
    o.alpha = int(alpha)
    o.block_threads = int(block_threads)
    return o


def profile_capacity(opts):
    mi = opts.npoints // 2 if opts.max_interior_pts < 0 else opts.max_interior_pts
    return opts.npoints + mi


def _as_device(torch, x, device, shape=None):
    if isinstance(x, torch.Tensor):
        t = x.to(device=device, dtype=torch.float64, non_blocking=True)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(device, non_blocking=True)
    if shape is not None:
        t = t.expand(shape) if t.dim() == 0 or tuple(t.shape) != tuple(shape) else t
    return t.contiguous()


def launch_bounce(kind, d_params, d_phi_abs, d_phi_meta, opts, out, profiles=None, stream=None):
    """Lowest-level call: device tensors in, device tensors (pre-allocated dict `out`) filled.
    Asynchronous on `stream` (default: torch's current stream)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    n = d_phi_abs.numel()
    dev = d_phi_abs.device
    tabs = _tables.device_tables(dev) if kind != _lib.POT_POLY4 else None
    bo = _lib.BounceOut()
    for k in _F64_OUT + _I32_OUT:
        setattr(bo, "d_" + k, out[k].data_ptr() if k in out else None)
    if profiles is not None:
        bo.d_prof_R, bo.d_prof_Phi, bo.d_prof_dPhi = (profiles[k].data_ptr() for k in ("R", "Phi", "dPhi"))
        bo.prof_stride = profiles["R"].shape[1]
    st = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
    with torch.cuda.device(dev):
        _lib.check(lib.ctb_bounce_batch(kind, d_params.data_ptr(), d_phi_abs.data_ptr(), d_phi_meta.data_ptr(), n,
                                        opts, tabs.jb if tabs else None, tabs.jf if tabs else None, bo, st),
                   "ctb_bounce_batch")


def alloc_out(torch, n, device):
    out = {k: torch.empty(n, dtype=torch.float64, device=device) for k in _F64_OUT}
    out.update({k: torch.empty(n, dtype=torch.int32, device=device) for k in _I32_OUT})
    return out


def find_bounce_batch(potential_kind, params, phi_absMin, phi_metaMin, alpha=2, phi_eps=1e-3, device=None,
                      return_profiles=False, block_threads=0, **findProfile_knobs):
    """Solve N independent bounces.

    potential_kind : one of _lib.POT_* (or a DevicePotential subclass / instance, whose kind is used)
    params         : [N, K] tensor/array of potential coefficients (K = _lib.NPARAM[kind])
    phi_absMin, phi_metaMin : [N] (or scalars, broadcast)
    alpha, phi_eps : as SingleFieldInstanton.__init__ (tunneling1D.py:128-130)
    findProfile_knobs : xguess, xtol, phitol, thinCutoff, npoints, rmin, rmax, max_interior_pts
                        with the reference's defaults (tunneling1D.py:615-617)

    Returns a dict of tensors on the input's device (CPU inputs -> CPU outputs, copied through the
    current CUDA device): S, phi0, r0, rf, x, phi_bar, rscale (float64), status, n_shots, n_rkck,
    n_points (int32) and, if return_profiles, R / Phi / dPhi [N, npoints + max_interior].
    status codes: see include/ctbounce.h (0/1 = solved; 2/3 = PotentialError kinds; 4-6 = IntegrationError).
    """
    torch = _lib.require_cuda()
    kind = getattr(potential_kind, "kind", potential_kind)
    if kind not in _lib.NPARAM:
        raise ValueError("unknown potential kind %r" % (potential_kind,))
    opts = make_opts(alpha=alpha, phi_eps=phi_eps, block_threads=block_threads, **findProfile_knobs)
    on_host = not (isinstance(params, torch.Tensor) and params.is_cuda)
    if device is None:
        device = params.device if not on_host else torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    d_params = _as_device(torch, params, device).reshape(-1, _lib.NPARAM[kind])
    n = d_params.shape[0]
    d_abs = _as_device(torch, phi_absMin, device, (n,))
    d_meta = _as_device(torch, phi_metaMin, device, (n,))
    out = alloc_out(torch, n, device)
    prof = None
    if return_profiles:
        cap = profile_capacity(opts)
        prof = {k: torch.full((n, cap), float("nan"), dtype=torch.float64, device=device) for k in ("R", "Phi", "dPhi")}
    if n:
        launch_bounce(kind, d_params, d_abs, d_meta, opts, out, prof)
    if prof is not None:
        out.update(prof)
    if on_host:
        out = {k: v.cpu() for k, v in out.items()}
    return out
