"""Batched entry point: SingleFieldInstanton(...).findProfile() + .findAction() for N independent
instances in one kernel launch (tunneling1D.py:128-791; behaviour spec in SURVEY.md Appendix A)."""
import math
import os

import numpy as np

from . import _lib, _tables

FINDPROFILE_DEFAULTS = dict(xguess=None, xtol=1e-4, phitol=1e-4, thinCutoff=.01, npoints=500, rmin=1e-4, rmax=1e4,
                            max_interior_pts=None)

_F64_OUT = ["S", "phi0", "r0", "rf", "x", "phi_bar", "rscale"]
_I32_OUT = ["status", "n_shots", "n_rkck", "n_points"]


_KERNELS = {"auto": _lib.KERNEL_AUTO, "thread": _lib.KERNEL_THREAD, "warp": _lib.KERNEL_WARP}


def make_opts(alpha=2, phi_eps=1e-3, block_threads=0, kernel="auto", **knobs):
    kw = dict(FINDPROFILE_DEFAULTS)
    for k, v in knobs.items():
        if k not in kw:
            raise TypeError("unexpected findProfile argument %r" % k)
        kw[k] = v
    # alpha in {1, 2, 3, 4}: the specialised kernels (closed-form / I0-I1 Bessel start); any other real alpha in
    # [0, 8]: the general-order start of tunneling1D.py:336-372 (thread kernel, polynomial / tabulated potentials)
    alpha = float(alpha)
    if not (0.0 <= alpha <= 8.0):
        raise ValueError("alpha must be a real number in [0, 8] (d = alpha + 1 dimensions; the general-order Bessel "
                         "start is validated for nu = (alpha - 1) / 2 <= 7/2)")
    o = _lib.Opts()
    o.xtol, o.phitol, o.thinCutoff, o.rmin, o.rmax = kw["xtol"], kw["phitol"], kw["thinCutoff"], kw["rmin"], kw["rmax"]
    o.phi_eps = phi_eps
    o.xguess = math.nan if kw["xguess"] is None else float(kw["xguess"])
    o.npoints = int(kw["npoints"])
    o.max_interior_pts = -1 if kw["max_interior_pts"] is None else int(kw["max_interior_pts"])
    o.alpha = int(alpha) if alpha in (1.0, 2.0, 3.0, 4.0) else 0
    o.alpha_real = alpha
    o.block_threads = int(block_threads)
    o.kernel = _KERNELS[kernel]
    return o


def exact_solution_batch(alpha, r, phi0, dV, d2V):
    """SingleFieldInstanton.exactSolution (tunneling1D.py:294-373) for n argument sets on the device (parity aid for
    the Bessel start conditions): arrays r, phi0, dV, d2V [n] -> (phi [n], dphi [n]) numpy arrays."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    rows = np.stack([np.asarray(a, dtype=np.float64).reshape(-1) for a in (r, phi0, dV, d2V)], axis=1)
    d_in = torch.from_numpy(np.ascontiguousarray(rows)).to(dev)
    d_out = torch.empty((rows.shape[0], 2), dtype=torch.float64, device=dev)
    _lib.check(lib.ctb_exact_solution(float(alpha), d_in.data_ptr(), d_out.data_ptr(), rows.shape[0],
                                      torch.cuda.current_stream(dev).cuda_stream), "ctb_exact_solution")
    o = d_out.cpu().numpy()
    return o[:, 0], o[:, 1]


def profile_capacity(opts):
    mi = opts.npoints // 2 if opts.max_interior_pts < 0 else opts.max_interior_pts
    return opts.npoints + mi


def _as_device(torch, x, device, shape=None):
    if isinstance(x, torch.Tensor):
        t = x.to(device=device, dtype=torch.float64, non_blocking=True)
    else:
        a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        if not a.flags.writeable:     # (e.g. arrays out of an .npz: torch wants a writable buffer behind from_numpy)
            a = a.copy()
        t = torch.from_numpy(a).to(device, non_blocking=True)
    if shape is not None:
        t = t.expand(shape) if t.dim() == 0 or tuple(t.shape) != tuple(shape) else t
    return t.contiguous()


def _bounce_in(d_params, d_phi_abs, d_phi_meta, phi_bar=None, rscale=None, order=None, scratch=None):
    bi = _lib.BounceIn()
    bi.d_scratch = scratch.data_ptr() if scratch is not None else None
    bi.d_params, bi.d_phi_abs, bi.d_phi_meta = d_params.data_ptr(), d_phi_abs.data_ptr(), d_phi_meta.data_ptr()
    bi.d_phi_bar = phi_bar.data_ptr() if phi_bar is not None else None
    bi.d_rscale = rscale.data_ptr() if rscale is not None else None
    bi.d_order = order.data_ptr() if order is not None else None
    bi.n = d_phi_abs.numel()
    return bi


def _bounce_out(out, profiles=None):
    bo = _lib.BounceOut()
    for k in _F64_OUT + _I32_OUT:
        setattr(bo, "d_" + k, out[k].data_ptr() if k in out else None)
    if profiles is not None:
        bo.d_prof_R, bo.d_prof_Phi, bo.d_prof_dPhi = (profiles[k].data_ptr() for k in ("R", "Phi", "dPhi"))
        bo.prof_stride = profiles["R"].shape[1]
    return bo


# batches at least this large are solved in work-sorted order (two extra tiny launches + a sort)
SORT_THRESHOLD = 4096


def launch_bounce(kind, d_params, d_phi_abs, d_phi_meta, opts, out, profiles=None, schedule="auto",
                  phi_bar=None, rscale=None, split="auto"):
    """Lowest-level call: device tensors in, device tensors (pre-allocated dict `out`) filled.
    Asynchronous on torch's CURRENT stream of the tensors' device: the temporaries (sort keys, permutation,
    hand-over scratch) are allocated and the sort runs on that stream too, so everything is ordered on one
    stream.  To use another stream, wrap the call in `with torch.cuda.stream(s):` (a foreign cudaStream_t:
    torch.cuda.ExternalStream).

    schedule: "natural" = thread t solves instance t;  "sorted" = a setup kernel computes phi_bar,
    rscale and a work predictor per instance, the instances are sorted heavy-first (so that lanes of a
    warp follow similar trajectories and the longest-running warps start first), and the solve kernel
    walks that permutation.  Results do not depend on the schedule (instances are independent).
    "auto" = sorted for batches of SORT_THRESHOLD or more.
    split: run the thread-per-instance solve as two kernels (shooting, then dense output + action) handing
    over through a 3 n-double scratch buffer; "auto" = for batches of SORT_THRESHOLD or more.  Same results."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    n = d_phi_abs.numel()
    dev = d_phi_abs.device
    custom = getattr(kind, "custom_entry", None)   # a run-time compiled functor (jit.LineFunctor): its own entry points
    if custom is not None:
        jb, jf = custom.tables(_tables.device_tables(dev))
        setup = custom.setup
        solve = custom.solve
    else:
        tabs = _tables.device_tables(dev) if kind in _lib.THERMAL_KINDS else None
        jb, jf = (tabs.jb, tabs.jf) if tabs else (None, None)
        setup = lambda *a: lib.ctb_bounce_setup(kind, *a)
        solve = lambda *a: lib.ctb_bounce_batch(kind, *a)
    bo = _bounce_out(out, profiles)
    st = torch.cuda.current_stream(dev).cuda_stream
    if schedule == "auto":
        schedule = "sorted" if n >= SORT_THRESHOLD else "natural"
    with torch.cuda.device(dev):
        order = None
        if schedule == "sorted":
            key = torch.empty(n, dtype=torch.float32, device=dev)
            bi = _bounce_in(d_params, d_phi_abs, d_phi_meta, phi_bar, rscale)
            _lib.check(setup(bi, jb, jf, bo, key.data_ptr(), st), "ctb_bounce_setup")
            order = torch.argsort(key, descending=True)
            # the solve kernel reuses what the setup kernel found (NaN where __init__ failed: the solve
            # kernel then re-derives the same PotentialError status)
            phi_bar, rscale = out["phi_bar"], out["rscale"]
        elif schedule != "natural":
            raise ValueError("schedule must be 'auto', 'sorted' or 'natural'")
        scratch = None
        if split == "auto":
            split = n >= SORT_THRESHOLD
        if split and profiles is None and opts.kernel != _lib.KERNEL_WARP:
            scratch = torch.empty(3 * n, dtype=torch.float64, device=dev)
        bi = _bounce_in(d_params, d_phi_abs, d_phi_meta, phi_bar, rscale, order, scratch)
        _lib.check(solve(bi, opts, jb, jf, bo, st), "ctb_bounce_batch")


def bounce_setup(kind, params, phi_absMin, phi_metaMin, phi_bar=None, rscale=None):
    """SingleFieldInstanton.__init__ for one instance: returns (status, phi_bar, rscale)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    d_params = _as_device(torch, params, dev).reshape(-1, _lib.NPARAM[kind])
    n = d_params.shape[0]
    d_abs, d_meta = _as_device(torch, phi_absMin, dev, (n,)), _as_device(torch, phi_metaMin, dev, (n,))
    d_bar = _as_device(torch, [phi_bar], dev, (n,)) if phi_bar is not None else None
    d_rs = _as_device(torch, [rscale], dev, (n,)) if rscale is not None else None
    out = alloc_out(torch, n, dev)
    tabs = _tables.device_tables(dev) if kind in _lib.THERMAL_KINDS else None
    _lib.check(lib.ctb_bounce_setup(kind, _bounce_in(d_params, d_abs, d_meta, d_bar, d_rs),
                                    tabs.jb if tabs else None, tabs.jf if tabs else None, _bounce_out(out), None,
                                    torch.cuda.current_stream(dev).cuda_stream), "ctb_bounce_setup")
    return int(out["status"][0]), float(out["phi_bar"][0]), float(out["rscale"][0])


def find_wall_batch(potential_kind, params, phi_absMin, phi_metaMin, phi_eps=1e-3, device=None, return_profiles=False,
                    Fguess=None, Ftol=1e-4, phitol=1e-4, npoints=500, rmin=1e-4, rmax=1e4, phi0_rel=1e-3):
    """WallWithConstFriction(...).findProfile(...) (tunneling1D.py:829-999) for N independent instances.

    Returns a dict of tensors: F (the friction that lets the wall interpolate between the minima), rf, phi0, phi_bar,
    rscale, status, n_shots, n_rkck, n_points and, if return_profiles, R / Phi / dPhi [N, npoints]."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    kind = getattr(potential_kind, "kind", potential_kind)
    if kind not in _lib.NPARAM:
        raise ValueError("unknown potential kind %r" % (potential_kind,))
    on_host = not (isinstance(params, torch.Tensor) and params.is_cuda)
    if device is None:
        device = params.device if not on_host else torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    d_params = _as_device(torch, params, device).reshape(-1, _lib.NPARAM[kind])
    n = d_params.shape[0]
    d_abs, d_meta = _as_device(torch, phi_absMin, device, (n,)), _as_device(torch, phi_metaMin, device, (n,))
    o = _lib.WallOpts()
    o.Fguess = math.nan if Fguess is None else float(Fguess)
    o.Ftol, o.phitol, o.rmin, o.rmax, o.phi0_rel, o.phi_eps = Ftol, phitol, rmin, rmax, phi0_rel, phi_eps
    o.npoints = int(npoints)
    out = alloc_out(torch, n, device)
    prof = None
    if return_profiles:
        prof = {k: torch.full((n, o.npoints), float("nan"), dtype=torch.float64, device=device)
                for k in ("R", "Phi", "dPhi")}
    tabs = _tables.device_tables(device) if kind in _lib.THERMAL_KINDS else None
    if n:
        with torch.cuda.device(device):
            _lib.check(lib.ctb_wall_batch(kind, _bounce_in(d_params, d_abs, d_meta), o, tabs.jb if tabs else None,
                                          tabs.jf if tabs else None, _bounce_out(out, prof),
                                          torch.cuda.current_stream(device).cuda_stream), "ctb_wall_batch")
    out["F"] = out.pop("x")
    del out["S"]
    if prof is not None:
        out.update(prof)
    if on_host:
        out = {k: v.cpu() for k, v in out.items()}
    return out


def find_action_batch(potential_kind, params, phi_metaMin, R, Phi, dPhi, n_points=None, alpha=2, device=None):
    """SingleFieldInstanton.findAction(profile) (tunneling1D.py:763-791) for N given profiles.

    params [N, K], phi_metaMin [N]; R, Phi, dPhi [N, P] (row i: Profile1D.R / .Phi / .dPhi of instance i, the first
    n_points[i] entries are used; default: all P).  alpha may be any real >= 0.  Returns S [N] (on the device of R if R
    is a CUDA tensor, else a CPU tensor)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    kind = getattr(potential_kind, "kind", potential_kind)
    if kind not in _lib.NPARAM:
        raise ValueError("unknown potential kind %r" % (potential_kind,))
    on_host = not (isinstance(R, torch.Tensor) and R.is_cuda)
    if device is None:
        device = R.device if not on_host else torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    d_R = _as_device(torch, R, device)
    d_R = d_R.reshape(1, -1) if d_R.dim() == 1 else d_R
    n, stride = d_R.shape
    d_Phi, d_dPhi = _as_device(torch, Phi, device, (n, stride)), _as_device(torch, dPhi, device, (n, stride))
    d_params = _as_device(torch, params, device).reshape(-1, _lib.NPARAM[kind])
    if d_params.shape[0] != n:
        d_params = d_params.expand(n, -1).contiguous()
    d_meta = _as_device(torch, phi_metaMin, device, (n,))
    if n_points is None:
        d_np = torch.full((n,), stride, dtype=torch.int32, device=device)
    else:
        d_np = torch.as_tensor(n_points, dtype=torch.int32).to(device).expand(n).contiguous()
    S = torch.empty(n, dtype=torch.float64, device=device)
    tabs = _tables.device_tables(device) if kind in _lib.THERMAL_KINDS else None
    if n:
        with torch.cuda.device(device):
            _lib.check(lib.ctb_find_action(kind, d_params.data_ptr(), tabs.jb if tabs else None, tabs.jf if tabs else None,
                                           d_meta.data_ptr(), d_R.data_ptr(), d_Phi.data_ptr(), d_dPhi.data_ptr(), stride,
                                           d_np.data_ptr(), n, float(alpha), S.data_ptr(),
                                           torch.cuda.current_stream(device).cuda_stream), "ctb_find_action")
    return S.cpu() if on_host else S


def alloc_out(torch, n, device):
    out = {k: torch.empty(n, dtype=torch.float64, device=device) for k in _F64_OUT}
    out.update({k: torch.empty(n, dtype=torch.int32, device=device) for k in _I32_OUT})
    return out


def find_bounce_batch(potential_kind, params, phi_absMin, phi_metaMin, alpha=2, phi_eps=1e-3, device=None,
                      return_profiles=False, block_threads=0, schedule="auto", phi_bar=None, rscale=None,
                      kernel="auto", split="auto", **findProfile_knobs):
    """Solve N independent bounces.

    potential_kind : one of _lib.POT_* (or a DevicePotential subclass / instance, whose kind is used), or the line
                     functor of a run-time compiled model (jit.CompiledModel.line_functor())
    params         : [N, K] tensor/array of potential coefficients (K = _lib.NPARAM[kind])
    phi_absMin, phi_metaMin : [N] (or scalars, broadcast)
    alpha, phi_eps : as SingleFieldInstanton.__init__ (tunneling1D.py:128-130)
    phi_bar, rscale : optional [N] overrides, as in __init__ (NaN entries = compute)
    schedule       : "auto" | "sorted" | "natural" (see launch_bounce); does not change results
    kernel         : "auto" | "thread" | "warp" -- one thread or one warp per instance (include/ctbounce.h);
                     same shooting sequence, S equal up to the summation order of the quadrature
    findProfile_knobs : xguess, xtol, phitol, thinCutoff, npoints, rmin, rmax, max_interior_pts
                        with the reference's defaults (tunneling1D.py:615-617)

    Returns a dict of tensors on the input's device (CPU inputs -> CPU outputs, copied through the
    current CUDA device): S, phi0, r0, rf, x, phi_bar, rscale (float64), status, n_shots, n_rkck,
    n_points (int32) and, if return_profiles, R / Phi / dPhi [N, npoints + max_interior].
    status codes: see include/ctbounce.h (0/1 = solved; 2/3 = PotentialError kinds; 4-6 = IntegrationError).
    """
    torch = _lib.require_cuda()
    custom = getattr(potential_kind, "custom_entry", None)
    kind = custom if custom is not None else getattr(potential_kind, "kind", potential_kind)
    if custom is None and kind not in _lib.NPARAM:
        raise ValueError("unknown potential kind %r" % (potential_kind,))
    nparam = custom.nparam if custom is not None else _lib.NPARAM[kind]
    opts = make_opts(alpha=alpha, phi_eps=phi_eps, block_threads=block_threads, kernel=kernel, **findProfile_knobs)
    on_host = not (isinstance(params, torch.Tensor) and params.is_cuda)
    if device is None:
        device = params.device if not on_host else torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    d_params = _as_device(torch, params, device).reshape(-1, nparam)
    n = d_params.shape[0]
    d_abs = _as_device(torch, phi_absMin, device, (n,))
    d_meta = _as_device(torch, phi_metaMin, device, (n,))
    out = alloc_out(torch, n, device)
    prof = None
    if return_profiles:
        cap = profile_capacity(opts)
        prof = {k: torch.full((n, cap), float("nan"), dtype=torch.float64, device=device) for k in ("R", "Phi", "dPhi")}
    d_bar = _as_device(torch, phi_bar, device, (n,)) if phi_bar is not None else None
    d_rs = _as_device(torch, rscale, device, (n,)) if rscale is not None else None
    if n:
        launch_bounce(kind, d_params, d_abs, d_meta, opts, out, prof, schedule=schedule, phi_bar=d_bar, rscale=d_rs,
                      split=split)
    if prof is not None:
        out.update(prof)
    if on_host:
        out = {k: v.cpu() for k, v in out.items()}
    return out


class BouncePipeline:
    """Streaming front end for many same-sized batches coming from HOST memory (parameter scans that are generated
    or read on the CPU): `depth` slots, each with its own CUDA stream, device buffers and pinned result buffers, so
    that the host->device copy of batch k+1 and the device->host copy of batch k-1 run on the copy engines while the
    kernels of batch k occupy the SMs.  Results are identical to find_bounce_batch (instances are independent).

        pipe = BouncePipeline(_lib.POT_POLY4, n, alpha=2)
        for k, (params, phi_abs, phi_meta) in enumerate(batches):     # pinned CPU tensors for truly async copies
            if k >= pipe.depth:
                consume(pipe.collect(k - pipe.depth))                 # dict of pinned CPU tensors (valid until reuse)
            pipe.submit(params, phi_abs, phi_meta)
        ... collect the last `depth` tickets
    """

    def __init__(self, potential_kind, n, depth=3, device=None, outputs=("S", "status"), alpha=2, phi_eps=1e-3,
                 schedule="auto", split="auto", block_threads=0, kernel="auto", **findProfile_knobs):
        torch = _lib.require_cuda()
        self._torch = torch
        self.kind = getattr(potential_kind, "kind", potential_kind)
        self.n, self.depth = int(n), int(depth)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.opts = make_opts(alpha=alpha, phi_eps=phi_eps, block_threads=block_threads, kernel=kernel, **findProfile_knobs)
        self.schedule, self.split, self.outputs = schedule, split, tuple(outputs)
        K = _lib.NPARAM[self.kind]
        self._slots = []
        for _ in range(self.depth):
            slot = dict(stream=torch.cuda.Stream(self.device), done=torch.cuda.Event(),
                        params=torch.empty((self.n, K), dtype=torch.float64, device=self.device),
                        phi_abs=torch.empty(self.n, dtype=torch.float64, device=self.device),
                        phi_meta=torch.empty(self.n, dtype=torch.float64, device=self.device),
                        out=alloc_out(torch, self.n, self.device))
            slot["host"] = {k: torch.empty(self.n, dtype=slot["out"][k].dtype).pin_memory() for k in self.outputs}
            self._slots.append(slot)
        self._submitted = 0

    def submit(self, params, phi_absMin, phi_metaMin):
        """Queue one batch (CPU tensors / arrays of n rows); returns its ticket.  Does not block on the GPU, except
        that a slot is reused only after its previous results have been copied out.  Pinned input tensors are
        copied asynchronously: leave them unchanged until collect(ticket) has returned."""
        torch = self._torch
        k = self._submitted
        slot = self._slots[k % self.depth]
        slot["done"].synchronize()          # previous occupant of the slot (no-op for a fresh event)
        as_t = lambda x: x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        with torch.cuda.device(self.device), torch.cuda.stream(slot["stream"]):
            slot["params"].copy_(as_t(params).reshape(self.n, -1), non_blocking=True)
            slot["phi_abs"].copy_(as_t(phi_absMin).expand(self.n) if as_t(phi_absMin).dim() == 0 else as_t(phi_absMin),
                                  non_blocking=True)
            slot["phi_meta"].copy_(as_t(phi_metaMin).expand(self.n) if as_t(phi_metaMin).dim() == 0
                                   else as_t(phi_metaMin), non_blocking=True)
            launch_bounce(self.kind, slot["params"], slot["phi_abs"], slot["phi_meta"], self.opts, slot["out"],
                          schedule=self.schedule, split=self.split)
            for name in self.outputs:
                slot["host"][name].copy_(slot["out"][name], non_blocking=True)
            slot["done"].record(slot["stream"])
        self._submitted += 1
        return k

    def collect(self, ticket):
        """Wait for batch `ticket` and return its results (pinned CPU tensors owned by the pipeline: copy them if
        they must outlive the next `depth` submissions)."""
        if not (self._submitted - self.depth <= ticket < self._submitted):
            raise ValueError("ticket %d is not in flight" % ticket)
        slot = self._slots[ticket % self.depth]
        slot["done"].synchronize()
        return slot["host"]
