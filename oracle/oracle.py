"""ctypes front end of the CPU oracle (oracle/ctb_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libctb_oracle.so")

POT_POLY4, POT_MODEL1_LINE, POT_MODEL1F, POT_SPLINE, POT_THERMAL1F = 0, 1, 2, 3, 4
NPARAM = {POT_POLY4: 5, POT_MODEL1_LINE: 13, POT_MODEL1F: 6, POT_SPLINE: None, POT_THERMAL1F: 64}  # spline: 1 + 2 n_knots

STATUS_NAMES = {
    0: "converged", 1: "xtol", 2: "stable, not metastable", 3: "no barrier", 4: "r > rmax",
    5: "dr < drmin", 6: "step underflow", 7: "nonfinite", 8: "brentq sign", 9: "first shot nonfinite",
    10: "bad argument",
}


def build(force=False):
    """Compile oracle/ctb_oracle.c -> oracle/libctb_oracle.so (gcc, see oracle/Makefile)."""
    src = os.path.join(_HERE, "ctb_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libctb_oracle.so"])
    return _LIB_PATH


class _Config(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("nparam", C.c_int), ("alpha", C.c_int), ("npoints", C.c_int),
        ("max_interior_pts", C.c_int),
        ("xtol", C.c_double), ("phitol", C.c_double), ("thinCutoff", C.c_double), ("rmin", C.c_double),
        ("rmax", C.c_double), ("phi_eps", C.c_double), ("xguess", C.c_double),
        ("jb_t", C.c_void_p), ("jb_c", C.c_void_p), ("jb_n", C.c_int), ("jb_xmin", C.c_double),
        ("jb_xmax", C.c_double),
        ("jf_t", C.c_void_p), ("jf_c", C.c_void_p), ("jf_n", C.c_int), ("jf_xmin", C.c_double),
        ("jf_xmax", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.ctbo_brentq_cubic.restype = C.c_double
        _lib.ctbo_brentq_poly.restype = C.c_double
        _lib.ctbo_fminbound_poly.restype = C.c_double
        _lib.ctbo_simpson.restype = C.c_double
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


class Tables:
    """Jb/Jf spline tables (t, c) + clamps, as produced by the reference's finiteT module."""

    def __init__(self, tb, cb, tf, cf, xbmin=-3.72402637, xbmax=1.41e3, xfmin=-6.82200203, xfmax=1.35e3):
        self.tb = np.ascontiguousarray(tb, dtype=np.float64)
        self.cb = np.ascontiguousarray(cb, dtype=np.float64)
        self.tf = np.ascontiguousarray(tf, dtype=np.float64)
        self.cf = np.ascontiguousarray(cf, dtype=np.float64)
        self.xbmin, self.xbmax, self.xfmin, self.xfmax = xbmin, xbmax, xfmin, xfmax

    @classmethod
    def from_npz(cls, path):
        d = np.load(path)
        return cls(d["tb"], d["cb"], d["tf"], d["cf"])


def spline_params(tck):
    """Parameter row of POT_SPLINE for the oracle: [n, t[n], c[n]] (cubic tck as returned by splrep)."""
    t, c, k = tck
    assert k == 3
    return np.concatenate([[len(t)], np.asarray(t, float), np.asarray(c, float)])


def make_config(kind=POT_POLY4, alpha=2, tables=None, xtol=1e-4, phitol=1e-4, thinCutoff=.01, npoints=500,
                rmin=1e-4, rmax=1e4, max_interior_pts=None, phi_eps=1e-3, xguess=None, nparam=None):
    cfg = _Config()
    cfg.kind, cfg.nparam, cfg.alpha, cfg.npoints = kind, nparam or NPARAM[kind], int(alpha), int(npoints)
    cfg.max_interior_pts = -1 if max_interior_pts is None else int(max_interior_pts)
    cfg.xtol, cfg.phitol, cfg.thinCutoff, cfg.rmin, cfg.rmax = xtol, phitol, thinCutoff, rmin, rmax
    cfg.phi_eps = phi_eps
    cfg.xguess = float("nan") if xguess is None else float(xguess)
    cfg._keep = tables
    if tables is not None:
        cfg.jb_t, cfg.jb_c, cfg.jb_n = _dp(tables.tb), _dp(tables.cb), len(tables.tb)
        cfg.jb_xmin, cfg.jb_xmax = tables.xbmin, tables.xbmax
        cfg.jf_t, cfg.jf_c, cfg.jf_n = _dp(tables.tf), _dp(tables.cf), len(tables.tf)
        cfg.jf_xmin, cfg.jf_xmax = tables.xfmin, tables.xfmax
    return cfg


def quartic_params(c):
    """POLY4 coefficients [c0..c4] of V = phi^4/4 - (1+c)/3 phi^3 + c/2 phi^2 (dV = phi(phi-c)(phi-1))."""
    c = np.atleast_1d(np.asarray(c, dtype=np.float64))
    p = np.zeros((len(c), 5))
    p[:, 2] = c / 2
    p[:, 3] = -(1 + c) / 3
    p[:, 4] = 0.25
    return p


def bounce_batch(cfg, params, phi_abs, phi_meta, nthreads=1):
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, cfg.nparam)
    n = params.shape[0]
    phi_abs = np.ascontiguousarray(np.broadcast_to(np.asarray(phi_abs, dtype=np.float64), (n,)))
    phi_meta = np.ascontiguousarray(np.broadcast_to(np.asarray(phi_meta, dtype=np.float64), (n,)))
    dnames = ["S", "phi0", "r0", "rf", "x", "phi_bar", "rscale"]
    inames = ["status", "n_shots", "n_rkck", "n_exact", "n_points", "ctype"]
    out = {k: np.full(n, np.nan) for k in dnames}
    out.update({k: np.zeros(n, dtype=np.int32) for k in inames})
    lib().ctbo_bounce_batch(C.byref(cfg), _dp(params), _dp(phi_abs), _dp(phi_meta), C.c_int64(n),
                            C.c_int(int(nthreads)), *[_dp(out[k]) for k in dnames + inames])
    return out


def set_overrides(phi_bar=None, rscale=None):
    """phi_bar= / rscale= of SingleFieldInstanton.__init__ for the following calls (None = compute)."""
    nan = float("nan")
    lib().ctbo_set_overrides(C.c_double(nan if phi_bar is None else phi_bar),
                             C.c_double(nan if rscale is None else rscale))


def bounce_one(cfg, params, phi_abs, phi_meta, prof_cap=4096, shot_cap=256):
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(cfg.nparam)
    out9 = np.zeros(9)
    iout6 = np.zeros(6, dtype=np.int32)
    R, Phi, dPhi = np.zeros(prof_cap), np.zeros(prof_cap), np.zeros(prof_cap)
    sx = np.zeros(shot_cap)
    stype = np.zeros(shot_cap, dtype=np.int32)
    n = lib().ctbo_bounce_one(C.byref(cfg), _dp(params), C.c_double(phi_abs), C.c_double(phi_meta), _dp(out9),
                              _dp(iout6), _dp(R), _dp(Phi), _dp(dPhi), C.c_int(prof_cap), _dp(sx), _dp(stype),
                              C.c_int(shot_cap))
    names9 = ["S", "phi0", "r0", "rf", "x", "phi_bar", "rscale", "yf0", "yf1"]
    names6 = ["status", "n_shots", "n_rkck", "n_exact", "n_points", "ctype"]
    res = {k: float(v) for k, v in zip(names9, out9)}
    res.update({k: int(v) for k, v in zip(names6, iout6)})
    ns = res["n_shots"]
    res.update(R=R[:n].copy(), Phi=Phi[:n].copy(), dPhi=dPhi[:n].copy(), shot_x=sx[:ns].copy(),
               shot_type=stype[:ns].copy())
    return res


def jspline(t, c, xmin, xmax, theta, der=0):
    t = np.ascontiguousarray(t, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    th = np.ascontiguousarray(theta, dtype=np.float64)
    out = np.empty_like(th)
    lib().ctbo_jspline(_dp(t), _dp(c), C.c_int(len(t)), C.c_double(xmin), C.c_double(xmax), _dp(th), _dp(out),
                       C.c_int64(th.size), C.c_int(der))
    return out


def potential(cfg, params, phi):
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(cfg.nparam)
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    out = np.empty_like(phi)
    lib().ctbo_potential(C.byref(cfg), _dp(params), _dp(phi), _dp(out), C.c_int64(phi.size))
    return out


def jseries(which, approx, deriv, nterms, lowcoef, x):
    """which: 0 Jb / 1 Jf; approx: 'high' | 'low' (finiteT.py:225-296)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    lc = np.ascontiguousarray(lowcoef, dtype=np.float64)
    out = np.empty_like(x)
    lib().ctbo_jseries(C.c_int(which), C.c_int(1 if approx == "low" else 0), C.c_int(deriv), C.c_int(nterms), _dp(lc),
                       _dp(x), _dp(out), C.c_int64(x.size))
    return out


def potential_rows(cfg, params, phi):
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, cfg.nparam)
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    out = np.empty_like(phi)
    lib().ctbo_potential_rows(C.byref(cfg), _dp(params), _dp(phi), _dp(out), C.c_int64(phi.size))
    return out


def minimize_1d(cfg, params, lo, hi, xtol_rel):
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, cfg.nparam)
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    x, v = np.empty_like(lo), np.empty_like(lo)
    lib().ctbo_minimize_1d(C.byref(cfg), _dp(params), _dp(lo), _dp(hi), C.c_double(xtol_rel), C.c_int64(lo.size),
                           _dp(x), _dp(v))
    return x, v


def vtot_model1_grid(cfg, params, s, T):
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(13)
    s = np.ascontiguousarray(s, dtype=np.float64)
    T = np.ascontiguousarray(T, dtype=np.float64)
    out = np.empty((s.size, T.size))
    lib().ctbo_vtot_model1_grid(C.byref(cfg), _dp(params), _dp(s), C.c_int64(s.size), _dp(T), C.c_int64(T.size),
                                _dp(out))
    return out


def wall_one(cfg, params, phi_abs, phi_meta, Fguess=None, Ftol=1e-4, phi0_rel=1e-3, prof_cap=4096):
    """WallWithConstFriction(phi_abs, phi_meta, V, ...).findProfile(Fguess, Ftol, phitol, npoints, rmin, rmax, phi0_rel)
    (tunneling1D.py:829-999); phitol / npoints / rmin / rmax / phi_eps come from cfg."""
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(cfg.nparam)
    dout, iout = np.zeros(6), np.zeros(4, dtype=np.int32)
    R, Phi, dPhi = (np.zeros(prof_cap) for _ in range(3))
    lib().ctbo_wall_one(C.byref(cfg), _dp(params), C.c_double(phi_abs), C.c_double(phi_meta),
                        C.c_double(float("nan") if Fguess is None else Fguess), C.c_double(Ftol), C.c_double(phi0_rel),
                        _dp(dout), _dp(iout), _dp(R), _dp(Phi), _dp(dPhi), C.c_int(prof_cap))
    n = int(iout[3])
    return dict(F=dout[0], rf=dout[1], phi_bar=dout[2], rscale=dout[3], phi0=dout[4], dphi0=dout[5], status=int(iout[0]),
                n_shots=int(iout[1]), n_rkck=int(iout[2]), n_points=n, R=R[:n], Phi=Phi[:n], dPhi=dPhi[:n])


def find_action(cfg, params, phi_meta, alpha, R, Phi, dPhi):
    """SingleFieldInstanton.findAction for a given profile (tunneling1D.py:763-791)."""
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(cfg.nparam)
    R, Phi, dPhi = (np.ascontiguousarray(a, dtype=np.float64) for a in (R, Phi, dPhi))
    f = lib().ctbo_find_action
    f.restype = C.c_double
    return f(C.byref(cfg), _dp(params), C.c_double(phi_meta), C.c_double(alpha), _dp(R), _dp(Phi), _dp(dPhi),
             C.c_int(len(R)))


def jexact(which, mode, x):
    """finiteT.Jb / Jf from the defining integrals, restated call for call: scipy.integrate.quad (QUADPACK qagi /
    qags, default tolerances) on the reference's integrands.  which 0 = Jb, 1 = Jf; mode 'x' (finiteT.py:40-52,
    69-81), 'dx' (finiteT.py:104-111), 'theta' (finiteT.py:54-66,84-96).  Elementwise over x (finiteT.py:114-130)."""
    from scipy import integrate
    sg, pre = (-1.0, 1.0) if which == 0 else (1.0, -1.0)
    trig = np.sin if which == 0 else np.cos

    def one(v):
        if mode == "dx":
            f = lambda y: y * y * (np.exp(np.sqrt(y * y + v * v)) + sg) ** -1 * v / np.sqrt(y * y + v * v)
            return integrate.quad(f, 0, np.inf)[0]
        theta = v * v if mode == "x" else v
        with np.errstate(all="ignore"):
            if theta >= 0:
                f = lambda y: pre * y * y * np.log(1 + sg * np.exp(-np.sqrt(y * y + theta)))
                return integrate.quad(f, 0, np.inf)[0]
            f = lambda y: (pre * y * y * np.log(1 + sg * np.exp(-np.sqrt(y * y + theta + 0j)))).real
            f1 = lambda y: pre * y * y * np.log(2 * abs(trig(np.sqrt(-theta - y * y) / 2)))
            return integrate.quad(f, abs(theta) ** .5, np.inf)[0] + integrate.quad(f1, 0, abs(theta) ** .5)[0]

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.array([one(float(v)) for v in np.atleast_1d(np.asarray(x, dtype=np.float64))])


def brentq_poly(coefs_high_to_low, a, b):
    p = np.array([len(coefs_high_to_low) - 1] + list(coefs_high_to_low), dtype=np.float64)
    err, nf = C.c_int(0), C.c_int(0)
    x = lib().ctbo_brentq_poly(_dp(p), C.c_double(a), C.c_double(b), C.byref(err), C.byref(nf))
    return x, err.value, nf.value


def fminbound_poly(coefs_high_to_low, a, b, xatol):
    p = np.array([len(coefs_high_to_low) - 1] + list(coefs_high_to_low), dtype=np.float64)
    nf = C.c_int(0)
    x = lib().ctbo_fminbound_poly(_dp(p), C.c_double(a), C.c_double(b), C.c_double(xatol), C.byref(nf))
    return x, nf.value


def simpson(y, x):
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    return lib().ctbo_simpson(_dp(y), _dp(x), C.c_int(len(y)))


def exact_solution(alpha, r, phi0, dV, d2V):
    out = np.zeros(2)
    lib().ctbo_exact_solution(C.c_int(alpha), C.c_double(r), C.c_double(phi0), C.c_double(dV), C.c_double(d2V),
                              _dp(out))
    return out[0], out[1]


def rkqs_poly4(coef5, alpha, y, t, dt_try, epsfrac, epsabs):
    coef5 = np.ascontiguousarray(coef5, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    ef = np.ascontiguousarray(epsfrac, dtype=np.float64)
    ea = np.ascontiguousarray(epsabs, dtype=np.float64)
    out = np.zeros(4)
    st = lib().ctbo_rkqs_poly4(_dp(coef5), C.c_int(alpha), _dp(y), C.c_double(t), C.c_double(dt_try), _dp(ef),
                               _dp(ea), _dp(out))
    return st, out
