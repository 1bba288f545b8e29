#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ (and the product's Jb/Jf table file) by RUNNING
THE REFERENCE ITSELF in the build container.

TEST INFRASTRUCTURE.  Run once, here:   python oracle/gen_golden.py
The reference (/root/reference, pure Python) is imported from a writable scratch copy because its
finiteT module writes two cache files next to itself on first import (finiteT.py:153-163,183-193).
It does not travel to the GPU box, so everything the tests need is frozen into small files:

  cosmotransitions_b200/data/finiteT_tables.npz  the 2x1000 exact-integral samples (xb,yb,xf,yf)
                                                 and scipy's splrep knots/coefficients (tb,cb,tf,cf)
  tests/golden/jspline.npz      Jb_spline / Jf_spline (finiteT.py:167-204), deriv 0..3
  tests/golden/jseries.npz      Jb / Jf with approx='high' (deriv 0..3) and 'low' (finiteT.py:225-375)
  tests/golden/quartic.json     SingleFieldInstanton findProfile/findAction on quartic potentials
  tests/golden/profiles.npz     full R/Phi/dPhi profiles for three cases
  tests/golden/units.npz        exactSolution / rkqs unit vectors
  tests/golden/model1.npz       generic_potential.Vtot for examples/testModel1.py + finite-T bounces
  tests/golden/model1f.json     one-field thermal model sweep
  tests/golden/thermal1f.json   general one-field generic_potential (bosons + fermions, Jb and Jf): Vtot, bounces
  tests/golden/splinepath.npz   every 1-D solve inside pathDeformation.fullTunneling (examples/fullTunneling.py,
                                model1.findAllTransitions): the SplinePath potential spline, profile, action
  tests/golden/alpha14.json     alpha = 1 and 4 (d = 2, 5): quartic bounces and exactSolution unit vectors
  tests/golden/findaction.json  findAction on given (truncated / strided) profiles, alpha in {1, 2, 2.5, 3, 4}
  tests/golden/jexact.npz       Jb / Jf with approx='exact' (deriv 0, 1), Jb_exact2 / Jf_exact2 incl. theta < 0, + mpmath values
  tests/golden/wall.json        WallWithConstFriction.findProfile on the quartic family (F, rf, sampled profile, errors)
  tests/golden/model1_derivs.npz gradV / d2V / dgradV_dT / energyDensity / massSqMatrix of examples/testModel1.py
  tests/golden/VERSIONS.json    numpy / scipy versions (scipy is part of the oracle)
"""
import json
import os
import shutil
import sys
import tempfile
import types
import warnings

import numpy as np
import scipy

warnings.filterwarnings("ignore")

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CTB_REFERENCE", "/root/reference")
GOLD = os.path.join(REPO, "tests", "golden")
DATA = os.path.join(REPO, "cosmotransitions_b200", "data")


def import_reference():
    tmp = tempfile.mkdtemp(prefix="ctref_")
    shutil.copytree(os.path.join(REF, "cosmoTransitions"), os.path.join(tmp, "cosmoTransitions"))
    shutil.copytree(os.path.join(REF, "examples"), os.path.join(tmp, "examples"))
    sys.path.insert(0, tmp)
    sys.path.insert(0, os.path.join(tmp, "examples"))
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
    import cosmoTransitions.tunneling1D as t1
    import cosmoTransitions.helper_functions as hf
    import cosmoTransitions.finiteT as fT
    import cosmoTransitions.generic_potential as gp
    return t1, hf, fT, gp


t1, hf, fT, gp = import_reference()

# ---- instrumentation -------------------------------------------------------------------------
_counts = {"rkck": 0, "exact": 0}
_orig_rkck = hf._rkck


def _rkck_counted(*a, **k):
    _counts["rkck"] += 1
    return _orig_rkck(*a, **k)


hf._rkck = _rkck_counted


class SFI(t1.SingleFieldInstanton):
    def __init__(self, *a, **k):
        self.shots = []
        super().__init__(*a, **k)

    def exactSolution(self, *a):
        _counts["exact"] += 1
        return super().exactSolution(*a)

    def integrateProfile(self, r0, y0, *a):
        rv = super().integrateProfile(r0, y0, *a)
        self.shots.append([float(r0), float(y0[0]), float(y0[1]), float(rv.r), float(rv.y[0]), float(rv.y[1]),
                           {"converged": 0, "undershoot": 1, "overshoot": 2}[rv.convergence_type]])
        return rv


def run_case(V, dV, d2V, phi_abs, phi_meta, alpha, init_kw=None, prof_kw=None, keep_profile=False):
    _counts["rkck"] = _counts["exact"] = 0
    out = {}
    try:
        o = SFI(phi_abs, phi_meta, V, dV, d2V, alpha=alpha, **(init_kw or {}))
        out.update(phi_bar=float(o.phi_bar), rscale=float(o.rscale))
        p = o.findProfile(**(prof_kw or {}))
        S = o.findAction(p)
        last = o.shots[-1]
        out.update(S=float(S), phi0=float(p.Phi[0]), R0=float(p.R[0]), r0=last[0], rf=float(p.R[-1]),
                   npts=int(len(p.R)), n_shots=len(o.shots), shot_types=[s[6] for s in o.shots],
                   shot_rf=[s[3] for s in o.shots], n_rkck=_counts["rkck"], n_exact=_counts["exact"],
                   error=None)
        if keep_profile:
            out["_profile"] = p
    except t1.PotentialError as e:
        out["error"] = "PotentialError:" + e.args[1]
    except hf.IntegrationError as e:
        out["error"] = "IntegrationError:" + str(e.args[0])
    except ValueError as e:
        out["error"] = "ValueError:" + str(e)[:40]
    except AssertionError as e:
        out["error"] = "AssertionError"
    return out


def poly_funcs(coef_low_to_high):
    c = np.asarray(coef_low_to_high, dtype=float)
    p = c[::-1]
    dp = np.polyder(p)
    d2p = np.polyder(dp)
    return (lambda x: np.polyval(p, x)), (lambda x: np.polyval(dp, x)), (lambda x: np.polyval(d2p, x))


# ---- 1. tables ------------------------------------------------------------------------------
def gen_tables():
    os.makedirs(DATA, exist_ok=True)
    tb, cb, kb = fT._tckb
    tf, cf, kf = fT._tckf
    assert kb == 3 and kf == 3
    # low-x series coefficients (finiteT.py:207-222): [a, b, c, d, logab, g_1..g_50] and [a, b, d, logaf, g_1..g_50]
    ab, bb, cbb, db, logab, _, gb = fT.lowCoef_b
    af, bf, df, logaf, _, gf = fT.lowCoef_f
    np.savez(os.path.join(DATA, "finiteT_tables.npz"), xb=fT._xb, yb=fT._yb, xf=fT._xf, yf=fT._yf, tb=tb, cb=cb,
             tf=tf, cf=cf, xbmin=fT._xbmin, xbmax=fT._xbmax, xfmin=fT._xfmin, xfmax=fT._xfmax,
             low_b=np.concatenate([[ab, bb, cbb, db, logab], gb]), low_f=np.concatenate([[af, bf, df, logaf], gf]))


def gen_jspline():
    rng = np.random.default_rng(7)
    th = np.concatenate([
        rng.uniform(-10, 2000, 600), rng.uniform(-7, 7, 500), rng.normal(0, 1e-3, 50), 10 ** rng.uniform(-8, 3, 300),
        fT._xb[::37], fT._xf[::41], fT._xb[[0, 1, 2, 3, -4, -3, -2, -1]], fT._xf[[0, 1, 2, 3, -4, -3, -2, -1]],
        np.array([fT._xbmin, fT._xbmax, fT._xfmin, fT._xfmax, 0.0, -0.0, 1410.0, 1350.0, 1409.9999, 1410.0001,
                  np.nextafter(fT._xbmin, -np.inf), np.nextafter(fT._xbmin, np.inf), -4.84123428, -4.85, -8.868, -9,
                  1e-300, -1e-300, 5e3, 1e6, -1e6]),
    ])
    out = {"theta": th}
    for n in range(4):
        out["Jb%d" % n] = fT.Jb_spline(th, n)
        out["Jf%d" % n] = fT.Jf_spline(th, n)
    np.savez(os.path.join(GOLD, "jspline.npz"), **out)


def gen_jseries():
    """finiteT.Jb / Jf with approx='high' (deriv 0..3, n=8 and n=3) and approx='low' (n=20, n=5)  (finiteT.py:225-375)"""
    rng = np.random.default_rng(17)
    x = np.concatenate([[0.0, 1e-8, 1e-3, 0.05, 0.5, 1.0, 2.0, 3.5, 7.0, 15.0, 40.0, 90.0, 200.0, 800.0],
                        10 ** rng.uniform(-4, 2, 200), rng.uniform(0, 12, 200)])
    out = {"x": x, "xneg": -x[:40]}
    with np.errstate(all="ignore"):
        for nt in (8, 3):
            for d in range(4):
                out["Jb_high_n%d_d%d" % (nt, d)] = fT.Jb(x, approx="high", deriv=d, n=nt)
                out["Jf_high_n%d_d%d" % (nt, d)] = fT.Jf(x, approx="high", deriv=d, n=nt)
        for d in range(4):
            out["Jb_high_neg_d%d" % d] = fT.Jb(-x[:40], approx="high", deriv=d, n=8)
            out["Jf_high_neg_d%d" % d] = fT.Jf(-x[:40], approx="high", deriv=d, n=8)
        for nt in (20, 5, 50):
            out["Jb_low_n%d" % nt] = fT.Jb(x, approx="low", n=nt)
            out["Jf_low_n%d" % nt] = fT.Jf(x, approx="low", n=nt)
    np.savez(os.path.join(GOLD, "jseries.npz"), **out)


# ---- 2. quartic bounces --------------------------------------------------------------------
def quartic_lambdas(c):
    V = lambda p: 0.25 * p ** 4 - (1 + c) / 3 * p ** 3 + c / 2 * p ** 2
    dV = lambda p: p * (p - c) * (p - 1)
    d2V = lambda p: 3 * p * p - 2 * (1 + c) * p + c
    return V, dV, d2V


def gen_quartic():
    cases = []
    cs = list(np.round(0.02 + 0.475 * (np.arange(40) + 0.5) / 40, 12)) + [0.01, 0.2, 0.47, 0.4975, 0.498, 0.499]
    for alpha in (2, 3):
        for c in cs:
            V, dV, d2V = quartic_lambdas(c)
            r = run_case(V, dV, d2V, 1.0, 0.0, alpha)
            r.update(kind="family", c=float(c), alpha=alpha, coef=[0, 0, c / 2, -(1 + c) / 3, 0.25], phi_abs=1.0,
                     phi_meta=0.0, opts={})
            cases.append(r)
    # non-default knobs
    knob_sets = [
        dict(phitol=1e-6, xtol=1e-6), dict(npoints=200), dict(npoints=301, thinCutoff=0.05),
        dict(max_interior_pts=0), dict(max_interior_pts=40), dict(rmin=1e-3), dict(xguess=3.0),
        dict(phitol=1e-5, xtol=1e-5, npoints=1000),
    ]
    for alpha in (2, 3):
        for c in (0.1, 0.3, 0.45, 0.49):
            for kw in knob_sets:
                V, dV, d2V = quartic_lambdas(c)
                r = run_case(V, dV, d2V, 1.0, 0.0, alpha, prof_kw=kw)
                r.update(kind="knobs", c=float(c), alpha=alpha, coef=[0, 0, c / 2, -(1 + c) / 3, 0.25], phi_abs=1.0,
                         phi_meta=0.0, opts=kw)
                cases.append(r)
    # phi_eps knob (enters dV_from_absMin)
    for c in (0.2, 0.47):
        V, dV, d2V = quartic_lambdas(c)
        r = run_case(V, dV, d2V, 1.0, 0.0, 2, init_kw=dict(phi_eps=1e-2))
        r.update(kind="knobs", c=float(c), alpha=2, coef=[0, 0, c / 2, -(1 + c) / 3, 0.25], phi_abs=1.0, phi_meta=0.0,
                 opts=dict(phi_eps=1e-2))
        cases.append(r)
    # phi_bar / rscale overrides of __init__ (tunneling1D.py:140-147)
    for c, ikw in ((0.3, dict(rscale=2.0)), (0.45, dict(phi_bar=0.8)), (0.2, dict(phi_bar=0.35, rscale=2.5))):
        V, dV, d2V = quartic_lambdas(c)
        r = run_case(V, dV, d2V, 1.0, 0.0, 2, init_kw=ikw)
        r.update(kind="init", c=float(c), alpha=2, coef=[0, 0, c / 2, -(1 + c) / 3, 0.25], phi_abs=1.0, phi_meta=0.0,
                 opts={}, init=ikw)
        cases.append(r)
    # shifted / scaled / mirrored quartics, driven through np.polyval so that V, dV, d2V come from
    # the same coefficient vector the device functor receives
    rng = np.random.default_rng(11)
    for k in range(16):
        c = float(rng.uniform(0.05, 0.49))
        pm, pa = float(rng.uniform(-5, 5)), float(rng.uniform(-5, 5))
        if abs(pm - pa) < 0.3:
            pa = pm + 1.7
        lam = float(10 ** rng.uniform(-2, 2))
        # V(phi) = lam * q(u), u = (phi - pm)/(pa - pm), q(u) = u^4/4 - (1+c)/3 u^3 + c/2 u^2
        u = np.poly1d([1.0 / (pa - pm), -pm / (pa - pm)])
        q = 0.25 * u ** 4 - (1 + c) / 3 * u ** 3 + c / 2 * u ** 2
        coef = (lam * q).coeffs[::-1]
        coef = np.concatenate([coef, np.zeros(5 - len(coef))])
        V, dV, d2V = poly_funcs(coef)
        alpha = 2 + (k % 2)
        r = run_case(V, dV, d2V, pa, pm, alpha)
        r.update(kind="affine", c=c, alpha=alpha, coef=[float(x) for x in coef], phi_abs=pa, phi_meta=pm, opts={})
        cases.append(r)
    # error kinds
    for name, coef, pa, pm in [
        ("stable", [0, 0, 0.1, -0.4, 0.25], 0.0, 1.0),       # minima swapped
        ("nobarrier", [0, 0, -0.1, -0.8 / 3, 0.25], 1.0, 0.0),  # c = -0.2: phi=0 is a maximum
        ("nobarrier2", [0, -0.05, 0.1, -0.4, 0.25], 1.0, 0.0),  # tilted
    ]:
        V, dV, d2V = poly_funcs(coef)
        r = run_case(V, dV, d2V, pa, pm, 2)
        r.update(kind="error", name=name, alpha=2, coef=[float(x) for x in coef], phi_abs=pa, phi_meta=pm, opts={})
        cases.append(r)
    with open(os.path.join(GOLD, "quartic.json"), "w") as f:
        json.dump(cases, f, indent=0)
    return cases


def gen_profiles():
    out = {}
    for name, c, alpha in [("thin2", 0.47, 2), ("thick2", 0.2, 2), ("mid3", 0.35, 3)]:
        V, dV, d2V = quartic_lambdas(c)
        # the docstring examples pass only V and dV; d2V then falls back to the FD stencil, but it
        # only enters through phi0-dependent start data, so pass the analytic one like the batch does
        r = run_case(V, dV, d2V, 1.0, 0.0, alpha, keep_profile=True)
        p = r["_profile"]
        out[name + "_R"], out[name + "_Phi"], out[name + "_dPhi"] = p.R, p.Phi, p.dPhi
        out[name + "_S"] = r["S"]
        out[name + "_c"] = c
        out[name + "_alpha"] = alpha
    # docstring form: V and dV only (d2V by finite differences, tunneling1D.py:112-122)
    for name, c in [("doc_thin", 0.47), ("doc_thick", 0.2)]:
        V, dV, _ = quartic_lambdas(c)
        o = t1.SingleFieldInstanton(1.0, 0.0, V, dV)
        p = o.findProfile()
        out[name + "_S"] = o.findAction(p)
    np.savez(os.path.join(GOLD, "profiles.npz"), **out)


# ---- 3. unit vectors ------------------------------------------------------------------------
def gen_units():
    rng = np.random.default_rng(3)
    out = {}
    rows, vals = [], []
    for alpha in (2, 3):
        o = t1.SingleFieldInstanton(1.0, 0.0, *quartic_lambdas(0.3), alpha=alpha)
        for k in range(400):
            r = 10 ** rng.uniform(-5, 3)
            phi0 = rng.uniform(-2, 2)
            dV = rng.normal() * 10 ** rng.uniform(-20, 1)
            d2V = rng.normal() * 10 ** rng.uniform(-3, 1)
            with np.errstate(all="ignore"):
                ph, dph = o.exactSolution(r, phi0, dV, d2V)
            rows.append([alpha, r, phi0, dV, d2V])
            vals.append([ph, dph])
    out["exact_in"], out["exact_out"] = np.array(rows), np.array(vals)
    # rkqs on the quartic EOM
    rows, vals = [], []
    for alpha in (2, 3):
        for k in range(200):
            c = rng.uniform(0.05, 0.49)
            V, dV, d2V = quartic_lambdas(c)
            o = t1.SingleFieldInstanton(1.0, 0.0, V, dV, d2V, alpha=alpha)
            y = np.array([rng.uniform(-0.1, 1.1), rng.normal() * 0.3])
            t = 10 ** rng.uniform(-3, 2)
            dt = 10 ** rng.uniform(-4, 1)
            epsfrac = np.array([1e-4, 1e-4])
            epsabs = np.array([1e-4, 1e-4 / rng.uniform(0.5, 5)])
            f = lambda y_, r_: o.equationOfMotion(y_, r_)
            dy, dtu, dtn = hf.rkqs(y, f(y, t), t, f, dt, epsfrac, epsabs)
            rows.append([alpha, c, y[0], y[1], t, dt, epsabs[0], epsabs[1]])
            vals.append([dy[0], dy[1], dtu, dtn])
    out["rkqs_in"], out["rkqs_out"] = np.array(rows), np.array(vals)
    np.savez(os.path.join(GOLD, "units.npz"), **out)


# ---- 4. model1 (examples/testModel1.py) ----------------------------------------------------
def gen_model1():
    from scipy import optimize
    import testModel1
    m = testModel1.model1()
    v2 = testModel1.v2
    params = dict(l1=m.l1, l2=m.l2, mu2=m.mu2, Y1=m.Y1, Y2=m.Y2, n=float(m.n), renorm=m.renormScaleSq, v2=v2)
    x_low = np.array([286.39824989, 382.19727238])
    x_high = np.array([231.10296383, -136.82260069])
    L = np.linalg.norm(x_high - x_low)
    nhat = (x_high - x_low) / L
    s = np.linspace(-0.2 * L, 1.2 * L, 96)
    T = np.concatenate([np.linspace(1, 250, 40), [0.0, 1e-3, 84.2381996653]])
    X = x_low[None, :] + s[:, None] * nhat[None, :]
    # (the reference's in-place `y +=` needs X already broadcast to the full grid shape)
    Vg = m.Vtot(np.broadcast_to(X[:, None, :], (len(s), len(T), 2)), T[None, :])
    out = dict(grid_s=s, grid_T=T, grid_V=Vg, grid_x0=x_low, grid_nhat=nhat, **{"p_" + k: v for k, v in params.items()})
    # scattered 2-D points
    rng = np.random.default_rng(5)
    Xs = rng.uniform(-400, 400, (300, 2))
    Ts = rng.uniform(0, 300, 300)
    out["pts_X"], out["pts_T"], out["pts_V"] = Xs, Ts, m.Vtot(Xs, Ts)
    out["pts_V1T"] = m.V1T_from_X(Xs, Ts)           # generic_potential.py:298-310
    out["pts_DV"] = m.DVtot(Xs, Ts)                  # generic_potential.py:339-345
    # finite-T bounces along the straight line between the two phases' minima at each T
    bT, bxl, bxh, res = [], [], [], []
    xl, xh = x_low.copy(), x_high.copy()
    for Tb in [84.2381996653, 78.0, 80.0, 90.0, 95.0, 100.0, 105.0]:
        xl = optimize.fmin(m.Vtot, xl if Tb >= 84 else x_low, args=(Tb,), xtol=1e-8, ftol=1e-12, disp=0)
        xh = optimize.fmin(m.Vtot, xh if Tb >= 84 else x_high, args=(Tb,), xtol=1e-8, ftol=1e-12, disp=0)
        Lb = np.linalg.norm(xh - xl)
        nb = (xh - xl) / Lb
        Vline = lambda phi, xl=xl, nb=nb, Tb=Tb: m.Vtot(xl + np.asarray(phi)[..., None] * nb, Tb)
        r = run_case(Vline, None, None, 0.0, Lb, 2)
        r.update(T=Tb, L=Lb)
        bT.append(Tb); bxl.append(xl.copy()); bxh.append(xh.copy()); res.append(r)
        print("model1 bounce T=%g L=%g" % (Tb, Lb), {k: r.get(k) for k in ("S", "npts", "n_shots", "n_rkck", "error")})
    out["b_T"], out["b_xlow"], out["b_xhigh"] = np.array(bT), np.array(bxl), np.array(bxh)
    out["b_json"] = json.dumps(res)
    np.savez(os.path.join(GOLD, "model1.npz"), **out)


class SMlike(gp.generic_potential):
    """One-field model with the particle content layout of CTB_POT_THERMAL1F: a Higgs-like field with gauge
    bosons, Goldstones, a heavy portal scalar multiplet with a thermal mass, and a top-like fermion -- written
    against the real generic_potential so that V1 / V1T / Jb_spline / Jf_spline are the reference's own code."""
    v2 = 246. ** 2

    def init(self, lam=0.13, g2=0.42, gz2=0.55, yt2=0.98, Yp=0.9, np_=12., tp=0.15):
        self.Ndim = 1
        self.renormScaleSq = self.v2
        self.lam, self.g2, self.gz2, self.yt2, self.Yp, self.np_, self.tp = lam, g2, gz2, yt2, Yp, np_, tp
        self.mu2 = lam * self.v2

    def V0(self, X):
        phi = np.asanyarray(X)[..., 0]
        return -.5 * self.mu2 * phi ** 2 + .25 * self.lam * phi ** 4

    def bosons_spec(self):   # (a, b, t, dof, c)
        return [(-self.mu2, 3 * self.lam, 0.0, 1., 1.5), (-self.mu2, self.lam, 0.0, 3., 1.5),
                (0.0, self.g2 / 4, 0.0, 6., 5. / 6), (0.0, self.gz2 / 4, 0.0, 3., 5. / 6),
                (0.0, self.Yp, self.tp, self.np_, 1.5)]

    def fermions_spec(self):  # (a, b, dof)
        return [(0.0, self.yt2 / 2, 12.)]

    def boson_massSq(self, X, T):
        phi = np.asanyarray(X)[..., 0]
        T = np.asanyarray(T, dtype=float)
        spec = self.bosons_spec()
        M = np.array([a + b * phi * phi + t * T * T for (a, b, t, n, c) in spec])
        M = np.rollaxis(M, 0, len(M.shape))
        return M, np.array([s_[3] for s_ in spec]), np.array([s_[4] for s_ in spec])

    def fermion_massSq(self, X):
        phi = np.asanyarray(X)[..., 0]
        spec = self.fermions_spec()
        M = np.array([a + b * phi * phi for (a, b, n) in spec])
        M = np.rollaxis(M, 0, len(M.shape))
        return M, np.array([s_[2] for s_ in spec])


def gen_thermal1f():
    from scipy import optimize
    m = SMlike()
    V = lambda phi, T: m.Vtot(np.asarray(phi, dtype=float)[..., None], T)
    rng = np.random.default_rng(31)
    phis, Ts = rng.uniform(-50, 400, 400), rng.uniform(0, 250, 400)
    out = dict(V0_coefs=[0.0, 0.0, -.5 * m.mu2, 0.0, .25 * m.lam], renorm=m.renormScaleSq, bosons=m.bosons_spec(),
               fermions=m.fermions_spec(), pts_phi=phis.tolist(), pts_T=Ts.tolist(),
               pts_V=[float(V(p_, t_)) for p_, t_ in zip(phis, Ts)])
    # temperature window: phi = 0 metastable, broken minimum deeper
    def curv0(T):
        return float(V(0.5, T) - V(0.0, T))
    def broken_min(T):
        return float(optimize.minimize_scalar(lambda p: float(V(p, T)), bounds=(30, 500), method="bounded",
                                              options=dict(xatol=1e-10)).x)
    T0 = optimize.brentq(curv0, 5, 400)
    Tc = optimize.brentq(lambda T: float(V(broken_min(T), T) - V(0.0, T)), T0 + 0.2, 400)
    out.update(T0=float(T0), Tc=float(Tc))
    cases = []
    for frac in (0.2, 0.5, 0.75, 0.92):
        T = T0 + frac * (Tc - T0)
        pa = broken_min(T)
        r = run_case(lambda phi, T=T: V(phi, T), None, None, pa, 0.0, 2)
        r.update(T=float(T), phi_abs=pa)
        cases.append(r)
        print("thermal1f T=%.3f" % T, {q: r.get(q) for q in ("S", "npts", "n_shots", "n_rkck", "error")})
    out["bounces"] = cases
    # the generic_potential surface on (phi, T) grids with broadcasting, radiation terms on and off
    m2 = SMlike()
    m2.num_boson_dof, m2.num_fermion_dof = 40., 90.
    gphi, gT = np.linspace(-20, 320, 18), np.array([0.0, 35.0, 90.0, 160.0, 240.0])
    X = np.broadcast_to(gphi[:, None, None], (18, 5, 1))   # (the reference's in-place `y +=` needs the full shape)
    gp_ = dict(phi=gphi.tolist(), T=gT.tolist(), num_boson_dof=40., num_fermion_dof=90.)
    gp_["Vtot"] = m2.Vtot(X, gT[None, :]).tolist()
    gp_["Vtot_norad"] = m2.Vtot(X, gT[None, :], False).tolist()
    gp_["V1T"] = m2.V1T_from_X(X, gT[None, :]).tolist()
    gp_["DVtot"] = m2.DVtot(X, gT[None, :]).tolist()
    gp_["gradV"] = m2.gradV(X, gT[None, :]).tolist()
    gp_["d2V"] = m2.d2V(X, gT[None, :]).tolist()
    gp_["dgradV_dT"] = m2.dgradV_dT(X, gT[None, :]).tolist()
    gp_["energyDensity"] = m2.energyDensity(X, gT[None, :]).tolist()
    gp_["Vtot_point"] = float(m2.Vtot(np.array([123.0]), 77.0))
    out["generic_potential"] = gp_
    with open(os.path.join(GOLD, "thermal1f.json"), "w") as f:
        json.dump(out, f, indent=0)


def gen_splinepath():
    """The reference's OWN multi-field flow: pathDeformation.fullTunneling on examples/fullTunneling.py
    (BASELINE.json configs[0]) and model1.findAllTransitions().  Every time fullTunneling instantiates the
    tunneling class it passes SplinePath.V / dV / d2V (pathDeformation.py:959-964); the spline (t, c, k),
    the profile, evenlySpacedPhi's output and the action are recorded."""
    import cosmoTransitions.pathDeformation as pd
    import fullTunneling as ex
    import testModel1
    rec = []

    class Recorder(t1.SingleFieldInstanton):
        def __init__(self, phi_absMin, phi_metaMin, V, dV, d2V, **kw):
            super().__init__(phi_absMin, phi_metaMin, V, dV, d2V, **kw)
            self._rec = dict(tck=V.__self__._V_tck, phi_abs=phi_absMin, phi_meta=phi_metaMin,
                             phi_bar=self.phi_bar, rscale=self.rscale)
            rec.append(self._rec)

        def findProfile(self, **kw):
            p = super().findProfile(**kw)
            self._rec.update(R=p.R, Phi=p.Phi, dPhi=p.dPhi, S=self.findAction(p))
            return p

        def evenlySpacedPhi(self, phi, dphi, **kw):
            out = super().evenlySpacedPhi(phi, dphi, **kw)
            self._rec.update(esp_kw=kw, esp_p=out[0], esp_dp=out[1])
            return out

    tags = []
    for fy in (2.0, 80.0):
        n0 = len(rec)
        m = ex.Potential(c=5., fx=0., fy=fy)
        Y = pd.fullTunneling([[1, 1.], [0, 0]], m.V, m.dV, tunneling_class=Recorder)
        tags += ["fullTunneling_fy%g" % fy] * (len(rec) - n0)
        print("fullTunneling fy=%g: %d 1-D solves, action %.9f" % (fy, len(rec) - n0, Y.action))
    n0 = len(rec)
    m1 = testModel1.model1()
    m1.findAllTransitions(tunnelFromPhase_args=dict(fullTunneling_params=dict(tunneling_class=Recorder)))
    tags += ["model1_findAllTransitions"] * (len(rec) - n0)
    print("model1.findAllTransitions: %d 1-D solves; Tnuc %.6f action %.6f" % (
        len(rec) - n0, m1.TnTrans[0]["Tnuc"], m1.TnTrans[0]["action"]))
    out = {"n": len(rec), "tags": np.array(tags)}
    for i, r in enumerate(rec):
        t, c, k = r["tck"]
        assert k == 3
        out["t%d" % i], out["c%d" % i] = t, c
        out["scal%d" % i] = np.array([r["phi_abs"], r["phi_meta"], r["phi_bar"], r["rscale"], r.get("S", np.nan)])
        for key in ("R", "Phi", "dPhi", "esp_p", "esp_dp"):
            if key in r:
                out["%s%d" % (key, i)] = r[key]
        if "esp_kw" in r:
            out["espkw%d" % i] = np.array([r["esp_kw"].get("npoints", 100), r["esp_kw"].get("k", 1),
                                           1 if r["esp_kw"].get("fixAbs", True) else 0])
    np.savez_compressed(os.path.join(GOLD, "splinepath.npz"), **out)


class Model1F(gp.generic_potential):
    """One-field model1-style thermal potential, written against the real
    generic_potential so that Vtot / V1 / V1T / Jb_spline are the reference's own code."""

    def init(self, lam=0.12, Y=0.35, n=30., v2=246. ** 2):
        self.Ndim = 1
        self.renormScaleSq = v2
        self.lam, self.Y, self.n, self.v2 = lam, Y, n, v2

    def V0(self, X):
        phi = np.asanyarray(X)[..., 0]
        return .25 * self.lam * (phi * phi - self.v2) ** 2

    def boson_massSq(self, X, T):
        phi = np.asanyarray(X)[..., 0]
        T = np.asanyarray(T, dtype=float)
        M = np.array([self.lam * (3 * phi * phi - self.v2) + 0 * T, self.Y * phi * phi + 0 * T])
        M = np.rollaxis(M, 0, len(M.shape))
        return M, np.array([1., self.n]), np.array([1.5, 1.5])


def gen_model1f():
    from scipy import optimize
    rng = np.random.default_rng(20260101)
    cases = []
    for k in range(6):
        lam, Y, n = (0.12, 0.35, 30.) if k == 0 else (0.12 * rng.uniform(.8, 1.2), 0.35 * rng.uniform(.8, 1.2),
                                                      30. * rng.uniform(.8, 1.2))
        m = Model1F(lam=lam, Y=Y, n=n)
        V = lambda phi, T: m.Vtot(np.asarray(phi, dtype=float)[..., None], T)
        # T0: phi=0 turns into a local minimum; Tc: degenerate with the broken minimum
        def curv0(T):
            h = 0.5
            return float(V(h, T) - V(0.0, T))
        T0 = optimize.brentq(curv0, 20, 300)
        def gap(T):
            pa = optimize.minimize_scalar(lambda p: float(V(p, T)), bounds=(50, 500), method="bounded",
                                          options=dict(xatol=1e-10)).x
            return float(V(pa, T) - V(0.0, T))
        Tc = optimize.brentq(gap, T0 + 0.5, 300)
        for frac in (0.15, 0.4, 0.65, 0.85, 0.95):
            T = T0 + frac * (Tc - T0)
            pa = float(optimize.minimize_scalar(lambda p: float(V(p, T)), bounds=(50, 500), method="bounded",
                                                options=dict(xatol=1e-10)).x)
            Vl = lambda phi, T=T: V(phi, T)
            r = run_case(Vl, None, None, pa, 0.0, 2)
            r.update(lam=lam, Y=Y, n=n, v2=246. ** 2, T=float(T), phi_abs=pa, phi_meta=0.0, T0=float(T0), Tc=float(Tc),
                     Vsamples=[float(V(x, T)) for x in (0.0, 0.3 * pa, 0.7 * pa, pa, 1.1 * pa)])
            cases.append(r)
            print("model1f", k, "T=%.3f" % T, {q: r.get(q) for q in ("S", "npts", "n_shots", "n_rkck", "error")})
        # nucleation temperature the reference's way: brentq over T of S3(T)/T - 140
        # (transitionFinder.py:773, 836-837), phi_absMin(T) re-minimised at every T
        def nucl(T):
            pa_ = float(optimize.minimize_scalar(lambda p: float(V(p, T)), bounds=(50, 500), method="bounded",
                                                 options=dict(xatol=1e-10)).x)
            o = t1.SingleFieldInstanton(pa_, 0.0, lambda phi: V(phi, T))
            return o.findAction(o.findProfile()) / T - 140.0
        Tn = optimize.brentq(nucl, T0 + 0.05 * (Tc - T0), T0 + 0.98 * (Tc - T0), xtol=1e-6)
        for c_ in cases[-5:]:
            c_["Tnuc"] = float(Tn)
        print("model1f", k, "Tnuc = %.6f" % Tn)
    with open(os.path.join(GOLD, "model1f.json"), "w") as f:
        json.dump(cases, f, indent=0)


def gen_findaction():
    """findAction on GIVEN profiles (tunneling1D.py:763-791): the three profiles of profiles.npz, truncated to
    odd / even / tiny lengths, with alpha in {1, 2, 2.5, 3, 4} -> tests/golden/findaction.json"""
    g = np.load(os.path.join(GOLD, "profiles.npz"))
    cases = []
    for name in ("thin2", "thick2", "mid3"):
        c = float(g[name + "_c"])
        V, dV, d2V = quartic_lambdas(c)
        R, Phi, dPhi = g[name + "_R"], g[name + "_Phi"], g[name + "_dPhi"]
        for alpha in (1, 2, 2.5, 3, 4):
            o = t1.SingleFieldInstanton(1.0, 0.0, V, dV, d2V, alpha=alpha)
            for sl in (slice(None), slice(0, 2), slice(0, 3), slice(0, 4), slice(0, 5), slice(0, 100), slice(3, 104),
                       slice(250, None), slice(0, None, 7)):
                p = o.profile_rval(R[sl], Phi[sl], dPhi[sl], None)
                cases.append(dict(name=name, c=c, alpha=alpha, sl=[sl.start, sl.stop, sl.step], n=int(len(p.R)),
                                  S=float(o.findAction(p))))
    with open(os.path.join(GOLD, "findaction.json"), "w") as f:
        json.dump(cases, f, indent=0)
    print(len(cases), "findAction cases")


def gen_jexact():
    """finiteT.Jb_exact / Jf_exact / dJb_exact / dJf_exact / Jb_exact2 / Jf_exact2 (finiteT.py:40-147, QUADPACK through
    scipy.integrate.quad, tolerance 1.49e-8) + 30-digit mpmath values of the same integrals (independent check of the
    device quadrature, which is more accurate than the reference's) -> tests/golden/jexact.npz"""
    import mpmath as mp
    mp.mp.dps = 30
    rng = np.random.default_rng(23)
    x = np.concatenate([[0.0, 1e-8, 1e-3, 0.05, 0.5, 1.0, 2.0, 3.5, 7.0, 15.0, 30.0], 10 ** rng.uniform(-3, 1.5, 60)])
    th = np.concatenate([[0.0, -1e-6, -0.01, -1.0, -3.72402637, -6.82200203, -9.8, -30.0, -45.0, 1e-4, 2.0, 100.0, 1400.0],
                         rng.uniform(-8, 8, 60)])
    out = dict(x=x, theta=th, xneg=-x[:12])
    with np.errstate(all="ignore"):
        out["Jb_exact"], out["Jf_exact"] = fT.Jb(x, approx="exact"), fT.Jf(x, approx="exact").real
        out["dJb_exact"], out["dJf_exact"] = fT.Jb(x, approx="exact", deriv=1), fT.Jf(x, approx="exact", deriv=1)
        out["Jb_exact_neg"], out["dJb_exact_neg"] = fT.Jb_exact(-x[:12]), fT.dJb_exact(-x[:12])
        out["Jf_exact_neg"], out["dJf_exact_neg"] = fT.Jf_exact(-x[:12]).real, fT.dJf_exact(-x[:12])
        out["Jb_exact2"], out["Jf_exact2"] = fT.Jb_exact2(th), fT.Jf_exact2(th)
        ximag = 1j * np.array([0.1, 1.0, 1.9, 2.6])
        out["ximag"] = ximag.imag
        out["Jb_exact_imag"], out["Jf_exact_imag"] = fT.Jb_exact(ximag), fT.Jf_exact(ximag).real

    def mpJ(theta, which):
        theta = mp.mpf(float(theta))
        sg = -1 if which == 0 else 1
        pre = 1 if which == 0 else -1
        if theta == 0:
            return -mp.pi ** 4 / 45 if which == 0 else -7 * mp.pi ** 4 / 360
        if theta >= 0:
            f = lambda y: pre * y * y * mp.log(1 + sg * mp.exp(-mp.sqrt(y * y + theta)))
            return mp.quad(f, [0, mp.sqrt(theta) + mp.mpf("1e-30"), 1, 10, 50, mp.inf])
        a = mp.sqrt(-theta)
        f = lambda y: pre * y * y * mp.log(1 + sg * mp.exp(-mp.sqrt(y * y - a * a)))
        if which == 0:
            f1 = lambda y: y * y * mp.log(2 * abs(mp.sin(mp.sqrt(a * a - y * y) / 2)))
            sing = [mp.sqrt(a * a - (2 * mp.pi * k) ** 2) for k in range(0, int(a / (2 * mp.pi)) + 1)]
        else:
            f1 = lambda y: -y * y * mp.log(2 * abs(mp.cos(mp.sqrt(a * a - y * y) / 2)))
            sing = [mp.sqrt(a * a - (mp.pi * (2 * k + 1)) ** 2) for k in range(0, 50) if a > mp.pi * (2 * k + 1)]
        pts = sorted(set([mp.mpf(0)] + sing + [a]))
        return mp.quad(f1, pts) + mp.quad(f, [a, a + 1, a + 10, a + 50, mp.inf])

    def mpdJ(x, which):
        x = mp.mpf(float(x))
        sg = -1 if which == 0 else 1
        f = lambda y: y * y / (mp.exp(mp.sqrt(y * y + x * x)) + sg) * x / mp.sqrt(y * y + x * x)
        return mp.quad(f, [0, x + mp.mpf("1e-30"), 1, 10, 50, mp.inf])

    out["mp_theta"] = th[:40]
    out["mp_Jb"] = np.array([float(mpJ(t, 0)) for t in th[:40]])
    out["mp_Jf"] = np.array([float(mpJ(t, 1)) for t in th[:40]])
    out["mp_x"] = x[1:31]
    out["mp_dJb"] = np.array([float(mpdJ(t, 0)) for t in x[1:31]])
    out["mp_dJf"] = np.array([float(mpdJ(t, 1)) for t in x[1:31]])
    np.savez(os.path.join(GOLD, "jexact.npz"), **out)
    print("jexact: reference quad vs mpmath, max abs diff Jb", np.max(np.abs(out["Jb_exact2"][:40] - out["mp_Jb"])),
          "Jf", np.max(np.abs(out["Jf_exact2"][:40] - out["mp_Jf"])))


def gen_wall():
    """WallWithConstFriction.findProfile (tunneling1D.py:829-999) on the quartic family, default and non-default
    knobs, error cases -> tests/golden/wall.json (profiles sampled every 20th point)"""
    class Wall(t1.WallWithConstFriction):
        nshots = 0

        def integrateProfile(self, *a):
            self.nshots += 1
            return super().integrateProfile(*a)

    cases = []
    runs = [(c, {}) for c in (0.05, 0.1, 0.15, 0.2, 0.25, 0.3, 0.35, 0.4, 0.45, 0.48, 0.49, 0.499)]
    runs += [(0.3, dict(phitol=1e-6, Ftol=1e-6)), (0.4, dict(npoints=201)), (0.25, dict(Fguess=0.5)),
             (0.45, dict(phi0_rel=1e-2, rmin=1e-3)), (0.2, dict(rmax=2.0)), (0.35, dict(npoints=2))]
    for c, kw in runs:
        V, dV, d2V = quartic_lambdas(c)
        _counts["rkck"] = 0
        r = dict(c=float(c), coef=[0, 0, c / 2, -(1 + c) / 3, 0.25], phi_abs=1.0, phi_meta=0.0, opts=kw, error=None)
        try:
            w = Wall(1.0, 0.0, V, dV, d2V)
            r.update(phi_bar=float(w.phi_bar), rscale=float(w.rscale))
            p = w.findProfile(**kw)
            r.update(F=float(p.F), rf=float(p.R[-1]), npts=int(len(p.R)), n_shots=w.nshots, n_rkck=_counts["rkck"],
                     R=[float(x) for x in p.R[::20]], Phi=[float(x) for x in p.Phi[::20]],
                     dPhi=[float(x) for x in p.dPhi[::20]], Phi_last=float(p.Phi[-1]), dPhi_last=float(p.dPhi[-1]))
        except t1.PotentialError as e:
            r["error"] = "PotentialError:" + e.args[1]
        except hf.IntegrationError as e:
            r["error"] = "IntegrationError:" + str(e.args[0])
        cases.append(r)
        print("wall", c, kw, {q: r.get(q) for q in ("F", "rf", "n_shots", "n_rkck", "error")})
    for name, coef, pa, pm in [("stable", [0, 0, 0.1, -0.4, 0.25], 0.0, 1.0)]:
        V, dV, d2V = poly_funcs(coef)
        r = dict(name=name, coef=[float(x) for x in coef], phi_abs=pa, phi_meta=pm, opts={}, error=None)
        try:
            Wall(pa, pm, V, dV, d2V).findProfile()
        except t1.PotentialError as e:
            r["error"] = "PotentialError:" + e.args[1]
        except hf.IntegrationError as e:
            r["error"] = "IntegrationError:" + str(e.args[0])
        cases.append(r)
        print("wall", name, r["error"])
    with open(os.path.join(GOLD, "wall.json"), "w") as f:
        json.dump(cases, f, indent=0)


def gen_model1_derivs():
    """generic_potential.gradV / d2V / dgradV_dT / energyDensity / massSqMatrix for examples/testModel1.py
    (finite differences of Vtot, generic_potential.py:347-456) -> tests/golden/model1_derivs.npz"""
    import testModel1
    m = testModel1.model1()
    rng = np.random.default_rng(55)
    X = rng.uniform(-350, 350, (48, 2))
    T = rng.uniform(5, 250, 48)
    out = dict(X=X, T=T, gradV=m.gradV(X, T), d2V=m.d2V(X, T), dgradV_dT=m.dgradV_dT(X, T),
               energyDensity=m.energyDensity(X, T), massSqMatrix=m.massSqMatrix(X),
               gradV_point=m.gradV(np.array([200., 150.]), 90.0), d2V_point=m.d2V(np.array([200., 150.]), 90.0))
    # broadcasting: 6 field points against 4 temperatures
    Xg = np.broadcast_to(X[:6, None, :], (6, 4, 2))
    out["grid_T"] = T[:4]
    out["grid_gradV"] = m.gradV(Xg, T[None, :4])
    np.savez(os.path.join(GOLD, "model1_derivs.npz"), **out)
    print({k: np.asarray(v).shape for k, v in out.items()})


def gen_phases():
    """Minima of examples/testModel1.py's Vtot followed in temperature for 10 parameter points (SURVEY s.8f row 1):
    phase A from approxZeroTMin()[0] at T = 0 upwards until it ends, phase B from the point the minimiser lands on
    downwards (a phase has ended when the minimiser lands more than 0.3 (|X_start| + 1) away from the previous
    minimum: the spinodal end moves the minimum by ~0.1 of that per 4 K step, the neighbouring phase is ~1 away) -- each
    point by the reference's own tool, scipy.optimize.fmin on m.Vtot (generic_potential.py:477-484,
    transitionFinder.py:127-129, 658-664) with tight tolerances, then polished by Newton steps on the reference's own
    finite-difference gradV / d2V (generic_potential.py:347-442).  Plus getPhases() of the default model."""
    from scipy import optimize
    import testModel1
    rng = np.random.default_rng(77)
    base = dict(m1=120., m2=50., mu=25., Y1=.1, Y2=.15, n=30)
    pts = [dict(base)]
    for _ in range(9):
        pts.append({k: (v * rng.uniform(.85, 1.15) if k != "n" else float(int(v * rng.uniform(.85, 1.15)))) for k, v in base.items()})
    T = np.linspace(0.0, 240.0, 61)

    def tight(m, x, t):
        x = optimize.fmin(m.Vtot, x, args=(t,), xtol=1e-9, ftol=1e-13, maxiter=4000, maxfun=8000, disp=0)
        xn = np.array(x, float)
        for _ in range(4):
            H = m.d2V(xn, t)
            if np.linalg.eigvalsh(H)[0] <= 0:
                break
            xn = xn - np.linalg.solve(H, m.gradV(xn, t))
        return np.array(x, float), xn

    model8, XA, XB, XAn, XBn, VA, VB, eA = [], [], [], [], [], [], [], []
    for kw in pts:
        m = testModel1.model1(**kw)
        model8.append([m.l1, m.l2, m.mu2, m.Y1, m.Y2, float(m.n), m.renormScaleSq, testModel1.v2])
        xa = np.full((len(T), 2), np.nan); xan = xa.copy(); xb = xa.copy(); xbn = xa.copy()
        x = np.array(m.approxZeroTMin()[0], float)
        scale = np.linalg.norm(x) + 1.0
        end = len(T)
        for k, t in enumerate(T):
            xf, xn = tight(m, x if k < 2 else 2 * xan[k - 1] - xan[k - 2], t)
            if k and np.linalg.norm(xn - xan[k - 1]) > 0.3 * scale:
                xf, xn = tight(m, xan[k - 1], t)     # the extrapolated start may overshoot: retry from the last minimum
            if k and np.linalg.norm(xn - xan[k - 1]) > 0.3 * scale:
                end = k
                xb[k], xbn[k] = xf, xn
                break
            xa[k], xan[k] = xf, xn
            x = xn
        if end < len(T):
            for k in range(end - 1, -1, -1):
                xf, xn = tight(m, xbn[k + 1] if k > end - 3 else 2 * xbn[k + 1] - xbn[k + 2], T[k])
                H = m.d2V(xn, T[k])
                if np.linalg.norm(xn - xbn[k + 1]) > 0.3 * (np.linalg.norm(xbn[end]) + 1.0) or np.linalg.eigvalsh(H)[0] <= 0:
                    break
                xb[k], xbn[k] = xf, xn
        XA.append(xa); XAn.append(xan); XB.append(xb); XBn.append(xbn); eA.append(end)
        with np.errstate(invalid="ignore"):
            VA.append(np.array([m.Vtot(xan[k], T[k]) if np.isfinite(xan[k, 0]) else np.nan for k in range(len(T))]))
            VB.append(np.array([m.Vtot(xbn[k], T[k]) if np.isfinite(xbn[k, 0]) else np.nan for k in range(len(T))]))
        print("phases", kw, "A ends at index", end, "B from", int(np.argmax(np.isfinite(xb[:, 0]))),
              "max |fmin - newton| A %.2e B %.2e" % (np.nanmax(np.abs(xa - xan)), np.nanmax(np.abs(xb - xbn)) if np.isfinite(xb).any() else 0))
    out = dict(model8=np.array(model8), T=T, XA_fmin=np.array(XA), XA=np.array(XAn), XB_fmin=np.array(XB), XB=np.array(XBn),
               VA=np.array(VA), VB=np.array(VB), endA=np.array(eA))
    # the reference's own phase tracing of the default model (transitionFinder.traceMultiMin through getPhases)
    m = testModel1.model1()
    ph = m.getPhases()
    for key, p in ph.items():
        out["ref_phase%d_T" % key], out["ref_phase%d_X" % key] = np.array(p.T), np.array(p.X)
    np.savez(os.path.join(GOLD, "phases.npz"), **out)


def gen_alpha14():
    """alpha = 1 and 4 (d = 2 and 5 dimensions, Bessel order nu = 0 and 3/2): quartic family, a few knob sets,
    and exactSolution unit vectors -> tests/golden/alpha14.json"""
    cases = []
    cs = list(np.round(0.02 + 0.475 * (np.arange(16) + 0.5) / 16, 12)) + [0.01, 0.47, 0.495]
    for alpha in (1, 4):
        for c in cs:
            V, dV, d2V = quartic_lambdas(c)
            r = run_case(V, dV, d2V, 1.0, 0.0, alpha)
            r.update(kind="family", c=float(c), alpha=alpha, coef=[0, 0, c / 2, -(1 + c) / 3, 0.25], phi_abs=1.0,
                     phi_meta=0.0, opts={})
            cases.append(r)
        for c in (0.1, 0.45):
            for kw in (dict(phitol=1e-6, xtol=1e-6), dict(npoints=301, thinCutoff=0.05), dict(max_interior_pts=0)):
                V, dV, d2V = quartic_lambdas(c)
                r = run_case(V, dV, d2V, 1.0, 0.0, alpha, prof_kw=kw)
                r.update(kind="knobs", c=float(c), alpha=alpha, coef=[0, 0, c / 2, -(1 + c) / 3, 0.25], phi_abs=1.0,
                         phi_meta=0.0, opts=kw)
                cases.append(r)
    rng = np.random.default_rng(14)
    rows, vals = [], []
    for alpha in (1, 4):
        o = t1.SingleFieldInstanton(1.0, 0.0, *quartic_lambdas(0.3), alpha=alpha)
        for k in range(400):
            r = 10 ** rng.uniform(-5, 3)
            phi0 = rng.uniform(-2, 2)
            dV = rng.normal() * 10 ** rng.uniform(-20, 1)
            d2V = rng.normal() * 10 ** rng.uniform(-3, 1)
            with np.errstate(all="ignore"):
                ph, dph = o.exactSolution(r, phi0, dV, d2V)
            rows.append([alpha, r, phi0, dV, d2V])
            vals.append([float(ph), float(dph)])
    with open(os.path.join(GOLD, "alpha14.json"), "w") as f:
        json.dump(dict(cases=cases, exact_in=rows, exact_out=vals), f, indent=0)
    for c_ in cases:
        print("alpha14", c_["alpha"], c_["c"], c_["opts"], {q: c_.get(q) for q in ("S", "npts", "n_shots", "error")})


class XSM(gp.generic_potential):
    """A second multi-field model, written against the real generic_potential: Higgs h + Z2 singlet s with the scalar
    mass matrix eigenvalues, Goldstones, W, Z and a top quark -- the particle content a user's subclass would have
    (generic_potential.py:132-238).  Mirrored in CUDA by tests/test_gpu_jit.py through jit.CompiledModel."""
    P = dict(lh=0.129, ls=0.5, lhs=0.6, mus2=100. ** 2, g2=0.4225, gz2=0.55, yt2=0.98, v2=246. ** 2)

    def init(self):
        self.Ndim = 2
        self.renormScaleSq = self.P["v2"]
        self.muh2 = self.P["lh"] * self.P["v2"]

    def V0(self, X):
        X = np.asanyarray(X)
        h, s_ = X[..., 0], X[..., 1]
        P = self.P
        return (-.5 * self.muh2 * h * h + .25 * P["lh"] * h ** 4 - .5 * P["mus2"] * s_ * s_ + .25 * P["ls"] * s_ ** 4
                + .5 * P["lhs"] * h * h * s_ * s_)

    def boson_massSq(self, X, T):
        X = np.asanyarray(X)
        h, s_ = X[..., 0], X[..., 1]
        P = self.P
        a = -self.muh2 + 3 * P["lh"] * h * h + P["lhs"] * s_ * s_
        b = -P["mus2"] + 3 * P["ls"] * s_ * s_ + P["lhs"] * h * h
        c = 2 * P["lhs"] * h * s_
        A, B = .5 * (a + b), np.sqrt(.25 * (a - b) ** 2 + c * c)
        gold = -self.muh2 + P["lh"] * h * h + P["lhs"] * s_ * s_
        mW, mZ = .25 * P["g2"] * h * h, .25 * P["gz2"] * h * h
        M = np.array([A + B, A - B, gold, mW, mZ])
        M = np.rollaxis(M, 0, len(M.shape))
        return M, np.array([1., 1., 3., 6., 3.]), np.array([1.5, 1.5, 1.5, 5. / 6, 5. / 6])

    def fermion_massSq(self, X):
        X = np.asanyarray(X)
        h = X[..., 0]
        M = np.array([.5 * self.P["yt2"] * h * h])
        M = np.rollaxis(M, 0, len(M.shape))
        return M, np.array([12.])


def gen_xsm():
    from scipy import optimize
    m = XSM()
    rng = np.random.default_rng(88)
    X = np.concatenate([rng.uniform(-350, 350, (300, 2)), [[0., 0.], [246., 0.], [0., 141.4], [1e-3, -1e-3]]])
    T = np.concatenate([rng.uniform(0, 250, 300), [0., 50., 120., 200.]])
    out = dict(X=X, T=T, V=m.Vtot(X, T), params=np.array([XSM.P[k] for k in ("lh", "ls", "lhs", "mus2", "g2", "gz2", "yt2", "v2")]))
    sel = slice(0, 60)
    out["grad"] = m.gradV(X[sel], T[sel])
    out["hess"] = m.d2V(X[sel], T[sel])
    # minima: the electroweak vacuum (h, 0) and the singlet vacuum (0, s) followed in T by the reference's own fmin
    Tm = np.linspace(0., 150., 31)
    mins = {}
    for name, seed in (("ew", np.array([246., 0.])), ("singlet", np.array([0., 141.4]))):
        x = seed.copy()
        rows = []
        for t in Tm:
            x = optimize.fmin(m.Vtot, x, args=(t,), xtol=1e-9, ftol=1e-13, maxiter=4000, maxfun=8000, disp=0)
            xn = np.array(x, float)
            for _ in range(4):
                H = m.d2V(xn, t)
                if np.linalg.eigvalsh(H)[0] <= 0:
                    break
                xn = xn - np.linalg.solve(H, m.gradV(xn, t))
            ok = np.linalg.eigvalsh(m.d2V(xn, t))[0] > 0 and np.linalg.norm(xn - x) < 1.0
            rows.append(xn if ok else [np.nan, np.nan])
            if not ok:
                break
            x = xn
        rows += [[np.nan, np.nan]] * (len(Tm) - len(rows))
        mins[name] = np.array(rows, float)
        print("xsm", name, "followed to T =", Tm[np.isfinite(mins[name][:, 0])][-1], mins[name][0], mins[name][np.isfinite(mins[name][:, 0])][-1])
    out.update(Tm=Tm, min_ew=mins["ew"], min_singlet=mins["singlet"])
    # bounces along the straight line between the two vacua (the 1-D solve of the reference on Vtot restricted to the
    # line, finite-difference dV / d2V): from the higher (metastable) to the lower one, wherever both exist
    res = []
    for k, t in enumerate(Tm):
        xe, xs = mins["ew"][k], mins["singlet"][k]
        if not (np.isfinite(xe[0]) and np.isfinite(xs[0])) or t == 0.0:
            continue
        ve, vs = float(m.Vtot(xe, t)), float(m.Vtot(xs, t))
        x_abs, x_meta = (xe, xs) if ve < vs else (xs, xe)
        L = float(np.linalg.norm(x_meta - x_abs))
        nh = (x_meta - x_abs) / L
        Vline = lambda phi, x_abs=x_abs, nh=nh, t=t: m.Vtot(x_abs + np.asarray(phi)[..., None] * nh, t)
        r = run_case(Vline, None, None, 0.0, L, 2)
        r.update(T=float(t), L=L, x_abs=[float(q) for q in x_abs], x_meta=[float(q) for q in x_meta])
        res.append(r)
        print("xsm bounce T=%g" % t, {q: r.get(q) for q in ("S", "npts", "n_shots", "n_rkck", "error")})
    out["bounces_json"] = json.dumps(res)
    np.savez(os.path.join(GOLD, "xsm.npz"), **out)


def gen_alpha_real():
    """Real-valued alpha (d = alpha + 1 dimensions, any real alpha >= 0): the reference's general-order Bessel start
    conditions (tunneling1D.py:336-372, scipy.special.iv / jv / gamma) -- exactSolution on random arguments, and full
    bounces of the quartic family."""
    alphas = [0.0, 0.5, 1.5, 2.5, 3.3, 5.0, 7.25]
    cases = []
    for alpha in alphas:
        for c in (0.05, 0.2, 0.35, 0.47):
            V, dV, d2V = quartic_lambdas(c)
            r = run_case(V, dV, d2V, 1.0, 0.0, alpha)
            r.update(kind="quartic", c=float(c), alpha=alpha, coef=[0, 0, c / 2, -(1 + c) / 3, 0.25], phi_abs=1.0,
                     phi_meta=0.0, opts={})
            cases.append(r)
            print("alpha_real", alpha, c, {q: r.get(q) for q in ("S", "npts", "n_shots", "n_rkck", "error")})
    rng = np.random.default_rng(1414)
    rows, vals = [], []
    for alpha in alphas + [1.0, 2.0, 3.0, 4.0]:
        o = t1.SingleFieldInstanton(1.0, 0.0, *quartic_lambdas(0.3), alpha=alpha)
        for k in range(300):
            r = 10 ** rng.uniform(-5, 2.7)
            phi0 = rng.uniform(-2, 2)
            dV = rng.normal() * 10 ** rng.uniform(-20, 1)
            d2V = rng.normal() * 10 ** rng.uniform(-3, 1)
            with np.errstate(all="ignore"):
                ph, dph = o.exactSolution(r, phi0, dV, d2V)
            rows.append([alpha, r, phi0, dV, d2V])
            vals.append([float(ph), float(dph)])
    with open(os.path.join(GOLD, "alpha_real.json"), "w") as f:
        json.dump(dict(cases=cases, exact_in=rows, exact_out=vals), f, indent=0)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["tables", "jspline", "jseries", "quartic", "profiles", "units", "model1", "model1f",
                               "thermal1f", "splinepath", "alpha14", "findaction", "jexact", "wall", "model1_derivs", "phases", "alpha_real", "xsm"]
    for w in which:
        print("generating", w)
        globals()["gen_" + w]()
    with open(os.path.join(GOLD, "VERSIONS.json"), "w") as f:
        json.dump(dict(numpy=np.__version__, scipy=scipy.__version__, python=sys.version.split()[0],
                       reference="clwainwright/CosmoTransitions setup.py version 2.0.7"), f)
