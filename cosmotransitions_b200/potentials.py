"""Device-functor potentials: the objects that stand in for the Python callables V, dV, d2V of
SingleFieldInstanton(phi_absMin, phi_metaMin, V, dV, d2V) (tunneling1D.py:128-130).

A functor names a potential the CUDA kernels know how to evaluate (include/ctbounce.h CTB_POT_*)
plus one row of parameters.  Calling it evaluates V(phi) ON THE GPU (ctb_potential)."""
import numpy as np

from . import _lib, _tables


class DevicePotential:
    kind = None

    def __init__(self, params):
        self.params = np.ascontiguousarray(params, dtype=np.float64).reshape(_lib.NPARAM[self.kind])

    needs_tables = False

    def __call__(self, phi):
        torch = _lib.require_cuda()
        lib = _lib.load()
        dev = torch.device("cuda", torch.cuda.current_device())
        x = np.asarray(phi, dtype=np.float64)
        d_phi = torch.from_numpy(np.ascontiguousarray(x.reshape(-1))).to(dev)
        d_par = torch.from_numpy(self.params).to(dev)
        out = torch.empty_like(d_phi)
        tabs = _tables.device_tables(dev) if self.needs_tables else None
        jb = tabs.jb if tabs else None
        jf = tabs.jf if tabs else None
        _lib.check(lib.ctb_potential(self.kind, d_par.data_ptr(), jb, jf, d_phi.data_ptr(), out.data_ptr(),
                                     d_phi.numel(), torch.cuda.current_stream().cuda_stream), "ctb_potential")
        res = out.cpu().numpy().reshape(x.shape)
        return float(res) if res.ndim == 0 else res


class Poly4Potential(DevicePotential):
    """V = c0 + c1 phi + c2 phi^2 + c3 phi^3 + c4 phi^4 with analytic dV, d2V."""
    kind = _lib.POT_POLY4

    def __init__(self, c0=0.0, c1=0.0, c2=0.0, c3=0.0, c4=0.0):
        super().__init__([c0, c1, c2, c3, c4])


def quartic_family_params(c):
    """Rows [c0..c4] of V = phi^4/4 - (1+c)/3 phi^3 + c/2 phi^2, i.e. dV = phi (phi-c) (phi-1)
    (the family of the reference's docstring examples, tunneling1D.py:112-122; minima at 0 and 1)."""
    c = np.atleast_1d(np.asarray(c, dtype=np.float64))
    p = np.zeros((c.size, 5))
    p[:, 2] = c / 2
    p[:, 3] = -(1 + c) / 3
    p[:, 4] = 0.25
    return p


class QuarticPotential(Poly4Potential):
    def __init__(self, c):
        DevicePotential.__init__(self, quartic_family_params(c)[0])


class SplinePotential(DevicePotential):
    """Tabulated potential: the cubic B-spline (t, c, k) that scipy.interpolate.splrep returns -- what
    pathDeformation.SplinePath builds from V_spline_samples samples and hands to the tunneling class as
    V, dV, d2V = splev(x, tck, der=0|1|2) (pathDeformation.py:801-834).  Converted once, on the host, to
    one cubic per knot interval."""
    kind = _lib.POT_SPLINE

    def __init__(self, tck):
        t, c, k = tck
        if k > 3:
            raise ValueError("spline degree must be <= 3")
        brk, coef = _tables.bspline_to_pp(np.asarray(t, float), np.asarray(c, float), int(k))
        nint = coef.shape[0]
        if nint > _lib.SPLINE_MAX_PIECES:
            raise ValueError("spline has %d pieces; at most %d are supported" % (nint, _lib.SPLINE_MAX_PIECES))
        row = np.zeros(_lib.NPARAM[self.kind])
        row[0] = nint
        row[1:1 + nint + 1] = brk
        c4 = np.zeros((nint, 4))
        c4[:, :coef.shape[1]] = coef
        off = 1 + _lib.SPLINE_MAX_PIECES + 1
        row[off:off + 4 * nint] = c4.reshape(-1)
        self.tck = tck
        super().__init__(row)

    @classmethod
    def from_callable(cls, V):
        """If V is a bound method of an object that carries a `_V_tck` spline (the reference's
        SplinePath.V / .dV / .d2V), wrap that spline; otherwise None."""
        tck = getattr(getattr(V, "__self__", None), "_V_tck", None)
        return cls(tck) if tck is not None else None


class SampledPotential(SplinePotential):
    """Arbitrary Python callables V (and dV) as a device functor: what the reference accepts as
    SingleFieldInstanton(phi_absMin, phi_metaMin, V, dV, d2V) (tunneling1D.py:112-149) and what
    pathDeformation.fullTunneling passes when V_spline_samples is None (raw path.V, path.dV, pathDeformation.py:
    959-964).  A Python function cannot run inside a CUDA kernel, so it is SAMPLED ONCE on the host -- `pieces` + 1
    equidistant points covering [phi_lo, phi_hi] -- and handed to the tabulated functor (CTB_POT_SPLINE) as a piecewise
    cubic: the Hermite interpolant of (V, dV) when dV is given (C1, exact V and dV at the knots, error
    h^4 max|d4V| / 384 between them), otherwise scipy's not-a-knot C2 spline of the V samples.  V, dV and d2V inside
    the solver are that cubic and its derivatives: SPLINE-ACCURATE, NOT TRAJECTORY-EXACT -- the action converges like
    pieces^-4, but step and shot counts may differ from a run on the callable itself.  The integration itself
    (everything after the sampling) runs on the GPU."""

    def __init__(self, V, phi_lo, phi_hi, dV=None, pieces=None):
        from scipy import interpolate
        pieces = _lib.SPLINE_MAX_PIECES if pieces is None else int(pieces)
        if not (4 <= pieces <= _lib.SPLINE_MAX_PIECES):
            raise ValueError("pieces must be in [4, %d]" % _lib.SPLINE_MAX_PIECES)
        x = np.linspace(float(phi_lo), float(phi_hi), pieces + 1)

        def sample(f):
            try:
                y = np.asarray(f(x), dtype=np.float64)
                if y.shape == x.shape:
                    return y
            except Exception:
                pass
            return np.array([float(f(xi)) for xi in x], dtype=np.float64)

        v = sample(V)
        if not np.isfinite(v).all():
            raise ValueError("V is not finite on [%g, %g]" % (phi_lo, phi_hi))
        pp = (interpolate.CubicHermiteSpline(x, v, sample(dV)) if dV is not None
              else interpolate.CubicSpline(x, v, bc_type="not-a-knot"))
        row = np.zeros(_lib.NPARAM[self.kind])
        row[0] = pieces
        row[1:1 + pieces + 1] = pp.x
        off = 1 + _lib.SPLINE_MAX_PIECES + 1
        row[off:off + 4 * pieces] = pp.c[::-1, :].T.reshape(-1)      # c0..c3 per piece (PPoly stores descending powers)
        self.tck = None
        self.pieces = pieces
        DevicePotential.__init__(self, row)

    @classmethod
    def for_instanton(cls, V, dV, phi_absMin, phi_metaMin, pieces=None, pad=0.05):
        """Sampling window of a bounce between the two minima: the interval between them plus `pad` of its length on
        either side (the 5-point stencils and an overshooting trajectory step slightly past the ends; farther out the
        end pieces extrapolate)."""
        lo, hi = min(phi_absMin, phi_metaMin), max(phi_absMin, phi_metaMin)
        d = hi - lo
        return cls(V, lo - pad * d, hi + pad * d, dV=dV, pieces=pieces)


def thermal1f_params(V0_coefs, renormScaleSq, T, bosons=(), fermions=()):
    """Rows of CTB_POT_THERMAL1F (include/ctbounce.h): a one-field generic_potential with
    V0 = sum_k V0_coefs[k] phi^k (k <= 4), bosons = [(a, b, t, dof, c), ...] with m^2 = a + b phi^2 + t T^2,
    fermions = [(a, b, dof), ...] with m^2 = a + b phi^2.  T may be an array: one row per temperature."""
    T = np.atleast_1d(np.asarray(T, dtype=np.float64))
    if len(bosons) > 8 or len(fermions) > 4:
        raise ValueError("at most 8 bosons and 4 fermions")
    row = np.zeros(_lib.NPARAM[_lib.POT_THERMAL1F])
    c = np.zeros(5)
    c[:len(V0_coefs)] = V0_coefs
    row[0:5] = c
    row[5], row[7], row[8] = renormScaleSq, len(bosons), len(fermions)
    for i, b in enumerate(bosons):
        row[9 + 5 * i:14 + 5 * i] = b
    for j, f in enumerate(fermions):
        row[49 + 3 * j:52 + 3 * j] = f
    rows = np.tile(row, (T.size, 1))
    rows[:, 6] = T
    return rows


class Thermal1FPotential(DevicePotential):
    """General one-field thermal potential (see thermal1f_params) at one temperature."""
    kind = _lib.POT_THERMAL1F
    needs_tables = True

    def __init__(self, V0_coefs, renormScaleSq, T, bosons=(), fermions=()):
        super().__init__(thermal1f_params(V0_coefs, renormScaleSq, T, bosons, fermions)[0])


def model1_model8(m1=120., m2=50., mu=25., Y1=.1, Y2=.15, n=30, v2=246. ** 2):
    """(l1, l2, mu2, Y1, Y2, n, renormScaleSq, v2) of examples/testModel1.py:19-47."""
    return np.array([.5 * m1 ** 2 / v2, .5 * m2 ** 2 / v2, mu ** 2, Y1, Y2, float(n), v2, v2])


class Model1LinePotential(DevicePotential):
    """model1.Vtot along X(phi) = x0 + phi * nhat at temperature T; derivatives by finite differences."""
    kind = _lib.POT_MODEL1_LINE
    needs_tables = True

    def __init__(self, x0, nhat, T, model8=None):
        m8 = model1_model8() if model8 is None else np.asarray(model8, dtype=np.float64)
        super().__init__(np.concatenate([m8, [T], np.asarray(x0, float), np.asarray(nhat, float)]))


class Model1FPotential(DevicePotential):
    """One-field model1-style thermal potential (see include/ctbounce.h CTB_POT_MODEL1F)."""
    kind = _lib.POT_MODEL1F
    needs_tables = True

    def __init__(self, lam, Y, n, T, v2=246. ** 2, renormScaleSq=None):
        super().__init__([lam, v2, Y, n, v2 if renormScaleSq is None else renormScaleSq, T])
