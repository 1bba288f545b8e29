"""The oracle restates four scipy routines the reference calls on the hot path.
scipy itself is installed (it is a third-party dependency of the reference, not the reference), so
the restatements are checked against the real thing on random inputs.  CPU only."""
import numpy as np
import pytest
from scipy import integrate, interpolate, optimize


def test_brentq_matches_scipy(oracle):
    rng = np.random.default_rng(0)
    n = 0
    for k in range(400):
        deg = int(rng.integers(1, 6))
        coefs = rng.normal(size=deg + 1)
        a, b = np.sort(rng.uniform(-3, 3, 2))
        fa, fb = np.polyval(coefs, a), np.polyval(coefs, b)
        if fa * fb >= 0:
            continue
        n += 1
        calls = [0]

        def f(x):
            calls[0] += 1
            y = coefs[0]
            for c in coefs[1:]:
                y = y * x + c
            return y
        x_ref, r = optimize.brentq(f, a, b, full_output=True, disp=False)
        x, err, nf = oracle.brentq_poly(coefs, a, b)
        assert err == 0
        assert x == x_ref, (x, x_ref)
        assert nf == r.function_calls
    assert n > 100


def test_brentq_sign_error(oracle):
    x, err, nf = oracle.brentq_poly([1.0, 0.0, 1.0], -1.0, 1.0)
    assert err == 1


def test_fminbound_matches_scipy(oracle):
    rng = np.random.default_rng(1)
    for k in range(300):
        deg = int(rng.integers(2, 6))
        coefs = rng.normal(size=deg + 1)
        a, b = np.sort(rng.uniform(-3, 3, 2))
        xatol = 10 ** rng.uniform(-9, -3) * (b - a)

        def f(x):
            y = coefs[0]
            for c in coefs[1:]:
                y = y * x + c
            return y
        x_ref, fval, ierr, numfunc = optimize.fminbound(f, a, b, xtol=xatol, full_output=True, disp=0)
        x, nf = oracle.fminbound_poly(coefs, a, b, xatol)
        assert x == x_ref
        assert nf == numfunc


@pytest.mark.parametrize("n", [2, 3, 4, 5, 10, 11, 500, 501, 750, 751])
def test_simpson_matches_scipy(oracle, n):
    rng = np.random.default_rng(n)
    x = np.cumsum(rng.uniform(0.01, 1.0, n))
    y = rng.normal(size=n) * np.exp(rng.normal(size=n))
    ref = integrate.simpson(y, x=x)
    got = oracle.simpson(y, x)
    assert abs(got - ref) <= 2e-14 * np.sum(np.abs(y)) * (x[-1] - x[0]) / n + 1e-300


def test_splev_matches_scipy(oracle, tables):
    rng = np.random.default_rng(2)
    for t, c, lo, hi in ((tables.tb, tables.cb, tables.tb[0], tables.tb[-1]),
                         (tables.tf, tables.cf, tables.tf[0], tables.tf[-1])):
        x = np.concatenate([rng.uniform(lo, hi, 3000), rng.uniform(-3, 3, 2000), t[3:-3:17], [lo, hi]])
        for der in range(4):
            ref = interpolate.splev(x, (t, c, 3), der=der)
            got = oracle.jspline(t, c, -1e300, 1e300, x, der)   # clamps disabled
            scale = np.maximum(1.0, np.abs(ref))
            assert np.max(np.abs(got - ref) / scale) < 1e-12, der
