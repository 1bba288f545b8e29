"""Run-time compiled model functors (cosmotransitions_b200/jit.py): a new multi-field model as a CUDA string.

* model1 written as a compiled model against the built-in Model1XY kernels (same physics, independent code path);
* a second model the library has never seen -- Higgs + Z2 singlet with mass-matrix eigenvalues, Goldstones, W, Z and a
  top quark -- against goldens from the reference's own generic_potential subclass (oracle/gen_golden.py gen_xsm):
  Vtot, gradV, d2V, and both vacua followed in temperature by the reference's fmin;
* bounces along lines between minima through the sampled tabulated functor against the exact line functor.
Needs a B200 and nvcc (both are in the image): pytest -m gpu."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

XSM_BODY = '''
const double h = X[0], s = X[1];
const double muh2 = lh * v2;
const double a = -muh2 + 3 * lh * h * h + lhs * s * s, b = -mus2 + 3 * ls * s * s + lhs * h * h, c = 2 * lhs * h * s;
const double A = .5 * (a + b), B = sqrt(.25 * (a - b) * (a - b) + c * c);
const double gold = -muh2 + lh * h * h + lhs * s * s, mW = .25 * g2 * h * h, mZ = .25 * gz2 * h * h, mt = .5 * yt2 * h * h;
double v = -.5 * muh2 * h * h + .25 * lh * sq(h * h) - .5 * mus2 * s * s + .25 * ls * sq(s * s) + .5 * lhs * h * h * s * s;
v += (CW(A + B, 1.5, v2) + CW(A - B, 1.5, v2) + 3 * CW(gold, 1.5, v2) + 6 * CW(mW, 5. / 6, v2) + 3 * CW(mZ, 5. / 6, v2)
      - 12 * CW(mt, 1.5, v2)) / (64 * PI * PI);
v += (Jb((A + B) / T2) + Jb((A - B) / T2) + 3 * Jb(gold / T2) + 6 * Jb(mW / T2) + 3 * Jb(mZ / T2) + 12 * Jf(mt / T2))
     * T4 / (2 * PI * PI);
return v;
'''
XSM_PARAMS = ("lh", "ls", "lhs", "mus2", "g2", "gz2", "yt2", "v2")


@pytest.fixture(scope="module")
def jit():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from cosmotransitions_b200 import jit as J
    return J


def test_compiled_model1_matches_builtin(jit, gold_dir):
    from cosmotransitions_b200 import _lib, generic_potential as gp, phases
    m = jit.model1_compiled()
    builtin = gp.model1()
    cm = gp.compiled_model(m, builtin.model8)
    rng = np.random.default_rng(4)
    X, T = rng.uniform(-400, 400, (5000, 2)), rng.uniform(0, 300, 5000)
    np.testing.assert_allclose(cm.Vtot(X, T), builtin.Vtot(X, T), rtol=1e-12, atol=1e-3)
    # broadcasting contract of the reference: X[..., 2] against T[...]
    assert cm.Vtot(X[:7, None, :], T[None, :5]).shape == (7, 5)
    g = np.load(os.path.join(gold_dir, "phases.npz"))
    v = np.sqrt(g["model8"][:, 7])
    a = phases.trace_minima(g["model8"], np.stack([v, v], -1), g["T"])
    b = phases.trace_minima(g["model8"], np.stack([v, v], -1), g["T"], model=m)
    assert np.array_equal(a["flag"].cpu().numpy(), b["flag"].cpu().numpy())
    ok = a["flag"].cpu().numpy() == _lib.MIN_OK
    np.testing.assert_allclose(b["X"].cpu().numpy()[ok], a["X"].cpu().numpy()[ok], rtol=0, atol=1e-6)


def test_new_model_vs_reference_subclass(jit, gold_dir):
    """The xSM-like model exists only as a CUDA string here and as a generic_potential subclass in the golden generator."""
    from cosmotransitions_b200 import _lib, generic_potential as gp, phases
    g = np.load(os.path.join(gold_dir, "xsm.npz"))
    model = jit.CompiledModel(2, XSM_PARAMS, XSM_BODY)
    assert model.uses_fermions
    cm = gp.compiled_model(model, g["params"])
    V = cm.Vtot(g["X"], g["T"])
    np.testing.assert_allclose(V, g["V"], rtol=1e-11, atol=1e-11 * np.abs(g["V"]).max())
    X, T = g["X"][:60], g["T"][:60]
    scale_g, scale_h = np.abs(g["grad"]).max(), np.abs(g["hess"]).max()
    np.testing.assert_allclose(cm.gradV(X, T), g["grad"], rtol=1e-6, atol=1e-7 * scale_g)
    # (second differences with step 1e-3 of a potential of size 1e10 carry rounding noise eps |V| / h^2 ~ 1..10 on both sides)
    np.testing.assert_allclose(cm.d2V(X, T), g["hess"], rtol=1e-4, atol=2e-4 * scale_h)
    for name, seed in (("ew", [246.0, 0.0]), ("singlet", [0.0, 141.4])):
        ref = g["min_" + name]
        have = np.isfinite(ref[:, 0])
        r = phases.trace_minima(g["params"], np.array([seed]), g["Tm"], model=model)
        Xd, flag = r["X"].cpu().numpy()[0], r["flag"].cpu().numpy()[0]
        n_ok = int(have.sum())
        assert (flag[:n_ok - 1] == _lib.MIN_OK).all(), (name, flag)
        err = np.linalg.norm(Xd[:n_ok - 1] - ref[:n_ok - 1], axis=-1) / (np.linalg.norm(ref[:n_ok - 1], axis=-1) + 1.0)
        assert err.max() < 2e-5, (name, err.max())
    # findMinimum through the generic_potential-style class
    Xm = cm.findMinimum(np.array([[240.0, 5.0], [3.0, 150.0]]), np.array([40.0, 40.0]))
    k = int(np.argmin(np.abs(g["Tm"] - 40.0)))
    np.testing.assert_allclose(np.abs(Xm[0]), np.abs(g["min_ew"][k]), rtol=0, atol=2e-5 * 250)
    np.testing.assert_allclose(np.abs(Xm[1]), np.abs(g["min_singlet"][k]), rtol=0, atol=2e-5 * 250)


def test_line_bounces_through_the_sampled_functor(jit, gold_dir):
    """phases.line_bounces (model-agnostic: the line potential is sampled by the compiled kernel) against the exact
    CTB_POT_MODEL1_LINE functor on the same lines, and the generic phase_scan against model1_phase_scan."""
    from cosmotransitions_b200 import phases
    g = np.load(os.path.join(gold_dir, "phases.npz"))
    m = jit.model1_compiled()
    m8, T = g["model8"][:4], np.linspace(70.0, 120.0, 26)
    ref = phases.model1_phase_scan(m8, T)
    v = np.sqrt(m8[:, 7])
    gen = phases.phase_scan(m, m8, T, np.stack([v, v], -1), method="sampled")
    st_r, st_g = ref["status"].cpu().numpy(), gen["status"].cpu().numpy()
    both = ((st_r == 0) | (st_r == 1)) & ((st_g == 0) | (st_g == 1))
    assert both.sum() >= 0.9 * ((st_r == 0) | (st_r == 1)).sum() and both.sum() > 20
    Sr, Sg = ref["S"].cpu().numpy()[both], gen["S"].cpu().numpy()[both]
    rel = np.abs(Sg - Sr) / np.abs(Sr)
    print("sampled-line bounces vs the exact line functor: max rel %.2e, median %.2e over %d bounces" % (rel.max(), np.median(rel), both.sum()))
    # (the solver's own bisection noise at the default xtol is ~1e-5; thin walls near T_crit are the most sensitive)
    assert np.median(rel) < 3e-5 and rel.max() < 1e-3
    np.testing.assert_allclose(gen["Tcrit"].cpu().numpy(), ref["Tcrit"].cpu().numpy(), rtol=0, atol=1e-6)
    tn_r, tn_g = ref["Tnuc"].cpu().numpy(), gen["Tnuc"].cpu().numpy()
    okn = np.isfinite(tn_r) & np.isfinite(tn_g)
    assert okn.any() and np.abs(tn_r[okn] - tn_g[okn]).max() < 0.05


def test_jit_errors_are_loud(jit):
    from cosmotransitions_b200 import _lib
    with pytest.raises(_lib.CtbError):
        jit.CompiledModel(2, ("a",), "return undefined_symbol * X[0];")
    with pytest.raises(ValueError):
        jit.CompiledModel(7, (), "return 0.0;")


def test_three_field_model_stencils_and_minima(jit):
    """ndim = 3: the Hessian stencil has 61 points, i.e. two passes of the warp -- checked on a polynomial model whose
    gradient / Hessian are known in closed form and whose minima scipy finds on the host from the same expression."""
    from scipy import optimize
    from cosmotransitions_b200 import _lib, generic_potential as gp, phases
    body = '''
    double v = 0.0;
    const double vv[3] = {v0, v1, v2};
    for (int i = 0; i < 3; i++) v += 0.25 * lam * sq(X[i] * X[i] - vv[i] * vv[i]) + 0.5 * cT * T * T * X[i] * X[i];
    return v + g * X[0] * X[1] * X[2];
    '''
    model = jit.CompiledModel(3, ("lam", "cT", "g", "v0", "v1", "v2"), body)
    par = np.array([0.2, 0.3, 0.05, 100.0, 140.0, 60.0])
    cm = gp.compiled_model(model, par)

    def V(x, T):
        x = np.asarray(x, float)
        vv = par[3:]
        return float((0.25 * par[0] * (x * x - vv * vv) ** 2 + 0.5 * par[1] * T * T * x * x).sum() + par[2] * x[0] * x[1] * x[2])

    rng = np.random.default_rng(9)
    X, T = rng.uniform(-150, 150, (64, 3)), rng.uniform(0, 60, 64)
    np.testing.assert_allclose(cm.Vtot(X, T), [V(x, t) for x, t in zip(X, T)], rtol=1e-13)
    vv = par[3:]
    grad = par[0] * X * (X * X - vv * vv) + par[1] * (T * T)[:, None] * X + par[2] * np.stack(
        [X[:, 1] * X[:, 2], X[:, 0] * X[:, 2], X[:, 0] * X[:, 1]], -1)
    np.testing.assert_allclose(cm.gradV(X, T), grad, rtol=1e-7, atol=1e-4)
    H = cm.d2V(X, T)
    for i in range(3):
        np.testing.assert_allclose(H[:, i, i], par[0] * (3 * X[:, i] ** 2 - vv[i] ** 2) + par[1] * T * T, rtol=1e-5, atol=0.5)
    np.testing.assert_allclose(H[:, 0, 1], par[2] * X[:, 2], rtol=1e-4, atol=0.5)
    np.testing.assert_allclose(H[:, 1, 2], par[2] * X[:, 0], rtol=1e-4, atol=0.5)
    # minima: 40 random starts near the eight vacua, at T = 10 and 40
    starts = np.sign(rng.uniform(-1, 1, (40, 3))) * vv * rng.uniform(0.8, 1.2, (40, 3))
    Tm = np.where(np.arange(40) % 2 == 0, 10.0, 40.0)
    r = phases.find_minimum(par, starts, Tm, model=model)
    assert (r["flag"].cpu().numpy() == _lib.MIN_OK).all()
    Xd = r["X"].cpu().numpy()
    for k in range(0, 40, 5):
        ref = optimize.minimize(V, starts[k], args=(Tm[k],), method="BFGS", options=dict(gtol=1e-10)).x
        ref = optimize.minimize(V, ref, args=(Tm[k],), method="Nelder-Mead", options=dict(xatol=1e-9, fatol=1e-14)).x
        assert np.linalg.norm(Xd[k] - ref) < 2e-5 * (np.linalg.norm(ref) + 1.0), (k, Xd[k], ref)


def test_compiled_line_functor_bounces_vs_reference(jit, gold_dir):
    """The bounce kernels instantiated for a compiled model (CompiledModel.line_functor): 25 finite-T bounces of the
    Higgs + singlet model along the line between its two vacua against the reference's SingleFieldInstanton on the same
    lines (finite-difference dV / d2V on both sides), and against the exact built-in functor for model1."""
    import json
    from cosmotransitions_b200 import _lib, batch, phases
    g = np.load(os.path.join(gold_dir, "xsm.npz"))
    cases = json.loads(str(g["bounces_json"]))
    model = jit.CompiledModel(2, XSM_PARAMS, XSM_BODY)
    f = model.line_functor()
    assert f.nparam == 8 + 1 + 4
    rows, L = [], []
    for c in cases:
        xa, xm = np.array(c["x_abs"]), np.array(c["x_meta"])
        rows.append(np.concatenate([g["params"], [c["T"]], xa, (xm - xa) / c["L"]]))
        L.append(c["L"])
    rows, L = np.array(rows), np.array(L)
    for kernel in ("warp", "thread"):
        r = batch.find_bounce_batch(f, rows, 0.0, L, alpha=2, kernel=kernel)
        S, st = r["S"].numpy(), r["status"].numpy()
        Sref = np.array([c["S"] for c in cases])
        assert (st <= 1).all(), st
        rel = np.abs(S - Sref) / Sref
        print("compiled line functor (%s kernel) vs the reference, 25 bounces: max rel %.2e, median %.2e" % (kernel, rel.max(), np.median(rel)))
        assert (rel < 1e-6).all(), rel            # BASELINE.json's bar; measured: 3.4e-10 max, 9e-12 median
        assert (r["n_shots"].numpy() == np.array([c["n_shots"] for c in cases])).all()
        assert (r["n_points"].numpy() == np.array([c["npts"] for c in cases])).all()
    # model1: compiled functor vs the built-in exact line functor on the lines of a phase scan
    gph = np.load(os.path.join(gold_dir, "phases.npz"))
    m1 = jit.model1_compiled()
    m8, T = gph["model8"][:3], np.linspace(75.0, 115.0, 17)
    ref = phases.model1_phase_scan(m8, T)
    v = np.sqrt(m8[:, 7])
    gen = phases.phase_scan(m1, m8, T, np.stack([v, v], -1), seed_B=np.stack([v, -v], -1))   # method="exact" is the default
    st_r, st_g = ref["status"].cpu().numpy(), gen["status"].cpu().numpy()
    assert np.array_equal((st_r == 0) | (st_r == 1), (st_g == 0) | (st_g == 1))
    ok = (st_r == 0) | (st_r == 1)
    rel = np.abs(gen["S"].cpu().numpy()[ok] - ref["S"].cpu().numpy()[ok]) / ref["S"].cpu().numpy()[ok]
    print("compiled vs built-in line functor on %d model1 bounces: max rel %.2e" % (ok.sum(), rel.max()))
    assert rel.max() < 1e-7
