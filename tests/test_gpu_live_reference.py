"""The CUDA path beside the UNMODIFIED reference, live (oracle/_ref: installed by oracle/make_ref.py in the build
container, travels to the GPU box with the snapshot; the product never imports it).

* the GPU class plugged into the reference's own multi-field driver, pathDeformation.fullTunneling(...,
  tunneling_class=...) (pathDeformation.py:845-1001), on examples/fullTunneling.py (BASELINE.json configs[0]);
* S of the batched GPU solve against the reference's SingleFieldInstanton on samples of C2 and C3;
* Jb / Jf spline and model1.Vtot against the reference's finiteT / generic_potential on fresh random inputs.
Needs a B200: pytest -m gpu."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref is not in this snapshot (python oracle/make_ref.py in the build container)")
    return ref_runner.import_reference()


@pytest.mark.parametrize("fy,S_expected", [(2.0, 1766.937676), (80.0, 4.503666131)])
def test_full_tunneling_with_gpu_tunneling_class(ref, fy, S_expected, capsys):
    """examples/fullTunneling.py:67-69,81-83 with tunneling_class = the GPU SingleFieldInstanton: every 1-D solve of
    the path deformation (SplinePath potentials) runs on the device; the final action must be the reference's.

    Conditioning.  The reference's own result is not a single number at the 1e-6 level: moving the start point of the
    initial path by 1e-14 makes it return 4.5036661 or 4.5036731 (fy = 80; 1766.93756 ... 1766.93768 for fy = 2) --
    findRScale's fminbound (xtol 1e-6) and the bisection in phi(0) (xtol 1e-4) take different discrete paths, and the
    path deformation (converged to fRatio ~ 2e-2 only) carries the difference into the next spline.  So the bar is:
    the reference reproduces SURVEY.md's value, and the GPU-class result lies within 1e-6 of the band the reference
    itself spans under such perturbations (and of one of its members)."""
    from cosmoTransitions.examples import fullTunneling as ft
    from cosmotransitions_b200.tunneling1D import SingleFieldInstanton as GpuSFI
    pd = ref["pd"]
    m = ft.Potential(c=5., fx=0., fy=fy)
    calls = []

    class Counting(GpuSFI):
        def findProfile(self, **kw):
            calls.append(1)
            return GpuSFI.findProfile(self, **kw)

    Y = pd.fullTunneling([[1, 1.], [0, 0]], m.V, m.dV, tunneling_class=Counting)
    assert len(calls) >= 2
    refs = [pd.fullTunneling([[1 + eps, 1.], [0, 0]], m.V, m.dV) for eps in (0, 1e-14, -1e-14, 1e-13, -1e-13, 1e-12, 1e-11)]
    Yref = refs[0]
    assert abs(Yref.action - S_expected) < 5e-7 * S_expected
    acts = np.array([float(r.action) for r in refs])
    rel_band = (acts.max() - acts.min()) / acts.min()
    nearest = np.abs(acts - Y.action).min() / acts.min()
    with capsys.disabled():
        print("\nfullTunneling fy=%g: GPU class S = %.10f; reference S = %.10f, its band under 1e-14..1e-11 perturbations "
              "[%.10f, %.10f] (%.1e rel.); distance to the nearest member %.1e"
              % (fy, Y.action, Yref.action, acts.min(), acts.max(), rel_band, nearest))
    assert acts.min() * (1 - 1e-6) <= Y.action <= acts.max() * (1 + 1e-6), (Y.action, acts)
    assert nearest < 1e-6
    assert len(Y.profile1D.R) == len(Yref.profile1D.R)
    j = int(np.abs(acts - Y.action).argmin())
    # (the radial grids differ with r_f, which moves by ~6e-4 between the members of the band)
    np.testing.assert_allclose(Y.profile1D.Phi, refs[j].profile1D.Phi, rtol=0, atol=2e-3 * abs(Yref.profile1D.Phi).max())
    assert np.abs(Y.Phi - refs[j].Phi).max() < 2e-3 * np.abs(Yref.Phi).max()


def _assert_parity_or_knife_edge(ref_runner, c, alpha, S, S_ref, rtol=1e-6, max_outliers=3):
    """The 1e-6 bar of BASELINE.json -- except on knife-edge instances, where the reference's OWN answer jumps by more
    than that when c moves in its 15th digit (a shot's converged / overshoot classification flips, which adds or drops
    one bisection step: measured 5.5e-6 at c = 0.35455224, alpha = 3).  There the GPU value has to lie inside the band
    the reference itself spans under relative perturbations of c up to 1e-13."""
    rel = np.abs(S - S_ref) / np.abs(S_ref)
    bad = np.nonzero(~(rel < rtol))[0]
    assert len(bad) <= max_outliers, (len(bad), rel[bad])
    for i in bad:
        lo, hi = ref_runner.quartic_band(float(c[i]), alpha)
        assert lo * (1 - rtol) <= S[i] <= hi * (1 + rtol), (c[i], S[i], S_ref[i], lo, hi)
        assert (hi - lo) > rtol * abs(S_ref[i])        # it IS a knife edge of the reference, not a GPU error


def test_knife_edge_instance_of_the_reference(ref):
    """c = 0.35455224 (C3 index 704320), alpha = 3: the reference returns 1670.5463 or 1670.5554 depending on the 15th
    digit of c; the GPU returns one of the two."""
    from oracle import ref_runner
    from cosmotransitions_b200 import _lib, batch, potentials
    c = 0.02 + 0.475 * (4402 * 160 + 0.5) / 1000000
    lo, hi = ref_runner.quartic_band(c, 3)
    assert (hi - lo) / lo > 4e-6
    S = float(batch.find_bounce_batch(_lib.POT_POLY4, potentials.quartic_family_params([c]), 1.0, 0.0, alpha=3)["S"][0])
    assert min(abs(S - lo), abs(S - hi)) < 1e-8 * S, (S, lo, hi)


@pytest.mark.parametrize("alpha,n", [(2, 100000), (3, 1000000)])
def test_batched_solve_vs_live_reference(ref, alpha, n):
    """~300 instances of C2 / C3 spread over the thin -> thick range: S rel. 1e-6 (BASELINE.json), live."""
    from oracle import ref_runner
    from cosmotransitions_b200 import _lib, batch, potentials
    idx = np.arange(11, n, n // 300)
    c = 0.02 + 0.475 * (idx + 0.5) / n
    S_ref, status, _ = ref_runner.quartic_pool(c, alpha, chunksize=5)
    assert set(status) == {"ok"}
    g = batch.find_bounce_batch(_lib.POT_POLY4, potentials.quartic_family_params(c), 1.0, 0.0, alpha=alpha)
    assert (g["status"].numpy() <= 1).all()
    # and through the thread-per-instance kernels the big batches use
    g2 = batch.find_bounce_batch(_lib.POT_POLY4, potentials.quartic_family_params(c), 1.0, 0.0, alpha=alpha,
                                 kernel="thread", schedule="sorted", split=True)
    for S in (g["S"].numpy(), g2["S"].numpy()):
        rel = np.abs(S - S_ref) / np.abs(S_ref)
        print("alpha=%d: max rel err in S vs the live reference: %.2e; above 1e-6: %d" % (alpha, rel.max(), (rel >= 1e-6).sum()))
        _assert_parity_or_knife_edge(ref_runner, c, alpha, S, S_ref)


def test_docstring_example_of_the_reference(ref):
    """tunneling1D.py:112-122: the thin- and thick-walled quartics, scalar API on both sides."""
    from cosmotransitions_b200 import potentials, tunneling1D as g1
    t1 = ref["t1"]
    for c in (0.47, 0.2):       # V = phi^4/4 - (1+c)/3 phi^3 + c/2 phi^2: .49 / .235 and .4 / .1
        a, b = (1 + c) / 3, c / 2
        # (analytic d2V on both sides; without it the reference differentiates V numerically, tunneling1D.py:191-201)
        inst = t1.SingleFieldInstanton(1.0, 0.0, lambda p: .25 * p ** 4 - a * p ** 3 + b * p ** 2,
                                       lambda p: p * (p - c) * (p - 1.0), lambda p: 3 * p * p - 2 * (1 + c) * p + c)
        prof = inst.findProfile()
        S_ref = inst.findAction(prof)
        ginst = g1.SingleFieldInstanton(1.0, 0.0, potentials.QuarticPotential(c))
        gprof = ginst.findProfile()
        assert abs(ginst.findAction(gprof) - S_ref) < 1e-9 * S_ref
        assert len(gprof.R) == len(prof.R)
        np.testing.assert_allclose(gprof.R, prof.R, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(gprof.Phi, prof.Phi, rtol=0, atol=1e-9)


def test_thermal_functions_vs_live_reference(ref):
    """finiteT.Jb_spline / Jf_spline (derivatives 0..3) and model1.Vtot on fresh random inputs (not the goldens)."""
    from cosmotransitions_b200 import finiteT as gfT, generic_potential as ggp
    fT = ref["fT"]
    rng = np.random.default_rng(77)
    th = np.concatenate([rng.uniform(-8, 1500, 20000), rng.normal(0, 2, 5000)])
    for n in range(4):
        assert np.abs(gfT.Jb_spline(th, n=n) - fT.Jb_spline(th, n=n)).max() < 1e-10
        assert np.abs(gfT.Jf_spline(th, n=n) - fT.Jf_spline(th, n=n)).max() < 1e-10
    from cosmoTransitions.examples import testModel1 as tm
    m_ref, m_gpu = tm.model1(), ggp.model1()
    X = rng.uniform(-350, 350, (3000, 2))
    T = rng.uniform(0, 250, 3000)
    v_ref, v_gpu = m_ref.Vtot(X, T), m_gpu.Vtot(X, T)
    assert np.max(np.abs(v_gpu - v_ref) / np.maximum(1.0, np.abs(v_ref))) < 1e-11


def test_full_tunneling_without_path_spline(ref):
    """pathDeformation.fullTunneling(..., V_spline_samples=None) hands the tunneling class the raw path.V / path.dV
    callables instead of a spline (pathDeformation.py:959-964); the GPU class samples them onto its tabulated functor.

    (i) One 1-D solve along the initial straight path (SplinePath with extend_to_minima, as fullTunneling builds it):
    the GPU class on the sampled callables against the reference's class on the callables themselves.  Measured with
    the reference's own solver on the same 255-piece table: -5e-7 relative (its bisection noise at xtol = 1e-4).
    (ii) The whole deformation loop.  Its stopping rule (fRatio < 2e-2) makes the final action depend on the 1-D
    profiles at the 1e-4 level -- the reference's own two branches (V_spline_samples = 100 vs None) differ by 4e-5 -- so
    the bar there is 5e-4."""
    from cosmoTransitions.examples import fullTunneling as ft
    from cosmotransitions_b200 import potentials
    from cosmotransitions_b200.tunneling1D import SingleFieldInstanton as GpuSFI
    pd, t1 = ref["pd"], ref["t1"]
    m = ft.Potential(c=5., fx=0., fy=2.)
    path = pd.SplinePath(np.array([[1, 1.], [0, 0]]), m.V, m.dV, V_spline_samples=None, extend_to_minima=True)
    robj = t1.SingleFieldInstanton(0.0, path.L, path.V, path.dV, None)
    S_ref = robj.findAction(robj.findProfile())
    gobj = GpuSFI(0.0, path.L, path.V, path.dV, None)
    assert isinstance(gobj.V, potentials.SampledPotential)
    S_gpu = gobj.findAction(gobj.findProfile())
    print("1-D solve on raw path callables: GPU (sampled) S = %.8f, reference S = %.8f, rel %.1e"
          % (S_gpu, S_ref, abs(S_gpu - S_ref) / S_ref))
    assert abs(S_gpu - S_ref) < 3e-6 * S_ref
    Yref = pd.fullTunneling([[1, 1.], [0, 0]], m.V, m.dV, V_spline_samples=None)
    Y = pd.fullTunneling([[1, 1.], [0, 0]], m.V, m.dV, V_spline_samples=None, tunneling_class=GpuSFI)
    print("fullTunneling(V_spline_samples=None): GPU class (sampled callables) S = %.8f, reference S = %.8f"
          % (Y.action, Yref.action))
    assert abs(Yref.action - 1766.86) < 0.05
    assert abs(Y.action - Yref.action) < 5e-4 * Yref.action, (Y.action, Yref.action)
