"""Real-valued alpha (d = alpha + 1 dimensions): the general-order Bessel start conditions of
tunneling1D.py:336-372 on the device against goldens from the live reference (tests/golden/alpha_real.json,
oracle/gen_golden.py gen_alpha_real).  Needs a B200: pytest -m gpu."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ERR2STATUS = {"PotentialError:stable, not metastable": 2, "PotentialError:no barrier": 3,
              "ValueError:f(a) and f(b) must have different signs": 8}


@pytest.fixture(scope="module")
def gold(gold_dir):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    with open(os.path.join(gold_dir, "alpha_real.json")) as f:
        return json.load(f)


def test_exact_solution_vs_reference(gold):
    """exactSolution(r, phi0, dV, d2V) for alpha in {0, 0.5, 1.5, 2.5, 3.3, 5, 7.25} (general order) and {1, 2, 3, 4}
    (the specialised forms the bounce kernels use): 300 random argument sets each, z = beta r from 1e-7 to 1e3."""
    from cosmotransitions_b200 import batch
    rows, ref = np.array(gold["exact_in"]), np.array(gold["exact_out"])
    for alpha in np.unique(rows[:, 0]):
        sel = rows[:, 0] == alpha
        a = rows[sel]
        phi, dphi = batch.exact_solution_batch(alpha, a[:, 1], a[:, 2], a[:, 3], a[:, 4])
        want = ref[sel]
        fin = np.isfinite(want).all(axis=1)
        assert fin.sum() > 200
        # non-finite in the reference (I_nu overflow) -> non-finite here
        assert (~np.isfinite(phi[~fin]) | ~np.isfinite(dphi[~fin])).all()
        ratio = np.abs(a[:, 3] / a[:, 4])                     # |V'/V''|: the scale of phi - phi0
        beta = np.sqrt(np.abs(a[:, 4]))
        tol_phi = 1e-9 * np.abs(want[:, 0] - a[:, 2]) + 1e-11 * ratio + 1e-15 * np.abs(a[:, 2])
        tol_dphi = 1e-9 * np.abs(want[:, 1]) + 1e-11 * ratio * beta
        assert (np.abs(phi - want[:, 0])[fin] <= tol_phi[fin]).all(), alpha
        assert (np.abs(dphi - want[:, 1])[fin] <= tol_dphi[fin]).all(), alpha


def test_bounces_real_alpha_vs_reference(gold):
    """Full bounces of the quartic family for seven real alpha, polynomial functor; S to 1e-6 (the stated bar), same
    shot counts; the reference's brentq failure at alpha = 0 (no friction term) comes back as its status."""
    from cosmotransitions_b200 import _lib, batch
    worst = 0.0
    for c in gold["cases"]:
        r = batch.find_bounce_batch(_lib.POT_POLY4, np.array([c["coef"]]), c["phi_abs"], c["phi_meta"], alpha=c["alpha"])
        st = int(r["status"][0])
        if c["error"]:
            assert st == ERR2STATUS[c["error"]], (c["alpha"], c["c"], st, c["error"])
            continue
        assert st in (0, 1), (c["alpha"], c["c"], st)
        rel = abs(float(r["S"][0]) - c["S"]) / abs(c["S"])
        worst = max(worst, rel)
        assert rel < 1e-6, (c["alpha"], c["c"], rel)
        assert int(r["n_shots"][0]) == c["n_shots"] and int(r["n_points"][0]) == c["npts"]
    assert worst < 1e-6


def test_real_alpha_matches_integer_kernels_and_scalar_api(gold):
    """The general-order path evaluated AT an integer alpha (forced through alpha_real by an offset of 1e-13) reproduces
    the specialised kernels; the scalar class accepts a real alpha; thermal functors refuse it loudly."""
    from cosmotransitions_b200 import _lib, batch, potentials, tunneling1D
    p = potentials.quartic_family_params([0.1, 0.3, 0.45])
    for alpha in (1, 2, 3, 4):
        a = batch.find_bounce_batch(_lib.POT_POLY4, p, 1.0, 0.0, alpha=alpha, kernel="thread")
        b = batch.find_bounce_batch(_lib.POT_POLY4, p, 1.0, 0.0, alpha=alpha + 1e-13, kernel="thread")
        np.testing.assert_allclose(b["S"].numpy(), a["S"].numpy(), rtol=1e-9)
        assert np.array_equal(a["n_shots"].numpy(), b["n_shots"].numpy())
    inst = tunneling1D.SingleFieldInstanton(1.0, 0.0, potentials.QuarticPotential(0.35), alpha=2.5)
    prof = inst.findProfile()
    case = [c for c in gold["cases"] if c["alpha"] == 2.5 and abs(c["c"] - 0.35) < 1e-12][0]
    assert abs(inst.findAction(prof) - case["S"]) < 1e-6 * case["S"] and len(prof.R) == case["npts"]
    with pytest.raises(_lib.CtbError):
        batch.find_bounce_batch(_lib.POT_MODEL1F, np.array([[0.12, 246.0 ** 2, 0.35, 30.0, 246.0 ** 2, 100.0]]), 240.0, 0.0,
                                alpha=2.5)
    with pytest.raises(ValueError):
        batch.make_opts(alpha=-1)


def test_fused_bessel_i0_i1_against_mpmath():
    """The fused I0 / I1 evaluation behind the alpha = 1, 3 start conditions (csrc/ctb_device.cuh bessel_i01), read out
    through ctb_exact_solution with V' = V'' = 1 (phi - phi0 = I0(r) - 1, phi' = I1(r) for alpha = 1; (2/r) I1 - 1 and
    (2/r) I2 for alpha = 3): a few ulp against 30-digit mpmath values, including the small-z end where the plain
    subtraction I0 - 1 would lose digits."""
    mp = pytest.importorskip("mpmath")
    from cosmotransitions_b200 import batch
    mp.mp.dps = 30
    z = np.concatenate([np.logspace(-2 + 1e-9, np.log10(7.75), 250), np.logspace(np.log10(7.7500001), np.log10(690.0), 250),
                        [7.75, 7.750000000000001, 100.0, 500.0]])
    one, zero = np.ones_like(z), np.zeros_like(z)
    p1, d1 = batch.exact_solution_batch(1, z, zero, one, one)
    p3, d3 = batch.exact_solution_batch(3, z, zero, one, one)
    ulp = lambda got, want: abs(got - float(want)) / np.spacing(abs(float(want)))
    worst = dict(i0m1=0.0, i1=0.0, f3=0.0, d3=0.0)
    for k, zk in enumerate(z):
        zz = mp.mpf(float(zk))
        i0, i1, i2 = mp.besseli(0, zz), mp.besseli(1, zz), mp.besseli(2, zz)
        worst["i0m1"] = max(worst["i0m1"], ulp(p1[k], i0 - 1))
        worst["i1"] = max(worst["i1"], ulp(d1[k], i1))
        worst["f3"] = max(worst["f3"], ulp(p3[k], 2 * i1 / zz - 1))
        worst["d3"] = max(worst["d3"], ulp(d3[k], 2 * i2 / zz))
    print("fused Bessel, worst ulp errors:", {k: round(v, 1) for k, v in worst.items()})
    assert worst["i0m1"] <= 8 and worst["i1"] <= 8 and worst["f3"] <= 12 and worst["d3"] <= 24
    # overflow behaviour: inf where I_nu overflows (as scipy's iv), finite just below
    p, d = batch.exact_solution_batch(1, np.array([705.0, 720.0, 1e5]), np.zeros(3), np.ones(3), np.ones(3))
    assert np.isfinite(p[0]) and np.isinf(p[1]) and np.isinf(p[2]) and np.isinf(d[2])
