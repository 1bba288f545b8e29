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
