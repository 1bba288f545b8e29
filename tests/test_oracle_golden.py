"""Pins the CPU oracle (oracle/ctb_oracle.c) against outputs of the reference itself, frozen by
oracle/gen_golden.py (tests/golden/*).  CPU only."""
import json
import os

import numpy as np
import pytest

S_RTOL = 1e-9      # required by BASELINE.json: 1e-6; the restatement is trajectory-faithful
ERR2STATUS = {"PotentialError:stable, not metastable": 2, "PotentialError:no barrier": 3,
              "IntegrationError:r > rmax": 4, "IntegrationError:dr < drmin": 5,
              "IntegrationError:Stepsize rounds down to zero.": 6}


def _status_for(err):
    if err in ERR2STATUS:
        return ERR2STATUS[err]
    if err.startswith("ValueError:The function value"):
        return 7
    if err.startswith("ValueError"):
        return 8
    if err.startswith("AssertionError"):
        return 9
    raise KeyError(err)


def _cfg(oracle, case, **kw):
    opts = dict(case.get("opts", {}))
    return oracle.make_config(alpha=case["alpha"], **opts, **kw)


def _load_cases(gold_dir, fname):
    cases = json.load(open(os.path.join(gold_dir, fname)))
    return cases["cases"] if isinstance(cases, dict) else cases


@pytest.mark.parametrize("fname,nmin", [("quartic.json", 150), ("alpha14.json", 40)])
def test_quartic_cases(oracle, gold_dir, fname, nmin):
    cases = _load_cases(gold_dir, fname)
    assert len(cases) > nmin
    worst = 0.0
    for case in cases:
        cfg = _cfg(oracle, case)
        oracle.set_overrides(**case.get("init", {}))
        r = oracle.bounce_one(cfg, case["coef"], case["phi_abs"], case["phi_meta"])
        oracle.set_overrides()
        if case["error"] is not None:
            assert r["status"] == _status_for(case["error"]), (case, r["status"])
            continue
        assert r["status"] in (0, 1), case
        rel = abs(r["S"] - case["S"]) / abs(case["S"])
        worst = max(worst, rel)
        assert rel < S_RTOL, (case["c"], case["alpha"], case["opts"], rel)
        assert r["n_points"] == case["npts"]
        assert r["n_shots"] == case["n_shots"]
        assert list(r["shot_type"]) == case["shot_types"]
        assert r["n_rkck"] == case["n_rkck"]
        # brentq inside initialConditions may take one evaluation more/less when sinh/exp differ
        # from scipy's iv() in the last ulp
        assert abs(r["n_exact"] - case["n_exact"]) <= 8
        # golden "phi0" is profile.Phi[0] (= phi(r=0) unless max_interior_pts=0)
        assert abs(r["Phi"][0] - case["phi0"]) <= 1e-12 * max(1.0, abs(case["phi0"]))
        assert abs(r["R"][0] - case["R0"]) <= 1e-9 * abs(case["R0"])
        # rf is where phi' (or phi - phi_meta) crosses zero in the flat tail: ill-conditioned, so
        # last-ulp differences in the start data are amplified (bar: solver tolerance 1e-4)
        assert abs(r["rf"] - case["rf"]) <= 1e-7 * abs(case["rf"])
        assert abs(r["r0"] - case["r0"]) <= 1e-9 * abs(case["r0"])
        assert abs(r["phi_bar"] - case["phi_bar"]) <= 1e-12 * max(1.0, abs(case["phi_bar"]))
        assert abs(r["rscale"] - case["rscale"]) <= 1e-9 * case["rscale"]
    print("worst S rel err", worst)


def test_profiles(oracle, gold_dir):
    g = np.load(os.path.join(gold_dir, "profiles.npz"))
    for name in ("thin2", "thick2", "mid3"):
        cfg = oracle.make_config(alpha=int(g[name + "_alpha"]))
        r = oracle.bounce_one(cfg, oracle.quartic_params(float(g[name + "_c"]))[0], 1.0, 0.0)
        assert len(r["R"]) == len(g[name + "_R"])
        np.testing.assert_allclose(r["R"], g[name + "_R"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(r["Phi"], g[name + "_Phi"], rtol=0, atol=1e-10)
        np.testing.assert_allclose(r["dPhi"], g[name + "_dPhi"], rtol=0, atol=1e-10)
        assert abs(r["S"] - float(g[name + "_S"])) < S_RTOL * abs(float(g[name + "_S"]))
    # docstring form (d2V by finite differences in the reference): within the 1e-6 bar
    for name, c in (("doc_thin", 0.47), ("doc_thick", 0.2)):
        cfg = oracle.make_config(alpha=2)
        r = oracle.bounce_one(cfg, oracle.quartic_params(c)[0], 1.0, 0.0)
        assert abs(r["S"] - float(g[name + "_S"])) < 1e-8 * abs(float(g[name + "_S"]))


def test_jexact_oracle_vs_reference(oracle, gold_dir):
    """The scipy.quad restatement of finiteT's exact integrals reproduces the reference's own output."""
    g = np.load(os.path.join(gold_dir, "jexact.npz"))
    for which, nm in ((0, "Jb"), (1, "Jf")):
        np.testing.assert_allclose(oracle.jexact(which, "x", g["x"]), g[nm + "_exact"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(oracle.jexact(which, "dx", g["x"]), g["d" + nm + "_exact"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(oracle.jexact(which, "theta", g["theta"]), g[nm + "_exact2"], rtol=1e-9, atol=1e-12)
        # and both sit within QUADPACK's tolerance of the 30-digit values (away from interior singularities)
        ok = g["mp_theta"] > (-39.0 if which == 0 else -9.8)
        assert np.max(np.abs(g[nm + "_exact2"][:40][ok] - g["mp_" + nm][ok])) < 1e-7


def test_wall_with_const_friction(oracle, gold_dir):
    """WallWithConstFriction.findProfile: oracle restatement vs the reference (F, rf, profile, shot / step counts)."""
    cases = json.load(open(os.path.join(gold_dir, "wall.json")))
    assert len(cases) >= 19
    for case in cases:
        kw = dict(case["opts"])
        wkw = {k: kw.pop(k) for k in ("Fguess", "Ftol", "phi0_rel") if k in kw}
        r = oracle.wall_one(oracle.make_config(**kw), case["coef"], case["phi_abs"], case["phi_meta"], **wkw)
        if case["error"] is not None:
            assert r["status"] == _status_for(case["error"]), (case, r["status"])
            continue
        assert r["status"] in (0, 1)
        assert r["n_shots"] == case["n_shots"] and r["n_rkck"] == case["n_rkck"] and r["n_points"] == case["npts"]
        assert abs(r["F"] - case["F"]) <= 1e-13 * case["F"]
        assert abs(r["rf"] - case["rf"]) <= 1e-8 * case["rf"]
        assert abs(r["rscale"] - case["rscale"]) <= 1e-12 * case["rscale"]
        assert abs(r["phi_bar"] - case["phi_bar"]) <= 1e-12
        np.testing.assert_allclose(r["R"][::20], case["R"], rtol=1e-8)
        np.testing.assert_allclose(r["Phi"][::20], case["Phi"], rtol=0, atol=1e-8)
        np.testing.assert_allclose(r["dPhi"][::20], case["dPhi"], rtol=0, atol=1e-8)
        assert abs(r["Phi"][-1] - case["Phi_last"]) < 1e-8 and abs(r["dPhi"][-1] - case["dPhi_last"]) < 1e-8


def test_thermal1f_generic_potential_grid(oracle, tables, gold_dir):
    """generic_potential.Vtot on a (phi, T) grid (radiation off) and the radiation constant (generic_potential.py:289-295)."""
    g = json.load(open(os.path.join(gold_dir, "thermal1f.json")))
    gp = g["generic_potential"]
    phi, T = np.array(gp["phi"]), np.array(gp["T"])
    P, TT = np.meshgrid(phi, T, indexing="ij")
    cfg = oracle.make_config(kind=oracle.POT_THERMAL1F, tables=tables)
    V = oracle.potential_rows(cfg, _thermal1f_rows(g, TT.reshape(-1)), P.reshape(-1)).reshape(P.shape)
    np.testing.assert_allclose(V, np.array(gp["Vtot_norad"]), rtol=1e-12, atol=1e-4)
    rad = (gp["num_boson_dof"] - sum(b[3] for b in g["bosons"])) * np.pi ** 4 / 45. \
        + (gp["num_fermion_dof"] - sum(f[2] for f in g["fermions"])) * 7 * np.pi ** 4 / 360.
    np.testing.assert_allclose(np.array(gp["Vtot"]) - np.array(gp["Vtot_norad"]),
                               np.broadcast_to(-rad * T ** 4 / (2 * np.pi ** 2), P.shape), rtol=1e-6, atol=1e-3)


def _profile_slice(g, case):
    sl = slice(*case["sl"])
    return tuple(g[case["name"] + k][sl] for k in ("_R", "_Phi", "_dPhi"))


def test_find_action_given_profiles(oracle, gold_dir):
    """findAction on given profiles (odd / even / tiny N, alpha incl. non-integer): oracle vs reference."""
    g = np.load(os.path.join(gold_dir, "profiles.npz"))
    cases = json.load(open(os.path.join(gold_dir, "findaction.json")))
    assert len(cases) > 100
    for case in cases:
        R, Phi, dPhi = _profile_slice(g, case)
        assert len(R) == case["n"]
        S = oracle.find_action(oracle.make_config(), oracle.quartic_params(case["c"])[0], 0.0, case["alpha"], R, Phi, dPhi)
        assert abs(S - case["S"]) <= 1e-12 * abs(case["S"]) + 1e-300, case


@pytest.mark.parametrize("fname", ["units.npz", "alpha14.json"])
def test_exact_solution_units(oracle, gold_dir, fname):
    if fname.endswith(".npz"):
        g = np.load(os.path.join(gold_dir, fname))
    else:
        g = {k: np.array(v) for k, v in json.load(open(os.path.join(gold_dir, fname))).items() if k != "cases"}
    bad = 0
    for row, ref in zip(g["exact_in"], g["exact_out"]):
        alpha, r, phi0, dV, d2V = row
        ph, dph = oracle.exact_solution(int(alpha), r, phi0, dV, d2V)
        for a, b in ((ph, ref[0]), (dph, ref[1])):
            if np.isnan(b):
                assert np.isnan(a)
            elif np.isinf(b):
                assert a == b
            else:
                # the reference's own formula cancels (I_{nu-1} - 2nu/z I_nu - I_{nu+1} ...): compare
                # against the magnitude of the uncancelled terms
                scale = max(abs(b), abs(dV / d2V) if d2V != 0 else 0.0, 1e-300)
                if not abs(a - b) <= 5e-10 * scale + 1e-13 * abs(phi0):
                    bad += 1
    assert bad == 0


def test_rkqs_units(oracle, gold_dir):
    g = np.load(os.path.join(gold_dir, "units.npz"))
    for row, ref in zip(g["rkqs_in"], g["rkqs_out"]):
        alpha, c, y0, y1, t, dt, ea0, ea1 = row
        st, out = oracle.rkqs_poly4(oracle.quartic_params(c)[0], int(alpha), [y0, y1], t, dt, [1e-4, 1e-4], [ea0, ea1])
        assert st == 0
        np.testing.assert_allclose(out[2:], ref[2:], rtol=1e-9)
        np.testing.assert_allclose(out[:2], ref[:2], rtol=1e-9, atol=1e-14)


def test_jspline(oracle, tables, gold_dir):
    g = np.load(os.path.join(gold_dir, "jspline.npz"))
    th = g["theta"]
    for der in range(4):
        jb = oracle.jspline(tables.tb, tables.cb, tables.xbmin, tables.xbmax, th, der)
        jf = oracle.jspline(tables.tf, tables.cf, tables.xfmin, tables.xfmax, th, der)
        tol = 1e-10 if der < 2 else 1e-9 * max(1.0, np.max(np.abs(g["Jb%d" % der])))
        np.testing.assert_allclose(jb, g["Jb%d" % der], rtol=0, atol=tol)
        np.testing.assert_allclose(jf, g["Jf%d" % der], rtol=0, atol=tol)


def test_jseries(oracle, gold_dir):
    import numpy as np
    g = np.load(os.path.join(gold_dir, "jseries.npz"))
    d = np.load(os.path.join(os.path.dirname(gold_dir), "..", "cosmotransitions_b200", "data", "finiteT_tables.npz"))
    x, xneg = g["x"], g["xneg"]
    for which, nm, lc in ((0, "Jb", d["low_b"]), (1, "Jf", d["low_f"])):
        for nt in (8, 3):
            for der in range(4):
                np.testing.assert_allclose(oracle.jseries(which, "high", der, nt, lc, x),
                                           g["%s_high_n%d_d%d" % (nm, nt, der)], rtol=1e-11, atol=1e-10)
        for der in range(4):
            ref = g["%s_high_neg_d%d" % (nm, der)]
            got = oracle.jseries(which, "high", der, 8, lc, xneg)
            assert np.array_equal(np.isnan(got), np.isnan(ref))
            np.testing.assert_allclose(got[~np.isnan(ref)], ref[~np.isnan(ref)], rtol=1e-11, atol=1e-10)
        for nt in (20, 5, 50):
            np.testing.assert_allclose(oracle.jseries(which, "low", 0, nt, lc, x), g["%s_low_n%d" % (nm, nt)],
                                       rtol=1e-11, atol=1e-10)


def _model1_params(g, T, x0, nhat):
    return np.array([g["p_l1"], g["p_l2"], g["p_mu2"], g["p_Y1"], g["p_Y2"], g["p_n"], g["p_renorm"], g["p_v2"], T,
                     x0[0], x0[1], nhat[0], nhat[1]], dtype=float)


def test_model1_vtot(oracle, tables, gold_dir):
    g = np.load(os.path.join(gold_dir, "model1.npz"))
    cfg = oracle.make_config(kind=oracle.POT_MODEL1_LINE, tables=tables)
    p = _model1_params(g, 0.0, g["grid_x0"], g["grid_nhat"])
    V = oracle.vtot_model1_grid(cfg, p, g["grid_s"], g["grid_T"])
    ref = g["grid_V"]
    np.testing.assert_allclose(V, ref, rtol=1e-12, atol=1e-9 * np.max(np.abs(ref)) * 1e-3)
    # scattered points: a "line" through each point with phi = 0
    for X, T, v in zip(g["pts_X"][:100], g["pts_T"][:100], g["pts_V"][:100]):
        p = _model1_params(g, T, X, [1.0, 0.0])
        out = oracle.potential(cfg, p, np.array([0.0]))[0]
        assert abs(out - v) <= 1e-12 * abs(v) + 1e-6


def test_model1_line_bounces(oracle, tables, gold_dir):
    g = np.load(os.path.join(gold_dir, "model1.npz"))
    res = json.loads(str(g["b_json"]))
    cfg = oracle.make_config(kind=oracle.POT_MODEL1_LINE, tables=tables, alpha=2)
    for T, xl, xh, ref in zip(g["b_T"], g["b_xlow"], g["b_xhigh"], res):
        L = np.linalg.norm(xh - xl)
        nhat = (xh - xl) / L
        assert abs(L - ref["L"]) < 1e-12 * L
        r = oracle.bounce_one(cfg, _model1_params(g, T, xl, nhat), 0.0, L)
        assert ref["error"] is None and r["status"] in (0, 1)
        assert abs(r["S"] - ref["S"]) < 1e-6 * abs(ref["S"]), (T, r["S"], ref["S"])
        assert r["n_points"] == ref["npts"]
        assert r["n_shots"] == ref["n_shots"]


def test_model1f_bounces(oracle, tables, gold_dir):
    cases = json.load(open(os.path.join(gold_dir, "model1f.json")))
    cfg = oracle.make_config(kind=oracle.POT_MODEL1F, tables=tables, alpha=2)
    for c in cases:
        p = np.array([c["lam"], c["v2"], c["Y"], c["n"], c["v2"], c["T"]])
        pa = c["phi_abs"]
        V = oracle.potential(cfg, p, np.array([0.0, 0.3 * pa, 0.7 * pa, pa, 1.1 * pa]))
        np.testing.assert_allclose(V, c["Vsamples"], rtol=1e-12)
        r = oracle.bounce_one(cfg, p, pa, 0.0)
        assert c["error"] is None and r["status"] in (0, 1)
        assert abs(r["S"] - c["S"]) < 1e-6 * abs(c["S"]), (c["T"], r["S"], c["S"])
        assert r["n_points"] == c["npts"]
        assert r["n_shots"] == c["n_shots"]


def test_splinepath_solves(oracle, gold_dir):
    """Every 1-D solve the reference performs inside pathDeformation.fullTunneling
    (examples/fullTunneling.py thin + thick, model1.findAllTransitions): potential = SplinePath's spline."""
    g = np.load(os.path.join(gold_dir, "splinepath.npz"))
    n = int(g["n"])
    assert n >= 20
    worst = 0.0
    for i in range(n):
        tck = (g["t%d" % i], g["c%d" % i], 3)
        pa, pm, bar, rs, S = g["scal%d" % i]
        params = oracle.spline_params(tck)
        cfg = oracle.make_config(kind=oracle.POT_SPLINE, alpha=2, nparam=len(params))
        r = oracle.bounce_one(cfg, params, pa, pm)
        assert r["status"] in (0, 1)
        assert abs(r["phi_bar"] - bar) <= 1e-12 * max(1.0, abs(bar))
        assert abs(r["rscale"] - rs) <= 1e-9 * rs
        rel = abs(r["S"] - S) / abs(S)
        worst = max(worst, rel)
        assert rel < S_RTOL, (i, str(g["tags"][i]), rel)
        assert len(r["R"]) == len(g["R%d" % i])
        np.testing.assert_allclose(r["Phi"], g["Phi%d" % i], rtol=0, atol=1e-9 * abs(pm - pa))
    print("splinepath worst rel err", worst)


def _thermal1f_rows(g, T):
    sys_path_hack = None
    from cosmotransitions_b200 import potentials
    return potentials.thermal1f_params(g["V0_coefs"], g["renorm"], T, g["bosons"], g["fermions"])


def test_thermal1f_general_model(oracle, tables, gold_dir):
    """General one-field generic_potential (5 bosons incl. a thermal mass, 1 fermion -> Jb AND Jf)."""
    g = json.load(open(os.path.join(gold_dir, "thermal1f.json")))
    cfg = oracle.make_config(kind=oracle.POT_THERMAL1F, tables=tables)
    rows = _thermal1f_rows(g, g["pts_T"])
    V = oracle.potential_rows(cfg, rows, g["pts_phi"])
    np.testing.assert_allclose(V, g["pts_V"], rtol=1e-12, atol=1e-4)
    for c in g["bounces"]:
        r = oracle.bounce_one(cfg, _thermal1f_rows(g, c["T"])[0], c["phi_abs"], 0.0)
        assert c["error"] is None and r["status"] in (0, 1)
        assert abs(r["S"] - c["S"]) < 1e-6 * abs(c["S"]), (c["T"], r["S"], c["S"])
        assert r["n_points"] == c["npts"] and r["n_shots"] == c["n_shots"]
