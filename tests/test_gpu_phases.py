"""Batched N-field minimum finder / phase follower (ctb_trace_minima) against goldens from
the live reference: minima of examples/testModel1.py found by scipy.optimize.fmin (the reference's findMinimum /
traceMinimum corrector) for 10 parameter points, and the reference's own getPhases() trace of the default model
(tests/golden/phases.npz, oracle/gen_golden.py gen_phases).  Needs a B200: pytest -m gpu."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

X_RTOL = 2e-5   # minima within 2e-5 (relative to |X| + 1) of the reference's fmin


@pytest.fixture(scope="module")
def ph():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from cosmotransitions_b200 import _lib, phases
    assert _lib.load().ctb_device_ok() == 0
    return phases


@pytest.fixture(scope="module")
def gold(gold_dir):
    return np.load(os.path.join(gold_dir, "phases.npz"))


def _close(X, ref):
    return np.linalg.norm(X - ref, axis=-1) <= X_RTOL * (np.linalg.norm(ref, axis=-1) + 1.0)


def test_phase_A_traced_upwards_vs_fmin(ph, gold):
    """10 parameter points x 61 temperatures in one launch, from approxZeroTMin()[0] at T = 0."""
    from cosmotransitions_b200 import _lib
    m8, T = gold["model8"], gold["T"]
    v = np.sqrt(m8[:, 7])
    r = ph.trace_minima(m8, np.stack([v, v], axis=-1), T)
    X, V, flag = (r[k].cpu().numpy() for k in ("X", "V", "flag"))
    for p in range(len(m8)):
        e = int(gold["endA"][p])
        assert (flag[p, :e] == _lib.MIN_OK).all(), (p, flag[p])
        for ref in (gold["XA"][p, :e], gold["XA_fmin"][p, :e]):
            assert _close(X[p, :e], ref).all(), (p, np.abs(X[p, :e] - ref).max())
        np.testing.assert_allclose(V[p, :e], gold["VA"][p, :e], rtol=1e-10)
        # the phase ends where the reference's minimiser falls into the other basin; the landing point is a minimum of B
        assert flag[p, e] == _lib.MIN_JUMPED and (flag[p, e + 1:] == _lib.MIN_NOT_TRACED).all()
        assert _close(X[p, e], gold["XB"][p, e]), (p, X[p, e], gold["XB"][p, e])
        assert np.isnan(X[p, e + 1:]).all()


def test_phase_B_traced_downwards_vs_fmin(ph, gold):
    """From the landing point of phase A downwards in T (per-instance start index, direction -1)."""
    from cosmotransitions_b200 import _lib
    m8, T, e = gold["model8"], gold["T"], gold["endA"].astype(int)
    seed = np.stack([gold["XB"][p, e[p]] for p in range(len(m8))])
    r = ph.trace_minima(m8, seed, T, t_begin=e, t_dir=-np.ones(len(m8)))
    X, flag = r["X"].cpu().numpy(), r["flag"].cpu().numpy()
    for p in range(len(m8)):
        have = np.isfinite(gold["XB"][p, :, 0])
        first = int(np.argmax(have))
        ok = flag[p] == _lib.MIN_OK
        # the same temperatures exist, up to one grid point at the spinodal end (where H -> 0 and "exists" is a matter
        # of tolerance in both codes)
        assert ok[first + 1:e[p] + 1].all() and not ok[:max(first - 1, 0)].any(), (p, first, flag[p])
        assert (flag[p, e[p] + 1:] == _lib.MIN_NOT_TRACED).all()
        both = ok & have
        assert _close(X[p][both], gold["XB"][p][both]).all()
        assert _close(X[p][both], gold["XB_fmin"][p][both]).all()


def test_find_minimum_batched(ph, gold):
    """generic_potential.findMinimum for 200 (X, T) pairs in one launch: displaced starts return to the golden minima."""
    from cosmotransitions_b200 import _lib
    rng = np.random.default_rng(11)
    m8, T = gold["model8"], gold["T"]
    rows, X0, Tq, ref = [], [], [], []
    for _ in range(200):
        p = int(rng.integers(len(m8)))
        t = int(rng.integers(0, gold["endA"][p] - 2))
        rows.append(m8[p]); Tq.append(T[t]); ref.append(gold["XA"][p, t])
        X0.append(gold["XA"][p, t] + rng.uniform(-10, 10, 2))
    r = ph.find_minimum(np.array(rows), np.array(X0), np.array(Tq))
    assert (r["flag"].cpu().numpy() == _lib.MIN_OK).all()
    assert _close(r["X"].cpu().numpy(), np.array(ref)).all()
    assert int(r["iters"].max()) <= 12


def test_trace_vs_reference_getPhases(ph, gold):
    """The reference's own traceMultiMin output for the default model (transitionFinder.py:371-553 through
    generic_potential.getPhases): the device follows each phase on the reference's own temperature samples, started
    from its first point.  The reference's points carry the tolerance of its fmin corrector (~1e-4 of |X|)."""
    from cosmotransitions_b200 import _lib
    m8 = gold["model8"][0]
    for key in (0, 1):
        T, X = gold["ref_phase%d_T" % key], gold["ref_phase%d_X" % key]
        sel = slice(1, -1) if key == 1 else slice(0, -1)   # the end points sit on the spinodals
        T, X = T[sel], X[sel]
        r = ph.trace_minima(m8, X[:1], T)
        Xd, flag = r["X"].cpu().numpy()[0], r["flag"].cpu().numpy()[0]
        assert (flag == _lib.MIN_OK).all(), (key, flag)
        err = np.linalg.norm(Xd - X, axis=-1) / (np.linalg.norm(X, axis=-1) + 1.0)
        assert err.max() < 2e-3 and np.median(err) < 2e-4, (key, err.max(), np.median(err))


def test_edge_cases(ph, gold):
    from cosmotransitions_b200 import _lib
    m8 = gold["model8"][0]
    # empty batch, empty grid
    r = ph.trace_minima(m8, np.zeros((0, 2)), gold["T"])
    assert r["X"].shape == (0, len(gold["T"]), 2)
    r = ph.trace_minima(m8, np.array([[246.0, 246.0]]), np.zeros(0))
    assert r["flag"].shape == (1, 0)
    # a start on the symmetric saddle at T = 0 stays there and is reported as "not a minimum", never as a minimum
    r = ph.find_minimum(m8, np.array([[0.0, 0.0]]), np.array([0.0]))
    assert int(r["flag"][0]) in (_lib.MIN_NOT_A_MINIMUM, _lib.MIN_NOT_CONVERGED)
    # NaN input terminates
    r = ph.find_minimum(m8, np.array([[np.nan, 1.0]]), np.array([50.0]))
    assert int(r["flag"][0]) == _lib.MIN_NOT_CONVERGED
    with pytest.raises(ValueError):
        ph.trace_minima(m8, np.array([[1.0, 1.0]]), gold["T"], t_begin=[len(gold["T"])])


def test_model1_phase_scan_two_field(ph, gold, oracle, tables):
    """BASELINE.json configs[4] on the real two-field model: phases on the device -> straight-line functor -> bounce.
    Critical temperatures against the golden V_A = V_B crossing, a sample of the bounces against the CPU oracle."""
    m8, T = gold["model8"], np.linspace(60.0, 125.0, 66)
    r = ph.model1_phase_scan(m8, T)
    S, st = r["S"].cpu().numpy(), r["status"].cpu().numpy()
    XA, XB = r["XA"].cpu().numpy(), r["XB"].cpu().numpy()
    assert r["n_bounces"] > 100 and ((st == 0) | (st == 1)).sum() > 0.9 * r["n_bounces"]
    # Tcrit: linear interpolation of the golden V_A - V_B on its own 4 K grid agrees to a fraction of that grid
    gT = gold["T"]
    for p in range(len(m8)):
        d = gold["VA"][p] - gold["VB"][p]
        k = np.where((d[:-1] < 0) & (d[1:] >= 0))[0]
        if len(k):
            Tc_gold = gT[k[-1]] + (0 - d[k[-1]]) * 4.0 / (d[k[-1] + 1] - d[k[-1]])
            assert abs(float(r["Tcrit"][p]) - Tc_gold) < 0.5, (p, float(r["Tcrit"][p]), Tc_gold)
    # default model: T_nuc of the straight-line path is the C4 value (83.6 K; the deformed path of the reference: 84.2 K)
    assert abs(float(r["Tnuc"][0]) - 83.6) < 0.3
    cfg = oracle.make_config(kind=oracle.POT_MODEL1_LINE, tables=tables)
    idx = np.argwhere((st == 0) | (st == 1))
    worst = 0.0
    for p, t in idx[:: max(1, len(idx) // 12)]:
        L = float(np.linalg.norm(XB[p, t] - XA[p, t]))
        row = np.concatenate([m8[p], [T[t]], XA[p, t], (XB[p, t] - XA[p, t]) / L])
        o = oracle.bounce_one(cfg, row, 0.0, L)
        assert o["status"] in (0, 1)
        worst = max(worst, abs(o["S"] - S[p, t]) / abs(o["S"]))
    assert worst < 1e-6, worst
