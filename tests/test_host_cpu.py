"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, the table conversion is exact, option marshalling, partitioning, loud failure without a GPU."""
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from cosmotransitions_b200 import _lib
    hdr = open(os.path.join(REPO, "include", "ctbounce.h")).read()
    declared = re.findall(r"CTB_API\s+(?:const\s+)?\w+\s*\*?\s*(ctb_\w+)\s*\(", hdr)
    assert len(declared) >= 10
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(declared) == sorted(_lib.EXPORTS)
    assert lib.ctb_abi_version() == _lib.ABI_VERSION
    assert lib.ctb_strerror(-1) == b"bad argument"
    assert lib.ctb_status_string(3) == b"no barrier"


def test_struct_layout_matches_header():
    import ctypes as C
    from cosmotransitions_b200 import _lib
    assert C.sizeof(_lib.JTable) == 64
    assert C.sizeof(_lib.Opts) == 7 * 8 + 6 * 4 + 8
    assert C.sizeof(_lib.BounceOut) == 15 * 8
    assert C.sizeof(_lib.BounceIn) == 8 * 8


def test_ctypes_structs_follow_header_field_order():
    """Field names and order of every ctypes.Structure must equal the typedef in include/ctbounce.h."""
    from cosmotransitions_b200 import _lib
    hdr = open(os.path.join(REPO, "include", "ctbounce.h")).read()

    def fields(name):
        m = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", hdr, re.S)
        body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            parts = decl.split(",")
            out.append(parts[0].split()[-1].lstrip("*"))
            out += [q.strip().lstrip("*") for q in parts[1:]]
        return out

    for cname, cls in (("ctb_jtable", _lib.JTable), ("ctb_opts", _lib.Opts), ("ctb_bounce_in", _lib.BounceIn),
                       ("ctb_bounce_out", _lib.BounceOut), ("ctb_wall_opts", _lib.WallOpts)):
        assert fields(cname) == [f[0] for f in cls._fields_], cname
    assert "#define CTB_ABI_VERSION %d" % _lib.ABI_VERSION in hdr


def test_integration_md_stub_matches_the_abi():
    """The reference-side ctypes stub printed in INTEGRATION.md declares the same structures as the product."""
    import ctypes as C
    from cosmotransitions_b200 import _lib
    md = open(os.path.join(REPO, "INTEGRATION.md")).read()
    code = md.split("```python")[1].split("```")[0]
    assert "ctb_abi_version() == %d" % _lib.ABI_VERSION in code
    classes = code[code.index("class JTable"):code.index("def bounce_actions_quartic")]
    ns = {"C": C}
    exec(classes, ns)
    for name in ("JTable", "Opts", "BounceIn", "BounceOut"):
        assert [f[0] for f in ns[name]._fields_] == [f[0] for f in getattr(_lib, name)._fields_], name
        assert C.sizeof(ns[name]) == C.sizeof(getattr(_lib, name)), name


def test_bspline_to_pp_matches_scipy():
    from scipy import interpolate
    from cosmotransitions_b200 import _tables
    h = _tables.host_tables()
    rng = np.random.default_rng(0)
    for t, c, brk, coef in ((h.tb, h.cb, h.brk_b, h.coef_b), (h.tf, h.cf, h.brk_f, h.coef_f)):
        assert len(brk) == coef.shape[0] + 1 == 998
        assert np.all(np.diff(brk) > 0)
        x = np.concatenate([rng.uniform(brk[0], brk[-1], 5000), rng.uniform(-5, 5, 5000), brk[:-1]])
        x = x[(x >= brk[0]) & (x < brk[-1])]
        j = np.searchsorted(brk, x, side="right") - 1
        dx = x - brk[j]
        for der in range(4):
            ref = interpolate.splev(x, (t, c, 3), der=der)
            cc = coef[j]
            if der == 0:
                got = cc[:, 0] + dx * (cc[:, 1] + dx * (cc[:, 2] + dx * cc[:, 3]))
            elif der == 1:
                got = cc[:, 1] + dx * (2 * cc[:, 2] + dx * 3 * cc[:, 3])
            elif der == 2:
                got = 2 * cc[:, 2] + dx * 6 * cc[:, 3]
            else:
                got = 6 * cc[:, 3]
            assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) < 2e-12, der


def test_opts_marshalling_and_defaults():
    from cosmotransitions_b200 import batch
    o = batch.make_opts()
    assert (o.xtol, o.phitol, o.thinCutoff, o.npoints, o.rmin, o.rmax) == (1e-4, 1e-4, .01, 500, 1e-4, 1e4)
    assert o.max_interior_pts == -1 and np.isnan(o.xguess) and o.alpha == 2 and o.phi_eps == 1e-3
    assert batch.profile_capacity(o) == 750
    o = batch.make_opts(alpha=3, npoints=200, max_interior_pts=0, xguess=2.5)
    assert batch.profile_capacity(o) == 200 and o.xguess == 2.5
    with pytest.raises(TypeError):
        batch.make_opts(bogus=1)
    # integer alpha 1..4: the specialised kernels; any other real alpha in [0, 8]: alpha = 0 + alpha_real
    assert batch.make_opts(alpha=1).alpha == 1 and batch.make_opts(alpha=4).alpha == 4
    o = batch.make_opts(alpha=2.5)
    assert o.alpha == 0 and o.alpha_real == 2.5 and batch.make_opts(alpha=5).alpha == 0
    for bad in (-0.5, 8.5, float("nan")):
        with pytest.raises(ValueError):
            batch.make_opts(alpha=bad)


def test_partition_roundtrip():
    import torch
    from cosmotransitions_b200 import multigpu
    rng = np.random.default_rng(0)
    for n in (0, 1, 255, 256, 257, 1003, 10000):
        for world in (1, 2, 3, 8):
            for block in (1, 7, 256):
                vals = rng.random((n, 3))
                out = np.full((n, 3), np.nan)
                shards = []
                for r in range(world):
                    lay = multigpu.ShardLayout(n, world, r, block)
                    idx = multigpu.partition(n, world, r, block)
                    assert lay.n_local == len(idx) == lay.nfull * block + lay.tail
                    shards.append(lay.take(vals))
                    assert np.array_equal(shards[-1], vals[idx])
                    lay.put(shards[-1], out)
                assert sum(len(s) for s in shards) == n and np.array_equal(out, vals)
                assert np.array_equal(multigpu.gather_blocks(shards, n, world, block), vals)
    assert np.array_equal(multigpu.partition(10, 3, 1), [1, 4, 7])      # block = 1: plain round-robin


def test_shared_host_arrays_roundtrip(tmp_path):
    """The shared-memory result segment of the one-process-per-GPU driver: two mappings see the same pages."""
    from cosmotransitions_b200 import multigpu
    path = str(tmp_path / "seg")
    spec = [("S_0", np.float64, 1000), ("status_0", np.int32, 1000)]
    a = multigpu.SharedHostArrays(path, spec, create=True, register=False)
    b = multigpu.SharedHostArrays(path, spec, create=False, register=False)
    a.arrays["S_0"][:] = np.arange(1000) * 0.5
    b.arrays["status_0"][7] = 3
    assert np.array_equal(b.arrays["S_0"], np.arange(1000) * 0.5) and a.arrays["status_0"][7] == 3
    assert a.arrays["status_0"].ctypes.data % 4096 == a.arrays["S_0"].ctypes.data % 4096      # page-aligned slots
    b.close()
    a.close()
    assert not os.path.exists(path)


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cosmotransitions_b200 import _lib, finiteT, find_bounce_batch, potentials
    with pytest.raises(_lib.CtbError):
        find_bounce_batch(_lib.POT_POLY4, potentials.quartic_family_params([0.2]), 1.0, 0.0)
    with pytest.raises(_lib.CtbError):
        finiteT.Jb_spline(np.array([0.0]))


def test_status_to_exception_mapping():
    from cosmotransitions_b200 import tunneling1D as t1, helper_functions as hf, _lib
    with pytest.raises(t1.PotentialError) as e:
        t1.raise_for_status(_lib.ST_STABLE)
    assert e.value.args[1] == "stable, not metastable"
    with pytest.raises(t1.PotentialError) as e:
        t1.raise_for_status(_lib.ST_NO_BARRIER)
    assert e.value.args[1] == "no barrier"
    for st, msg in ((_lib.ST_RMAX, "r > rmax"), (_lib.ST_DRMIN, "dr < drmin"),
                    (_lib.ST_STEP_UNDERFLOW, "Stepsize rounds down to zero.")):
        with pytest.raises(hf.IntegrationError) as e:
            t1.raise_for_status(st)
        assert e.value.args[0] == msg
    with pytest.raises(ValueError):
        t1.raise_for_status(_lib.ST_NONFINITE)
    t1.raise_for_status(0)
    t1.raise_for_status(1)


def test_evenly_spaced_phi_matches_reference_golden(gold_dir):
    """evenlySpacedPhi is host-side numpy (tunneling1D.py:793-826): checked on CPU against the reference's
    recorded calls (no GPU needed: the method does not touch the solver)."""
    from cosmotransitions_b200 import tunneling1D as t1
    g = np.load(os.path.join(gold_dir, "splinepath.npz"))
    done = 0
    for i in range(int(g["n"])):
        if "esp_p%d" % i not in g:
            continue
        pa, pm = g["scal%d" % i][:2]
        obj = t1.SingleFieldInstanton.__new__(t1.SingleFieldInstanton)
        obj.phi_absMin, obj.phi_metaMin = float(pa), float(pm)
        npts, k, fixAbs = (int(v) for v in g["espkw%d" % i])
        p, dp = obj.evenlySpacedPhi(g["Phi%d" % i], g["dPhi%d" % i], npoints=npts, k=k, fixAbs=bool(fixAbs))
        np.testing.assert_allclose(p, g["esp_p%d" % i], rtol=1e-14, atol=1e-14)
        np.testing.assert_allclose(dp, g["esp_dp%d" % i], rtol=1e-11, atol=1e-13)
        done += 1
    assert done >= 10


def test_spline_potential_row_layout():
    from scipy import interpolate
    from cosmotransitions_b200 import _lib, potentials
    x = np.linspace(-1, 3, 60)
    tck = interpolate.splrep(x, np.sin(x) * x, s=0)
    pot = potentials.SplinePotential(tck)
    row = pot.params
    nint = int(row[0])
    assert nint == 57 and len(row) == _lib.NPARAM[_lib.POT_SPLINE]
    brk = row[1:2 + nint]
    coef = row[1 + 256:1 + 256 + 4 * nint].reshape(nint, 4)
    xs = np.random.default_rng(0).uniform(-1, 3, 500)
    j = np.clip(np.searchsorted(brk, xs, side="right") - 1, 0, nint - 1)
    dx = xs - brk[j]
    v = coef[j, 0] + dx * (coef[j, 1] + dx * (coef[j, 2] + dx * coef[j, 3]))
    np.testing.assert_allclose(v, interpolate.splev(xs, tck), rtol=1e-12, atol=1e-13)


def _run_bench(*flags):
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"] + list(flags), stdout=subprocess.PIPE, text=True, timeout=600).stdout
    return json.loads(out.strip().splitlines()[-1])


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints one JSON line with the keys the driver reads.  `--port`: the C restatement
    with pthreads; default (oracle/_ref built by oracle/make_ref.py): the unmodified Python reference under
    multiprocessing.Pool.  Runs on CPU in well under a minute."""
    j = _run_bench("--port")
    assert j["impl"] == "reference" and j["unit"] == "bounces/s" and j["higher_is_better"] is True
    assert j["value"] > 1e3 and j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert j["config"]["workload"].startswith("C2") and j["solved_fraction"] == 1.0
    assert j["config"] == {"workload": j["config"]["workload"], "instances": 100000, "alpha": 2, "order": "natural"}
    if not os.path.exists(os.path.join(REPO, "oracle", "_ref", "cosmoTransitions", "tunneling1D.py")):
        pytest.skip("oracle/_ref not built here (python oracle/make_ref.py)")
    r = _run_bench("--ref-seconds", "1.0")
    assert r["cpu_baseline"]["kind"] == "reference" and r["config"] == j["config"]
    assert 1.0 < r["value"] < 1e5 and r["solved_fraction"] == 1.0 and "scipy" in r["versions"]


def test_live_reference_matches_oracle_and_goldens(oracle):
    """The unmodified reference shipped as oracle/_ref (oracle/make_ref.py), run live through oracle/ref_runner.py, against
    the C restatement on the same C2 instances: the pin of the oracle, re-checked wherever the suite runs."""
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref not built here (python oracle/make_ref.py)")
    from cosmotransitions_b200 import potentials
    n = 100000
    idx = np.arange(7, n, 2500)
    c = 0.02 + 0.475 * (idx + 0.5) / n
    for alpha in (2, 3):
        S, status, _ = ref_runner.quartic_pool(c, alpha, procs=min(8, len(os.sched_getaffinity(0))), chunksize=5)
        assert set(status) == {"ok"}
        o = oracle.bounce_batch(oracle.make_config(alpha=alpha), potentials.quartic_family_params(c), 1.0, 0.0)
        assert (o["status"] <= 1).all()
        assert np.max(np.abs(o["S"] - S) / np.abs(S)) < 1e-9


def test_monotonic_indices_matches_the_sequential_definition():
    """helper_functions.monotonicIndices (vectorised) against the point-by-point rule of helper_functions.py:57-75."""
    from cosmotransitions_b200.helper_functions import monotonicIndices

    def sequential(x):
        x = np.array(x)
        flipped = x[0] > x[-1]
        if flipped:
            x = x[::-1]
        kept = [0]
        for i in range(1, len(x) - 1):
            if x[kept[-1]] < x[i] < x[-1]:
                kept.append(i)
        kept = np.array(kept + [len(x) - 1])
        return len(x) - 1 - kept if flipped else kept

    rng = np.random.default_rng(0)
    for t in range(500):
        n = int(rng.integers(2, 40))
        x = np.cumsum(rng.normal(0.2 * rng.choice([-1, 1]), 1, n)) if t % 2 else rng.integers(0, 6, n).astype(float)
        assert np.array_equal(monotonicIndices(x), sequential(x))


def test_finite_difference_stencils_are_exact_on_polynomials():
    """generic_potential._stencil1 (the gradientFunction / hessianFunction weights, helper_functions.py:446-528):
    order 4 differentiates quartics exactly, order 2 quadratics; the radiation constant of V1T (generic_potential.py:289-295)."""
    from cosmotransitions_b200 import generic_potential as gpm
    p = np.poly1d([0.7, -1.3, 0.4, 2.0, -5.0])
    x0, eps = 0.83, 1e-2
    for order, poly in ((4, p), (2, np.poly1d([0.4, 2.0, -5.0]))):
        d1 = sum(w * poly(x0 + h) for h, w in gpm._stencil1(eps, order, 1))
        d2 = sum(w * poly(x0 + h) for h, w in gpm._stencil1(eps, order, 2))
        assert abs(d1 - poly.deriv(1)(x0)) < 1e-10 and abs(d2 - poly.deriv(2)(x0)) < 1e-8
    m = gpm.one_field_model.__new__(gpm.one_field_model)
    m.bosons, m.fermions, m.num_boson_dof, m.num_fermion_dof = [(0, 1, 0, 3., 1.5), (0, 1, 0, 1., 1.5)], [(0, 1, 12.)], 10., None
    assert m._rad(False) == 0.0 and abs(m._rad(True) - 6. * np.pi ** 4 / 45.) < 1e-12
    m.num_fermion_dof = 90.
    assert abs(m._rad(True) - (6. * np.pi ** 4 / 45. + 78. * 7 * np.pi ** 4 / 360.)) < 1e-10


def test_sampled_potential_rows_reproduce_the_callable():
    """potentials.SampledPotential (Python callables -> CTB_POT_SPLINE row): layout and interpolation error on the host."""
    from cosmotransitions_b200 import _lib, potentials
    c = 0.3
    V = lambda p: .25 * p ** 4 - (1 + c) / 3 * p ** 3 + c / 2 * p ** 2
    dV = lambda p: p * (p - c) * (p - 1.0)
    x = np.random.default_rng(0).uniform(-0.04, 1.04, 2000)
    for dv, pieces, tol in ((dV, 255, 1e-11), (None, 255, 2e-11), (dV, 32, 5e-8), (lambda p: float(dV(float(p))), 64, 5e-9)):
        sp = potentials.SampledPotential.for_instanton(V, dv, 1.0, 0.0, pieces=pieces)
        row = sp.params
        assert sp.kind == _lib.POT_SPLINE and row.shape == (_lib.NPARAM[_lib.POT_SPLINE],) and int(row[0]) == pieces
        brk = row[1:pieces + 2]
        assert abs(brk[0] + 0.05) < 1e-15 and abs(brk[-1] - 1.05) < 1e-15 and (np.diff(brk) > 0).all()
        coef = row[257:257 + 4 * pieces].reshape(pieces, 4)
        j = np.clip(np.searchsorted(brk, x, side="right") - 1, 0, pieces - 1)
        dx = x - brk[j]
        v = coef[j, 0] + dx * (coef[j, 1] + dx * (coef[j, 2] + dx * coef[j, 3]))
        assert np.abs(v - V(x)).max() < tol
        if dv is not None:   # Hermite: V and dV exact at the knots
            np.testing.assert_allclose(coef[:, 0], V(brk[:-1]), rtol=0, atol=1e-15)
            np.testing.assert_allclose(coef[:, 1], dV(brk[:-1]), rtol=0, atol=1e-15)
    with pytest.raises(ValueError):
        potentials.SampledPotential(V, 0.0, 1.0, pieces=1000)
    with pytest.raises(ValueError):
        potentials.SampledPotential(lambda p: np.log(p), -1.0, 1.0)


def test_compiled_model_builds_without_a_gpu():
    """jit.CompiledModel: the generated source compiles for sm_100a here (nvcc cross-compiles), exports the three entry
    points with the model's ndim / nparam, is cached by content, and reports nvcc errors with the source location."""
    import ctypes as C
    from cosmotransitions_b200 import _lib, jit
    m = jit.model1_compiled()
    assert (m.ndim, m.nparam, m.uses_fermions) == (2, 8, False) and os.path.exists(m.so_path)
    lib = C.CDLL(m.so_path)
    for name in ("ctbjit_info", "ctbjit_vtot", "ctbjit_trace_minima"):
        assert hasattr(lib, name)
    t0 = os.path.getmtime(m.so_path)
    assert jit.model1_compiled().so_path == m.so_path and os.path.getmtime(m.so_path) == t0      # cache hit
    m3 = jit.CompiledModel(3, ("k",), "return k * (X[0] * X[0] + X[1] * X[1] + X[2] * X[2]) + T * Jf(X[0] / T2);")
    assert m3.uses_fermions and m3.so_path != m.so_path
    with pytest.raises(_lib.CtbError) as e:
        jit.CompiledModel(2, ("a",), "return no_such_symbol * X[0];")
    assert "no_such_symbol" in str(e.value)
    with pytest.raises(_lib.CtbError):     # no GPU here: the calls fail loudly, there is no CPU fallback
        m.vtot(np.zeros(8), np.zeros((1, 2)), np.zeros(1))
    # the bounce kernels instantiated for the model's line potential (second library, built on first use)
    f = m.line_functor()
    assert f.nparam == 8 + 1 + 2 * 2 and f.custom_entry is f and os.path.exists(f.so_path) and f.so_path != m.so_path
    blib = C.CDLL(f.so_path)
    for name in ("ctbjit_bounce_info", "ctbjit_bounce_setup", "ctbjit_bounce_batch"):
        assert hasattr(blib, name)
    assert m.line_functor() is f
