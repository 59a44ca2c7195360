"""CPU: host-side logic -- radial splines -> per-surface tables (the geometry kernel's input contract),
synthetic equilibria, grid helpers, sharding arithmetic."""
import numpy as np
import pytest

from helpers import tables_from_fixture


def test_radial_splines_match_reference_fitpack(golden):
    """RadialSplines.evaluate (vectorised not-a-knot cubic) vs the tables the REFERENCE's FITPACK splines
    produced (utils.py:37-158, 311-357), stored in the fixtures."""
    from ideal_ballooning_solver_b200 import synthetic, tables
    for name, kind in (("synthetic_ncsx", "ncsx"), ("synthetic_d3d", "d3d"), ("synthetic_hberg", "hberg")):
        D = golden(name)
        st = tables.RadialSplines(synthetic.make_equilibrium(kind)).evaluate(D["surfaces"])
        for got, key in ((st.tab_mn, "tab_mn"), (st.tab_nyq, "tab_nyq"), (st.scal, "scal")):
            ref = D[key]
            assert got.shape == ref.shape
            assert np.max(np.abs(got - ref)) <= 1e-11 * max(1.0, np.max(np.abs(ref))), (name, key)
        assert np.array_equal(st.xm, D["xm"]) and np.array_equal(st.xn_nyq, D["xn_nyq"])


def test_synthetic_equilibria_have_the_device_mode_counts():
    from ideal_ballooning_solver_b200 import synthetic
    for kind, mn, mnq, nfp in (("d3d", 80, 84, 1), ("ncsx", 242, 392, 3), ("hberg", 242, 392, 2)):
        w = synthetic.make_equilibrium(kind, seed=1)
        assert w.rmnc.shape[0] == mn and w.bmnc.shape[0] == mnq and int(w.nfp) == nfp
        assert np.all(np.mod(w.xn, nfp) == 0)
        # deterministic in the seed
        w2 = synthetic.make_equilibrium(kind, seed=1)
        assert np.array_equal(w.rmnc, w2.rmnc)


def test_s_alpha_coefficients_formula():
    """bishop_ball_s-alpha.py:30-45."""
    from ideal_ballooning_solver_b200 import synthetic
    th = np.linspace(-3, 3, 11)
    g, c, f = synthetic.s_alpha_coefficients(0.7, 0.5, 0.2, th)
    lam = 0.7 * (th - 0.2) - 0.5 * (np.sin(th) - np.sin(0.2))
    np.testing.assert_allclose(g, 1 + lam ** 2)
    np.testing.assert_allclose(c, 0.5 * (np.cos(th) + lam * np.sin(th)))
    np.testing.assert_allclose(f, g)


def test_grid_spacing_is_the_reference_h():
    """utils.py:1567-1575: h = np.diff(theta_half)[2]."""
    from ideal_ballooning_solver_b200 import engine
    from oracle import ballooning_oracle as bo
    th = np.linspace(-4 * np.pi, 4 * np.pi, 969)
    one = np.ones_like(th)
    assert engine.grid_spacing(th) == bo.discretise(th, one, one, one)[0]


def test_shard_range_partitions():
    from ideal_ballooning_solver_b200 import scan
    for n in (0, 1, 5, 64, 129):
        for world in (1, 2, 3, 8):
            parts = [scan.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_scan_theta_grid_matches_reference_resolution():
    from ideal_ballooning_solver_b200 import scan
    assert len(scan.scan_theta_grid(80, 0)) == 641            # D3D: 2*80*4+1   (ball_scan.py:206-208)
    assert len(scan.scan_theta_grid(11, 11)) == 969           # NCSX: 2*11*11*4+1 (ball_scan.py:201-204)


def test_surface_tables_select_and_rows(golden):
    D = golden("ncsx_wout_op")
    st = tables_from_fixture(D)
    sub = st.select([2, 0])
    assert sub.ns == 2 and np.array_equal(sub.s, D["scal"][[2, 0], 0])
    assert sub.row("lmns").shape == (2, 242) and sub.row("bsubvmnc").shape == (2, 392)
    np.testing.assert_allclose(sub.row("shat"), (-2 * sub.row("s") / sub.row("iota")) * sub.row("d_iota_d_s"))


def test_read_wout_roundtrip(tmp_path):
    """NetCDF-3 reader used for real VMEC files."""
    from scipy.io import netcdf_file
    from ideal_ballooning_solver_b200 import synthetic, tables
    w = synthetic.make_equilibrium("d3d", seed=2)
    p = str(tmp_path / "wout_test.nc")
    with netcdf_file(p, "w") as f:
        ns, mn, mnq = int(w.ns), w.rmnc.shape[0], w.bmnc.shape[0]
        f.createDimension("radius", ns); f.createDimension("mn_mode", mn); f.createDimension("mn_mode_nyq", mnq)
        f.createDimension("n_tor", len(w.raxis_cc))
        for k in tables._TABLES_2D:
            a = getattr(w, k)
            v = f.createVariable(k, "d", ("radius", "mn_mode" if a.shape[0] == mn else "mn_mode_nyq"))
            v[:] = a.T
        for k in tables._TABLES_1D:
            a = np.asarray(getattr(w, k), float)
            dim = {ns: "radius", mn: "mn_mode", mnq: "mn_mode_nyq"}.get(len(a), "n_tor")
            if mn == mnq and k.endswith("_nyq"):
                dim = "mn_mode_nyq"
            v = f.createVariable(k, "d", (dim,)); v[:] = a
        for k in tables._SCALARS:
            val = getattr(w, k)
            v = f.createVariable(k, "d" if isinstance(val, float) else "i", ()); v[...] = val
    r = tables.read_wout(p)
    assert np.array_equal(r.rmnc, w.rmnc) and np.array_equal(r.bsubvmnc, w.bsubvmnc) and r.nfp == w.nfp
    a = tables.RadialSplines(r).evaluate([0.5, 0.9]); b = tables.RadialSplines(w).evaluate([0.5, 0.9])
    assert np.array_equal(a.tab_mn, b.tab_mn) and np.array_equal(a.scal, b.scal)


def test_save_results_append_semantics(tmp_path):
    """ball_scan.py:359-384 on the placeholder files of arr_create2.py:87-97."""
    from ideal_ballooning_solver_b200 import scan
    for nm in ("ball_gam", "ball_theta0", "ball_alpha"):
        np.save(str(tmp_path / f"{nm}3.npy"), np.empty([]))
    mk = lambda k: scan.BallScanResult(gamma=np.arange(5.0) + k, theta0=np.arange(5.0) * 0.1 + k, alpha=np.arange(5.0) * 0.2 + k,
                                       gamma_coarse=np.zeros((5, 2, 2)), X=np.zeros((5, 9)), refine=None)
    scan.save_results(str(tmp_path), 3, 0, mk(0))
    g = np.load(str(tmp_path / "ball_gam3.npy"))
    assert g.shape == (5,) and np.array_equal(g, np.arange(5.0))
    scan.save_results(str(tmp_path), 3, 1, mk(10))
    scan.save_results(str(tmp_path), 3, 2, mk(20))
    g = np.load(str(tmp_path / "ball_gam3.npy"))
    a = np.load(str(tmp_path / "ball_alpha3.npy"))
    assert g.shape == (3, 5) and np.array_equal(g[2], np.arange(5.0) + 20)
    assert np.array_equal(a[1], np.arange(5.0) * 0.2 + 10)


def test_chain_length_divides():
    from ideal_ballooning_solver_b200 import scan
    assert scan.chain_length(64) == 16 and scan.chain_length(15) == 15 and scan.chain_length(24) == 12
    assert scan.chain_length(1) == 1 and scan.chain_length(17) == 1 and scan.chain_length(32) == 16


def test_ballooning_penalty_matches_reference_formula():
    """sims_runner_NCSX.py:254-261, 313."""
    from ideal_ballooning_solver_b200 import penalty
    gam = np.array([-1e-3, -2e-4, -1e-4, 3e-4, 0.0])
    want = 50 * np.sum(np.maximum(gam - (-0.0002), 0.0))
    assert penalty.ballooning_penalty(gam) == want
    assert penalty.objective(0.25, gam) == np.sqrt(0.25 + want)
    assert penalty.objective(0.25, gam, converged=False) == np.sqrt(9999.0)
    gams = np.stack([gam, gam + 1e-5, gam - 2e-5])
    f, df = penalty.fd_jacobian([0.25, 0.26, 0.24], gams, np.array([0.0, 1e-3, 2e-3]))
    assert df.shape == (1, 3) and df[0, 0] == 0
    np.testing.assert_allclose(df[0, 1], (f[1] - f[0]) / 1e-3 * 0.5 / np.sqrt(f[0]))


def test_hf_jacobian_is_fd_jacobian_on_predicted_growth_rates():
    from ideal_ballooning_solver_b200 import penalty
    gam0 = np.array([-1e-3, 2e-4, 5e-4, -3e-4])
    dgam = np.array([[1e-5, -2e-5, 3e-5, 4e-4], [0.0, 1e-5, -1e-3, 0.0]])
    steps = np.array([0.0, 1e-3, 2e-3])
    f, df = penalty.hf_jacobian([0.25, 0.26, 0.24], gam0, dgam, steps)
    f2, df2 = penalty.fd_jacobian([0.25, 0.26, 0.24], [gam0, gam0 + dgam[0], gam0 + dgam[1]], steps)
    assert np.array_equal(f, f2) and np.array_equal(df, df2)
    # the threshold kink is honoured: surface 3 crosses the threshold under the first perturbation
    assert f[1] - f[0] > 0.01 + penalty.PREFAC_BALL * (1e-5 - 2e-5 + 3e-5)


def test_derm_dermv_facade_matches_reference_golden(golden):
    """The host facade's derm / dermv (reference_api.py) against the reference's own outputs: bit-exact."""
    from ideal_ballooning_solver_b200 import reference_api as ra
    G = golden("derm")
    n = 0
    for key in G.files:
        if not key.startswith("derm"):
            continue
        fn, dim, ch, par = key.split("_")
        a = G["a1"] if dim == "1d" else G["a2"]
        if fn == "derm":
            r = ra.derm(a, ch, par)
        else:
            r = ra.dermv(a, G["b1"] if dim == "1d" else (G["b2l"] if ch == "l" else G["b2r"]), ch, par)
        assert r.shape == G[key].shape and np.array_equal(r, G[key]), key
        n += 1
    assert n == 14
    with pytest.raises(NotImplementedError):
        ra.dermv(G["a1"], G["b1"], "r")


def test_batched_objective_failure_does_not_hang():
    """A failing batched evaluation (CUDA error, bad input, ...) must reach EVERY waiting surface thread as an exception;
    before the fix the failing thread left its entry pending and all the others waited forever."""
    import threading
    from ideal_ballooning_solver_b200 import scan

    obj = scan._BatchedObjective.__new__(scan._BatchedObjective)
    obj.cv = threading.Condition(); obj.active = set(); obj.pending = {}; obj.results = {}; obj.nbatches = 0; obj.nevals = 0
    calls = []

    def boom():
        calls.append(sorted(obj.pending))
        raise RuntimeError("injected batch failure")

    obj._run_batch = boom
    obj.start(range(4))
    errs = [None] * 4

    def worker(i):
        try:
            obj.evaluate(i, (0.1 * i, 0.2))
        except Exception as e:      # noqa: BLE001
            errs[i] = e
        finally:
            obj.finish(i)

    ts = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=20)
    assert not any(t.is_alive() for t in ts), "a surface thread is still waiting"
    assert all(isinstance(e, RuntimeError) for e in errs) and calls == [[0, 1, 2, 3]]
