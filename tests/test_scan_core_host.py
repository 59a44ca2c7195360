"""CPU check of the lane code of the scan solver (csrc/ibs_scan_core.cuh) through the g++ harness
tools/scan_core_host.cpp: the same per-lane arithmetic the CUDA kernel runs (division-free twisted recurrences,
multigrid start, streaming output passes) against the oracle.  The harness is test infrastructure, not a fallback."""
import os
import shutil
import sys

import numpy as np
import pytest

from helpers import LAM_RTOL, X_ATOL, s_alpha_base

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    import scan_host_check as shc
    return shc, shc.build(str(tmp_path_factory.mktemp("sch")))


@pytest.mark.parametrize("name,nth0", [("synthetic_d3d", 5), ("synthetic_ncsx", 3), ("synthetic_hberg", 2)])
def test_lane_code_matches_oracle(harness, golden, name, nth0):
    shc, lib = harness
    D = golden(name)
    ns, na = D["geo_bmag"].shape[:2]
    base = np.stack([D["geo_" + n] for n in shc.BASE_NAMES], axis=2).reshape(ns * na, 8, -1)
    th0 = np.tile(np.linspace(0.0, np.pi / 2, nth0), (ns * na, 1))
    theta = D["theta"]
    R = shc.host_scan_solve(lib, base, D["dPdrho"].reshape(-1), th0, theta[1] - theta[0], sigma=np.full(th0.size, 1.0))
    assert np.all((R["info"] >> 16) == 0)
    assert R["cost"] / th0.size < 12.0                    # fine-grid-equivalent passes per solve (multigrid start)
    for line in range(ns * na):
        i, j = divmod(line, na)
        for t in range(nth0):
            gam, X, dX, _ = shc.oracle_solve(D, i, j, th0[line, t])
            s = line * nth0 + t
            assert abs(R["lam"][line, t] - gam) <= LAM_RTOL * abs(gam)
            np.testing.assert_allclose(R["X"][s], X, rtol=0, atol=X_ATOL)
            np.testing.assert_allclose(R["dX"][s], dX, rtol=0, atol=10 * X_ATOL * max(1.0, np.max(np.abs(dX))))


def test_lane_code_fixture_values(harness, golden):
    """Against the stored converged-reference values (ARPACK tol=0 run of the reference itself)."""
    shc, lib = harness
    D = golden("synthetic_d3d")
    ns, na, nt = D["lam_conv"].shape
    base = np.stack([D["geo_" + n] for n in shc.BASE_NAMES], axis=2).reshape(ns * na, 8, -1)
    th0 = np.tile(D["theta0s"], (ns * na, 1))
    theta = D["theta"]
    R = shc.host_scan_solve(lib, base, D["dPdrho"].reshape(-1), th0, theta[1] - theta[0])
    np.testing.assert_allclose(R["lam"].reshape(-1), D["lam_conv"].reshape(-1), rtol=LAM_RTOL, atol=0)


def test_lane_code_flags_bad_input(harness, golden):
    shc, lib = harness
    D = golden("synthetic_d3d")
    base = np.stack([D["geo_" + n] for n in shc.BASE_NAMES], axis=2).reshape(-1, 8, len(D["theta"]))[:2].copy()
    base[1, 4, 100] = np.nan                               # gds2 of line 1
    th0 = np.tile(np.linspace(0.0, 1.0, 3), (2, 1))
    theta = D["theta"]
    R = shc.host_scan_solve(lib, base, D["dPdrho"].reshape(-1)[:2], th0, theta[1] - theta[0])
    assert np.all((R["info"][0] >> 16) == 0) and np.all(np.isfinite(R["lam"][0]))
    assert np.all((R["info"][1] >> 16) == 2) and np.all(np.isnan(R["lam"][1]))
    assert np.all(R["X"][3:] == 0.0)


def test_lane_code_off_centre_modes(harness):
    """s-alpha with theta0 up to 20: the eigenfunction peaks up to 600 rows away from the middle of the line
    (the matching row has to follow it)."""
    from oracle import ballooning_oracle as bo
    shc, lib = harness
    theta = np.linspace(-10 * np.pi, 10 * np.pi, 2049)
    th0 = np.linspace(0.0, 20.0, 5)
    base, dP = s_alpha_base(0.8, 0.8, theta)
    R = shc.host_scan_solve(lib, base[None], np.array([dP]), th0[None], theta[1] - theta[0])
    assert np.all((R["info"] >> 16) == 0)
    for t, t0 in enumerate(th0):
        gam, X, dX, *_ = bo.gamma_ball_full(dP, theta, base[0], base[1], base[2] + t0 * base[3],
                                            base[4] + 2 * t0 * base[5] + t0 ** 2 * base[6], method="lambda_max")
        X = X * np.sign(X[np.argmax(np.abs(X))])
        assert abs(R["lam"][0, t] - gam) <= LAM_RTOL * abs(gam)
        np.testing.assert_allclose(R["X"][t], X, rtol=0, atol=X_ATOL)


@pytest.mark.parametrize("n", [65, 129, 257])
def test_lane_code_small_grids(harness, n):
    """N = 65 has no coarse level (cold start on the fine grid), 129 one, 257 two."""
    from oracle import ballooning_oracle as bo
    shc, lib = harness
    theta = np.linspace(-2 * np.pi, 2 * np.pi, n)
    th0 = np.linspace(0.0, 1.0, 4)
    base, dP = s_alpha_base(0.8, 0.9, theta)
    R = shc.host_scan_solve(lib, base[None], np.array([dP]), th0[None], theta[1] - theta[0])
    assert lib.scan_host_num_levels(n) == {65: 0, 129: 1, 257: 2}[n]
    assert np.all((R["info"] >> 16) == 0)
    for t, t0 in enumerate(th0):
        gam, X, dX, *_ = bo.gamma_ball_full(dP, theta, base[0], base[1], base[2] + t0 * base[3],
                                            base[4] + 2 * t0 * base[5] + t0 ** 2 * base[6], method="lambda_max")
        X = X * np.sign(X[np.argmax(np.abs(X))])
        assert abs(R["lam"][0, t] - gam) <= LAM_RTOL * abs(gam)
        np.testing.assert_allclose(R["X"][t], X, rtol=0, atol=X_ATOL)


def test_lane_per_chain_code_matches_oracle_and_two_chain_code(harness):
    """The lane code of scan2_solve_kernel (one chain per lane, mirrored backward chain, chains joined at a row that is a
    multiple of 16) on D3D- and NCSX-like lines of the bench's synthetic equilibria: against the oracle (LAPACK route) and
    against the two-chains-per-lane code (same pencil, same shift sequence -> agreement to rounding)."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    import bench
    from oracle import ballooning_oracle as bo
    shc, lib = harness
    for wl, alpha0, N in (("d3d", 0.0, 1025), ("ncsx", 0.7, 513)):
        st, _, _, _ = bench.build_tables(wl, 1, 0)
        theta = np.linspace(-4 * np.pi, 4 * np.pi, N)
        assert lib.scan2_host_size_ok(N)
        rng = np.random.default_rng(1)
        lines = rng.choice(st.ns, 3, replace=False)
        base, dP = [], []
        for js in lines:
            fl = bo.fieldlines(st.select([js]), np.array([alpha0]), theta)
            base.append(np.stack([getattr(fl, n)[0][0] for n in shc.BASE_NAMES])); dP.append(bo.dpdrho_of(fl))
        base, dP = np.stack(base), np.array(dP)
        th0 = np.tile(np.array([0.0, 0.6, 1.3]), (len(lines), 1))
        h = theta[1] - theta[0]
        A = shc.host_scan_solve(lib, base, dP, th0, h, kernel="scan")
        B = shc.host_scan_solve(lib, base, dP, th0, h, kernel="scan2")
        assert np.all((B["info"] >> 16) == 0)
        np.testing.assert_allclose(B["lam"], A["lam"], rtol=1e-12)
        np.testing.assert_allclose(B["X"], A["X"], rtol=0, atol=1e-10)
        for li in range(len(lines)):
            for t in range(th0.shape[1]):
                cv = base[li][2] + th0[li, t] * base[li][3]
                gd = base[li][4] + 2 * th0[li, t] * base[li][5] + th0[li, t] ** 2 * base[li][6]
                gam, X, dX, *_ = bo.gamma_ball_full(dP[li], theta, base[li][0], base[li][1], cv, gd, method="lambda_max")
                sg = np.sign(X[np.argmax(np.abs(X))])
                s = li * th0.shape[1] + t
                assert abs(B["lam"][li, t] - gam) <= LAM_RTOL * abs(gam)
                np.testing.assert_allclose(B["X"][s], X * sg, rtol=0, atol=X_ATOL)
                np.testing.assert_allclose(B["dX"][s], dX * sg, rtol=0, atol=10 * X_ATOL * max(1.0, np.max(np.abs(dX))))
    assert not lib.scan2_host_size_ok(969) and not lib.scan2_host_size_ok(1024)


def test_lane_per_chain_warm_start_chain(harness, monkeypatch):
    """The warm start of scan2_solve_kernel (every line starts each level from the level eigenvalues of the previous line --
    the adjacent surface / alpha of a scan grid -- at the same theta0) must change nothing but the number of passes: lambda and
    X agree with the cold-started solves to rounding and with the oracle to the parity tolerances.  Adjacent surfaces of the
    tokamak case are ~1 % apart in lambda: the warm lines need far fewer passes; adjacent alphas of the stellarator case are
    ~30 % apart: the distance guard (WARM_FAR) turns the warm start off after the first line and nothing is lost."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    import bench
    from oracle import ballooning_oracle as bo
    shc, lib = harness
    for wl, N in (("d3d", 1025), ("ncsx", 513)):
        st, alphas, _, _ = bench.build_tables(wl, 1, 0)
        theta = np.linspace(-4 * np.pi, 4 * np.pi, N)
        if wl == "d3d":         # adjacent surfaces, one alpha
            fl = bo.fieldlines(st.select(np.arange(38, 43)), np.array([0.0]), theta)
        else:                   # one surface, adjacent alphas
            fl = bo.fieldlines(st.select([30]), alphas[8:13], theta)
        base = np.stack([getattr(fl, n) for n in shc.BASE_NAMES], axis=2).reshape(-1, len(shc.BASE_NAMES), N)
        ns_, na_ = fl.bmag.shape[:2]
        dP = np.array([bo.dpdrho_of(fl, js, ja) for js in range(ns_) for ja in range(na_)])
        nline = base.shape[0]
        th0 = np.tile(np.linspace(0.0, 0.5 * np.pi, 20), (nline, 1))
        h = theta[1] - theta[0]
        monkeypatch.setenv("IBS_SCAN_WARM", "1")
        W = shc.host_scan_solve(lib, base, dP, th0, h, kernel="scan2")
        monkeypatch.setenv("IBS_SCAN_WARM", "0")
        C = shc.host_scan_solve(lib, base, dP, th0, h, kernel="scan2")
        assert np.all((W["info"] >> 16) == 0) and np.all((C["info"] >> 16) == 0)
        if wl == "d3d":
            assert W["passes"] < 0.7 * C["passes"], (W["passes"], C["passes"])
        else:
            assert W["passes"] < 1.03 * C["passes"], (W["passes"], C["passes"])
        np.testing.assert_allclose(W["lam"], C["lam"], rtol=1e-12)
        # (the matrix eigenvalue is converged to ~1e-15 of the spectrum's scale U, whatever |lambda| is)
        np.testing.assert_allclose(W["lam_matrix"], C["lam_matrix"], rtol=1e-11, atol=1e-13 * np.abs(W["bounds"][:, 0]).max())
        np.testing.assert_allclose(W["X"], C["X"], rtol=0, atol=1e-9)
        for (li, t) in ((1, 0), (2, 7), (4, 19), (3, 11)):
            b = base[li]
            cv = b[2] + th0[li, t] * b[3]
            gd = b[4] + 2 * th0[li, t] * b[5] + th0[li, t] ** 2 * b[6]
            info = {}
            gam, X, dX, *_ = bo.gamma_ball_full(dP[li], theta, b[0], b[1], cv, gd, method="lambda_max", info=info)
            sg = np.sign(X[np.argmax(np.abs(X))])
            # (near marginal stability |lambda| << U: LAPACK itself carries ~eps U there, so the relative bar gets an absolute floor)
            assert abs(W["lam"][li, t] - gam) <= LAM_RTOL * abs(gam) + 1e-13 * abs(W["bounds"][li, 0])
            if info["gap"] > 1e-6:
                np.testing.assert_allclose(W["X"][li * th0.shape[1] + t], X * sg, rtol=0, atol=X_ATOL)
