"""GPU parity of K2+K3 (+K4) against the converged reference (golden fixtures) -- through the C ABI."""
import numpy as np
import pytest

from helpers import LAM_RTOL, X_ATOL, fixture_base, fixture_gcf, sign_normalise

pytestmark = pytest.mark.gpu


def _engine():
    from ideal_ballooning_solver_b200 import engine
    return engine


@pytest.mark.parametrize("nth", [1024, 2048, 512])
def test_s_alpha_gamma_ball_full(cuda_lib, golden, nth):
    import torch
    from ideal_ballooning_solver_b200 import synthetic
    eng = _engine()
    G = golden("s_alpha")
    theta = G[f"theta_{nth}"]
    cases = G["cases"]
    g, c, f = synthetic.s_alpha_coefficients(cases[:, 0], cases[:, 1], cases[:, 2], theta)
    sol = eng.solve_gcf_batch(torch.from_numpy(g).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda(),
                              eng.grid_spacing(theta), sigma=torch.full((len(cases),), 2.0, dtype=torch.float64))
    lam = sol.lam.cpu().numpy()
    ref = G[f"lam_conv_{nth}"]
    assert np.all(sol.flags.cpu().numpy() == 0), sol.info
    np.testing.assert_allclose(lam, ref, rtol=LAM_RTOL, atol=0)
    # stable / unstable classification is bit-exact
    assert np.array_equal(lam > 0, ref > 0)
    X = sol.X.cpu().numpy()
    assert np.all(X >= 0) and np.allclose(X.max(axis=1), 1.0, rtol=0, atol=0)
    np.testing.assert_allclose(X, sign_normalise(G[f"X_conv_{nth}"]), rtol=0, atol=X_ATOL)
    sgn = np.sign(G[f"X_conv_{nth}"][np.arange(len(cases)), np.argmax(np.abs(G[f"X_conv_{nth}"]), axis=1)])
    np.testing.assert_allclose(sol.dX.cpu().numpy(), G[f"dX_conv_{nth}"] * sgn[:, None], rtol=0, atol=10 * X_ATOL)
    assert sol.iterations.max().item() < 40


@pytest.mark.parametrize("name", ["ncsx_wout_op", "synthetic_ncsx", "synthetic_d3d", "synthetic_hberg"])
def test_equilibrium_fixture_solves(cuda_lib, golden, name):
    """gamma_ball_full on reference-generated geometry: both coefficient paths (explicit g,c,f and
    the fused base-array path) against the converged reference."""
    import torch
    eng = _engine()
    D = golden(name)
    theta = D["theta"]
    h = eng.grid_spacing(theta)
    ns, na, nt = D["lam_conv"].shape
    gs, cs, fs = [], [], []
    for i in range(ns):
        for j in range(na):
            for k in range(nt):
                g, c, f = fixture_gcf(D, i, j, k)
                gs.append(g); cs.append(c); fs.append(f)
    g, c, f = (torch.from_numpy(np.array(a)).cuda() for a in (gs, cs, fs))
    sol = eng.solve_gcf_batch(g, c, f, h)
    ref = D["lam_conv"].reshape(-1)
    assert np.all(sol.flags.cpu().numpy() == 0)
    np.testing.assert_allclose(sol.lam.cpu().numpy(), ref, rtol=LAM_RTOL, atol=0)
    Xref = sign_normalise(D["X_conv"].reshape(-1, len(theta)))
    np.testing.assert_allclose(sol.X.cpu().numpy(), Xref, rtol=0, atol=X_ATOL)
    # fused base-array path: bit-identical coefficients, identical results
    base = torch.from_numpy(fixture_base(D)).cuda()
    dP = torch.from_numpy(D["dPdrho"]).cuda()
    th0 = torch.from_numpy(np.tile(D["theta0s"], ns * na)).cuda()
    sol2 = eng.solve_base_batch(base, dP, th0, h, nth0=nt, want_gcf=True)
    assert torch.equal(sol2.g, g) and torch.equal(sol2.c, c) and torch.equal(sol2.f, f)
    assert torch.equal(sol2.lam, sol.lam) and torch.equal(sol2.X, sol.X)


def test_count_above_matches_reference_check_ball(cuda_lib, golden):
    """The reference's own s-alpha test (Newcomb shooting at lambda=0, check_ball / check_ball_long)."""
    import torch
    from ideal_ballooning_solver_b200 import synthetic
    eng = _engine()
    G = golden("s_alpha")
    pts, t0s = G["cb_points"], G["cb_theta0"]
    for key, span, n in (("check_ball", 61, 1601), ("check_ball_long", 20, 401)):
        theta = np.linspace(-span * np.pi, span * np.pi, n)
        sh = np.repeat(pts[:, 0], len(t0s)); al = np.repeat(pts[:, 1], len(t0s)); t0 = np.tile(t0s, len(pts))
        g, c, _ = synthetic.s_alpha_coefficients(sh, al, t0, theta)
        f = np.ones_like(g)
        cnt = eng.count_above_batch(torch.from_numpy(g).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda(),
                                    theta[1] - theta[0], torch.zeros(len(sh), dtype=torch.float64))
        unstable = (cnt.cpu().numpy() > 0).astype(int).reshape(len(pts), len(t0s))
        assert np.array_equal(unstable, G[key]), key


def test_count_above_is_sturm_count(cuda_lib, golden):
    import torch
    from scipy.linalg import eigh_tridiagonal
    from oracle import ballooning_oracle as bo
    from ideal_ballooning_solver_b200 import synthetic
    eng = _engine()
    rng = np.random.default_rng(3)
    theta = np.linspace(-6 * np.pi, 6 * np.pi, 700)       # even N, ragged chunking
    g, c, f = synthetic.s_alpha_coefficients(0.7, 0.9, 0.1, theta)
    _, _, _, fu, sub, diag, sup = bo.discretise(theta, g, c, f)
    fi = fu[1:-1]
    ev = eigh_tridiagonal(diag, sup * np.sqrt(fi[:-1] / fi[1:]), eigvals_only=True)
    lam = np.concatenate([rng.uniform(ev[-30], ev[-1] + 0.1, 200), ev[-6:] + 1e-9, ev[-6:] - 1e-9])
    rep = lambda a: torch.from_numpy(np.tile(a, (len(lam), 1))).cuda()
    cnt = eng.count_above_batch(rep(g), rep(c), rep(f), eng.grid_spacing(theta), torch.from_numpy(lam)).cpu().numpy()
    assert np.array_equal(cnt, (ev[None, :] > lam[:, None]).sum(axis=1))


@pytest.mark.parametrize("N", [5, 9, 33, 34, 65, 100, 257, 969, 1025, 1500, 2049, 4097, 8193])
def test_sizes_against_oracle(cuda_lib, N):
    """Edge sizes (tiny, ragged, even, every team width) against the LAPACK cross-check of the oracle."""
    import torch
    from oracle import ballooning_oracle as bo
    from ideal_ballooning_solver_b200 import synthetic
    eng = _engine()
    theta = np.linspace(-5 * np.pi, 5 * np.pi, N)
    cases = [(0.8, 0.8, 0.0), (1.2, 0.2, 0.1), (0.4, 0.5, 0.2)]
    g, c, f = synthetic.s_alpha_coefficients([a[0] for a in cases], [a[1] for a in cases], [a[2] for a in cases], theta)
    f = f * (1.5 + np.cos(theta))          # f != g
    sol = eng.solve_gcf_batch(torch.from_numpy(g).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda(),
                              eng.grid_spacing(theta))
    assert np.all(sol.flags.cpu().numpy() == 0)
    for k in range(len(cases)):
        one = np.ones_like(theta)
        # gamma_ball_full(dPdrho=-1, B=1, gradpar=1) gives g = gds2, c = cvdrift, f = gds2 -> feed (g,c,f) directly
        h, gu, cu, fu, sub, diag, sup = bo.discretise(theta, g[k], c[k], f[k])
        from scipy.linalg import eigh_tridiagonal
        fi = fu[1:-1]
        w, U = eigh_tridiagonal(diag, sup * np.sqrt(fi[:-1] / fi[1:]), select="i", select_range=(N - 3, N - 3))
        gam, X, dX = bo.postprocess(U[:, 0] / np.sqrt(fi), h, gu, cu, fu)
        # LAPACK's own eigenvalue error is ~eps*||S||; the flux-form recurrence is more accurate than that
        np.testing.assert_allclose(sol.lam_matrix[k].item(), w[0], rtol=1e-11, atol=8 * 2.2e-16 * np.abs(diag).max())
        if N >= 9:
            np.testing.assert_allclose(sol.lam[k].item(), gam, rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(sol.X[k].cpu().numpy(), sign_normalise(X), rtol=0, atol=1e-7)


def test_bad_input_is_flagged_not_silent(cuda_lib):
    import torch
    from ideal_ballooning_solver_b200 import synthetic
    eng = _engine()
    theta = np.linspace(-4 * np.pi, 4 * np.pi, 129)
    g, c, f = synthetic.s_alpha_coefficients([0.8, 0.8, 0.8], [0.8, 0.8, 0.8], [0.0, 0.0, 0.0], theta)
    f[1, 40] = -1.0
    g[2, 17] = np.nan
    sol = eng.solve_gcf_batch(torch.from_numpy(g).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda(),
                              eng.grid_spacing(theta))
    fl = sol.flags.cpu().numpy()
    assert fl[0] == 0 and fl[1] & eng.FLAG_BAD_INPUT and fl[2] & eng.FLAG_BAD_INPUT
    lam = sol.lam.cpu().numpy()
    assert np.isfinite(lam[0]) and np.isnan(lam[1]) and np.isnan(lam[2])


def test_empty_batch(cuda_lib):
    import torch
    eng = _engine()
    z = torch.zeros((0, 65), dtype=torch.float64, device="cuda")
    sol = eng.solve_gcf_batch(z, z, z, 0.1)
    assert sol.lam.numel() == 0
