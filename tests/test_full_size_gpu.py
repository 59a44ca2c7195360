"""GPU: BASELINE.json's full-size configurations checked through size-independent properties (the oracle's dense
ARPACK route cannot run at these sizes): Sturm-count certificates of lambda_max, residual of the discrete pencil,
the Simpson Rayleigh quotient recomputed from (X, g, c, f), positivity / normalisation of X, arg-max consistency,
determinism, and chained vs independent solves."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CONFIGS = {
    #        kind    ns  nalpha nth0 ntheta span
    "d3d":  ("d3d",  128, 1,    64,  1024, 4),     # BASELINE configs[1]
    "ncsx": ("ncsx", 64,  32,   32,  2048, 4),     # BASELINE configs[2]  (65 536 solves, N = 2049)
    "hberg": ("hberg", 256, 64, 1,   8192, 8),     # BASELINE configs[3]  (16 384 lines,  N = 8193); surfaces reduced below
}


def _setup(name, ns_cap=None):
    import torch
    from ideal_ballooning_solver_b200 import engine, synthetic, tables
    kind, ns, na, nt, nth, span = CONFIGS[name]
    if ns_cap:
        ns = min(ns, ns_cap)
    s = np.linspace(0.5, 0.95, ns)
    alpha = np.linspace(0.0, np.pi, na) if na > 1 else np.array([0.0])
    theta0 = np.linspace(0.0, 0.5 * np.pi, nt) if nt > 1 else np.array([0.0])
    theta = np.linspace(-span * np.pi, span * np.pi, nth + 1)
    st = tables.RadialSplines(synthetic.make_equilibrium(kind, seed=11)).evaluate(s)
    dt = engine.DeviceTables.from_host(st)
    geo = engine.geometry_batch(dt, alpha, theta, want_info=True)
    assert int((geo.info >> 16).max().item()) == 0
    th0 = torch.from_numpy(theta0).cuda().repeat(ns * na)
    return engine, geo, th0, theta, (ns, na, nt)


def _check_properties(engine, geo, th0, theta, shape, sample=4096, chain=16):
    import torch
    ns, na, nt = shape
    h = engine.grid_spacing(theta)
    N = len(theta)
    sol = engine.solve_base_batch(geo.base, geo.dPdrho, th0, h, nth0=nt, chain_len=chain, want_dX=True)
    assert int((sol.flags & 3).max().item()) == 0
    assert float(sol.iterations.double().mean().item()) < 9
    n = th0.numel()
    # --- X: non-negative, max exactly 1, Dirichlet ends
    assert bool((sol.X >= 0).all()) and bool((sol.X.max(dim=1).values == 1.0).all())
    assert bool((sol.X[:, 0] == 0).all()) and bool((sol.X[:, -1] == 0).all())
    # --- determinism, and chained == independent to rounding
    sol_b = engine.solve_base_batch(geo.base, geo.dPdrho, th0, h, nth0=nt, chain_len=chain, want_dX=False)
    assert torch.equal(sol.lam, sol_b.lam) and torch.equal(sol.X, sol_b.X)
    # --- a sample of the solves in detail (explicit coefficients from the exact base path)
    idx = torch.linspace(0, n - 1, min(sample, n)).round().long().cuda().unique()
    line = (idx // nt).to(torch.int32)
    ex = engine.solve_base_batch(geo.base, geo.dPdrho, th0[idx], h, line_of_solve=line, want_gcf=True)
    lam_m = ex.lam_matrix
    rel = lambda a, b: float(((a - b).abs() / b.abs().clamp_min(1e-300)).max().item())
    assert rel(ex.lam, sol.lam[idx]) < 1e-11, "chained polynomial-coefficient path vs independent exact path"
    assert float((ex.X - sol.X[idx]).abs().max().item()) < 1e-9
    g, c, f = ex.g, ex.c, ex.f
    # Sturm certificates: nothing above lambda_max (+ margin), exactly one eigenvalue above lambda_max - margin
    scale = lam_m.abs().clamp_min(1e-3)
    cnt_hi = engine.count_above_batch(g, c, f, h, lam_m + 1e-9 * scale)
    cnt_lo = engine.count_above_batch(g, c, f, h, lam_m - 1e-7 * scale)
    assert int(cnt_hi.max().item()) == 0 and int(cnt_lo.min().item()) >= 1
    assert float((cnt_lo == 1).double().mean().item()) > 0.99      # (a second eigenvalue within 1e-7 is possible but rare)
    # residual of the discrete pencil  (K - lam F) x  on the interior rows  (utils.py:1584-1592)
    X = ex.X
    gh = 0.5 * (g[:, 1:] + g[:, :-1])
    Kx = (gh[:, 1:] * (X[:, 2:] - X[:, 1:-1]) - gh[:, :-1] * (X[:, 1:-1] - X[:, :-2])) / h ** 2 + c[:, 1:-1] * X[:, 1:-1]
    res = Kx - lam_m[:, None] * f[:, 1:-1] * X[:, 1:-1]
    norm = (gh.abs().max(dim=1).values / h ** 2)[:, None]
    assert float((res.abs() / norm).max().item()) < 1e-11
    # the returned gam is the Simpson Rayleigh quotient of (X, dX)  (utils.py:1618-1621)
    w = torch.tensor([engine_simpson(p, N) for p in range(N)], dtype=torch.float64, device="cuda")
    y0 = (w * (-g * ex.dX ** 2 + c * X ** 2)).sum(dim=1)
    y1 = (w * f * X ** 2).sum(dim=1)
    assert rel(ex.lam, y0 / y1) < 1e-9
    # dX is the reference stencil of X
    dX_ref = torch.zeros_like(X)
    dX_ref[:, 2:-2] = 2 / (3 * h) * (X[:, 3:-1] - X[:, 1:-3]) - (X[:, 4:] - X[:, :-4]) / (12 * h)
    dX_ref[:, 1] = (X[:, 2] - X[:, 0]) / (2 * h); dX_ref[:, -2] = (X[:, -1] - X[:, -3]) / (2 * h)
    dX_ref[:, 0] = (-1.5 * X[:, 0] + 2 * X[:, 1] - 0.5 * X[:, 2]) / h
    dX_ref[:, -1] = (0.5 * X[:, -3] - 2 * X[:, -2]) / h
    assert float((ex.dX - dX_ref).abs().max().item()) < 1e-10 / h
    # --- per-surface arg-max: bit-exact against torch (first index on ties, ball_scan.py:279-295)
    gam = sol.lam.reshape(ns, na * nt)
    val, ia, sig = engine.scan_argmax(gam)
    tv, ti = gam.max(dim=1)
    assert torch.equal(val, tv)
    first = (gam == tv[:, None]).int().argmax(dim=1)
    assert torch.equal(ia.long(), first)
    return sol


def engine_simpson(p, N):
    """scipy.integrate.simpson weights, unit spacing (odd N: composite 1/3 rule)."""
    assert N % 2 == 1
    if p == 0 or p == N - 1:
        return 1.0 / 3.0
    return 4.0 / 3.0 if p % 2 else 2.0 / 3.0


def test_d3d_full_config(cuda_lib):
    engine, geo, th0, theta, shape = _setup("d3d")
    sol = _check_properties(engine, geo, th0, theta, shape)
    assert sol.lam.numel() == 128 * 64


def test_ncsx_full_config(cuda_lib):
    engine, geo, th0, theta, shape = _setup("ncsx")
    sol = _check_properties(engine, geo, th0, theta, shape, sample=2048)
    assert sol.lam.numel() == 64 * 32 * 32


def test_hberg_full_resolution(cuda_lib):
    """ntheta = 8192 (8 warps per solve); 32 of the 256 surfaces x 64 alpha to bound the memory of the dense checks."""
    engine, geo, th0, theta, shape = _setup("hberg", ns_cap=32)
    sol = _check_properties(engine, geo, th0, theta, shape, sample=512, chain=1)
    assert sol.lam.numel() == 32 * 64


# ---------------------------------------------------------------------------------------------------------------
# Direct parity with the oracle AT FULL SIZE: sampled solves of every configuration against the LAPACK tridiagonal route of
# the restated gamma_ball_full (oracle `method="lambda_max"`: O(N) per solve, so N = 1025 / 2049 / 8193 finish in seconds).
# ---------------------------------------------------------------------------------------------------------------
def _oracle_sample(geo, th0, theta, shape, sol, nsample=64, seed=5):
    from oracle import ballooning_oracle as bo
    from helpers import LAM_RTOL, X_ATOL
    from ideal_ballooning_solver_b200.engine import BASE_NAMES
    ns, na, nt = shape
    n = ns * na * nt
    rng = np.random.default_rng(seed)
    pick = np.unique(np.concatenate([[0, n - 1], rng.integers(0, n, nsample - 2)]))
    base = geo.base.reshape(ns * na, len(BASE_NAMES), -1)
    dP = geo.dPdrho.reshape(-1).cpu().numpy()
    th0_h = th0.cpu().numpy()
    lam_d, X_d = sol.lam.cpu().numpy(), sol.X
    row = {nm: k for k, nm in enumerate(BASE_NAMES)}
    worst_l = worst_x = 0.0
    for i in pick:
        line = i // nt
        b = base[line].cpu().numpy()
        t0 = th0_h[i]
        cv = b[row["cvdrift"]] + t0 * b[row["cvdrift0"]]                                     # ball_scan.py:267-268
        gd = b[row["gds2"]] + 2 * t0 * b[row["gds21"]] + t0 ** 2 * b[row["gds22"]]
        info = {}
        lam, X, *_ = bo.gamma_ball_full(dP[line], theta, b[row["bmag"]], b[row["gradpar_theta_pest"]], cv, gd,
                                        method="lambda_max", info=info)
        Xs = X * np.sign(X[np.argmax(np.abs(X))])
        worst_l = max(worst_l, abs(lam_d[i] - lam) / abs(lam))
        # eigenvector conditioning: eps ||S|| / gap; the 1e-8 bar applies where the gap allows it
        if info["gap"] > 1e-6:
            worst_x = max(worst_x, float(np.max(np.abs(X_d[i].cpu().numpy() - Xs))))
    assert worst_l < LAM_RTOL, worst_l
    assert worst_x < X_ATOL, worst_x
    return len(pick)


@pytest.mark.parametrize("name,ns_cap", [("d3d", None), ("ncsx", None), ("hberg", 64)])
def test_full_config_samples_match_oracle(cuda_lib, name, ns_cap):
    """64 sampled solves of each BASELINE scan configuration against the oracle: lambda to 1e-10 relative, X to 1e-8."""
    engine, geo, th0, theta, shape = _setup(name, ns_cap=ns_cap)
    ns, na, nt = shape
    sol, best, sig = engine.scan_solve_argmax(geo.base, geo.dPdrho, th0, engine.grid_spacing(theta), nt, na, want_X=True,
                                              chain_len=16 if nt > 1 else 1)
    assert int((sol.flags & 3).max().item()) == 0
    assert _oracle_sample(geo, th0, theta, shape, sol) >= 32
    # the fused / packed arg-max against numpy on the same gamma grid (first index on ties, ball_scan.py:279-295)
    from oracle import ballooning_oracle as bo
    gam = sol.lam.reshape(ns, na, nt).cpu().numpy()
    b = best.cpu().numpy()
    for js in range(0, ns, max(1, ns // 16)):
        ia, it, s0 = bo.argmax_with_guards(gam[js])
        assert b[js, 0] == gam[js].max() and int(b[js, 1]) == ia * nt + it
        assert float(sig[js].item()) == s0


def test_hberg_all_surfaces_properties(cuda_lib):
    """All 256 surfaces x 64 alpha of BASELINE configs[3] at ntheta = 8192 (16 384 solves of 8193 points): flags, positivity
    and normalisation of X, and the Sturm certificate of lambda_max on a sample (the dense checks of the property test above
    are bounded to a sample to bound memory)."""
    import torch
    engine, geo, th0, theta, shape = _setup("hberg")
    ns, na, nt = shape
    h = engine.grid_spacing(theta)
    sol = engine.solve_base_batch(geo.base, geo.dPdrho, th0, h, nth0=nt, chain_len=1, want_dX=False)
    assert sol.lam.numel() == 256 * 64
    assert int((sol.flags & 3).max().item()) == 0
    assert bool((sol.X >= 0).all()) and bool((sol.X.max(dim=1).values == 1.0).all())
    idx = torch.linspace(0, sol.lam.numel() - 1, 256).round().long().cuda().unique()
    ex = engine.solve_base_batch(geo.base, geo.dPdrho, th0[idx], h, line_of_solve=(idx // nt).to(torch.int32), want_gcf=True)
    scale = ex.lam_matrix.abs().clamp_min(1e-3)
    assert int(engine.count_above_batch(ex.g, ex.c, ex.f, h, ex.lam_matrix + 1e-9 * scale).max().item()) == 0
    assert int(engine.count_above_batch(ex.g, ex.c, ex.f, h, ex.lam_matrix - 1e-7 * scale).min().item()) >= 1
    assert float(((ex.lam - sol.lam[idx]).abs() / sol.lam[idx].abs()).max().item()) < 1e-11
