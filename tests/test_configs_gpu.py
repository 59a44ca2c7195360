"""GPU: BASELINE.json configs[0] (s-alpha scan) and configs[4] (adjoint gradient batch) at their stated sizes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_s_alpha_grid_classification(cuda_lib):
    """The 200 x 100 x 3 (shat, alpha, theta0) grid of bishop_ball_s-alpha.py:213-229, classified by the Sturm count at
    lambda = 0 on the reference's own grid (theta in +-61 pi, 1601 points): bit-exact against the restated check_ball on a
    sub-grid, the OR over theta0 of :282-289, the labelled regions of the Bishop/CHT figure (:297-299), and consistency
    with the sign of lambda_max from the eigen-solver."""
    import torch
    from oracle import ballooning_oracle as bo
    from ideal_ballooning_solver_b200 import engine
    shat = np.linspace(0.0, 2.0, 200); alpha = np.linspace(0.0, 1.2, 100); theta0 = np.array([0.0, 0.1, 0.2])
    theta = np.linspace(-61 * np.pi, 61 * np.pi, 1601)
    th = torch.from_numpy(theta).cuda()
    S, A, T = np.meshgrid(shat, alpha, theta0, indexing="ij")
    p = torch.from_numpy(np.stack([S.ravel(), A.ravel(), T.ravel()], 1)).cuda()
    sh, al, t0 = p[:, 0:1], p[:, 1:2], p[:, 2:3]
    L = sh * (th - t0) - al * (torch.sin(th) - torch.sin(t0))
    g = 1.0 + L * L
    c = al * (torch.cos(th) + torch.sin(th) * L)
    f = torch.ones_like(g)
    h = float(theta[1] - theta[0])
    cnt = engine.count_above_batch(g, c, f, h, torch.zeros(p.shape[0], dtype=torch.float64))
    unstable = (cnt > 0).reshape(200, 100, 3).cpu().numpy()
    # --- bit-exact against the reference's classifier on a sub-grid (every theta0)
    ii = np.linspace(1, 199, 12).astype(int); jj = np.linspace(1, 99, 8).astype(int)
    for i in ii:
        for j in jj:
            for k in range(3):
                assert int(unstable[i, j, k]) == bo.check_ball(shat[i], alpha[j], theta0[k]), (i, j, k)
    # --- OR over theta0 (bishop_ball_s-alpha.py:282-289) and the labelled regions of the figure (:297-299)
    region = unstable.any(axis=2)
    at = lambda s_, a_: region[np.argmin(np.abs(shat - s_)), np.argmin(np.abs(alpha - a_))]
    assert at(0.8, 0.8) and not at(1.2, 0.2) and not at(0.05, 1.0)
    assert 0.05 < region.mean() < 0.6
    # --- the eigen-solver's sign(lambda_max) is the same classifier (Sylvester: the count at 0 does not depend on f > 0)
    sub = torch.arange(0, p.shape[0], 7, device="cuda")
    sol = engine.solve_gcf_batch(g[sub], c[sub], g[sub], engine.grid_spacing(theta), want_dX=False)
    assert int((sol.flags & 3).max().item()) == 0
    lam = sol.lam_matrix
    clear = lam.abs() > 1e-9
    assert bool(((lam > 0) == (cnt[sub] > 0))[clear].all())


def test_adjoint_batch_config5(cuda_lib):
    """4096 NCSX-like points with (s, alpha, theta0) i.i.d. uniform (BASELINE configs[4], a slice of the per-GPU share):
    every point through K1 (three field lines) + K2/K3 + K4; eight of them against the restated obj_w_grad."""
    import torch
    from oracle import ballooning_oracle as bo
    from helpers import LAM_RTOL, X_ATOL
    from ideal_ballooning_solver_b200 import engine, synthetic, tables
    npts = 4096
    rng = np.random.default_rng(20261018 + 5)
    s = rng.uniform(0.5, 0.95, npts); al = rng.uniform(0, np.pi, npts); t0 = rng.uniform(0, 0.5 * np.pi, npts)
    theta = np.linspace(-4 * np.pi, 4 * np.pi, 1025)
    spl = tables.RadialSplines(synthetic.make_equilibrium("ncsx", seed=1))
    st = spl.evaluate(s)
    dt = engine.DeviceTables.from_host(st)
    d = 0.004
    alphas = torch.from_numpy(np.stack([al - 0.5 * d, al, al + 0.5 * d], axis=1)).cuda()
    geo = engine.geometry_batch(dt, alphas, theta, want_info=True)
    assert int((geo.info >> 16).max().item()) == 0
    h = engine.grid_spacing(theta)
    val, grad, X, dX, info = engine.obj_w_grad_batch(geo.base, geo.dPdrho, torch.from_numpy(t0).cuda(), h, del_alpha=d, want_X=True)
    assert int(((info >> 16) & 3).max().item()) == 0
    assert bool(torch.isfinite(val).all()) and bool(torch.isfinite(grad).all())
    # the objective is -lambda of the centre line (utils.py:1728)
    sol = engine.solve_base_batch(geo.base[:, 1], geo.dPdrho[:, 1], torch.from_numpy(t0).cuda(), h, nth0=1, want_dX=False)
    assert float(((val + sol.lam).abs() / sol.lam.abs()).max().item()) < 1e-11
    vguess = bo.default_vguess(theta)
    for i in np.linspace(0, npts - 1, 8).astype(int):
        st1 = st.select([i])
        v, gr = bo.obj_w_grad((al[i], t0[i]), lambda a: bo.fieldlines(st1, np.atleast_1d(a), theta), theta, vguess, 1.0,
                              method="lambda_max")
        assert abs(float(val[i].item()) - v) <= LAM_RTOL * abs(v) * 10 + 1e-14, (i, float(val[i].item()), v)
        assert np.max(np.abs(grad[i].cpu().numpy() - gr)) < X_ATOL, (i, grad[i].cpu().numpy(), gr)
