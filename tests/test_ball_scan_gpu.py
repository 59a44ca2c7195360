"""GPU: the drop-in scan driver (coarse grid -> guarded arg-max -> L-BFGS-B refine with the adjoint gradient ->
eigenpair at the optimum, ball_scan.py:196-339) against the same steps done with the CPU oracle + scipy."""
import numpy as np
import pytest

from helpers import tables_from_fixture

pytestmark = pytest.mark.gpu


def _oracle_ball_scan_surface(tab, theta, alpha_scan, theta0_scan):
    """ball_scan.py:248-339 for one surface with the oracle (LAPACK route = converged reference)."""
    from scipy.optimize import minimize
    from oracle import ballooning_oracle as bo
    fl_fn = lambda alphas: bo.fieldlines(tab, alphas, theta)
    gam = bo.coarse_scan_surface(fl_fn, theta, alpha_scan, theta0_scan, method="lambda_max")
    ia, it, sigma0 = bo.argmax_with_guards(gam)
    x0 = (0.0, 0.0) if ia < 0 else (alpha_scan[ia], theta0_scan[it])
    vg = bo.default_vguess(theta)
    obj = minimize(lambda x: bo.obj_w_grad(x, fl_fn, theta, vg, sigma0, method="lambda_max"), x0=x0, jac=True,
                   bounds=((0.0, np.pi), (0.0, 0.5 * np.pi)), options={"ftol": 5.0e-11, "gtol": 2.0e-08, "maxiter": 30})
    fl = fl_fn(np.array([obj.x[0]]))
    cv, gd = bo.theta0_shift(fl, obj.x[1])
    lam, X, *_ = bo.gamma_ball_full(bo.dpdrho_of(fl), theta, fl.bmag[0][0], fl.gradpar_theta_pest[0][0], cv, gd, method="lambda_max")
    return gam, obj, lam


def test_ball_scan_matches_oracle_driver(cuda_lib, golden):
    from ideal_ballooning_solver_b200 import scan
    D = golden("synthetic_d3d")
    st = tables_from_fixture(D)                                   # 3 surfaces
    theta = np.linspace(-3 * np.pi, 3 * np.pi, 193)
    na, nt = 5, 4
    alpha_scan, theta0_scan = np.linspace(0, np.pi, na), np.linspace(0, 0.5 * np.pi, nt)
    oracle = [_oracle_ball_scan_surface(st.select([js]), theta, alpha_scan, theta0_scan) for js in range(st.ns)]
    runs = {m: scan.ball_scan(st, theta=theta, nalpha_guess=na, ntheta0_guess=nt, refine_method=m) for m in ("device", "scipy")}
    for method, res in runs.items():
        assert res.refine is not None and res.refine.nbatches <= int(res.refine.nfev.max()) + 4     # evaluations were batched
        for js in range(st.ns):
            gam, obj, lam = oracle[js]
            np.testing.assert_allclose(res.gamma_coarse[js], gam, rtol=1e-9, atol=1e-13)
            # objective equal to ~1e-10 and the reference's stopping tolerances: the optimum agrees closely (scipy path: the same
            # optimiser as the oracle driver; device path: another quasi-Newton method under the same ftol / gtol / bounds)
            np.testing.assert_allclose(res.gamma[js], lam, rtol=1e-6, atol=1e-10)
            np.testing.assert_allclose([res.alpha[js], res.theta0[js]], obj.x, atol=2e-4 if method == "scipy" else 3e-3)
            assert res.gamma[js] >= gam.max() - 1e-12                  # the refinement never loses the coarse maximum
            assert 0.0 <= res.alpha[js] <= np.pi and 0.0 <= res.theta0[js] <= 0.5 * np.pi
    # the device-resident optimiser reaches at least the scipy path's optimum (to the reference's own ftol = 5e-11 on F)
    assert np.all(runs["device"].gamma >= runs["scipy"].gamma - 1e-10)


def test_refine_batches_all_surfaces(cuda_lib, golden):
    """64 surfaces refined together: a handful of batched GPU rounds instead of sum(nfev) launches."""
    from ideal_ballooning_solver_b200 import engine, scan, synthetic, tables
    st = tables.RadialSplines(synthetic.make_equilibrium("ncsx", seed=4)).evaluate(np.linspace(0.5, 0.95, 64))
    theta = np.linspace(-4 * np.pi, 4 * np.pi, 513)
    res = scan.ball_scan(st, theta=theta, nalpha_guess=6, ntheta0_guess=5)
    r = res.refine
    assert r.nbatches <= int(r.nfev.max()) + 4 and int(r.nfev.sum()) > 3 * r.nbatches
    assert np.all(r.success)
    assert np.all(res.gamma >= res.gamma_coarse.reshape(64, -1).max(axis=1) - 1e-12)
    assert np.all(np.isfinite(res.gamma)) and res.X.shape == (64, 513)
    # against the reference's optimiser (scipy L-BFGS-B, same options) on the same batched evaluations
    res_s = scan.ball_scan(st, theta=theta, nalpha_guess=6, ntheta0_guess=5, refine_method="scipy")
    assert np.all(res.gamma >= res_s.gamma - 1e-10 - 1e-6 * np.abs(res_s.gamma))
    assert np.median(np.abs(res.gamma - res_s.gamma) / np.abs(res_s.gamma)) < 1e-6


def test_hellmann_feynman_gamma_predicts_perturbed_growth_rates(cuda_lib):
    """f3, second half: the first-order change of each surface's growth rate under perturbed equilibria from one
    eigen-solve per surface (K1 on the same field lines + K4), against actually re-solving the perturbed problems."""
    import dataclasses
    import torch
    from ideal_ballooning_solver_b200 import engine, scan, synthetic, tables, penalty
    st = tables.RadialSplines(synthetic.make_equilibrium("ncsx", seed=11)).evaluate(np.linspace(0.55, 0.9, 6))
    theta = np.linspace(-3 * np.pi, 3 * np.pi, 513)
    rng = np.random.default_rng(5)
    a_star, t_star = rng.uniform(0.2, 2.5, st.ns), rng.uniform(0.0, 1.2, st.ns)
    eps = 1.0e-5
    perts = []
    for k in range(3):                      # smooth relative perturbations of the shape / field tables
        w_mn = 1.0 + eps * rng.standard_normal(st.tab_mn.shape[1:])[None] * np.exp(-0.3 * np.arange(st.tab_mn.shape[2]))[None, None]
        w_nq = 1.0 + eps * rng.standard_normal(st.tab_nyq.shape[1:])[None] * np.exp(-0.3 * np.arange(st.tab_nyq.shape[2]))[None, None]
        perts.append(dataclasses.replace(st, tab_mn=st.tab_mn * w_mn, tab_nyq=st.tab_nyq * w_nq))
    dt0 = engine.DeviceTables.from_host(st)
    dts = [engine.DeviceTables.from_host(p) for p in perts]
    sens = scan.hellmann_feynman_gamma(dt0, dts, a_star, t_star, theta)
    assert sens.dgamma.shape == (3, st.ns) and np.all(np.isfinite(sens.dgamma))
    h = engine.grid_spacing(theta)
    a = torch.from_numpy(a_star[:, None].copy()).cuda()
    t0 = torch.from_numpy(t_star.copy()).cuda()
    for k, dtp in enumerate(dts):
        geo = engine.geometry_batch(dtp, a, theta)
        lam_p = engine.solve_base_batch(geo.base, geo.dPdrho, t0, h, nth0=1, want_X=False, want_dX=False, want_matrix=False).lam.cpu().numpy()
        actual = lam_p - sens.gamma
        # first order in eps: the remainder is O(eps^2) relative to the change itself (plus rounding of the difference)
        np.testing.assert_allclose(sens.dgamma[k], actual, rtol=2e-3, atol=1e-12 * np.abs(sens.gamma).max() + 1e-14)
    f, df = penalty.hf_jacobian([0.3] * 4, sens.gamma, sens.dgamma, np.array([0.0, eps, eps, eps]), thresh=float(np.median(sens.gamma)))
    assert np.all(np.isfinite(df)) and df.shape == (1, 4)


def test_table_gradient_reverse_mode_k1(cuda_lib):
    """f3 proper: d gamma / d (Fourier table coefficient) by reverse mode through K1, against (i) the finite-perturbation
    Hellmann-Feynman prediction (K1 re-run per perturbed table set) and (ii) actually re-solving the perturbed problems, and
    (iii) central finite differences of single coefficients."""
    import dataclasses
    import torch
    from ideal_ballooning_solver_b200 import engine, scan, synthetic, tables
    st = tables.RadialSplines(synthetic.make_equilibrium("ncsx", seed=11)).evaluate(np.linspace(0.55, 0.9, 6))
    theta = np.linspace(-3 * np.pi, 3 * np.pi, 513)
    rng = np.random.default_rng(5)
    a_star, t_star = rng.uniform(0.2, 2.5, st.ns), rng.uniform(0.0, 1.2, st.ns)
    dt0 = engine.DeviceTables.from_host(st)
    tg = scan.table_gradient(dt0, a_star, t_star, theta)
    assert tg.grad_mn.shape == dt0.tab_mn.shape and tg.grad_nyq.shape == dt0.tab_nyq.shape
    assert bool(torch.isfinite(tg.grad_mn).all()) and bool(torch.isfinite(tg.grad_nyq).all())
    eps = 1.0e-5
    perts = []
    for k in range(3):
        w_mn = 1.0 + eps * rng.standard_normal(st.tab_mn.shape[1:])[None] * np.exp(-0.3 * np.arange(st.tab_mn.shape[2]))[None, None]
        w_nq = 1.0 + eps * rng.standard_normal(st.tab_nyq.shape[1:])[None] * np.exp(-0.3 * np.arange(st.tab_nyq.shape[2]))[None, None]
        perts.append(dataclasses.replace(st, tab_mn=st.tab_mn * w_mn, tab_nyq=st.tab_nyq * w_nq))
    dts = [engine.DeviceTables.from_host(p) for p in perts]
    pred = tg.predict(dt0, dts)                                                   # (3, ns) dot products
    hf = scan.hellmann_feynman_gamma(dt0, dts, a_star, t_star, theta)             # K1 per perturbed set + K4
    np.testing.assert_allclose(tg.gamma, hf.gamma, rtol=1e-12)
    # same first-order quantity: the analytic derivative vs a finite perturbation of size eps
    np.testing.assert_allclose(pred, hf.dgamma, rtol=5e-4, atol=1e-9 * np.abs(hf.dgamma).max())
    # (ii) against re-solving the perturbed problems (first order in eps)
    h = engine.grid_spacing(theta)
    a = torch.from_numpy(a_star[:, None].copy()).cuda()
    t0 = torch.from_numpy(t_star.copy()).cuda()
    for k, dtp in enumerate(dts):
        geo = engine.geometry_batch(dtp, a, theta)
        lam_p = engine.solve_base_batch(geo.base, geo.dPdrho, t0, h, nth0=1, want_X=False, want_dX=False, want_matrix=False).lam.cpu().numpy()
        np.testing.assert_allclose(pred[k], lam_p - tg.gamma, rtol=3e-3, atol=1e-12 * np.abs(tg.gamma).max() + 1e-14)
    # (iii) single coefficients: the reverse-mode chain against the same first-order quantity formed the long way (K1 on the
    # perturbed tables, exact delta g, c, f, then the Hellmann-Feynman sums) -- this isolates the adjoint of K1 from the
    # O(h^2) difference between the continuum Hellmann-Feynman formula (utils.py:1676-1680) and the discrete eigenvalue
    singles = (("mn", 0, 1), ("mn", 1, 14), ("mn", 2, 3), ("mn", 3, 1), ("mn", 4, 9), ("mn", 5, 20), ("mn", 2, 30),
               ("nyq", 1, 0), ("nyq", 0, 2), ("nyq", 5, 1), ("nyq", 4, 17), ("nyq", 2, 5), ("nyq", 3, 11), ("nyq", 6, 40))
    gmn, gnq = tg.grad_mn.cpu().numpy(), tg.grad_nyq.cpu().numpy()
    sets, steps, an = [], [], []
    for (which, row, mode) in singles:
        base_tab = st.tab_mn if which == "mn" else st.tab_nyq
        step = max(1e-4 * np.abs(base_tab[:, row, mode]).max(), 1e-5)      # (g, c, f carry ~1e-16 of rounding: keep noise / step small)
        for sgn in (+1.0, -1.0):
            tp = base_tab.copy()
            tp[:, row, mode] += sgn * step
            sets.append(dataclasses.replace(st, tab_mn=tp) if which == "mn" else dataclasses.replace(st, tab_nyq=tp))
        steps.append(step)
        an.append((gmn if which == "mn" else gnq)[:, row, mode])
    hf1 = scan.hellmann_feynman_gamma(dt0, [engine.DeviceTables.from_host(x) for x in sets], a_star, t_star, theta)
    for k, sgl in enumerate(singles):
        long_way = (hf1.dgamma[2 * k] - hf1.dgamma[2 * k + 1]) / (2.0 * steps[k])
        np.testing.assert_allclose(an[k], long_way, rtol=2e-4, atol=1e-4 * np.abs(long_way).max() + 1e-13, err_msg=str(sgl))
    # ... and central differences of lambda itself through the whole forward path for low-order coefficients (where the
    # continuum formula is accurate to ~1e-3 on this grid)
    def gamma_of(tab_mn, tab_nyq):
        dtx = engine.DeviceTables.from_host(dataclasses.replace(st, tab_mn=tab_mn, tab_nyq=tab_nyq))
        geo = engine.geometry_batch(dtx, a, theta)
        return engine.solve_base_batch(geo.base, geo.dPdrho, t0, h, nth0=1, want_X=False, want_dX=False, want_matrix=False).lam.cpu().numpy()
    for (which, row, mode) in (("mn", 0, 0), ("mn", 0, 1), ("nyq", 1, 0), ("nyq", 0, 0)):
        base_tab = st.tab_mn if which == "mn" else st.tab_nyq
        step = max(1e-6 * np.abs(base_tab[:, row, mode]).max(), 5e-6)      # lambda carries ~1e-14 of rounding: keep noise / step small
        tp, tm = base_tab.copy(), base_tab.copy()
        tp[:, row, mode] += step; tm[:, row, mode] -= step
        if which == "mn":
            fd = (gamma_of(tp, st.tab_nyq) - gamma_of(tm, st.tab_nyq)) / (2 * step)
            ana = gmn[:, row, mode]
        else:
            fd = (gamma_of(st.tab_mn, tp) - gamma_of(st.tab_mn, tm)) / (2 * step)
            ana = gnq[:, row, mode]
        np.testing.assert_allclose(ana, fd, rtol=1e-2, atol=5e-3 * np.abs(fd).max() + 1e-12, err_msg=f"{which} row {row} mode {mode}")
