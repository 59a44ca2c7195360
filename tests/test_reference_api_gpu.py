"""GPU: the reference-signature facade (reference_api.py) called exactly like ball_scan.py calls utils.py,
checked against the reference's own outputs (golden fixtures)."""
import numpy as np
import pytest

from helpers import LAM_RTOL, X_ATOL, sign_normalise, tables_from_fixture

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ncsx_wout_op", "synthetic_d3d"])
def test_scan_loop_like_ball_scan(cuda_lib, golden, name):
    """ball_scan.py:248-274 transcribed with the facade's functions."""
    from ideal_ballooning_solver_b200.reference_api import gamma_ball_full, vmec_fieldlines
    D = golden(name)
    theta = D["theta"]
    vs = tables_from_fixture(D).select([1])
    rho = D["surfaces"][1]
    vguess = (1 - np.tanh(theta[1:-1] / np.pi) ** 2) * np.cos(theta[1:-1] / 8)        # ball_scan.py:209
    for ja, al in enumerate(D["alphas"][:2]):
        fl = vmec_fieldlines(vs, rho, al, theta1d=theta)                                # ball_scan.py:251
        bmag, gbdrift, cvdrift = fl.bmag[0][0], fl.gbdrift[0][0], fl.cvdrift[0][0]
        cvdrift0, gds2, gds21, gds22 = fl.cvdrift0[0][0], fl.gds2[0][0], fl.gds21[0][0], fl.gds22[0][0]
        gradpar = fl.gradpar_theta_pest[0][0]
        dPdrho = -1.0 * 0.5 * np.mean((cvdrift - gbdrift) * bmag ** 2)                  # ball_scan.py:262
        np.testing.assert_allclose(dPdrho, D["dPdrho"][1, ja], rtol=1e-11)
        for jt, th0 in enumerate(D["theta0s"]):
            cvdrift_fth = cvdrift + th0 * cvdrift0                                      # ball_scan.py:267-268
            gds2_fth = gds2 + 2 * th0 * gds21 + th0 ** 2 * gds22
            gam, X, dX, g, c, f = gamma_ball_full(dPdrho, theta, bmag, gradpar, cvdrift_fth, gds2_fth, vguess, 1.0)
            vguess = X[1:-1]                                                             # ball_scan.py:272
            np.testing.assert_allclose(gam, D["lam_conv"][1, ja, jt], rtol=LAM_RTOL)
            np.testing.assert_allclose(X, sign_normalise(D["X_conv"][1, ja, jt]), rtol=0, atol=X_ATOL)
            # g, c, f are returned exactly as utils.py:1560-1562 forms them
            assert np.array_equal(g, np.abs(gradpar) * gds2_fth / bmag)
            assert np.array_equal(f, gds2_fth / bmag ** 2 * 1 / (np.abs(gradpar) * bmag))
            assert X.shape == dX.shape == theta.shape and X[0] == 0 and X[-1] == 0


def test_obj_w_grad_signature(cuda_lib, golden):
    from ideal_ballooning_solver_b200.reference_api import obj_w_grad
    D = golden("ncsx_wout_op")
    theta = D["theta"]
    for si, al, th0, val, ga, gt in D["grad_points"]:
        vs = tables_from_fixture(D).select([int(si)])
        v, grad = obj_w_grad((al, th0), vs, D["surfaces"][int(si)], theta, None, 1.0)   # ball_scan.py:308
        np.testing.assert_allclose(v, val, rtol=LAM_RTOL)
        np.testing.assert_allclose(grad, [ga, gt], rtol=0, atol=X_ATOL)
        assert grad.shape == (2,)


def test_argument_errors_match_reference(cuda_lib, golden):
    """utils.py:293-296: exactly one of theta1d / phi1d."""
    from ideal_ballooning_solver_b200.reference_api import vmec_fieldlines
    D = golden("synthetic_d3d")
    vs = tables_from_fixture(D).select([0])
    with pytest.raises(ValueError):
        vmec_fieldlines(vs, 0.5, 0.0)
    with pytest.raises(ValueError):
        vmec_fieldlines(vs, 0.5, 0.0, theta1d=D["theta"], phi1d=D["theta"])


def test_splines_from_wout_like_object(cuda_lib, golden):
    """vmec_splines + vmec_fieldlines from a wout-like object (what a simsopt Vmec exposes, utils.py:46-135)."""
    import types
    from ideal_ballooning_solver_b200 import synthetic
    from ideal_ballooning_solver_b200.reference_api import vmec_fieldlines, vmec_splines
    D = golden("synthetic_d3d")
    vmec = types.SimpleNamespace(wout=synthetic.make_equilibrium("d3d"))
    vs = vmec_splines(vmec)
    fl = vmec_fieldlines(vs, D["surfaces"], D["alphas"], theta1d=D["theta"])
    assert fl.bmag.shape == D["geo_bmag"].shape
    np.testing.assert_allclose(fl.gds2, D["geo_gds2"], rtol=1e-9)
    np.testing.assert_allclose(fl.theta_vmec, D["theta_vmec"], rtol=0, atol=1e-11)
