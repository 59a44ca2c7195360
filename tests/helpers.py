"""Shared helpers of the parity tests (CPU side)."""
import types

import numpy as np

from oracle import ballooning_oracle as bo

LAM_RTOL = 1e-10      # north star: lambda_max within 1e-10 relative
X_ATOL = 1e-8         # sign-normalised eigenvectors / gradients within 1e-8


def sign_normalise(X):
    """The reference's X has arbitrary sign (ARPACK); the engine returns X >= 0."""
    X = np.asarray(X)
    k = np.argmax(np.abs(X), axis=-1)
    s = np.sign(np.take_along_axis(X, k[..., None], axis=-1))
    return X * s


def tables_from_fixture(D):
    """SurfaceTables-like object from a golden .npz (tables evaluated by the REFERENCE's splines)."""
    from ideal_ballooning_solver_b200.tables import SurfaceTables
    return SurfaceTables(np.array(D["tab_mn"]), np.array(D["tab_nyq"]), np.array(D["scal"]), np.array(D["xm"]),
                         np.array(D["xn"]), np.array(D["xm_nyq"]), np.array(D["xn_nyq"]), float(D["phiedge"]),
                         float(D["Aminor_p"]), int(D["nfp"]))


def fixture_gcf(D, i, j, k):
    """g, c, f of fixture solve (surface i, alpha j, theta0 k), formed like utils.py:1560-1562."""
    th0 = D["theta0s"][k]
    cv = D["geo_cvdrift"][i, j] + th0 * D["geo_cvdrift0"][i, j]
    gd = D["geo_gds2"][i, j] + 2 * th0 * D["geo_gds21"][i, j] + th0 ** 2 * D["geo_gds22"][i, j]
    return bo.gcf(D["dPdrho"][i, j], D["geo_bmag"][i, j], D["geo_gradpar_theta_pest"][i, j], cv, gd)


def fixture_base(D):
    """(ns, nalpha, 8, nl) base array in the engine's row order from a golden fixture."""
    from ideal_ballooning_solver_b200.engine import BASE_NAMES
    return np.stack([D["geo_" + n] for n in BASE_NAMES], axis=2)


def s_alpha_base(shat, alpha, theta):
    """The s-alpha model (bishop_ball_s-alpha.py) as the eight base arrays of a field line, so that the theta0 dependence
    is the reference's (ball_scan.py:267-268): g = f = 1 + (shat (theta - theta0) - alpha sin theta)^2,
    c = alpha (cos theta + (...) sin theta).  Returns (base (8, N) in IBS_BASE_* order, dPdrho)."""
    L0 = shat * theta - alpha * np.sin(theta)
    one = np.ones_like(theta)
    base = np.stack([one, one, np.cos(theta) + L0 * np.sin(theta), -shat * np.sin(theta), 1 + L0 ** 2, -shat * L0,
                     shat ** 2 * one, one])
    return base, -alpha
