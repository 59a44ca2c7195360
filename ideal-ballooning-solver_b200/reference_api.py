"""Drop-in mirror of the reference's entry points for the ballooning hot path.

``ball_scan.py`` reaches the hot path through ``from utils import *`` (``ball_scan.py:19``); the four
call signatures below are that boundary.  Same names, argument meaning, return
tuples and error behaviour as ``/root/reference/utils.py``; numpy in, numpy out; the arithmetic runs
in the CUDA library (``include/ibs_b200.h``) on the current device.  There is no CPU fallback.

Differences that are deliberate and documented in DESIGN.md:
  * ``gamma_ball_full`` returns the *converged* eigenpair (the reference stops ARPACK at ``tol=5e-7``,
    ``utils.py:1597``); ``vguess`` is accepted and ignored (no Krylov start vector is needed); ``X`` is
    returned non-negative (the reference's sign is arbitrary, ``utils.py:1605``); ``sigma0`` keeps its
    meaning as a check: if the eigenvalue nearest ``sigma0`` would not be lambda_max a warning is issued.
  * ``vmec_fieldlines`` returns the quantities the ballooning path reads (``ball_scan.py:254-261``) plus
    ``theta_vmec``; the remaining ~90 gyrokinetic-geometry arrays of the reference Struct are out of
    scope.  The ``phi1d`` form raises ``NotImplementedError``.
"""
from __future__ import annotations

import types
import warnings

import numpy as np

from . import engine, tables

__all__ = ["vmec_splines", "vmec_fieldlines", "gamma_ball_full", "obj_w_grad", "Struct"]

DEL_ALPHA = 0.004          # utils.py:1639


class Struct:
    """Attribute bag, as in ``utils.py:31-34``."""


def vmec_splines(vmec):
    """``vmec_splines(vmec)`` (``utils.py:37-158``): radial splines of the Fourier tables.  ``vmec`` is any
    object with ``.wout`` (tables laid out ``(mn, ns)``) -- a ``simsopt`` ``Vmec`` or the result of
    ``tables.read_wout`` wrapped in a namespace; ``vmec.run()`` is called if present (``utils.py:46``)."""
    if hasattr(vmec, "run"):
        vmec.run()
    return tables.RadialSplines(vmec.wout if hasattr(vmec, "wout") else vmec)


def _as_splines(vs):
    if isinstance(vs, tables.RadialSplines):
        return vs
    if isinstance(vs, tables.SurfaceTables):
        return vs
    return vmec_splines(vs)            # utils.py:272-274: a Vmec object is converted on the fly


def vmec_fieldlines(vs, s, alpha, theta1d=None, phi1d=None, phi_center=0, plot=False, show=True):
    """``vmec_fieldlines`` (``utils.py:161-864``), hot-path subset, arrays shaped ``(ns, nalpha, nl)``."""
    vs = _as_splines(vs)
    s = np.atleast_1d(np.asarray(s, dtype=np.float64))             # utils.py:277-291
    alpha = np.atleast_1d(np.asarray(alpha, dtype=np.float64))
    if (theta1d is not None) and (phi1d is not None):
        raise ValueError("You cannot specify both theta and phi")   # utils.py:293-294
    if (theta1d is None) and (phi1d is None):
        raise ValueError("You must specify either theta or phi")    # utils.py:295-296
    if theta1d is None:
        raise NotImplementedError("vmec_fieldlines(phi1d=...) is outside the ballooning hot path")
    if plot:
        raise NotImplementedError("plotting is outside the ballooning hot path")
    theta1d = np.asarray(theta1d, dtype=np.float64)
    st = vs if isinstance(vs, tables.SurfaceTables) else vs.evaluate(s)
    dt = engine.DeviceTables.from_host(st)
    geo = engine.geometry_batch(dt, alpha, theta1d, phi_center=float(phi_center), want_theta_vmec=True, want_info=True)
    info = geo.info.cpu().numpy()
    if np.any(info >> 16):
        # the reference's scipy.optimize.newton raises RuntimeError when the root solve fails (utils.py:410)
        raise RuntimeError("Failed to converge: theta_vmec root solve")
    base = geo.base.cpu().numpy()
    out = Struct()
    out.ns, out.nalpha, out.nl = len(s), len(alpha), len(theta1d)
    out.s, out.alpha = s, alpha
    out.iota, out.d_iota_d_s = st.row("iota"), st.row("d_iota_d_s")
    out.d_pressure_d_s, out.shat = st.row("d_pressure_d_s"), st.row("shat")
    out.phi_center = phi_center
    out.theta_pest = np.broadcast_to(theta1d, (out.ns, out.nalpha, out.nl)).copy()
    out.phi = phi_center + (theta1d[None, None, :] - alpha[None, :, None]) / out.iota[:, None, None]   # utils.py:373
    out.theta_vmec = geo.theta_vmec.cpu().numpy()
    for k, name in enumerate(engine.BASE_NAMES):
        setattr(out, name, np.ascontiguousarray(base[:, :, k, :]))
    out.gbdrift0 = out.cvdrift0                                      # utils.py:720
    out.dPdrho = geo.dPdrho.cpu().numpy()                            # ball_scan.py:262 (convenience, not in the reference Struct)
    return out


def gamma_ball_full(dPdrho, theta_PEST, B, gradpar, cvdrift, gds2, vguess=None, sigma0=0.42):
    """``gamma_ball_full`` (``utils.py:1550-1624``): returns ``(gam, X, dX, g, c, f)``."""
    import torch
    theta = np.asarray(theta_PEST, dtype=np.float64)
    n = len(theta)
    # the reference re-grids everything onto a uniform theta grid first (utils.py:1564-1571)
    tu = np.linspace(theta[0], theta[-1], n)
    rows = {engine.BASE_NAMES.index("bmag"): B, engine.BASE_NAMES.index("gradpar_theta_pest"): gradpar,
            engine.BASE_NAMES.index("cvdrift"): cvdrift, engine.BASE_NAMES.index("gds2"): gds2}
    base = np.zeros((1, engine.NBASE, n))
    uniform = np.array_equal(tu, theta) or np.allclose(np.diff(theta), tu[1] - tu[0], rtol=1e-12, atol=0)
    for k, a in rows.items():
        a = np.asarray(a, dtype=np.float64)
        base[0, k] = a
    if not uniform:
        # On a non-uniform grid the reference interpolates g, c, f to a uniform grid but takes the half-point g from the
        # ORIGINAL grid (utils.py:1567-1576), which the kernels' (g_j + g_j+1) / 2 does not reproduce.  Every caller in the
        # reference passes np.linspace grids (ball_scan.py:203-208), so this is rejected rather than approximated.
        raise NotImplementedError("gamma_ball_full: theta_PEST must be equispaced (as in every reference call site)")
    sol = engine.solve_base_batch(torch.from_numpy(base).cuda(), torch.tensor([float(dPdrho)], dtype=torch.float64),
                                  torch.zeros(1, dtype=torch.float64), engine.grid_spacing(theta), nth0=1,
                                  sigma=torch.tensor([float(sigma0)], dtype=torch.float64), want_gcf=True)
    gcf = tuple(t[0].cpu().numpy() for t in (sol.g, sol.c, sol.f))
    flags = int(sol.flags[0].item())
    if flags & engine.FLAG_BAD_INPUT:
        raise ValueError("gamma_ball_full: non-finite input or g <= 0 / f <= 0 on the field line")
    if flags & engine.FLAG_NOT_CONVERGED:
        # scipy raises ArpackNoConvergence in the reference (utils.py:1597, uncaught)
        raise RuntimeError("gamma_ball_full: eigen-solve did not converge")
    if flags & engine.FLAG_SIGMA_NOT_MAX:
        warnings.warn("gamma_ball_full: the eigenvalue nearest sigma0 is not lambda_max; the reference's ARPACK "
                      "shift-invert (utils.py:1597) would have returned a different eigenpair", RuntimeWarning)
    return (float(sol.lam[0].item()), sol.X[0].cpu().numpy(), sol.dX[0].cpu().numpy(), *gcf)


def obj_w_grad(x0, vs, rho_val, theta, vguess00=None, sigma00=0.42):
    """``obj_w_grad`` (``utils.py:1632-1728``): ``(-gam, array([-dgam/dalpha, -dgam/dtheta0]))``."""
    import torch
    alpha_val, theta0_val = x0
    vs = _as_splines(vs)
    theta = np.asarray(theta, dtype=np.float64)
    st = vs if isinstance(vs, tables.SurfaceTables) else vs.evaluate(np.atleast_1d(rho_val))
    dt = engine.DeviceTables.from_host(st)
    alphas = np.array([[alpha_val - 0.5 * DEL_ALPHA, alpha_val, alpha_val + 0.5 * DEL_ALPHA]])    # utils.py:1641-1646
    geo = engine.geometry_batch(dt, torch.from_numpy(alphas).cuda(), theta)
    val, grad, _, _, info = engine.obj_w_grad_batch(geo.base, geo.dPdrho, torch.tensor([float(theta0_val)], dtype=torch.float64),
                                                    engine.grid_spacing(theta), del_alpha=DEL_ALPHA)
    if int(info[0].item()) >> 16:
        raise RuntimeError("obj_w_grad: eigen-solve failed (flags %d)" % (int(info[0].item()) >> 16))
    return float(val[0].item()), grad[0].cpu().numpy()


# ---------------------------------------------------------------------------------------------------------
# f4 (part): finite-difference helpers of the curvature penalty (host-side numpy in the reference as well)
# ---------------------------------------------------------------------------------------------------------
def _fd_layout(arr, ch):
    """Differencing axis last; 1-D inputs become (1, n) along a surface ('l') and (n, 1) across surfaces ('r'), which is
    also the shape the reference returns for them (``utils.py:1745-1779``)."""
    a = np.asarray(arr, dtype=np.float64)
    if a.ndim == 1:
        a = a[None, :] if ch == "l" else a[:, None]
    if a.ndim != 2 or ch not in ("l", "r"):
        raise ValueError("derm/dermv take 1-D or 2-D arrays and ch in {'l', 'r'}")
    return (a, 1) if ch == "l" else (a.T, 0)


def derm(arr, ch, par="e"):
    """``derm`` (``utils.py:1737-1807``): un-normalised central difference ``a[i+1] - a[i-1]`` along (``'l'``) or across
    (``'r'``) the flux surfaces; the ends are ``2 (a[1] - a[0])``, ``2 (a[-1] - a[-2])``, or zero along a surface for an
    even-parity array."""
    a, axis = _fd_layout(arr, ch)
    out = np.zeros_like(a)
    out[:, 1:-1] = np.diff(a[:, :-1], axis=1) + np.diff(a[:, 1:], axis=1)
    if not (ch == "l" and par == "e"):
        out[:, 0] = 2 * (a[:, 1] - a[:, 0])
        out[:, -1] = 2 * (a[:, -1] - a[:, -2])
    return out if axis == 1 else np.ascontiguousarray(out.T)


def dermv(arr, brr, ch, par="e"):
    """``dermv`` (``utils.py:1810-1943``): derivative of ``arr`` with respect to the non-uniform ``brr`` (weighted
    three-point formula inside; one-sided at the ends: second order for a 1-D odd-parity array along a surface, first
    order otherwise, zero for even parity along a surface).  A 1-D array with ``ch='r'`` raises: the reference stops in
    the debugger there (``utils.py:1866``)."""
    one_d = np.ndim(arr) == 1
    if one_d and ch != "l":
        raise NotImplementedError("dermv(1-D, 'r'): the reference enters pdb.set_trace() here (utils.py:1866)")
    a, axis = _fd_layout(arr, ch)
    b, _ = _fd_layout(brr, ch)
    out = np.zeros_like(a)
    h1 = b[:, 2:] - b[:, 1:-1]
    h0 = b[:, 1:-1] - b[:, :-2]
    out[:, 1:-1] = (a[:, 2:] / h1 ** 2 + a[:, 1:-1] * (1 / h0 ** 2 - 1 / h1 ** 2) - a[:, :-2] / h0 ** 2) / (1 / h1 + 1 / h0)
    if ch == "l" and par == "e":
        pass
    elif one_d:
        out[:, 0] = (4 * a[:, 1] - 3 * a[:, 0] - a[:, 2]) / (2 * (b[:, 1] - b[:, 0]))
        out[:, -1] = (-4 * a[:, -2] + 3 * a[:, -1] + a[:, -3]) / (2 * (b[:, -1] - b[:, -2]))
    else:
        out[:, 0] = 2 * (a[:, 1] - a[:, 0]) / (2 * (b[:, 1] - b[:, 0]))
        out[:, -1] = 2 * (a[:, -1] - a[:, -2]) / (2 * (b[:, -1] - b[:, -2]))
    return out if axis == 1 else np.ascontiguousarray(out.T)
