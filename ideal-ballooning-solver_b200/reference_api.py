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
  * ``vmec_fieldlines`` / ``vmec_fieldlines_axisym`` return the reference's whole Struct (``utils.py:723-864``,
    ``:1420-1542``): the per-point arrays come from the full-output geometry kernel (``ibs_geometry_full``), the
    per-surface scalars from the radial splines; ``theta_vmec`` is the converged root (the reference stops scipy's secant
    iteration at 1.48e-8).  ``plot=True`` raises ``NotImplementedError``.
"""
from __future__ import annotations

import types
import warnings

import numpy as np

from . import engine, tables

__all__ = ["vmec_splines", "vmec_fieldlines", "vmec_fieldlines_axisym", "gamma_ball_full", "obj_w_grad", "derm", "dermv", "Struct",
           "FULL_FIELDS"]

#: per-point arrays written by ``ibs_geometry_full``, in the order of the enum in ``csrc/ibs_geometry_full.cu``
FULL_FIELDS = (
    "phi", "theta_pest", "theta_vmec", "lambdas",
    "R", "d_R_d_s", "d_R_d_theta_vmec", "d_R_d_phi", "Z", "d_Z_d_s", "d_Z_d_theta_vmec", "d_Z_d_phi",
    "d_lambda_d_s", "d_lambda_d_theta_vmec", "d_lambda_d_phi",
    "sqrt_g_vmec", "modB", "d_B_d_s", "d_B_d_theta_vmec", "d_B_d_phi", "B_sup_theta_vmec", "B_sup_phi",
    "B_sub_s", "B_sub_theta_vmec", "B_sub_phi", "B_sup_theta_pest", "sqrt_g_vmec_alt", "sinphi", "cosphi",
    "d_X_d_theta_vmec", "d_X_d_phi", "d_X_d_s", "d_Y_d_theta_vmec", "d_Y_d_phi", "d_Y_d_s",
    "grad_s_X", "grad_s_Y", "grad_s_Z", "grad_theta_vmec_X", "grad_theta_vmec_Y", "grad_theta_vmec_Z",
    "grad_phi_X", "grad_phi_Y", "grad_phi_Z", "grad_psi_X", "grad_psi_Y", "grad_psi_Z",
    "grad_alpha_X", "grad_alpha_Y", "grad_alpha_Z", "grad_B_X", "grad_B_Y", "grad_B_Z", "B_X", "B_Y", "B_Z",
    "B_cross_grad_s_dot_grad_alpha", "B_cross_grad_s_dot_grad_alpha_alternate",
    "B_cross_grad_B_dot_grad_alpha", "B_cross_grad_B_dot_grad_alpha_alternate",
    "B_cross_grad_B_dot_grad_psi", "B_cross_kappa_dot_grad_psi", "B_cross_kappa_dot_grad_alpha",
    "grad_alpha_dot_grad_alpha", "grad_alpha_dot_grad_psi", "grad_psi_dot_grad_psi",
    "bmag", "gradpar_theta_pest", "gradpar_phi", "gds2", "gds21", "gds22", "gbdrift", "gbdrift0", "cvdrift", "cvdrift0",
    "B_p",
)

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


def _struct_common(out, vs, st, s, alpha, theta1d, phi1d, phi_center, fields, skip=()):
    """Per-surface scalars and bookkeeping entries of the Struct (``utils.py:723-760, 846-864``) + the per-point arrays."""
    out.ns, out.nalpha = len(s), len(alpha)
    out.nl = len(theta1d) if theta1d is not None else len(phi1d)
    out.s, out.alpha, out.theta1d, out.phi1d = s, alpha, theta1d, phi1d
    out.iota, out.d_iota_d_s = st.row("iota"), st.row("d_iota_d_s")
    out.d_pressure_d_s, out.shat = st.row("d_pressure_d_s"), st.row("shat")
    out.phi_center = phi_center
    psi_e = -st.phiedge / (2 * np.pi)                                 # utils.py:474
    out.edge_toroidal_flux_over_2pi = psi_e
    out.L_reference = st.Aminor_p                                     # utils.py:662-664
    out.B_reference = 2 * abs(psi_e) / (st.Aminor_p * st.Aminor_p)
    out.toroidal_flux_sign = np.sign(psi_e)
    pres = st.row("pressure")
    sqrt_s = np.sqrt(s)
    out.beta_N = 4 * np.pi * 1e-7 * pres / out.B_reference ** 2       # utils.py:668-676
    out.tprim = -1 * out.d_pressure_d_s * 2 * sqrt_s * 2 / 3 * 1 / pres
    out.fprim = -1 * out.d_pressure_d_s * 2 * sqrt_s * 1 / 3 * 1 / pres
    out.temp = pres ** (2 / 3)
    out.dens = pres ** (1 / 3)
    for k, name in enumerate(FULL_FIELDS):
        if name not in skip:
            setattr(out, name, np.ascontiguousarray(fields[:, :, k, :]))
    return out


def _surface_tables(vs, s):
    vs = _as_splines(vs)
    return vs if isinstance(vs, tables.SurfaceTables) else vs.evaluate(s)


def vmec_fieldlines(vs, s, alpha, theta1d=None, phi1d=None, phi_center=0, plot=False, show=True):
    """``vmec_fieldlines`` (``utils.py:161-864``): the whole Struct, arrays shaped ``(ns, nalpha, nl)``."""
    s = np.atleast_1d(np.asarray(s, dtype=np.float64))             # utils.py:277-291
    alpha = np.atleast_1d(np.asarray(alpha, dtype=np.float64))
    if (theta1d is not None) and (phi1d is not None):
        raise ValueError("You cannot specify both theta and phi")   # utils.py:293-294
    if (theta1d is None) and (phi1d is None):
        raise ValueError("You must specify either theta or phi")    # utils.py:295-296
    if plot:
        raise NotImplementedError("plotting is outside the ballooning hot path")
    st = _surface_tables(vs, s)
    grid = np.asarray(theta1d if theta1d is not None else phi1d, dtype=np.float64)
    if st.bsupumnc is None:
        # tables without bsupumnc (e.g. built by hand for the hot path): the eight hot-path arrays from K1
        if theta1d is None:
            raise ValueError("vmec_fieldlines(phi1d=...) needs the full tables (bsupumnc)")
        return _fieldlines_hot(st, s, alpha, grid, phi_center)
    fields, info = engine.geometry_full(st, alpha, grid, mode=0 if theta1d is not None else 1, phi_center=float(phi_center))
    if np.any(info.cpu().numpy() >> 16):
        # the reference's scipy.optimize.newton raises RuntimeError when the root solve fails (utils.py:410)
        raise RuntimeError("Failed to converge: theta_vmec root solve")
    out = _struct_common(Struct(), vs, st, s, alpha, None if theta1d is None else grid, None if phi1d is None else grid,
                         phi_center, fields.cpu().numpy(), skip=("lambdas", "B_p"))
    out.dPdrho = -1.0 * 0.5 * np.mean((out.cvdrift - out.gbdrift) * out.bmag ** 2, axis=-1)   # ball_scan.py:262 (convenience)
    return out


def _fieldlines_hot(st, s, alpha, theta1d, phi_center):
    """The hot-path subset through K1 (``ibs_geometry_batch``): the eight arrays ``ball_scan.py:254-261`` reads."""
    dt = engine.DeviceTables.from_host(st)
    geo = engine.geometry_batch(dt, alpha, theta1d, phi_center=float(phi_center), want_theta_vmec=True, want_info=True)
    if np.any(geo.info.cpu().numpy() >> 16):
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


def vmec_fieldlines_axisym(vs, s, alpha, theta1d=None, phi1d=None, phi_center=0, plot=False, show=True):
    """``vmec_fieldlines_axisym`` (``utils.py:872-1542``): the axisymmetric routine the boundary-curvature penalty uses
    (``Simsopt_runner.py:231-245``).  A uniform theta_vmec grid spanning ``theta1d`` (no root solve, ``:972-978``), the
    poloidal angle flipped by pi when the first step along the grid is not counter-clockwise from the inboard side
    (``:993-1010``), ``theta_pest = theta_vmec + lambda`` (``:1043``), and the ``*_1`` arrays interpolated back onto
    ``theta1d`` from the first surface / field line (``:1333-1369``)."""
    s = np.atleast_1d(np.asarray(s, dtype=np.float64))
    alpha = np.atleast_1d(np.asarray(alpha, dtype=np.float64))
    if (theta1d is not None) and (phi1d is not None):
        raise ValueError("You cannot specify both theta and phi")   # utils.py:901-902
    if (theta1d is None) and (phi1d is None):
        raise ValueError("You must specify either theta or phi")    # utils.py:903-904
    if theta1d is None:
        raise TypeError("vmec_fieldlines_axisym needs theta1d (the reference dereferences it unconditionally, utils.py:976)")
    if plot:
        raise NotImplementedError("plotting is outside the ballooning hot path")
    theta1d = np.asarray(theta1d, dtype=np.float64)
    st = _surface_tables(vs, s)
    if st.bsupumnc is None:
        raise ValueError("vmec_fieldlines_axisym needs the full tables (bsupumnc)")
    theta_v = np.linspace(np.min(theta1d), np.max(theta1d), len(theta1d))             # utils.py:972-978
    # orientation test on the first two points of the first surface (utils.py:984-1001), at phi = 0
    ang = st.xm[:, None] * theta_v[None, :2]
    R2 = st.tab_mn[0, 0] @ np.cos(ang)
    Z2 = st.tab_mn[0, 1] @ np.sin(ang)
    flipit = bool(R2[0] > R2[1] or Z2[1] > Z2[0])
    fields, _ = engine.geometry_full(st, alpha, theta_v, mode=2, theta_shift=np.pi if flipit else 0.0, zero_xn_nyq=True,
                                     phi_center=float(phi_center))
    F = fields.cpu().numpy()
    out = _struct_common(Struct(), vs, st, s, alpha, theta1d, phi1d, phi_center, F, skip=("lambdas", "B_p"))
    B_p = F[:, :, FULL_FIELDS.index("B_p"), :]                       # utils.py:1282-1286
    # the `_1` arrays: first surface, first field line, interpolated from theta_pest back to theta1d (utils.py:1333-1369)
    tp = out.theta_pest[0][0]
    on_grid = lambda a: np.interp(theta1d, tp, a[0][0])
    out.cvdrift0_1 = on_grid(out.cvdrift0)
    out.gbdrift0_1 = on_grid(out.cvdrift0)
    out.gradpar_theta_pest_1 = on_grid(out.gradpar_theta_pest)
    out.bmag_1, out.B_p_1 = on_grid(out.bmag), on_grid(B_p)
    out.cvdrift_1, out.gbdrift_1 = on_grid(out.cvdrift), on_grid(out.gbdrift)
    out.gds21_1, out.gds22_1, out.gds2_1 = on_grid(out.gds21), on_grid(out.gds22), on_grid(out.gds2)
    out.R_1, out.Z_1 = on_grid(out.R), on_grid(out.Z)
    psi_e = out.edge_toroidal_flux_over_2pi
    out.Rprime_1 = on_grid(out.d_R_d_s) * 1 / psi_e * 1 / out.iota * out.R_1 * out.B_p_1
    out.Zprime_1 = on_grid(out.d_Z_d_s) * 1 / psi_e * 1 / out.iota * out.R_1 * out.B_p_1
    out.loc_shr = out.gds21_1 * 0                                    # utils.py:1392
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
