"""TEST INFRASTRUCTURE ONLY -- CPU (numpy/scipy) restatement of the reference hot path.

This file restates, function by function, the algorithm of the reference
(``/root/reference/utils.py`` and the scan loops of ``/root/reference/ball_scan.py``) so
that the CUDA engine can be checked on machines where ``/root/reference`` does not exist
(the GPU box).  It is the *checker*, never the thing shipped or measured: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import it.

Parity pin: the reference has no numerical golden vectors of its own for this path (both of
its tests are visual).  The pin is therefore the reference itself, run in
the build container through ``oracle/ref_shim.py``: ``tests/golden/make_golden.py`` stores
its outputs in ``tests/golden/*.npz`` and ``tests/test_oracle.py`` checks this restatement
against them (and live against the reference when ``/root/reference`` is present).

Third-party arithmetic the reference delegates to and this file calls the same way:
``scipy.optimize.newton`` (secant, ``utils.py:410``), ARPACK shift-invert through
``scipy.sparse.linalg.eigs`` (``utils.py:1597``), ``scipy.integrate.simpson`` (the reference's
``simps``, ``utils.py:1621``), ``scipy.optimize.minimize`` L-BFGS-B (``ball_scan.py:307``).
No version pins exist in the reference (no lock file); scipy 1.18.1 / numpy 2.3.5 here.
"""
from __future__ import annotations

import types

import numpy as np
from scipy.integrate import simpson
from scipy.optimize import newton

MU0 = 4 * np.pi * 1.0e-7

HOT_FIELDS = ("bmag", "gradpar_theta_pest", "cvdrift", "cvdrift0", "gds2", "gds21", "gds22",
              "gbdrift")


# ---------------------------------------------------------------------------------------
# a1-a4: field-line geometry (utils.py:161-864, hot-path subset)
# ---------------------------------------------------------------------------------------
def fieldlines(tab, alpha, theta1d, phi_center=0.0, full=False):
    """Restatement of ``vmec_fieldlines(vs, s, alpha, theta1d=...)`` (``utils.py:161``) for
    tables already evaluated at the requested surfaces.

    ``tab`` is a ``SurfaceTables``-like object with ``row(name) -> (ns, nmodes)`` /
    ``(ns,)`` arrays, ``xm, xn, xm_nyq, xn_nyq, phiedge, Aminor_p``.  Returns a namespace
    whose arrays are shaped ``(ns, nalpha, nl)`` exactly like the reference Struct; only the
    quantities the ballooning path reads are produced (plus the 19 mode sums if ``full``).
    """
    alpha = np.atleast_1d(np.asarray(alpha, dtype=float))
    theta1d = np.asarray(theta1d, dtype=float)
    s = tab.row("s")
    iota, d_iota = tab.row("iota"), tab.row("d_iota_d_s")
    dpds, shat = tab.row("d_pressure_d_s"), tab.row("shat")
    ns, nalpha, nl = s.size, alpha.size, theta1d.size
    xm, xn, xmq, xnq = tab.xm, tab.xn, tab.xm_nyq, tab.xn_nyq
    lmns = tab.row("lmns")

    # utils.py:371-373 -- theta_pest given, phi follows the field line
    theta_pest = np.broadcast_to(theta1d, (ns, nalpha, nl)).copy()
    phi = phi_center + (theta1d[None, None, :] - alpha[None, :, None]) / iota[:, None, None]

    # utils.py:391-416 -- theta_vmec from theta_pest: vectorised secant (scipy newton, no fprime)
    theta_vmec = np.empty((ns, nalpha, nl))
    for js in range(ns):
        def residual(tv, phi0, target):
            return target - (tv + np.sum(lmns[js, :, None] * np.sin(xm[:, None] * tv - xn[:, None] * phi0),
                                         axis=0))
        for ja in range(nalpha):
            guess = theta_pest[js, ja]
            theta_vmec[js, ja] = newton(residual, x0=guess, x1=guess + 0.1,
                                        args=(phi[js, ja], theta_pest[js, ja]))

    # utils.py:420-468 -- the mode sums the hot path needs (19 of the reference's 21)
    def sums(xm_, xn_, cos_rows, sin_rows):
        ang = xm_[:, None, None, None] * theta_vmec[None] - xn_[:, None, None, None] * phi[None]
        ca, sa = np.cos(ang), np.sin(ang)
        out = {}
        for name, coef, weight in cos_rows:
            wgt = 1.0 if weight is None else weight[:, None, None, None]
            out[name] = np.einsum("ij,jikl->ikl", tab.row(coef), wgt * ca)
        for name, coef, weight in sin_rows:
            wgt = 1.0 if weight is None else weight[:, None, None, None]
            out[name] = np.einsum("ij,jikl->ikl", tab.row(coef), wgt * sa)
        return out

    q = sums(xm, xn,
             cos_rows=[("R", "rmnc", None), ("R_s", "d_rmnc_d_s", None), ("Z_t", "zmns", xm),
                       ("Z_p", "zmns", -xn), ("L_t", "lmns", xm), ("L_p", "lmns", -xn)],
             sin_rows=[("R_t", "rmnc", -xm), ("R_p", "rmnc", xn), ("Z_s", "d_zmns_d_s", None),
                       ("L_s", "d_lmns_d_s", None)])
    q.update(sums(xmq, xnq,
                  cos_rows=[("sqrtg", "gmnc", None), ("B", "bmnc", None), ("B_s", "d_bmnc_d_s", None),
                            ("Bsup_p", "bsupvmnc", None), ("Bsub_t", "bsubumnc", None),
                            ("Bsub_p", "bsubvmnc", None)],
                  sin_rows=[("B_t", "bmnc", -xmq), ("B_p", "bmnc", xnq), ("Bsub_s", "bsubsmns", None)]))
    R, R_s, R_t, R_p = q["R"], q["R_s"], q["R_t"], q["R_p"]
    Z_s, Z_t, Z_p = q["Z_s"], q["Z_t"], q["Z_p"]
    L_s, L_t, L_p = q["L_s"], q["L_t"], q["L_p"]
    sqrtg, modB = q["sqrtg"], q["B"]

    psi_e = -tab.phiedge / (2 * np.pi)                      # utils.py:474
    io = iota[:, None, None]
    # utils.py:480-508 -- Cartesian dual basis
    sp, cp = np.sin(phi), np.cos(phi)
    X_t, X_p, X_s = R_t * cp, R_p * cp - R * sp, R_s * cp
    Y_t, Y_p, Y_s = R_t * sp, R_p * sp + R * cp, R_s * sp
    gs = [(Y_t * Z_p - Z_t * Y_p) / sqrtg, (Z_t * X_p - X_t * Z_p) / sqrtg, (X_t * Y_p - Y_t * X_p) / sqrtg]
    gt = [(Y_p * Z_s - Z_p * Y_s) / sqrtg, (Z_p * X_s - X_p * Z_s) / sqrtg, (X_p * Y_s - Y_p * X_s) / sqrtg]
    gp = [(Y_s * Z_t - Z_s * Y_t) / sqrtg, (Z_s * X_t - X_s * Z_t) / sqrtg, (X_s * Y_t - Y_s * X_t) / sqrtg]
    # utils.py:515-538 -- grad psi, grad alpha
    gpsi = [c * psi_e for c in gs]
    a_s = L_s - (phi - phi_center) * d_iota[:, None, None]
    galpha = []
    for k in range(3):
        v = a_s * gs[k]
        v = v + ((1 + L_t) * gt[k] + (-io + L_p) * gp[k])
        galpha.append(v)
    # utils.py:603-618, 646-658 -- drifts from covariant B components
    BxgB_ga = (0 + (q["Bsub_s"] * q["B_t"] * (L_p - io) + q["Bsub_t"] * q["B_p"] * a_s
                    + q["Bsub_p"] * q["B_s"] * (1 + L_t) - q["Bsub_p"] * q["B_t"] * a_s
                    - q["Bsub_t"] * q["B_s"] * (L_p - io) - q["Bsub_s"] * q["B_p"] * (1 + L_t)) / sqrtg)
    ga_ga = galpha[0] * galpha[0] + galpha[1] * galpha[1] + galpha[2] * galpha[2]
    ga_gpsi = galpha[0] * gpsi[0] + galpha[1] * gpsi[1] + galpha[2] * gpsi[2]
    gpsi_gpsi = gpsi[0] * gpsi[0] + gpsi[1] * gpsi[1] + gpsi[2] * gpsi[2]
    BxgB_gpsi = (q["Bsub_t"] * q["B_p"] - q["Bsub_p"] * q["B_t"]) / sqrtg * psi_e
    # utils.py:662-720 -- GS2 normalisations
    L_ref = tab.Aminor_p
    B_ref = 2 * abs(psi_e) / (L_ref * L_ref)
    sgn = np.sign(psi_e)
    sqrt_s = np.sqrt(s)[:, None, None]
    s3, sh = s[:, None, None], shat[:, None, None]
    out = types.SimpleNamespace(ns=ns, nalpha=nalpha, nl=nl, theta_vmec=theta_vmec, phi=phi,
                                theta_pest=theta_pest)
    out.bmag = modB / B_ref
    out.gradpar_theta_pest = L_ref * (io * q["Bsup_p"]) / modB
    out.gds2 = ga_ga * L_ref * L_ref * s3
    out.gds21 = ga_gpsi * sh / B_ref
    out.gds22 = gpsi_gpsi * sh * sh / (L_ref * L_ref * B_ref * B_ref * s3)
    out.gbdrift = -1.0 * 2 * B_ref * L_ref * L_ref * sqrt_s * BxgB_ga / (modB * modB * modB) * sgn
    out.gbdrift0 = -1.0 * BxgB_gpsi * 2 * sh / (modB * modB * modB * sqrt_s) * sgn
    out.cvdrift = 1.0 * out.gbdrift - 2 * B_ref * L_ref * L_ref * sqrt_s * MU0 * dpds[:, None, None] * sgn / (
        psi_e * modB * modB)
    out.cvdrift0 = out.gbdrift0
    if full:
        out.sums = q
    return out


# ---------------------------------------------------------------------------------------
# f4: the whole Struct of vmec_fieldlines (utils.py:161-864) and of vmec_fieldlines_axisym (utils.py:872-1542)
# ---------------------------------------------------------------------------------------
def fieldlines_full(tab, alpha, theta1d=None, phi1d=None, phi_center=0.0, axisym=False):
    """Restatement of the full-output geometry: every per-point array of the Struct the reference returns
    (``utils.py:723-864``), for ``vmec_fieldlines(theta1d=...)``, ``vmec_fieldlines(phi1d=...)`` (``:364-373``) and -- with
    ``axisym=True`` -- ``vmec_fieldlines_axisym`` (uniform theta_vmec, orientation flip, ``theta_pest = theta_vmec +
    lambda``, the ``*_1`` arrays interpolated back to ``theta1d``: ``:972-1046, 1282-1369``).  ``tab`` must carry the
    ``bsupumnc`` table.  Vectors are kept as ``(3, ns, nalpha, nl)`` arrays."""
    alpha = np.atleast_1d(np.asarray(alpha, dtype=float))
    s, iota, d_iota = tab.row("s"), tab.row("iota"), tab.row("d_iota_d_s")
    dpds, shat = tab.row("d_pressure_d_s"), tab.row("shat")
    io, dio = iota[:, None, None], d_iota[:, None, None]
    ns, na = s.size, alpha.size
    xm, xn, xmq = tab.xm, tab.xn, tab.xm_nyq
    xnq = np.zeros_like(tab.xn_nyq) if axisym else tab.xn_nyq            # utils.py:913
    lmns = tab.row("lmns")
    if axisym:
        theta1d = np.asarray(theta1d, dtype=float)
        nl = theta1d.size
        theta_vmec = np.broadcast_to(np.linspace(theta1d.min(), theta1d.max(), nl), (ns, na, nl)).copy()     # utils.py:972-978
        phi_sum = np.zeros((ns, na, nl))                                  # phi is still zero when the angles are formed
        a0 = xm[:, None] * theta_vmec[0, 0, :2][None]
        R2, Z2 = tab.row("rmnc")[0] @ np.cos(a0), tab.row("zmns")[0] @ np.sin(a0)
        shift = np.pi if (R2[0] > R2[1] or Z2[1] > Z2[0]) else 0.0        # utils.py:993-1010
    else:
        shift = 0.0
        if theta1d is not None:
            theta1d = np.asarray(theta1d, dtype=float)
            nl = theta1d.size
            theta_pest = np.broadcast_to(theta1d, (ns, na, nl)).copy()
            phi = phi_center + (theta1d[None, None, :] - alpha[None, :, None]) / io                          # utils.py:371-373
        else:
            phi1d = np.asarray(phi1d, dtype=float)
            nl = phi1d.size
            phi = np.broadcast_to(phi1d, (ns, na, nl)).copy()
            theta_pest = alpha[None, :, None] + io * (phi1d[None, None, :] - phi_center)                     # utils.py:364-369
        theta_vmec = np.empty((ns, na, nl))
        for js in range(ns):                                              # utils.py:391-416
            res = lambda tv, p0, tgt: tgt - (tv + np.sum(lmns[js, :, None] * np.sin(xm[:, None] * tv - xn[:, None] * p0), axis=0))
            for ja in range(na):
                theta_vmec[js, ja] = newton(res, x0=theta_pest[js, ja], x1=theta_pest[js, ja] + 0.1, args=(phi[js, ja], theta_pest[js, ja]))
        phi_sum = phi

    def modesum(coef, m, n, kind, weight=None):
        ang = m[:, None, None, None] * (theta_vmec[None] + shift) - n[:, None, None, None] * phi_sum[None]
        basis = np.cos(ang) if kind == "c" else np.sin(ang)
        if weight is not None:
            basis = weight[:, None, None, None] * basis
        return np.einsum("ij,jikl->ikl", tab.row(coef) if isinstance(coef, str) else coef, basis)

    o = types.SimpleNamespace(ns=ns, nalpha=na, nl=nl, theta_vmec=theta_vmec)
    # utils.py:432-468
    o.R, o.d_R_d_s = modesum("rmnc", xm, xn, "c"), modesum("d_rmnc_d_s", xm, xn, "c")
    o.d_R_d_theta_vmec, o.d_R_d_phi = -modesum("rmnc", xm, xn, "s", xm), modesum("rmnc", xm, xn, "s", xn)
    o.Z, o.d_Z_d_s = modesum("zmns", xm, xn, "s"), modesum("d_zmns_d_s", xm, xn, "s")
    o.d_Z_d_theta_vmec, o.d_Z_d_phi = modesum("zmns", xm, xn, "c", xm), -modesum("zmns", xm, xn, "c", xn)
    lambdas = modesum("lmns", xm, xn, "s")
    o.d_lambda_d_s = modesum("d_lmns_d_s", xm, xn, "s")
    o.d_lambda_d_theta_vmec, o.d_lambda_d_phi = modesum("lmns", xm, xn, "c", xm), -modesum("lmns", xm, xn, "c", xn)
    if axisym:
        theta_pest = theta_vmec + lambdas                                 # utils.py:1043
        phi = phi_center + (theta_pest - alpha[None, :, None]) / io       # utils.py:1046
    o.theta_pest, o.phi = theta_pest, phi
    o.sqrt_g_vmec, o.modB, o.d_B_d_s = modesum("gmnc", xmq, xnq, "c"), modesum("bmnc", xmq, xnq, "c"), modesum("d_bmnc_d_s", xmq, xnq, "c")
    o.d_B_d_theta_vmec, o.d_B_d_phi = -modesum("bmnc", xmq, xnq, "s", xmq), modesum("bmnc", xmq, xnq, "s", xnq)
    o.B_sup_theta_vmec, o.B_sup_phi = modesum(tab.bsupumnc, xmq, xnq, "c"), modesum("bsupvmnc", xmq, xnq, "c")
    o.B_sub_s, o.B_sub_theta_vmec, o.B_sub_phi = modesum("bsubsmns", xmq, xnq, "s"), modesum("bsubumnc", xmq, xnq, "c"), modesum("bsubvmnc", xmq, xnq, "c")
    o.B_sup_theta_pest = io * o.B_sup_phi
    o.sqrt_g_vmec_alt = o.R * (o.d_Z_d_s * o.d_R_d_theta_vmec - o.d_R_d_s * o.d_Z_d_theta_vmec)
    psi_e = -tab.phiedge / (2 * np.pi)
    o.edge_toroidal_flux_over_2pi = psi_e
    # utils.py:480-508: covariant basis in Cartesian components, dual relations
    o.sinphi, o.cosphi = np.sin(phi), np.cos(phi)
    e_t = np.stack([o.d_R_d_theta_vmec * o.cosphi, o.d_R_d_theta_vmec * o.sinphi, o.d_Z_d_theta_vmec])
    e_p = np.stack([o.d_R_d_phi * o.cosphi - o.R * o.sinphi, o.d_R_d_phi * o.sinphi + o.R * o.cosphi, o.d_Z_d_phi])
    e_s = np.stack([o.d_R_d_s * o.cosphi, o.d_R_d_s * o.sinphi, o.d_Z_d_s])
    for nm, v in (("theta_vmec", e_t), ("phi", e_p), ("s", e_s)):
        setattr(o, f"d_X_d_{nm}", v[0]); setattr(o, f"d_Y_d_{nm}", v[1])
    cr = lambda a, b: np.cross(a, b, axis=0)
    grad_s, grad_t, grad_p = cr(e_t, e_p) / o.sqrt_g_vmec, cr(e_p, e_s) / o.sqrt_g_vmec, cr(e_s, e_t) / o.sqrt_g_vmec
    grad_psi = grad_s * psi_e
    a_s = o.d_lambda_d_s - (phi - phi_center) * dio                       # utils.py:519-538
    a_t, a_p = 1 + o.d_lambda_d_theta_vmec, -io + o.d_lambda_d_phi
    grad_alpha = a_s * grad_s + (a_t * grad_t + a_p * grad_p)
    grad_B = o.d_B_d_s * grad_s + o.d_B_d_theta_vmec * grad_t + o.d_B_d_phi * grad_p
    Bvec = psi_e * (a_t * e_p + (io - o.d_lambda_d_phi) * e_t) / o.sqrt_g_vmec                               # utils.py:555-577
    for nm, v in (("grad_s", grad_s), ("grad_theta_vmec", grad_t), ("grad_phi", grad_p), ("grad_psi", grad_psi),
                  ("grad_alpha", grad_alpha), ("grad_B", grad_B), ("B", Bvec)):
        for k, ax in enumerate("XYZ"):
            setattr(o, f"{nm}_{ax}", v[k])
    dt = lambda a, b: np.sum(a * b, axis=0)
    lpi = o.d_lambda_d_phi - io
    o.B_cross_grad_s_dot_grad_alpha = (o.B_sub_phi * a_t - o.B_sub_theta_vmec * lpi) / o.sqrt_g_vmec          # utils.py:588-591
    o.B_cross_grad_s_dot_grad_alpha_alternate = dt(Bvec, cr(grad_s, grad_alpha))
    o.B_cross_grad_B_dot_grad_alpha = (o.B_sub_s * o.d_B_d_theta_vmec * lpi + o.B_sub_theta_vmec * o.d_B_d_phi * a_s
                                       + o.B_sub_phi * o.d_B_d_s * a_t - o.B_sub_phi * o.d_B_d_theta_vmec * a_s
                                       - o.B_sub_theta_vmec * o.d_B_d_s * lpi - o.B_sub_s * o.d_B_d_phi * a_t) / o.sqrt_g_vmec
    o.B_cross_grad_B_dot_grad_alpha_alternate = dt(Bvec, cr(grad_B, grad_alpha))
    o.grad_alpha_dot_grad_alpha, o.grad_alpha_dot_grad_psi, o.grad_psi_dot_grad_psi = dt(grad_alpha, grad_alpha), dt(grad_alpha, grad_psi), dt(grad_psi, grad_psi)
    o.B_cross_grad_B_dot_grad_psi = (o.B_sub_theta_vmec * o.d_B_d_phi - o.B_sub_phi * o.d_B_d_theta_vmec) / o.sqrt_g_vmec * psi_e
    o.B_cross_kappa_dot_grad_psi = o.B_cross_grad_B_dot_grad_psi / o.modB
    o.B_cross_kappa_dot_grad_alpha = o.B_cross_grad_B_dot_grad_alpha / o.modB + MU0 * dpds[:, None, None] / psi_e
    # utils.py:662-720
    L_ref = tab.Aminor_p
    B_ref = 2 * abs(psi_e) / (L_ref * L_ref)
    sgn, sqrt_s, s3, sh = np.sign(psi_e), np.sqrt(s)[:, None, None], s[:, None, None], shat[:, None, None]
    o.L_reference, o.B_reference, o.toroidal_flux_sign = L_ref, B_ref, sgn
    o.bmag, o.gradpar_theta_pest, o.gradpar_phi = o.modB / B_ref, L_ref * o.B_sup_theta_pest / o.modB, L_ref * o.B_sup_phi / o.modB
    o.gds2 = o.grad_alpha_dot_grad_alpha * L_ref * L_ref * s3
    o.gds21 = o.grad_alpha_dot_grad_psi * sh / B_ref
    o.gds22 = o.grad_psi_dot_grad_psi * sh * sh / (L_ref * L_ref * B_ref * B_ref * s3)
    o.gbdrift = -1.0 * 2 * B_ref * L_ref * L_ref * sqrt_s * o.B_cross_grad_B_dot_grad_alpha / o.modB ** 3 * sgn
    o.gbdrift0 = -1.0 * o.B_cross_grad_B_dot_grad_psi * 2 * sh / (o.modB ** 3 * sqrt_s) * sgn
    o.cvdrift = o.gbdrift - 2 * B_ref * L_ref * L_ref * sqrt_s * MU0 * dpds[:, None, None] * sgn / (psi_e * o.modB * o.modB)
    o.cvdrift0 = o.gbdrift0
    if axisym:                                                           # utils.py:1282-1286, 1333-1392
        B_p = np.sqrt(o.B_sub_theta_vmec * abs(psi_e) * io)
        back = lambda a: np.interp(theta1d, theta_pest[0][0], a[0][0])
        for nm in ("gradpar_theta_pest", "bmag", "cvdrift", "gbdrift", "gds21", "gds22", "gds2", "R", "Z"):
            setattr(o, nm + "_1", back(getattr(o, nm)))
        o.cvdrift0_1 = o.gbdrift0_1 = back(o.cvdrift0)
        o.B_p_1 = back(B_p)
        o.Rprime_1 = back(o.d_R_d_s) / psi_e / iota * o.R_1 * o.B_p_1
        o.Zprime_1 = back(o.d_Z_d_s) / psi_e / iota * o.R_1 * o.B_p_1
    return o


def dpdrho_of(fl, js=0, ja=0):
    """``dPdrho = -0.5*mean((cvdrift-gbdrift)*bmag**2)`` (``ball_scan.py:262``, ``utils.py:1657``)."""
    return -1.0 * 0.5 * np.mean((fl.cvdrift[js][ja] - fl.gbdrift[js][ja]) * fl.bmag[js][ja] ** 2)


def theta0_shift(fl, theta0, js=0, ja=0):
    """``cvdrift_fth, gds2_fth`` (``ball_scan.py:267-268``, ``utils.py:1659-1660``)."""
    cv = fl.cvdrift[js][ja] + theta0 * fl.cvdrift0[js][ja]
    gd = fl.gds2[js][ja] + 2 * theta0 * fl.gds21[js][ja] + theta0 ** 2 * fl.gds22[js][ja]
    return cv, gd


# ---------------------------------------------------------------------------------------
# a5-a8: coefficients, discretisation, eigen-solve, post-processing (utils.py:1550-1624)
# ---------------------------------------------------------------------------------------
def gcf(dPdrho, B, gradpar, cvdrift, gds2):
    """``g, c, f`` (``utils.py:1560-1562``)."""
    g = np.abs(gradpar) * gds2 / (B)
    c = -1 * dPdrho * cvdrift * 1 / (np.abs(gradpar) * B)
    f = gds2 / B ** 2 * 1 / (np.abs(gradpar) * B)
    return g, c, f


def discretise(theta, g, c, f):
    """Half-grid g, spacing h and the three diagonals of ``A = F^-1 (D g D + c)``
    (``utils.py:1564-1592``).  Returns ``(h, g_u, c_u, f_u, sub, diag, sup)``; row ``r`` of A
    belongs to grid point ``r+1``; ``sub[k] = A[k+1,k]``, ``sup[k] = A[k,k+1]``."""
    n = len(g)
    tu = np.linspace(theta[0], theta[-1], n)
    g_u, c_u, f_u = np.interp(tu, theta, g), np.interp(tu, theta, c), np.interp(tu, theta, f)
    th_half = (tu[:-1] + tu[1:]) / 2
    h = np.diff(th_half)[2]
    gh = np.interp(th_half, theta, g)
    sub = gh[1:-1] / f_u[2:-1] * 1 / h ** 2
    diag = -(gh[1:] + gh[:-1]) / f_u[1:-1] * 1 / h ** 2 + c_u[1:-1] / f_u[1:-1]
    sup = gh[1:-1] / f_u[1:-2] * 1 / h ** 2
    return h, g_u, c_u, f_u, sub, diag, sup


def postprocess(v, h, g_u, c_u, f_u):
    """Max-abs normalise, Dirichlet pad, dX stencil and the Simpson Rayleigh quotient
    (``utils.py:1602-1621``).  ``v`` is the interior eigenvector (length N-2)."""
    n = len(g_u)
    X = np.zeros((n,))
    dX = np.zeros((n,))
    X[1:-1] = np.reshape(v, (-1,)) / np.max(np.abs(v))
    dX[0] = (-1.5 * X[0] + 2 * X[1] - 0.5 * X[2]) / h
    dX[1] = (X[2] - X[0]) / (2 * h)
    dX[-2] = (X[-1] - X[-3]) / (2 * h)
    dX[-1] = (0.5 * X[-3] - 2 * X[-2] + 1.5 * 0.0) / (h)
    dX[2:-2] = 2 / (3 * h) * (X[3:-1] - X[1:-3]) - (X[4:] - X[0:-4]) / (12 * h)
    Y0 = -g_u * dX ** 2 + c_u * X ** 2
    Y1 = f_u * X ** 2
    gam = simpson(Y0) / simpson(Y1)
    return gam, X, dX


def gamma_ball_full(dPdrho, theta_PEST, B, gradpar, cvdrift, gds2, vguess=None, sigma0=0.42,
                    tol=5.0e-7, method="arpack", info=None):
    """Restatement of ``gamma_ball_full`` (``utils.py:1550-1624``).

    ``method="arpack"``  -- the reference's own route: dense ``A`` + ARPACK shift-invert
                           ``eigs(A, 1, sigma=sigma0, v0=vguess, tol=tol, OPpart='r')``
                           (``utils.py:1582-1597``); ``tol=0`` is the *converged* reference
                           the CUDA engine is compared with.
    ``method="tridiag"`` -- independent cross-check: LAPACK on the symmetrised pencil
                           ``S = F^-1/2 K F^-1/2`` (same spectrum, ``x = F^-1/2 u``), picking
                           the eigenvalue nearest ``sigma0`` like shift-invert does.
    ``method="lambda_max"`` -- as "tridiag" but picking the largest eigenvalue (what the
                           engine computes; identical whenever ``sigma0 > (l1+l2)/2``).
    """
    g, c, f = gcf(dPdrho, B, gradpar, cvdrift, gds2)
    h, g_u, c_u, f_u, sub, diag, sup = discretise(theta_PEST, g, c, f)
    if method == "arpack":
        from scipy.sparse.linalg import eigs
        A = np.diag(sub, -1) + np.diag(diag, 0) + np.diag(sup, 1)
        w, v = eigs(A, 1, sigma=sigma0, v0=vguess, tol=tol, OPpart="r")
        vec, lam = v[:, 0].real, w[0].real
    else:
        from scipy.linalg import eigh_tridiagonal
        fi = f_u[1:-1]
        e = sup * np.sqrt(fi[:-1] / fi[1:])              # = gh_j / (h^2 sqrt(f_j f_{j+1}))
        if method == "lambda_max":
            m = len(diag)
            lam_all, U = eigh_tridiagonal(diag, e, select="i", select_range=(m - 2, m - 1))
            k = 1
        else:
            lam_all, U = eigh_tridiagonal(diag, e)
            k = int(np.argmin(np.abs(lam_all - sigma0)))
        lam = lam_all[k]
        vec = U[:, k] / np.sqrt(fi)
        if info is not None:
            info["gap"] = float(lam_all[k] - lam_all[k - 1]) if k > 0 else np.inf
    if info is not None:
        info["lambda_matrix"] = float(lam)
        info["h"] = float(h)
    gam, X, dX = postprocess(vec, h, g_u, c_u, f_u)
    return gam, X, dX, g_u, c_u, f_u


# ---------------------------------------------------------------------------------------
# a9: adjoint gradient (utils.py:1632-1728)
# ---------------------------------------------------------------------------------------
DEL_ALPHA = 0.004                                           # utils.py:1639


def adjoint_terms(lam, X, dX, f_arr, g_p, c_p, f_p):
    """Hellmann-Feynman contraction ``[int c_p X^2 - int g_p dX^2 - lam int f_p X^2]/int f X^2``
    (``utils.py:1676-1680``, ``:1721-1725``)."""
    Y1 = simpson(f_arr * X ** 2)
    return simpson(c_p * X ** 2) / Y1 - simpson(g_p * dX ** 2) / Y1 - lam * simpson(f_p * X ** 2) / Y1


def obj_w_grad(x0, fieldline_fn, theta, vguess00, sigma00=0.42, tol=5.0e-7, method="arpack"):
    """Restatement of ``obj_w_grad`` (``utils.py:1632-1728``).  ``fieldline_fn(alphas)`` must
    return the geometry namespace of :func:`fieldlines` for one surface and the given alphas
    (the reference calls ``vmec_fieldlines(vs, rho_val, [a-d/2, a, a+d/2], theta1d=theta)``)."""
    alpha_val, theta0_val = x0
    f1 = fieldline_fn(np.array([alpha_val - 0.5 * DEL_ALPHA, alpha_val, alpha_val + 0.5 * DEL_ALPHA]))

    def coeffs(ja):
        dP = dpdrho_of(f1, 0, ja)
        cv, gd = theta0_shift(f1, theta0_val, 0, ja)
        return dP, gcf(dP, f1.bmag[0][ja], f1.gradpar_theta_pest[0][ja], cv, gd)

    dP, _ = coeffs(1)
    bmag, gradpar = f1.bmag[0][1], f1.gradpar_theta_pest[0][1]
    cv, gd = theta0_shift(f1, theta0_val, 0, 1)
    lam, X, dX, g_arr, c_arr, f_arr = gamma_ball_full(dP, theta, bmag, gradpar, cv, gd, vguess00,
                                                      sigma00, tol=tol, method=method)
    # utils.py:1669-1673 -- analytic theta0 derivatives
    d_gds2 = 2 * f1.gds21[0][1] + 2 * theta0_val * f1.gds22[0][1]
    g_t0 = np.abs(gradpar) * d_gds2 / (bmag)
    c_t0 = -1 * dP * f1.cvdrift0[0][1] * 1 / (np.abs(gradpar) * bmag)
    f_t0 = d_gds2 / bmag ** 2 * 1 / (np.abs(gradpar) * bmag)
    jac_theta0 = adjoint_terms(lam, X, dX, f_arr, g_t0, c_t0, f_t0)
    # utils.py:1683-1718 -- central differences in alpha
    (_, (g_r, c_r, f_r)), (_, (g_l, c_l, f_l)) = coeffs(2), coeffs(0)
    jac_alpha = adjoint_terms(lam, X, dX, f_arr, (g_r - g_l) / DEL_ALPHA, (c_r - c_l) / DEL_ALPHA,
                              (f_r - f_l) / DEL_ALPHA)
    return -1 * lam, np.array([-1 * jac_alpha, -1 * jac_theta0])


# ---------------------------------------------------------------------------------------
# a10: coarse scan + argmax guards (ball_scan.py:196-295)
# ---------------------------------------------------------------------------------------
def scan_theta_grid(mpol, ntor, theta_fac=4):
    """``ntheta`` and the theta grid of ``ball_scan.py:201-208``."""
    ntheta = int(2 * mpol * theta_fac) + 1 if ntor == 0 else int(2 * mpol * ntor * theta_fac) + 1
    return np.linspace(-theta_fac * np.pi, theta_fac * np.pi, ntheta)


def default_vguess(theta, theta_fac=4):
    """Gaussian-like starting vector (``ball_scan.py:209,233``)."""
    return (1 - np.tanh(theta[1:-1] / np.pi) ** 2) * np.cos(theta[1:-1] / (2 * theta_fac))


def coarse_scan_surface(fieldline_fn, theta, alpha_scan, theta0_scan, tol=5.0e-7, method="arpack",
                        sigma=1.0):
    """The (alpha, theta0) grid of one surface with the warm start chained through both
    loops (``ball_scan.py:248-274``).  Returns ``gamma_scan (nalpha, ntheta0)``."""
    vguess = default_vguess(theta)
    gam = np.zeros((len(alpha_scan), len(theta0_scan)))
    for i, al in enumerate(alpha_scan):
        fl = fieldline_fn(np.array([al]))
        dP = dpdrho_of(fl)
        for j, th0 in enumerate(theta0_scan):
            cv, gd = theta0_shift(fl, th0)
            lam, X, *_ = gamma_ball_full(dP, theta, fl.bmag[0][0], fl.gradpar_theta_pest[0][0], cv, gd,
                                         vguess, sigma, tol=tol, method=method)
            vguess = X[1:-1]
            gam[i, j] = lam
    return gam


def argmax_with_guards(gamma_scan):
    """``ball_scan.py:279-295``: returns ``(ia, it, sigma0)``; ``ia = it = -1`` encodes the
    all-zero guard (alpha_guess = theta0_guess = 0, sigma0 = 0.05); ties take the first index
    in row-major order (``np.where(...)[k][0]``)."""
    gmax = np.max(gamma_scan)
    if gmax == 0.0:
        return -1, -1, 0.05
    idx = np.where(gamma_scan == gmax)
    ia, it = int(idx[0][0]), int(idx[1][0])
    return ia, it, float(1.3 * np.abs(gamma_scan[ia, it]) + 0.05)


# ---------------------------------------------------------------------------------------
# s-alpha marginal-stability classifier (tests/shifted-circle-s-alpha/bishop_ball_s-alpha.py:20-115)
# ---------------------------------------------------------------------------------------
def check_ball(shat_n, alpha_n, theta0, span=61, ntheta=1601):
    """Newcomb/shooting zero-crossing test at lambda = 0, restating ``check_ball``
    (``bishop_ball_s-alpha.py:20-115``; with its ``diff = 0`` all the c2/f2 terms vanish).
    ``span, ntheta = 20, 401`` gives ``check_ball_long`` (``:118-210``)."""
    th = np.linspace(-span * np.pi, span * np.pi, ntheta)
    dth = np.diff(th)
    lam = shat_n * (th - theta0) - alpha_n * (np.sin(th) - np.sin(theta0))
    g = 1 + lam ** 2
    c = alpha_n * (np.cos(th) + np.sin(th) * lam)
    gh = np.zeros(ntheta)
    gh[1:] = 0.5 * (g[1:] + g[:-1])
    c1 = np.zeros(ntheta)
    c1[1:-1] = 0.5 * (-dth[1:] * c[1:-1] - dth[:-1] * c[1:-1])
    g1 = np.zeros(ntheta)
    g2 = np.zeros(ntheta)
    g1[1:] = gh[1:] / dth
    g2[1:] = 1.0 / (gh[1:] / dth)
    psi = np.zeros(ntheta)
    psi[1] = dth[0]
    psi_prime = (psi[1] / g2[1]) * 0.5
    for ig in range(1, ntheta - 1):
        psi_prime = psi_prime + c1[ig] * psi[ig]
        psi[ig + 1] = (g1[ig + 1] * psi[ig] + psi_prime) * g2[ig + 1]
    return int(np.any(psi[1:-1] * psi[2:] <= 0))


# ---------------------------------------------------------------------------------------
# f4 (part): the finite-difference helpers of the curvature penalty (utils.py:1737-1943)
# ---------------------------------------------------------------------------------------
def derm(arr, ch, par="e"):
    """Restatement of ``derm`` (``utils.py:1737-1807``): un-normalised central differences ``a[i+1] - a[i-1]`` along
    (``ch='l'``) or across (``ch='r'``) flux surfaces; end points ``2 (a[1] - a[0])`` / ``2 (a[-1] - a[-2])`` -- or 0 along
    a surface when the input has even parity.  Shapes as the reference returns them: a 1-D input comes back as
    ``(1, n)`` for ``'l'`` and ``(n, 1)`` for ``'r'``."""
    a = np.asarray(arr, dtype=float)
    if a.ndim == 1:
        a = a.reshape(1, -1) if ch == "l" else a.reshape(-1, 1)
    axis = 1 if ch == "l" else 0
    a = np.moveaxis(a, axis, -1)
    d = np.zeros_like(a)
    d[..., 1:-1] = (a[..., 1:-1] - a[..., :-2]) + (a[..., 2:] - a[..., 1:-1])      # np.diff + np.diff, as the reference adds them
    if not (ch == "l" and par == "e"):
        d[..., 0] = 2 * (a[..., 1] - a[..., 0])
        d[..., -1] = 2 * (a[..., -1] - a[..., -2])
    return np.moveaxis(d, -1, axis)


def dermv(arr, brr, ch, par="e"):
    """Restatement of ``dermv`` (``utils.py:1810-1943``): derivative of ``arr`` with respect to the non-uniformly spaced
    ``brr``; interior ``(a[i+1]/h1^2 + a[i] (1/h0^2 - 1/h1^2) - a[i-1]/h0^2) / (1/h1 + 1/h0)``; end points one-sided
    (second order for a 1-D odd-parity array along a surface, first order otherwise; 0 for even parity along a
    surface).  The reference's 1-D ``'r'`` branch stops in ``pdb.set_trace()`` (``utils.py:1866``) and is not restated."""
    a, b = np.asarray(arr, dtype=float), np.asarray(brr, dtype=float)
    one_d = a.ndim == 1
    if one_d:
        if ch != "l":
            raise NotImplementedError("dermv(1-D, 'r') enters the debugger in the reference (utils.py:1866)")
        a, b = a.reshape(1, -1), b.reshape(1, -1)
    axis = 1 if ch == "l" else 0
    a, b = np.moveaxis(a, axis, -1), np.moveaxis(b, axis, -1)
    d = np.zeros_like(a)
    h1 = b[..., 2:] - b[..., 1:-1]
    h0 = b[..., 1:-1] - b[..., :-2]
    d[..., 1:-1] = (a[..., 2:] / h1 ** 2 + a[..., 1:-1] * (1 / h0 ** 2 - 1 / h1 ** 2) - a[..., :-2] / h0 ** 2) / (1 / h1 + 1 / h0)
    if ch == "l" and par == "e":
        pass                                                                      # even parity: zero slope at both ends
    elif one_d:
        d[..., 0] = (4 * a[..., 1] - 3 * a[..., 0] - a[..., 2]) / (2 * (b[..., 1] - b[..., 0]))
        d[..., -1] = (-4 * a[..., -2] + 3 * a[..., -1] + a[..., -3]) / (2 * (b[..., -1] - b[..., -2]))
    else:
        d[..., 0] = 2 * (a[..., 1] - a[..., 0]) / (2 * (b[..., 1] - b[..., 0]))
        d[..., -1] = 2 * (a[..., -1] - a[..., -2]) / (2 * (b[..., -1] - b[..., -2]))
    return np.moveaxis(d, -1, axis)
