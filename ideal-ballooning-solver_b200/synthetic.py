"""Synthetic VMEC-shaped equilibria ("smooth random Fourier spectra").

There is no VMEC in this environment, so the benchmark configurations of
BASELINE.json are driven by synthetic ``wout``-like objects that have exactly
the attributes the reference reads from ``simsopt.mhd.vmec.Vmec.wout``
(``/root/reference/utils.py:58-135``): 2-D tables laid out ``(mn, ns)``, mode
numbers in VMEC order, 1-D profiles and a few scalars.  Mode counts follow the
reference's input templates (``input.template_D3D:5-16``,
``input.template_NCSX:5-11``, ``input.template_HBERG:5-12``) and the Nyquist
rule seen in ``tests/comparn_w_COBRAVMEC/wout_NCSX_op.nc`` (mpol+3 / ntor+2).

The surfaces (R, Z), the stream function lambda, iota and the pressure are
drawn as smooth random spectra; the Jacobian and the magnetic field tables
(gmnc, bmnc, bsup*, bsub*) are then *derived* from them on a collocation grid
and projected back on the Nyquist modes, so the geometry is self consistent
(sqrt(g) != 0, |B| > 0, 1 + dlambda/dtheta > 0) although it is not a
force-balanced MHD equilibrium.  Parity is code-vs-code, so that is enough.
"""
from __future__ import annotations

import types

import numpy as np

MU0 = 4 * np.pi * 1.0e-7

#: mode counts of the three device templates (reference input.template_*)
KINDS = {
    #            nfp mpol ntor ns   R0    a     kappa  B0   iota0  iota1  beta0
    "d3d":   dict(nfp=1, mpol=80, ntor=0, ns=64, R0=1.70, a=0.60, kappa=1.60, B0=2.0,
                  iota0=0.95, iota1=-0.68, beta0=0.035, eps3d=0.0),
    "ncsx":  dict(nfp=3, mpol=11, ntor=11, ns=64, R0=1.42, a=0.32, kappa=1.45, B0=1.6,
                  iota0=0.40, iota1=0.25, beta0=0.05, eps3d=0.10),
    "hberg": dict(nfp=2, mpol=11, ntor=11, ns=128, R0=1.00, a=0.17, kappa=1.25, B0=1.0,
                  iota0=0.42, iota1=0.12, beta0=0.04, eps3d=0.08),
}


def vmec_mode_table(mpol_hi: int, ntor: int, nfp: int):
    """(xm, xn) in VMEC order: m=0 has n=0..ntor, m>0 has n=-ntor..ntor;
    ``xn`` is already multiplied by ``nfp`` as in the wout files."""
    xm, xn = [], []
    for m in range(mpol_hi + 1):
        for n in range(0 if m == 0 else -ntor, ntor + 1):
            xm.append(m)
            xn.append(n * nfp)
    return np.array(xm, dtype=float), np.array(xn, dtype=float)


def _radial(m, s, b):
    """profile s^{m/2}(1+b s) and its s-derivative (regular for the s used)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        p = s ** (0.5 * m) * (1.0 + b * s)
        dp = np.where(s > 0,
                      0.5 * m * s ** (0.5 * m - 1.0) * (1.0 + b * s) + s ** (0.5 * m) * b,
                      0.0)
    return p, dp


def make_equilibrium(kind: str = "ncsx", seed: int = 0, ns: int | None = None):
    """Return a wout-like namespace for ``kind`` in {"d3d", "ncsx", "hberg"}."""
    cfg = dict(KINDS[kind])
    if ns is not None:
        cfg["ns"] = ns
    rng = np.random.default_rng(20261018 + 7919 * seed + sum(map(ord, kind)))
    nfp, mpol, ntor, ns = cfg["nfp"], cfg["mpol"], cfg["ntor"], cfg["ns"]
    R0, a, kappa, B0 = cfg["R0"], cfg["a"], cfg["kappa"], cfg["B0"]

    xm, xn = vmec_mode_table(mpol - 1, ntor, nfp)
    mpol_nyq = mpol + 3
    ntor_nyq = ntor + 2 if ntor > 0 else 0
    xm_nyq, xn_nyq = vmec_mode_table(mpol_nyq, ntor_nyq, nfp)
    mnmax, mnmax_nyq = len(xm), len(xm_nyq)

    s_full = np.linspace(0.0, 1.0, ns)
    ds = s_full[1] - s_full[0]
    s_half = s_full[1:] - 0.5 * ds

    # ---- smooth random spectra for R, Z, lambda ------------------------------------
    decay = 0.7
    nn = np.abs(xn) / nfp
    env = np.exp(-decay * (xm + nn))
    amp_r = 0.18 * a * env * rng.standard_normal(mnmax)
    amp_z = 0.18 * a * env * rng.standard_normal(mnmax)
    amp_l = 0.05 * env * rng.standard_normal(mnmax)
    if ntor > 0:
        three_d = np.where(xn != 0, cfg["eps3d"] / 0.18, 1.0)
        amp_r *= three_d
        amp_z *= three_d
    else:
        amp_l *= 0.5
    b_r = 0.3 * rng.standard_normal(mnmax)
    b_z = 0.3 * rng.standard_normal(mnmax)
    b_l = 0.3 * rng.standard_normal(mnmax)
    i00 = int(np.where((xm == 0) & (xn == 0))[0][0])
    i10 = int(np.where((xm == 1) & (xn == 0))[0][0])
    amp_r[i00], b_r[i00] = R0, -0.04 * a / R0          # Shafranov-like shift
    amp_z[i00] = 0.0
    amp_l[i00] = 0.0
    amp_r[i10], b_r[i10] = a, 0.05
    amp_z[i10], b_z[i10] = kappa * a, -0.03

    def tables(s):
        out = []
        for amp, b in ((amp_r, b_r), (amp_z, b_z), (amp_l, b_l)):
            p, dp = _radial(xm[:, None], s[None, :], b[:, None])
            out.append(amp[:, None] * p)
            out.append(amp[:, None] * dp)
        return out  # r, r_s, z, z_s, l, l_s each (mnmax, len(s))

    iota_f = lambda s: cfg["iota0"] + cfg["iota1"] * s
    p0 = cfg["beta0"] * B0 * B0 / (2 * MU0)
    pres_f = lambda s: p0 * (1.0 - 0.9 * s - 0.1 * s * s) ** 2
    phiedge = np.pi * a * a * B0                         # > 0, so psi_edge < 0 and sqrt(g) < 0
    psi_e = -phiedge / (2 * np.pi)

    # ---- collocation grid over one field period ---------------------------------------
    nth = 4 * (mpol_nyq + 1)
    nze = 4 * (ntor_nyq + 1) if ntor > 0 else 1
    th = (np.arange(nth) + 0.5) * 2 * np.pi / nth
    ze = np.arange(nze) * 2 * np.pi / (nfp * nze)
    TH, ZE = np.meshgrid(th, ze, indexing="ij")
    TH, ZE = TH.ravel(), ZE.ravel()
    ang = xm[:, None] * TH[None, :] - xn[:, None] * ZE[None, :]
    cosb, sinb = np.cos(ang), np.sin(ang)
    ang_n = xm_nyq[:, None] * TH[None, :] - xn_nyq[:, None] * ZE[None, :]
    cosn, sinn = np.cos(ang_n), np.sin(ang_n)
    norm = np.where((xm_nyq == 0) & (xn_nyq == 0), 1.0, 2.0) / TH.size

    def fields(s):
        """sqrt(g), |B|, B^theta, B^phi, B_s, B_theta, B_phi on the grid, each (len(s), npts)."""
        r, r_s, z, z_s, l, l_s = tables(s)
        R = r.T @ cosb
        R_s = r_s.T @ cosb
        R_t = -(r * xm[:, None]).T @ sinb
        R_p = (r * xn[:, None]).T @ sinb
        Z_s = z_s.T @ sinb
        Z_t = (z * xm[:, None]).T @ cosb
        Z_p = -(z * xn[:, None]).T @ cosb
        L_t = (l * xm[:, None]).T @ cosb
        L_p = -(l * xn[:, None]).T @ cosb
        sqrtg = R * (Z_s * R_t - R_s * Z_t)
        iota = iota_f(s)[:, None]
        Bt = psi_e * (iota - L_p) / sqrtg
        Bp = psi_e * (1.0 + L_t) / sqrtg
        g_tt = R_t * R_t + Z_t * Z_t
        g_tp = R_t * R_p + Z_t * Z_p
        g_pp = R_p * R_p + Z_p * Z_p + R * R
        g_st = R_s * R_t + Z_s * Z_t
        g_sp = R_s * R_p + Z_s * Z_p
        B_t = Bt * g_tt + Bp * g_tp
        B_p = Bt * g_tp + Bp * g_pp
        B_s = Bt * g_st + Bp * g_sp
        modB = np.sqrt(Bt * B_t + Bp * B_p)
        assert np.all(1.0 + L_t > 0.05), "synthetic lambda too large: 1+dlambda/dtheta <= 0"
        return sqrtg, modB, Bt, Bp, B_s, B_t, B_p

    def project(F, basis):
        return (F @ basis.T) * norm[None, :]            # (len(s), mnmax_nyq)

    sqrtg, modB, Bt, Bp, _, B_t, B_p = fields(s_half)
    assert np.all(sqrtg < 0), "synthetic surfaces self-intersect (sqrt(g) changes sign)"
    pad = lambda T: np.concatenate([np.zeros((T.shape[1], 1)), T.T], axis=1)   # (mn, ns), col 0 unused
    gmnc = pad(project(sqrtg, cosn))
    bmnc = pad(project(modB, cosn))
    bsupumnc = pad(project(Bt, cosn))
    bsupvmnc = pad(project(Bp, cosn))
    bsubumnc = pad(project(B_t, cosn))
    bsubvmnc = pad(project(B_p, cosn))
    # B_s lives on the full mesh (utils.py:95-98); the axis value is extrapolated.
    B_s_full = fields(s_full[1:])[4]
    bs = project(B_s_full, sinn).T                      # (mn_nyq, ns-1)
    bsubsmns = np.concatenate([2 * bs[:, :1] - bs[:, 1:2], bs], axis=1)

    r, _, z, _, _, _ = tables(s_full)
    lh = tables(s_half)[4]
    w = types.SimpleNamespace()
    w.rmnc, w.zmns = r, z
    w.lmns = np.concatenate([np.zeros((mnmax, 1)), lh], axis=1)
    w.gmnc, w.bmnc = gmnc, bmnc
    w.bsupumnc, w.bsupvmnc = bsupumnc, bsupvmnc
    w.bsubsmns, w.bsubumnc, w.bsubvmnc = bsubsmns, bsubumnc, bsubvmnc
    w.pres = np.concatenate([[0.0], pres_f(s_half)])
    w.iotas = np.concatenate([[0.0], iota_f(s_half)])
    w.chi = np.concatenate([[0.0], psi_e * np.cumsum(iota_f(s_half)) * ds])
    w.phi = phiedge * s_full
    w.xm, w.xn, w.xm_nyq, w.xn_nyq = xm, xn, xm_nyq, xn_nyq
    w.raxis_cc = np.array([r[k, 0] for k in range(mnmax) if xm[k] == 0])
    w.Aminor_p = float(a * np.sqrt(kappa))
    w.mnmax, w.mnmax_nyq = int(mnmax), int(mnmax_nyq)
    w.nfp, w.ns, w.mpol, w.ntor = int(nfp), int(ns), int(mpol), int(ntor)
    w.kind = kind
    return w


# ---------------------------------------------------------------------------------------
# analytic s-alpha model (reference tests/shifted-circle-s-alpha/bishop_ball_s-alpha.py:30-45)
# ---------------------------------------------------------------------------------------
def s_alpha_coefficients(shat, alpha, theta0, theta):
    """g, c, f of the shifted-circle model on ``theta``; broadcasting over leading dims.

    g = f = 1 + Lambda^2, c = alpha (cos th + Lambda sin th),
    Lambda = shat (th - th0) - alpha (sin th - sin th0).
    Fed to ``gamma_ball_full`` as B = gradpar = 1, gds2 = g, cvdrift = c/alpha, dPdrho = -alpha.
    """
    shat = np.asarray(shat, dtype=float)[..., None]
    alpha = np.asarray(alpha, dtype=float)[..., None]
    theta0 = np.asarray(theta0, dtype=float)[..., None]
    lam = shat * (theta - theta0) - alpha * (np.sin(theta) - np.sin(theta0))
    g = 1.0 + lam * lam
    c = alpha * (np.cos(theta) + np.sin(theta) * lam)
    return g, c, g.copy()
