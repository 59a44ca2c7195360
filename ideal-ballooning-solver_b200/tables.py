"""Radial splines of the VMEC Fourier tables -> per-surface tables (host-side precompute).

This is the step immediately upstream of the geometry kernel.
It mirrors ``vmec_splines`` (``/root/reference/utils.py:37-158``) and the
per-surface evaluation at the top of ``vmec_fieldlines`` (``utils.py:311-357``), but
vectorised: one not-a-knot cubic ``BSpline`` per *table* (all modes at once) instead of
one FITPACK ``InterpolatedUnivariateSpline`` object per mode (5 372 Python objects and
4 596 ``splev`` calls per field line in the reference).  Both are the unique C2 cubic
interpolant with not-a-knot end conditions, so they agree to rounding.

The output layout IS the geometry kernel's input contract (``include/ibs_b200.h``):

  ``tab_mn``  float64 ``(ns, 6, mnmax)``      rows: rmnc, zmns, lmns, d_rmnc_d_s, d_zmns_d_s, d_lmns_d_s
  ``tab_nyq`` float64 ``(ns, 7, mnmax_nyq)``  rows: gmnc, bmnc, d_bmnc_d_s, bsupvmnc, bsubsmns, bsubumnc, bsubvmnc
  ``scal``    float64 ``(ns, 8)``             s, iota, d_iota_d_s, d_pressure_d_s, shat, pressure, 0, 0
"""
from __future__ import annotations

import dataclasses
import types

import numpy as np
from scipy.interpolate import make_interp_spline

TAB_MN_ROWS = ("rmnc", "zmns", "lmns", "d_rmnc_d_s", "d_zmns_d_s", "d_lmns_d_s")
TAB_NYQ_ROWS = ("gmnc", "bmnc", "d_bmnc_d_s", "bsupvmnc", "bsubsmns", "bsubumnc", "bsubvmnc")
SCAL_COLS = ("s", "iota", "d_iota_d_s", "d_pressure_d_s", "shat", "pressure")
NSCAL = 8


@dataclasses.dataclass
class SurfaceTables:
    """Per-surface Fourier tables + scalars, ready to be copied to the device."""
    tab_mn: np.ndarray      # (ns, 6, mnmax)
    tab_nyq: np.ndarray     # (ns, 7, mnmax_nyq)
    scal: np.ndarray        # (ns, 8)
    xm: np.ndarray
    xn: np.ndarray
    xm_nyq: np.ndarray
    xn_nyq: np.ndarray
    phiedge: float
    Aminor_p: float
    nfp: int
    bsupumnc: np.ndarray = None     # (ns, mnmax_nyq): only the full-output geometry needs it (B_sup_theta_vmec, utils.py:460)
    raxis_cc: np.ndarray = None     # magnetic-axis Fourier coefficients (utils.py:1032, axisymmetric routine)

    @property
    def ns(self):
        return self.tab_mn.shape[0]

    @property
    def s(self):
        return self.scal[:, 0]

    def row(self, name):
        if name in TAB_MN_ROWS:
            return self.tab_mn[:, TAB_MN_ROWS.index(name), :]
        if name in TAB_NYQ_ROWS:
            return self.tab_nyq[:, TAB_NYQ_ROWS.index(name), :]
        return self.scal[:, SCAL_COLS.index(name)]

    def select(self, idx):
        idx = np.atleast_1d(idx)
        return dataclasses.replace(self, tab_mn=np.ascontiguousarray(self.tab_mn[idx]),
                                   tab_nyq=np.ascontiguousarray(self.tab_nyq[idx]),
                                   scal=np.ascontiguousarray(self.scal[idx]),
                                   bsupumnc=None if self.bsupumnc is None else np.ascontiguousarray(self.bsupumnc[idx]))


class RadialSplines:
    """Vectorised equivalent of the reference's ``vmec_splines`` Struct."""

    def __init__(self, wout):
        w = wout
        ns = int(w.ns)
        self.s_full = np.linspace(0.0, 1.0, ns)
        ds = self.s_full[1] - self.s_full[0]
        self.s_half = self.s_full[1:] - 0.5 * ds
        full = lambda T: make_interp_spline(self.s_full, np.asarray(T, float).T, k=3)
        half = lambda T: make_interp_spline(self.s_half, np.asarray(T, float).T[1:], k=3)
        # utils.py:58-70 (full mesh: rmnc, zmns; half mesh: lmns)
        self.rmnc, self.zmns, self.lmns = full(w.rmnc), full(w.zmns), half(w.lmns)
        # utils.py:82-107 (bsubsmns is on the full mesh, the others on the half mesh)
        self.gmnc, self.bmnc = half(w.gmnc), half(w.bmnc)
        self.bsupvmnc = half(w.bsupvmnc)
        self.bsupumnc = half(w.bsupumnc) if hasattr(w, "bsupumnc") else None
        self.raxis_cc = np.asarray(w.raxis_cc, float) if hasattr(w, "raxis_cc") else None
        self.bsubsmns = full(w.bsubsmns)
        self.bsubumnc, self.bsubvmnc = half(w.bsubumnc), half(w.bsubvmnc)
        # utils.py:110-119
        self.pressure = half(w.pres)
        self.iota = half(w.iotas)
        self.phiedge = float(w.phi[-1])                  # utils.py:122
        self.Aminor_p = float(w.Aminor_p)
        self.nfp = int(w.nfp)
        self.xm, self.xn = np.asarray(w.xm, float), np.asarray(w.xn, float)
        self.xm_nyq, self.xn_nyq = np.asarray(w.xm_nyq, float), np.asarray(w.xn_nyq, float)

    def evaluate(self, s) -> SurfaceTables:
        s = np.atleast_1d(np.asarray(s, dtype=float))
        ns = s.size
        tab_mn = np.empty((ns, 6, self.xm.size))
        tab_mn[:, 0], tab_mn[:, 1], tab_mn[:, 2] = self.rmnc(s), self.zmns(s), self.lmns(s)
        tab_mn[:, 3] = self.rmnc(s, 1)
        tab_mn[:, 4] = self.zmns(s, 1)
        tab_mn[:, 5] = self.lmns(s, 1)
        tab_nyq = np.empty((ns, 7, self.xm_nyq.size))
        tab_nyq[:, 0], tab_nyq[:, 1] = self.gmnc(s), self.bmnc(s)
        tab_nyq[:, 2] = self.bmnc(s, 1)
        tab_nyq[:, 3], tab_nyq[:, 4] = self.bsupvmnc(s), self.bsubsmns(s)
        tab_nyq[:, 5], tab_nyq[:, 6] = self.bsubumnc(s), self.bsubvmnc(s)
        scal = np.zeros((ns, NSCAL))
        iota, diota = self.iota(s), self.iota(s, 1)
        scal[:, 0], scal[:, 1], scal[:, 2] = s, iota, diota
        scal[:, 3] = self.pressure(s, 1)
        scal[:, 4] = (-2 * s / iota) * diota            # shat, utils.py:316
        scal[:, 5] = self.pressure(s)
        return SurfaceTables(tab_mn, tab_nyq, scal, self.xm, self.xn, self.xm_nyq, self.xn_nyq,
                             self.phiedge, self.Aminor_p, self.nfp,
                             bsupumnc=None if self.bsupumnc is None else np.ascontiguousarray(self.bsupumnc(s)),
                             raxis_cc=self.raxis_cc)


_TABLES_2D = ["rmnc", "zmns", "lmns", "gmnc", "bmnc", "bsupumnc", "bsupvmnc",
              "bsubsmns", "bsubumnc", "bsubvmnc"]
_TABLES_1D = ["pres", "chi", "iotas", "phi", "xm", "xn", "xm_nyq", "xn_nyq", "raxis_cc"]
_SCALARS = ["Aminor_p", "mnmax", "mnmax_nyq", "nfp", "ns", "mpol", "ntor"]


def read_wout(path):
    """Read a VMEC ``wout_*.nc`` (NetCDF-3 classic) with ``scipy.io.netcdf_file``.

    Files store 2-D tables ``(radius, mn)``; they are transposed to the ``(mn, radius)``
    layout of ``simsopt``'s ``vmec.wout`` that the reference indexes (``utils.py:60``).
    """
    from scipy.io import netcdf_file

    w = types.SimpleNamespace()
    with netcdf_file(path, "r", mmap=False) as f:
        for k in _TABLES_2D:
            setattr(w, k, np.array(f.variables[k][:], dtype=float).T.copy())
        for k in _TABLES_1D:
            setattr(w, k, np.array(f.variables[k][:], dtype=float))
        for k in _SCALARS:
            setattr(w, k, f.variables[k][()].item())
    return w
