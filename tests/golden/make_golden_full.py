#!/usr/bin/env python
"""Golden fixtures of the FULL Struct of the reference's ``vmec_fieldlines`` (theta1d and phi1d forms) and of
``vmec_fieldlines_axisym``, generated from the UNMODIFIED reference.

    python tests/golden/make_golden_full.py        # needs /root/reference; writes full_struct.npz

Every array / scalar attribute of the returned Struct is stored under ``<case>__<name>``; the inputs (per-surface tables
evaluated with the reference's own FITPACK splines, mode numbers, grids) under ``<case>__in_*``."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from oracle import ref_shim  # noqa: E402
from ideal_ballooning_solver_b200 import synthetic  # noqa: E402
from ideal_ballooning_solver_b200.tables import TAB_MN_ROWS, TAB_NYQ_ROWS  # noqa: E402

u = ref_shim.load_reference_utils()


def ref_tables(vs, s):
    s = np.atleast_1d(np.asarray(s, float))
    tab_mn = np.stack([np.stack([spl(s) for spl in getattr(vs, name)], axis=1) for name in TAB_MN_ROWS], axis=1)
    tab_nyq = np.stack([np.stack([spl(s) for spl in getattr(vs, name)], axis=1) for name in TAB_NYQ_ROWS], axis=1)
    bsupu = np.stack([spl(s) for spl in vs.bsupumnc], axis=1)
    iota, diota = vs.iota(s), vs.d_iota_d_s(s)
    scal = np.zeros((s.size, 8))
    scal[:, 0], scal[:, 1], scal[:, 2] = s, iota, diota
    scal[:, 3] = vs.d_pressure_d_s(s)
    scal[:, 4] = (-2 * s / iota) * diota
    scal[:, 5] = vs.pressure(s)
    return tab_mn, tab_nyq, bsupu, scal


def dump(out, case, vs, s, alpha, res, **grids):
    tab_mn, tab_nyq, bsupu, scal = ref_tables(vs, s)
    out.update({f"{case}__in_tab_mn": tab_mn, f"{case}__in_tab_nyq": tab_nyq, f"{case}__in_bsupumnc": bsupu, f"{case}__in_scal": scal,
                f"{case}__in_xm": vs.xm, f"{case}__in_xn": vs.xn, f"{case}__in_xm_nyq": vs.xm_nyq, f"{case}__in_xn_nyq": vs.xn_nyq,
                f"{case}__in_phiedge": vs.phiedge, f"{case}__in_Aminor_p": vs.Aminor_p, f"{case}__in_nfp": vs.nfp,
                f"{case}__in_raxis_cc": np.asarray(vs.raxis_cc, float), f"{case}__in_s": np.atleast_1d(s), f"{case}__in_alpha": np.atleast_1d(alpha)})
    for k, v in grids.items():
        out[f"{case}__in_{k}"] = np.asarray(v, float)
    for name, val in vars(res).items():
        if val is None:
            continue
        a = np.asarray(val)
        if a.dtype.kind in "fiub":
            out[f"{case}__{name}"] = a.astype(float) if a.dtype.kind != "f" else a


def main():
    out = {}
    # ---- 3-D equilibrium: vmec_fieldlines, theta1d and phi1d forms
    vs = u.vmec_splines(ref_shim.FakeVmec(synthetic.make_equilibrium("ncsx", seed=3)))
    s, alpha = [0.55, 0.8], [0.0, 1.1]
    theta = np.linspace(-np.pi, np.pi, 33)
    dump(out, "fl_theta", vs, s, alpha, u.vmec_fieldlines(vs, s, alpha, theta1d=theta, phi_center=0.2), theta1d=theta, phi_center=0.2)
    phi = np.linspace(-0.9, 1.3, 29)
    dump(out, "fl_phi", vs, s, alpha, u.vmec_fieldlines(vs, s, alpha, phi1d=phi), phi1d=phi, phi_center=0.0)
    # ---- axisymmetric equilibrium: vmec_fieldlines_axisym (one surface: the routine's `_1` arrays only broadcast for ns = 1)
    vsa = u.vmec_splines(ref_shim.FakeVmec(synthetic.make_equilibrium("d3d", seed=2)))
    th2 = np.linspace(-np.pi, np.pi, 65)
    dump(out, "axisym", vsa, [0.7], [0.0], u.vmec_fieldlines_axisym(vsa, [0.7], [0.0], theta1d=th2), theta1d=th2, phi_center=0.0)
    np.savez_compressed(os.path.join(HERE, "full_struct.npz"), **out)
    print("wrote full_struct.npz:", len(out), "entries")


if __name__ == "__main__":
    main()
