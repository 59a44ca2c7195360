"""GPU: the whole Struct of ``vmec_fieldlines`` (theta1d and phi1d forms) and of ``vmec_fieldlines_axisym`` through the facade
(``reference_api``) and the full-output geometry kernel, field by field against fixtures generated from the unmodified
reference (tests/golden/make_golden_full.py).  SURVEY.md section 8 row f4."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tables(D, case):
    from ideal_ballooning_solver_b200.tables import SurfaceTables
    g = lambda k: np.array(D[f"{case}__in_{k}"])
    return SurfaceTables(g("tab_mn"), g("tab_nyq"), g("scal"), g("xm"), g("xn"), g("xm_nyq"), g("xn_nyq"), float(g("phiedge")),
                         float(g("Aminor_p")), int(g("nfp")), bsupumnc=g("bsupumnc"), raxis_cc=g("raxis_cc"))


def _compare(D, case, res, skip=()):
    names = [k[len(case) + 2:] for k in D.files if k.startswith(case + "__") and not k.startswith(case + "__in_")]
    assert len(names) >= 90
    checked = 0
    for name in names:
        if name in skip:
            continue
        want = np.asarray(D[f"{case}__{name}"])
        assert hasattr(res, name), f"Struct field {name} missing"
        got = np.asarray(getattr(res, name), dtype=float)
        assert got.shape == want.shape, (name, got.shape, want.shape)
        scale = max(float(np.max(np.abs(want))), 1e-300)
        if name[-2:] in ("_X", "_Y", "_Z"):
            # a Cartesian component that vanishes analytically (e.g. grad_phi_Z) is rounding noise in the reference: the
            # scale of a vector field is the largest of its three components
            scale = max(float(np.max(np.abs(D[f"{case}__{name[:-1]}{ax}"]))) for ax in "XYZ")
        # the two "_alternate" triple products cancel to ~1e-16 of their terms; everything else is a plain sum / product
        tol = 2e-9 if name.endswith("_alternate") else 5e-10
        err = float(np.max(np.abs(got - want))) / scale
        assert err < tol, (case, name, err)
        checked += 1
    return checked


@pytest.mark.parametrize("case,grid", [("fl_theta", "theta1d"), ("fl_phi", "phi1d")])
def test_vmec_fieldlines_full_struct(cuda_lib, golden, case, grid):
    from ideal_ballooning_solver_b200 import reference_api as api
    D = golden("full_struct")
    st = _tables(D, case)
    kw = {grid: np.array(D[f"{case}__in_{grid}"])}
    res = api.vmec_fieldlines(st, np.array(D[f"{case}__in_s"]), np.array(D[f"{case}__in_alpha"]),
                              phi_center=float(D[f"{case}__in_phi_center"]), **kw)
    n = _compare(D, case, res, skip=("phi1d",) if grid == "theta1d" else ("theta1d",))
    assert n >= 90
    # the hot-path kernel (K1) gives the same eight arrays
    if grid == "theta1d":
        import dataclasses
        hot = api.vmec_fieldlines(dataclasses.replace(st, bsupumnc=None), np.array(D[f"{case}__in_s"]), np.array(D[f"{case}__in_alpha"]),
                                  theta1d=kw["theta1d"], phi_center=float(D[f"{case}__in_phi_center"]))
        for name in ("bmag", "gradpar_theta_pest", "cvdrift", "cvdrift0", "gds2", "gds21", "gds22", "gbdrift", "theta_vmec"):
            a, b = getattr(hot, name), getattr(res, name)
            assert np.max(np.abs(a - b)) <= 2e-11 * np.max(np.abs(b)), name


def test_vmec_fieldlines_axisym_struct(cuda_lib, golden):
    from ideal_ballooning_solver_b200 import reference_api as api
    D = golden("full_struct")
    st = _tables(D, "axisym")
    res = api.vmec_fieldlines_axisym(st, np.array(D["axisym__in_s"]), np.array(D["axisym__in_alpha"]), theta1d=np.array(D["axisym__in_theta1d"]))
    assert _compare(D, "axisym", res) >= 105
    with pytest.raises(ValueError):
        api.vmec_fieldlines_axisym(st, [0.7], [0.0])
    with pytest.raises(ValueError):
        api.vmec_fieldlines(st, [0.7], [0.0], theta1d=[0.0], phi1d=[0.0])
