"""CPU check of the per-problem logic of the batched refine (csrc/ibs_refine_core.cuh, the code the CUDA step kernel runs)
through the g++ harness tools/refine_core_host.cpp, against scipy's L-BFGS-B with the reference's options
(ball_scan.py:305-314) on smooth 2-D test functions with the reference's bounds."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
BOUNDS = ((0.0, np.pi), (0.0, 0.5 * np.pi))
OPT = dict(ftol=5.0e-11, gtol=2.0e-08, maxiter=30)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("rch") / "refine_core_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tools", "refine_core_host.cpp")])
    L = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    L.refine_host_init.argtypes = [dp] + [ctypes.c_double] * 6
    L.refine_host_consume.argtypes = [dp, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_double,
                                      ctypes.c_double, ctypes.c_int]
    return L


def run(lib, fun, x0, fail_at=None, max_rounds=200):
    n = lib.refine_host_nstate()
    st = np.zeros(n)
    p = st.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    lib.refine_host_init(p, x0[0], x0[1], BOUNDS[0][0], BOUNDS[0][1], BOUNDS[1][0], BOUNDS[1][1])
    rounds = 0
    while int(st[18]) != 2 and rounds < max_rounds:
        xt = st[11:13].copy()
        assert BOUNDS[0][0] <= xt[0] <= BOUNDS[0][1] and BOUNDS[1][0] <= xt[1] <= BOUNDS[1][1]      # trial points stay in the box
        f, g = fun(xt)
        failed = fail_at is not None and rounds == fail_at
        lib.refine_host_consume(p, f, g[0], g[1], int(failed), OPT["ftol"], OPT["gtol"], OPT["maxiter"])
        rounds += 1
    return st, rounds


def growth_like(c, w, amp=1e-3):
    """F = -lambda with a smooth bump of height `amp` centred at c (the scale of the real objective: lambda ~ 1e-3)."""
    def fun(x):
        d = (np.asarray(x) - c) / w
        e = np.exp(-0.5 * np.dot(d, d))
        return -amp * e, amp * e * d / w
    return fun


@pytest.mark.parametrize("c,x0", [((1.0, 0.6), (0.8, 0.5)), ((2.9, 0.2), (2.6, 0.4)), ((3.5, -0.3), (2.8, 0.3)),     # interior; near a bound; optimum outside the box
                                   ((0.4, 1.9), (0.5, 1.2)), ((1.5, 0.8), (1.5, 0.8))])                              # outside in theta0; start at the optimum
def test_matches_scipy_lbfgsb(lib, c, x0):
    from scipy.optimize import minimize
    fun = growth_like(np.array(c), np.array([0.7, 0.5]))
    ref = minimize(fun, x0=x0, jac=True, bounds=BOUNDS, options=OPT)
    st, rounds = run(lib, fun, x0)
    assert int(st[18]) == 2 and int(st[19]) in (1, 2, 3, 4)
    # the optimum reached is at least as good as scipy's (to the reference's own ftol) and lies at the same place
    assert st[2] <= ref.fun + 5e-11
    assert np.max(np.abs(st[0:2] - ref.x)) < 2e-3
    assert rounds <= 3 * max(ref.nfev, 4) + 4          # comparable number of objective evaluations
    # monotone: the accepted value never exceeds the starting value
    assert st[2] <= fun(np.clip(x0, [b[0] for b in BOUNDS], [b[1] for b in BOUNDS]))[0] + 1e-18


def test_rosenbrock_like_valley(lib):
    """A curved valley inside the box (exercises the BFGS update and the backtracking)."""
    from scipy.optimize import minimize
    def fun(x):
        a, b = x[0] - 1.0, x[1] - 0.5
        f = 1e-3 * ((1 - a) ** 2 + 20 * (b - a * a) ** 2) * 0.1
        g = 1e-4 * np.array([-2 * (1 - a) - 80 * a * (b - a * a), 40 * (b - a * a)])
        return f, g
    x0 = (0.3, 1.2)
    ref = minimize(fun, x0=x0, jac=True, bounds=BOUNDS, options=dict(OPT, maxiter=200))
    st, rounds = run(lib, lambda x: fun(x), x0, max_rounds=600)
    assert int(st[18]) == 2
    # both stop on the reference's ftol/maxiter well inside the valley; compare the objective reached under the same options
    ref30 = minimize(fun, x0=x0, jac=True, bounds=BOUNDS, options=OPT)
    assert st[2] <= ref30.fun * 1.5 + 1e-9


def test_failed_evaluation_freezes_the_problem(lib):
    fun = growth_like(np.array([1.0, 0.6]), np.array([0.7, 0.5]))
    st, rounds = run(lib, fun, (0.8, 0.5), fail_at=0)
    assert int(st[18]) == 2 and int(st[19]) == 5 and rounds == 1
    # a failure inside a line search is treated as a rejected step, not as the end
    st, rounds = run(lib, fun, (0.8, 0.5), fail_at=2)
    assert int(st[18]) == 2 and int(st[19]) in (1, 2, 3, 4)
