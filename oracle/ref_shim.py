"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference.

This module makes ``/root/reference/utils.py`` importable in THIS container so
that (a) the numpy restatement in ``oracle/ballooning_oracle.py`` can be
validated against the real thing and (b) golden vectors can be generated
(``tests/golden/make_golden.py``).  ``/root/reference`` does not exist on the
GPU box; there the only copy of the reference is ``oracle/_ref/utils.py`` (made
byte for byte by ``oracle/make_ref.py``, git-ignored), which ``bench.py``'s CPU
arm times as the real reference.  The ``-m gpu`` tests and ``smoke()`` use the
committed fixtures and the restatement, not this file.

Nothing under ``oracle/`` is product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference arm may
import it, and only as the checker.

Shims applied:
  1. ``scipy.integrate.simps`` was removed from scipy; the reference imports it
     at ``utils.py:14``.  We alias it to ``scipy.integrate.simpson`` (identical
     for an odd number of samples, which is all the reference ever uses).
  2. ``simsopt`` is not installed; the reference does
     ``from simsopt.mhd.vmec import Vmec`` (``utils.py:18``) and only touches
     ``vmec.run()``, ``vmec.wout.<tables>``, ``vmec.s_full_grid`` and
     ``vmec.s_half_grid`` (``utils.py:46-135``).  ``FakeVmec`` provides exactly
     those from a wout-like object whose 2-D tables are laid out ``(mn, ns)``
     as ``utils.py:60`` indexes them.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference_root():
    """``$IBS_REFERENCE_ROOT``, else ``/root/reference`` (build container), else ``oracle/_ref`` (the byte-for-byte copy of
    ``utils.py`` made by ``oracle/make_ref.py``; the only form in which the reference exists on the GPU box)."""
    env = os.environ.get("IBS_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isfile(os.path.join(cand, "utils.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_reference_root()

_TABLES_2D = ["rmnc", "zmns", "lmns", "gmnc", "bmnc", "bsupumnc", "bsupvmnc",
              "bsubsmns", "bsubumnc", "bsubvmnc"]
_TABLES_1D = ["pres", "chi", "iotas", "phi", "xm", "xn", "xm_nyq", "xn_nyq",
              "raxis_cc"]
_SCALARS = ["Aminor_p", "mnmax", "mnmax_nyq", "nfp", "ns", "mpol", "ntor"]


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils.py"))


class FakeVmec:
    """Stand-in for ``simsopt.mhd.vmec.Vmec`` (see module docstring)."""

    def __init__(self, wout):
        self.wout = wout
        ns = int(wout.ns)
        self.s_full_grid = np.linspace(0, 1, ns)
        ds = self.s_full_grid[1] - self.s_full_grid[0]
        self.s_half_grid = self.s_full_grid[1:] - 0.5 * ds

    def run(self):  # the reference calls vmec.run() at utils.py:46
        pass


def wout_from_netcdf(path):
    """Read a VMEC ``wout_*.nc`` (NetCDF-3) into a wout-like namespace.

    The file stores 2-D tables as ``(radius, mn)``; the reference indexes
    ``vmec.wout.rmnc[jmn, :]`` (``utils.py:60``), so they are transposed here.
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


_utils_mod = None


def load_reference_utils():
    """Import the unmodified ``/root/reference/utils.py`` and return the module."""
    global _utils_mod
    if _utils_mod is not None:
        return _utils_mod
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    import scipy.integrate as si

    if not hasattr(si, "simps"):
        si.simps = si.simpson
    for name in ("simsopt", "simsopt.mhd", "simsopt.mhd.vmec"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["simsopt.mhd.vmec"].Vmec = FakeVmec
    import importlib.util

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        spec = importlib.util.spec_from_file_location(
            "_ibs_reference_utils", os.path.join(REFERENCE_ROOT, "utils.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    _utils_mod = mod
    return mod


@contextlib.contextmanager
def converged_arpack(utils_mod=None, tol=0.0, maxiter=None):
    """Run the reference with ARPACK ``tol`` overridden (``utils.py:1597`` hard
    codes ``tol=5e-7``, which is 3-5 orders looser than the parity target)."""
    u = utils_mod or load_reference_utils()
    orig = u.eigs

    def eigs_tight(*a, **kw):
        kw["tol"] = tol
        if maxiter is not None:
            kw["maxiter"] = maxiter
        return orig(*a, **kw)

    u.eigs = eigs_tight
    try:
        yield u
    finally:
        u.eigs = orig


def load_s_alpha_checkers():
    """``check_ball`` / ``check_ball_long`` from the reference's s-alpha test,
    extracted with ``ast`` because the module's top level imports matplotlib and
    starts a 40-process pool (``bishop_ball_s-alpha.py:17,213-275``)."""
    import ast

    path = os.path.join(REFERENCE_ROOT, "tests", "shifted-circle-s-alpha",
                        "bishop_ball_s-alpha.py")
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef)
            and n.name in ("check_ball", "check_ball_long")]
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns["check_ball"], ns["check_ball_long"]
