"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports exactly the symbols that
include/ibs_b200.h declares (no compute calls -- there is no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ibs_b200.h")


def _declared():
    """{name: (return type, number of parameters)} parsed from the header."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|const char\s*\*)\s+(ibs_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(2)] = (m.group(1), n)
    return out


@pytest.fixture(scope="module")
def lib():
    from ideal_ballooning_solver_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    from ideal_ballooning_solver_b200 import _lib
    decl = _declared()
    assert len(decl) >= 12
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    for name, (ret, nargs) in decl.items():
        assert len(_lib.SIGNATURES[name][1]) == nargs, name
        getattr(lib, name)          # exported


def test_exported_symbols_are_plain_c(lib):
    from ideal_ballooning_solver_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(_declared()) <= exported
    # no torch / python dependency in the product library
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "python" not in ldd


def test_built_for_sm_100a(lib):
    from ideal_ballooning_solver_b200 import _lib
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_version_and_argument_errors_without_gpu(lib):
    assert lib.ibs_version() >= 100
    # argument validation happens before any CUDA call and reports through ibs_last_error()
    rc = lib.ibs_solve_gcf_batch(None, None, None, 4, 2, 0.1, None, None, 1, None, None, None, None, None, None)
    assert rc == 1 and b"N >= 5" in lib.ibs_last_error()
    rc = lib.ibs_solve_gcf_batch(None, None, None, 4, 65, 0.1, None, None, 1, None, None, None, None, None, None)
    assert rc == 1 and b"null" in lib.ibs_last_error()
    assert lib.ibs_solve_gcf_batch(None, None, None, 0, 65, 0.1, None, None, 1, None, None, None, None, None, None) == 0


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point of the host layer must fail loudly."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ideal_ballooning_solver_b200 import _lib, engine
    z = np.ones((2, 33))
    with pytest.raises(_lib.IbsError):
        engine.solve_gcf_batch(z, z, z, 0.1)
    with pytest.raises(_lib.IbsError):
        engine.scan_argmax(torch.zeros(2, 3, dtype=torch.float64))


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, "ideal-ballooning-solver_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn
                assert "/root/reference" not in txt or fn.endswith((".py", ".cu", ".cuh")), fn
    # bench/smoke/tests may; the package must not read /root/reference at run time either
    for fn in ("engine.py", "scan.py", "_lib.py", "reference_api.py"):
        p = os.path.join(pkg, fn)
        if os.path.isfile(p):
            code = "\n".join(ln for ln in open(p).read().splitlines() if "open(" in ln or "import" in ln)
            assert "/root/reference" not in code


def test_layout_constants_match_the_python_mirrors():
    """Pure host queries of the library (no GPU needed): the field list of the full-output geometry and the size of the refine state."""
    from ideal_ballooning_solver_b200 import _lib, reference_api
    lib = _lib.load(build_if_missing=True)
    assert lib.ibs_geometry_full_nfields() == len(reference_api.FULL_FIELDS)
    assert len(set(reference_api.FULL_FIELDS)) == len(reference_api.FULL_FIELDS)
    assert lib.ibs_refine_state_doubles() == 30
