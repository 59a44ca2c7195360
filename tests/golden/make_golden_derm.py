"""Golden vectors of the reference's finite-difference helpers ``derm`` / ``dermv`` (``utils.py:1737-1943``), produced by
the UNMODIFIED reference imported through ``oracle/ref_shim.py``.  Run in the build container (``/root/reference``):

    python tests/golden/make_golden_derm.py        ->  tests/golden/derm.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

u = ref_shim.load_reference_utils()
rng = np.random.default_rng(20261018)
n1, n2 = 9, 17
a1 = rng.standard_normal(n2)
b1 = np.cumsum(rng.uniform(0.5, 1.5, n2))
a2 = rng.standard_normal((n1, n2))
b2l = np.cumsum(rng.uniform(0.5, 1.5, (n1, n2)), axis=1)
b2r = np.cumsum(rng.uniform(0.5, 1.5, (n1, n2)), axis=0)
out = dict(a1=a1, b1=b1, a2=a2, b2l=b2l, b2r=b2r)
for ch in "lr":
    for par in "eo":
        out[f"derm_1d_{ch}_{par}"] = u.derm(a1.copy(), ch, par)
        out[f"derm_2d_{ch}_{par}"] = u.derm(a2.copy(), ch, par)
        out[f"dermv_2d_{ch}_{par}"] = u.dermv(a2.copy(), (b2l if ch == "l" else b2r).copy(), ch, par)
    # (dermv on a 1-D array with ch='r' stops in pdb.set_trace() in the reference: not generated)
for par in "eo":
    out[f"dermv_1d_l_{par}"] = u.dermv(a1.copy(), b1.copy(), "l", par)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "derm.npz"), **out)
print("wrote derm.npz:", sorted(out))
