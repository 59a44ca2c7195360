#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- recipe that makes the UNMODIFIED reference travel to the GPU box.

The reference is plain Python with no ``setup.py`` / ``pyproject.toml`` (flat scripts), so there is nothing to
``pip install`` or compile: its hot path is the single file ``/root/reference/utils.py``.  This recipe copies that file,
byte for byte, into the git-ignored ``oracle/_ref/`` (never into history; ``oracle/_ref/`` is NOT in ``.gpurunignore``, so
it is sent to the GPU box like the built ``.so``) together with a manifest holding its sha256, so that

  * ``bench.py --impl reference`` / ``cpu_baseline`` can time the reference's OWN ``vmec_fieldlines`` +
    ``gamma_ball_full`` (``cpu_baseline.kind = "reference"``) instead of the oracle port, and
  * ``tests/test_oracle.py`` can check the restatement against the real thing on the GPU box too.

``__graft_entry__.build()`` runs it whenever ``/root/reference`` is present.  The two import shims the file needs
(``scipy.integrate.simps`` alias, stand-in for ``simsopt.mhd.vmec.Vmec``) live in ``oracle/ref_shim.py``.

    python oracle/make_ref.py            # -> oracle/_ref/utils.py + oracle/_ref/MANIFEST.json
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("IBS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["utils.py"]          # the whole hot path of the reference


def make(verbose=True) -> bool:
    if not os.path.isfile(os.path.join(SRC_ROOT, FILES[0])):
        if verbose:
            print("reference not present at", SRC_ROOT, "- keeping whatever oracle/_ref already holds")
        return os.path.isfile(os.path.join(DST, FILES[0]))
    os.makedirs(DST, exist_ok=True)
    manifest = {"source_root": SRC_ROOT, "files": {}}
    for name in FILES:
        shutil.copyfile(os.path.join(SRC_ROOT, name), os.path.join(DST, name))
        manifest["files"][name] = hashlib.sha256(open(os.path.join(DST, name), "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    if verbose:
        print("oracle/_ref:", manifest["files"])
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
