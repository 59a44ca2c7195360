"""Import alias: ``ideal_ballooning_solver_b200`` -> ``ideal-ballooning-solver_b200/``.

The product directory carries the reference's (hyphenated) repository name, which is not a
valid Python identifier; this two-line package points its ``__path__`` there and runs the real
``__init__``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "ideal-ballooning-solver_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
