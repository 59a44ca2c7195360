"""B200-native ideal-ballooning stability engine (hot path only).

Geometry assembly -> tridiagonal pencil -> lambda_max + eigenfunction -> adjoint gradient,
as hand-written sm_100a CUDA kernels behind a C-ABI shared library
(``include/ibs_b200.h``), with a host-side mirror of the reference's entry points
(``vmec_fieldlines``, ``gamma_ball_full``, ``obj_w_grad`` and the ``ball_scan`` loops).

The directory is called ``ideal-ballooning-solver_b200`` (not importable as written);
``import ideal_ballooning_solver_b200`` resolves to it through the alias package at
the repo root.  There is no CPU fallback: every compute entry point raises if the CUDA
library is missing or no GPU is visible.
"""
__version__ = "0.1.0"

from . import synthetic, tables  # noqa: F401  (pure-numpy host helpers)
