"""Growth-rate scan over (surface s, field-line label alpha, ballooning angle theta0).

Batched re-expression of the loops of ``/root/reference/ball_scan.py:196-339``: the reference runs
one MPI group per surface and one rank per alpha stripe (``ball_scan.py:172-194,248-252``) and
gathers three doubles per surface at the end (``:341-347``).  Here every (s, alpha) field line and
every (s, alpha, theta0) solve is one element of a batched launch; surfaces are sharded in
contiguous blocks over ranks (one process per GPU) and the only exchange is one all-gather of the
per-surface maxima.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np
import torch

from . import engine
from .tables import SurfaceTables

# reference defaults (ball_scan.py:197,201,223-226; utils.py:1639; ball_scan.py:313)
THETA_FAC = 4
NTHETA0_GUESS = 15
NALPHA_GUESS = 24
DEL_ALPHA = 0.004


def scan_theta_grid(mpol: int, ntor: int, theta_fac: int = THETA_FAC) -> np.ndarray:
    """The theta grid of ``ball_scan.py:201-208`` (``ntheta`` odd)."""
    ntheta = int(2 * mpol * theta_fac) + 1 if ntor == 0 else int(2 * mpol * ntor * theta_fac) + 1
    return np.linspace(-theta_fac * np.pi, theta_fac * np.pi, ntheta)


def chain_length(nth0: int, cap: int = 16) -> int:
    """Warm-start run length over consecutive theta0 of one field line (``ball_scan.py:265-274`` chains its
    start vector through the same loop): the largest divisor of ``nth0`` not above ``cap``."""
    k = min(int(nth0), cap)
    while k > 1 and nth0 % k:
        k -= 1
    return max(k, 1)


def shard_range(n: int, rank: int, world: int):
    """Contiguous block of ``range(n)`` owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@dataclasses.dataclass
class ScanResult:
    gamma: torch.Tensor            # (ns_local, nalpha, nth0)
    val: torch.Tensor              # (ns_local,)   max over the grid (ball_scan.py:279-295)
    idx: torch.Tensor              # (ns_local,)   flat (alpha, theta0) index of the maximum, -1 = all-zero guard
    sigma0: torch.Tensor           # (ns_local,)   1.3 |max| + 0.05
    alpha_guess: torch.Tensor      # (ns_local,)
    theta0_guess: torch.Tensor     # (ns_local,)
    X: Optional[torch.Tensor]      # (ns_local, nalpha, nth0, nl) eigenfunctions (if requested)
    info: torch.Tensor
    geometry: engine.Geometry


def coarse_scan(tables: engine.DeviceTables, alpha, theta0, theta, want_X: bool = False,
                lam0=None) -> ScanResult:
    """Coarse (alpha, theta0) grid of every surface in ``tables`` (``ball_scan.py:248-295``): K1 for
    ``ns x nalpha`` field lines, K2+K3 for ``ns x nalpha x nth0`` solves, guarded arg-max."""
    dev = tables.tab_mn.device
    alpha_t = torch.as_tensor(np.asarray(alpha, dtype=np.float64)).to(dev) if not isinstance(alpha, torch.Tensor) else alpha
    theta0_t = torch.as_tensor(np.asarray(theta0, dtype=np.float64)).to(dev) if not isinstance(theta0, torch.Tensor) else theta0
    theta_np = theta.cpu().numpy() if isinstance(theta, torch.Tensor) else np.asarray(theta, dtype=np.float64)
    geo = engine.geometry_batch(tables, alpha_t, theta_np)
    ns, na, nt = tables.ns, alpha_t.shape[-1], theta0_t.numel()
    th0 = theta0_t.repeat(ns * na)
    sol = engine.solve_base_batch(geo.base, geo.dPdrho, th0, engine.grid_spacing(theta_np), nth0=nt, lam0=lam0,
                                  want_X=want_X, want_dX=False, want_matrix=False, chain_len=chain_length(nt))
    gamma = sol.lam.reshape(ns, na, nt)
    val, idx, sig = engine.scan_argmax(gamma)
    safe = idx.clamp(min=0).long()
    ia, it = safe // nt, safe % nt
    if alpha_t.dim() == 2:
        a_guess = alpha_t.gather(1, ia[:, None])[:, 0]
    else:
        a_guess = alpha_t[ia]
    t_guess = theta0_t[it]
    zero = idx < 0          # ball_scan.py:279-282: alpha_guess = theta0_guess = 0
    a_guess = torch.where(zero, torch.zeros_like(a_guess), a_guess)
    t_guess = torch.where(zero, torch.zeros_like(t_guess), t_guess)
    X = sol.X.reshape(ns, na, nt, -1) if want_X else None
    return ScanResult(gamma, val, idx, sig, a_guess, t_guess, X, sol.info, geo)


def gather_surface_maxima(val: torch.Tensor, idx: torch.Tensor, ns_total: int, group=None):
    """The one exchange step of the scan (replaces the three ``MPI.Gather`` of ``ball_scan.py:345-347``):
    all-gather the per-surface ``(val, idx)`` pairs of every rank's contiguous block of surfaces.  Works
    with NCCL (GPU tensors) and gloo (CPU tensors)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return val, idx
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(ns_total, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    packed = torch.zeros((nmax, 2), dtype=torch.float64, device=val.device)
    lo, hi = sizes[rank]
    packed[: hi - lo, 0] = val
    packed[: hi - lo, 1] = idx.to(torch.float64)
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed, group=group)
    vals = torch.cat([o[: h - l, 0] for o, (l, h) in zip(out, sizes)])
    idxs = torch.cat([o[: h - l, 1] for o, (l, h) in zip(out, sizes)]).to(torch.int32)
    return vals, idxs


def sharded_coarse_scan(st: SurfaceTables, alpha, theta0, theta, device=None, want_X=False, group=None):
    """Shard the surfaces of ``st`` over the ranks of ``group`` (or run everything when not
    distributed), scan the local block and all-gather the per-surface maxima.  Returns
    ``(local ScanResult, (lo, hi), val_all, idx_all)``."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_range(st.ns, rank, world)
    device = device or torch.device("cuda", torch.cuda.current_device())
    dt = engine.DeviceTables.from_host(st.select(np.arange(lo, hi)), device)
    res = coarse_scan(dt, alpha, theta0, theta, want_X=want_X)
    val_all, idx_all = gather_surface_maxima(res.val, res.idx, st.ns, group)
    return res, (lo, hi), val_all, idx_all
