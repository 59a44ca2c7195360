#!/usr/bin/env python
"""CPU: iteration passes per level of the lane-per-chain scan solver (tools/scan2_core_host.cpp) on a bench-shaped workload,
per solve and per warp (the kernel repeats a level's pass until ALL 16 solves of the warp are done: a warp pays the maximum).
Test tooling only (geometry from the oracle's numpy restatement of K1).

    python tools/eval_hist_host.py [d3d|ncsx] [nsurf]
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench  # noqa: E402
import scan_host_check as shc  # noqa: E402
from oracle import ballooning_oracle as bo  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "d3d"
    nsurf = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    st, alpha, theta0, theta = bench.build_tables(wl, 1, 0)
    pick = np.arange(nsurf) + (st.ns - nsurf) // 2                      # ADJACENT surfaces: the kernel warm-starts a line from the previous one
    st = st.select(pick)
    alpha = alpha[:: max(1, alpha.size // 4)][:4]
    fl = bo.fieldlines(st, alpha, theta)
    names = shc.BASE_NAMES
    base = np.stack([getattr(fl, n) for n in names], axis=2).reshape(-1, len(names), theta.size)       # (ns*na, 8, N)
    dP = np.array([bo.dpdrho_of(fl, js, ja) for js in range(st.ns) for ja in range(alpha.size)])
    nline = base.shape[0]
    t0 = np.tile(theta0, (nline, 1))
    lib = shc.build()
    lib.scan2_host_eval_counts.restype = ctypes.c_long
    lib.scan2_host_eval_counts.argtypes = [ctypes.POINTER(ctypes.c_int)]
    res = shc.host_scan_solve(lib, base, dP, t0, theta[1] - theta[0], want_X=False, want_dX=False, kernel="scan2")
    cnt = np.zeros(nline * theta0.size * 8, dtype=np.int32)
    lib.scan2_host_eval_counts(cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    cnt = cnt.reshape(nline, theta0.size, 8)
    nlev = lib.scan_host_num_levels(theta.size)
    print(f"{wl}: {nline} lines x {theta0.size} theta0, N = {theta.size}, levels 0..{nlev}; info iters mean {np.mean(res['info'] & 0xffff):.3f}")
    groups = -(-theta0.size // 16)

    def warp_max(c):          # warp (line, grp) holds 16 consecutive theta0
        pad = np.concatenate([c, np.repeat(c[:, -1:], groups * 16 - theta0.size, axis=1)], axis=1)
        return pad.reshape(nline, groups, 16).max(axis=2)

    tot_s = tot_w = 0.0
    for lev in range(nlev, -1, -1):
        c = cnt[:, :, lev]
        g = warp_max(c)
        print(f"  level {lev} (1/{1 << lev}): per solve mean {c.mean():.3f} hist {np.bincount(c.ravel())[:10]}   per warp (max of 16) mean {g.mean():.3f} hist {np.bincount(g.ravel())[:10]}")
        tot_s += c.mean() / (1 << lev); tot_w += g.mean() / (1 << lev)
    o1 = cnt[:, :, 7]
    g = warp_max(o1)
    print(f"  first output passes on the fine level: per solve mean {o1.mean():.3f} hist {np.bincount(o1.ravel())[:6]}, per warp mean {g.mean():.3f}")
    print(f"  fine-equivalent iteration passes: per solve {tot_s:.3f}, per warp {tot_w:.3f}")


if __name__ == "__main__":
    main()
