#!/usr/bin/env python
"""GPU: time each stage of one scan step back to back (host running ahead), to separate kernel time from launch gaps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ideal_ballooning_solver_b200 import engine, scan
wl = sys.argv[1] if len(sys.argv) > 1 else "d3d"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 16
st, alpha, theta0, theta = bench.build_tables(wl, E, 0)
kind, ns1, na, nt, nth, span = bench.WORKLOADS[wl]
dt = engine.DeviceTables.from_host(st)
a_d, th_d = torch.from_numpy(alpha).cuda(), torch.from_numpy(theta).cuda()
nl = st.ns * na
t0_d = torch.from_numpy(theta0).cuda().repeat(nl)
h = engine.grid_spacing(theta)
ev = lambda: torch.cuda.Event(enable_timing=True)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = ev(), ev(); t = time.perf_counter(); a.record()
    for _ in range(reps): fn()
    b.record(); th = time.perf_counter() - t; torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, 1e3 * th / reps
geo = engine.geometry_batch(dt, a_d, th_d)
print("geometry_batch      gpu %.3f ms  host-issue %.3f ms" % timeit(lambda: engine.geometry_batch(dt, a_d, th_d)))
ch = scan.chain_length(nt)
for wantX in (True, False):
    print("solve (X=%s, chain %d) gpu %.3f ms  host-issue %.3f ms" % ((wantX, ch) + timeit(lambda: engine.solve_base_batch(geo.base, geo.dPdrho, t0_d, h, nth0=nt, want_X=wantX, want_dX=False, want_matrix=False, chain_len=ch))))
sol = engine.solve_base_batch(geo.base, geo.dPdrho, t0_d, h, nth0=nt, want_X=False, want_dX=False, want_matrix=False, chain_len=ch)
print("argmax              gpu %.3f ms  host-issue %.3f ms" % timeit(lambda: engine.scan_argmax(sol.lam.reshape(st.ns, na * nt))))
print("solves", st.ns * na * nt, "lines", nl, "N", len(theta), "mean its", float((sol.info & 0xffff).double().mean()))
