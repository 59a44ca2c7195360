#!/usr/bin/env python
"""GPU: histogram of solver evaluations per solve for a bench workload (cold = first of each chain run)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ideal_ballooning_solver_b200 import engine, scan
wl = sys.argv[1] if len(sys.argv) > 1 else "d3d"
st, alpha, theta0, theta = bench.build_tables(wl, 2, 0)
kind, ns1, na, nt, nth, span = bench.WORKLOADS[wl]
dt = engine.DeviceTables.from_host(st)
geo = engine.geometry_batch(dt, torch.from_numpy(alpha).cuda(), torch.from_numpy(theta).cuda())
t0 = torch.from_numpy(theta0).cuda().repeat(st.ns * na)
ch = scan.chain_length(nt if nt > 1 else na)
sol = engine.solve_base_batch(geo.base, geo.dPdrho, t0, engine.grid_spacing(theta), nth0=nt, chain_len=ch, want_dX=False)
it = (sol.info & 0xffff).cpu().numpy()
pos = np.arange(it.size) % ch
print("chain", ch, "mean", it.mean(), "cold mean", it[pos == 0].mean(), "2nd", it[pos == 1].mean(), "3rd", it[pos == 2].mean(), "warm(>=3) mean", it[pos >= 3].mean())
print("hist warm(>=3):", np.bincount(it[pos >= 3])[:12])
print("hist cold:", np.bincount(it[pos == 0])[:24])
ref = engine.solve_base_batch(geo.base, geo.dPdrho, t0, engine.grid_spacing(theta), nth0=nt, chain_len=1, want_dX=False)
print("max rel dlam vs unchained", float(((sol.lam - ref.lam).abs() / ref.lam.abs()).max()), "max dX", float((sol.X - ref.X).abs().max()))
