#!/usr/bin/env python
"""BASELINE.json configs[4] (per-GPU share): lambda_max + eigenvector + (dlam/dalpha, dlam/dtheta0) for NCSX-like field
lines with (s, alpha, theta0) i.i.d. uniform in [0.5,0.95] x [0,pi] x [0,pi/2], ntheta = 1024.
    python tools/bench_adjoint.py [npoints_per_gpu=16384] [steps=5]
Stages per step: K1 (three field lines per point: alpha -+ del/2) -> K3 (centre line) -> K4 (obj_w_grad contraction)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ideal_ballooning_solver_b200 import engine, synthetic, tables
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rng = np.random.default_rng(20261018 + 5)
s = rng.uniform(0.5, 0.95, npts); al = rng.uniform(0, np.pi, npts); t0 = rng.uniform(0, 0.5 * np.pi, npts)
theta = np.linspace(-4 * np.pi, 4 * np.pi, 1025)
st = tables.RadialSplines(synthetic.make_equilibrium("ncsx", seed=1)).evaluate(s)          # one table set per point
dt = engine.DeviceTables.from_host(st)
d = 0.004
alphas = torch.from_numpy(np.stack([al - 0.5 * d, al, al + 0.5 * d], axis=1)).cuda()
th_d, t0_d = torch.from_numpy(theta).cuda(), torch.from_numpy(t0).cuda()
h = engine.grid_spacing(theta)
ev = lambda: torch.cuda.Event(enable_timing=True)
def step(tm=None):
    if tm: tm[0].record()
    geo = engine.geometry_batch(dt, alphas, th_d)
    if tm: tm[1].record()
    out = engine.obj_w_grad_batch(geo.base, geo.dPdrho, t0_d, h, del_alpha=d, want_X=True)
    if tm: tm[2].record()
    return out
for _ in range(3): out = step()
torch.cuda.synchronize()
tms = [[ev() for _ in range(3)] for _ in range(steps)]
for k in range(steps): step(tms[k])
torch.cuda.synchronize()
geo_ms = np.mean([t[0].elapsed_time(t[1]) for t in tms]); sol_ms = np.mean([t[1].elapsed_time(t[2]) for t in tms])
val, grad, X, dX, info = out
bad = int(((info >> 16) & 3).count_nonzero().item())
print(json.dumps({"workload": "adjoint batch (BASELINE configs[4]), per-GPU share", "points": npts, "N": 1025,
                  "adjoint_solves_per_s": npts / ((geo_ms + sol_ms) * 1e-3), "geometry_ms": geo_ms, "solve_plus_adjoint_ms": sol_ms,
                  "mean_iterations": float((info & 0xffff).double().mean().item()), "bad": bad,
                  "grad_finite": bool(torch.isfinite(grad).all().item())}))
