#!/usr/bin/env python
"""GPU: does the time of the lane-per-chain solver depend on the low-order bits of its input?  (K1's one-reciprocal epilogue changed
the base arrays by 2-3 ulp and the solver ran 8-10 % slower with identical instruction counts.)  Default bench batch; the base
arrays as K1 wrote them, with +-1 ulp of noise, and with the low 12 mantissa bits cleared."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ideal_ballooning_solver_b200 import engine

st, alpha, theta0, theta = bench.build_tables("d3d", 37, 0)
dt = engine.DeviceTables.from_host(st)
geo = engine.geometry_batch(dt, torch.from_numpy(alpha).cuda(), torch.from_numpy(theta).cuda())
th0 = torch.from_numpy(theta0).cuda().repeat(st.ns)
h = engine.grid_spacing(theta)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")


import subprocess, tempfile, time


def timeit(base, tag):
    # clocks / power while this variant runs alone for ~1.5 s
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown",
                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=f, stderr=subprocess.DEVNULL)
    time.sleep(0.5)
    for _ in range(400):
        engine.scan_solve_argmax(base, geo.dPdrho, th0, h, theta0.size, 1, want_X=True)
    torch.cuda.synchronize()
    p.terminate(); p.wait()
    rows = [r.strip().split(",") for r in open(f.name) if r.strip()][6:]
    if rows:
        sm = [float(r[0]) for r in rows]; pw = [float(r[2]) for r in rows]
        print(f"{tag:28s} clocks.sm median {np.median(sm):.0f} min {min(sm):.0f}  power median {np.median(pw):.0f} W max {max(pw):.0f} W  power-cap flags {sum('Active' in r[3] for r in rows)}/{len(rows)}", flush=True)
    ts = []
    for _ in range(30):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sol, best, sig = engine.scan_solve_argmax(base, geo.dPdrho, th0, h, theta0.size, 1, want_X=True)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    it = float((sol.info & 0xFFFF).double().mean().item())
    print(f"{tag:28s} solve {np.median(ts):.3f} ms (min {min(ts):.3f})  iters {it:.5f}", flush=True)


base = geo.base
timeit(base, "as K1 wrote them")
g = torch.Generator(device="cuda").manual_seed(1)
bits = base.view(torch.int64)
noise = torch.randint(-1, 2, base.shape, generator=g, device="cuda", dtype=torch.int64)
timeit((bits + noise).view(torch.float64), "+-1 ulp of noise")
timeit((bits & ~0xFFF).view(torch.float64), "low 12 mantissa bits cleared")
timeit((bits & ~0xFFFFFF).view(torch.float64), "low 24 mantissa bits cleared")
timeit(base, "as K1 wrote them (again)")
