"""GPU: per-solve iteration counts of the default batch on the original and on the bit-truncated base arrays; dumps the inputs of the
slowest solves (tools/data_sensitivity.py found an 8 % tail effect)."""
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
out = {}
for tag, base in (("orig", geo.base), ("trunc", (geo.base.view(torch.int64) & ~0xFFF).view(torch.float64))):
    sol, best, sig = engine.scan_solve_argmax(base, geo.dPdrho, th0, h, theta0.size, 1, want_X=True)
    it = (sol.info & 0xFFFF).cpu().numpy().reshape(st.ns, theta0.size)
    fl = (sol.info >> 16).cpu().numpy()
    print(tag, "mean", it.mean(), "max", it.max(), "hist", np.bincount(it.ravel())[:40], "flags", np.unique(fl))
    worst = np.argsort(it.ravel())[-6:][::-1]
    for w in worst:
        print("   line", w // theta0.size, "theta0 idx", w % theta0.size, "iters", it.ravel()[w])
    lines = sorted(set(int(w // theta0.size) for w in worst[:3]))
    for l in lines:
        out[f"{tag}_base_{l}"] = base.reshape(st.ns, 8, -1)[l].cpu().numpy(); out[f"{tag}_dP_{l}"] = geo.dPdrho.reshape(-1)[l].item()
        if l > 0:
            out[f"{tag}_base_{l-1}"] = base.reshape(st.ns, 8, -1)[l - 1].cpu().numpy(); out[f"{tag}_dP_{l-1}"] = geo.dPdrho.reshape(-1)[l - 1].item()
out["theta0"] = theta0; out["theta"] = theta
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/stragglers.npz", **out)
