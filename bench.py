#!/usr/bin/env python
"""Benchmark of the ideal-ballooning hot path: field-line solves/sec (fp64, lambda_max + eigenvector).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload d3d|ncsx|hberg]

One *step* = one pass of the hot path over one batch of synthetic equilibria:
K1 geometry (ns x nalpha field lines) -> K2+K3 (ns x nalpha x nth0 solves, lambda + eigenvector) ->
guarded per-surface arg-max (-> one all-gather of the per-surface maxima when N > 1).
Default workload = BASELINE.json configs[1]: D3D-like (nfp=1, 80/84 modes) scan of 128 surfaces x 64
theta0 at ntheta=1024, for ``--equilibria`` independently perturbed equilibria per step (the reference
scans totalndofs+1 perturbed equilibria per outer iteration, sims_runner_D3D.py:109).  With N GPUs
every rank scans its own equilibria (weak scaling, no data-path collective but the final gather).

``value``   device-timed throughput with the Fourier tables already resident in HBM.
``e2e``     the same metric through the host-buffer C-ABI call ``ibs_scan_host`` (pinned host tables in,
            gamma grid + arg-max + eigenfunction at each surface's maximum out), copies inside the timing.
``roofline``  the solver kernel (K2+K3): algorithmic bytes 32*N per solve / its CUDA-event time vs the
            measured HBM copy bandwidth of MEASURED_PEAKS.json.
``cpu_baseline`` / ``--impl reference``: the oracle port of the reference's numpy/scipy path (dense matrix +
            ARPACK shift-invert, shipped tol) on the host cores, one process per core, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    #          kind    ns   nalpha nth0 ntheta span(pi)
    "d3d":   ("d3d",   128, 1,     64,  1024,  4),
    "ncsx":  ("ncsx",  64,  32,    32,  2048,  4),
    "hberg": ("hberg", 256, 64,    1,   8192,  8),
}
KERNELS_PER_STEP = 7        # pack_mn, pack_nyq, geometry, dpdrho, scan_prep (or poly_prep), scan_solve (or solve), argmax  (profiles/launches_*.csv)


def solver_kernel_name(nth0, N):
    """Which K2+K3 kernel the library dispatches a scan-shaped batch to (mirrors scan_solver_eligible, ibs_scan_solver.cu)."""
    scan_ok = os.environ.get("IBS_SCAN", "1") != "0" and nth0 >= 4 and (N & 1) == 1 and N >= 65
    return "scan_solve_kernel (K2+K3, lane per solve)" if scan_ok else "solve_kernel (K2+K3, team per solve)"


def workload_grids(name):
    kind, ns, na, nt, nth, span = WORKLOADS[name]
    s = np.linspace(0.5, 0.95, ns)                                   # ball_scan.py:197
    alpha = np.linspace(0.0, np.pi, na) if na > 1 else np.array([0.0])
    theta0 = np.linspace(0.0, 0.5 * np.pi, nt) if nt > 1 else np.array([0.0])
    theta = np.linspace(-span * np.pi, span * np.pi, nth + 1)        # ntheta intervals -> N = ntheta+1 points
    return kind, s, alpha, theta0, theta


def build_tables(name, equilibria, seed0):
    """Host-side setup (not timed): synthetic equilibria -> radial splines -> per-surface tables."""
    import dataclasses
    from ideal_ballooning_solver_b200 import synthetic, tables
    kind, s, alpha, theta0, theta = workload_grids(name)
    parts = [tables.RadialSplines(synthetic.make_equilibrium(kind, seed=seed0 + e)).evaluate(s) for e in range(equilibria)]
    st = dataclasses.replace(parts[0], tab_mn=np.concatenate([p.tab_mn for p in parts]),
                             tab_nyq=np.concatenate([p.tab_nyq for p in parts]),
                             scal=np.concatenate([p.scal for p in parts]))
    return st, alpha, theta0, theta


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        if os.environ.get("IBS_BENCH_NO_SAMPLER"):       # diagnostic: is a slow step caused by the nvidia-smi queries?
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's numpy/scipy path
# ---------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One field line: geometry (vmec_fieldlines) + `nsolve` gamma_ball_full calls, reference algorithm."""
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"     # slurm_ball_scan_template.sl:10
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import ballooning_oracle as bo
    st_one, alpha, theta0s, theta = args
    t0 = time.perf_counter()
    fl = bo.fieldlines(st_one, np.array([alpha]), theta)
    dP = bo.dpdrho_of(fl)
    vg = bo.default_vguess(theta, theta_fac=int(round(theta[-1] / np.pi)))
    n = 0
    for th0 in theta0s:
        cv, gd = bo.theta0_shift(fl, th0)
        lam, X, *_ = bo.gamma_ball_full(dP, theta, fl.bmag[0][0], fl.gradpar_theta_pest[0][0], cv, gd, vg, 1.0,
                                        tol=5.0e-7, method="arpack")
        vg = X[1:-1]
        n += 1
    return n, time.perf_counter() - t0


def cpu_reference_rate(workload, lines_per_core=1, solves_per_line=6, cores=None):
    """Solves/s of the reference algorithm on the host cores for a bounded sample of `workload`."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    st, alpha, theta0, theta = build_tables(workload, 1, 12345)
    nlines = cores * lines_per_core
    rng = np.random.default_rng(0)
    tasks = []
    for k in range(nlines):
        js = int(rng.integers(0, st.ns))
        tasks.append((st.select([js]), float(alpha[k % len(alpha)]),
                      np.asarray(theta0[np.linspace(0, len(theta0) - 1, min(solves_per_line, len(theta0))).astype(int)]), theta))
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, tasks, chunksize=1)
    wall = time.perf_counter() - t0
    nsolve = sum(r[0] for r in res)
    percore = nsolve / sum(r[1] for r in res)
    sample = (f"{nlines} field lines x {len(tasks[0][2])} theta0 of the {workload} workload (N={len(theta)} points): "
              f"reference algorithm (numpy geometry + dense matrix + ARPACK shift-invert, tol=5e-7), "
              f"{cores} processes x 1 BLAS thread; {percore:.2f} solves/s/core")
    return nsolve / wall, cores, sample, wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, s, alpha, theta0, theta = workload_grids(args.workload)
    vals = []
    t_all = time.perf_counter()
    steps = max(1, min(args.steps, 3))
    for _ in range(steps):
        v, cores, sample, wall = cpu_reference_rate(args.workload)
        vals.append(v)
    value = float(np.median(vals))
    line = {"metric": "field-line ballooning solves/sec (fp64, lambda_max+eigvec)", "value": value, "unit": "solves/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": steps, "warmup": 0,
            "ms_per_step": 1e3 * (time.perf_counter() - t_all) / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": describe(args.workload, args.equilibria)},
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def describe(workload, equilibria):
    kind, ns, na, nt, nth, span = WORKLOADS[workload]
    return (f"{kind.upper()}-like synthetic VMEC-shaped scan: {ns} surfaces x {na} alpha x {nt} theta0, ntheta={nth} "
            f"(N={nth + 1} points, theta in +-{span}pi), x {equilibria} equilibria per step per GPU")


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="d3d", choices=sorted(WORKLOADS))
    ap.add_argument("--equilibria", type=int, default=37,
                    help="independent equilibria batched per step per GPU (37 x 128 lines x 2 theta0 groups = 9472 items = 8 full rounds "
                         "of the 148 x 8 resident warps of the solver kernel; 16 leaves the last of 3.5 rounds half empty: -12%%)")
    ap.add_argument("--chain", type=int, default=-1, help="warm-start run length over theta0 (-1 = scan default, 1 = off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from ideal_ballooning_solver_b200 import engine, scan, _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line (NCCL prints its version banner)
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank if world > 1 else torch.cuda.current_device())
    torch.cuda.set_device(dev)
    _lib.load(build_if_missing=True)

    E = args.equilibria
    _, _, na_w, nt_w, _, _ = WORKLOADS[args.workload]
    # warm-start chain over theta0 (or over alpha when the scan has a single theta0), ball_scan.py:265-274
    chain = scan.chain_length(nt_w if nt_w > 1 else na_w) if args.chain < 0 else max(1, args.chain)
    st, alpha, theta0, theta = build_tables(args.workload, E, seed0=1000 * rank)
    kind, ns1, na, nt, nth, span = WORKLOADS[args.workload]
    N = nth + 1
    ns = st.ns
    nlines, nsolve = ns * na, ns * na * nt
    h = engine.grid_spacing(theta)
    dt = engine.DeviceTables.from_host(st, dev)
    alpha_d = torch.from_numpy(alpha).to(dev)
    theta_d = torch.from_numpy(theta).to(dev)
    th0_d = torch.from_numpy(theta0).to(dev).repeat(nlines)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)     # > 126 MB L2
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(timers=None):
        if timers is not None:
            timers[0].record()
        geo = engine.geometry_batch(dt, alpha_d, theta_d)
        if timers is not None:
            timers[1].record()
        sol = engine.solve_base_batch(geo.base, geo.dPdrho, th0_d, h, nth0=nt, want_X=True, want_dX=False,
                                      want_matrix=False, chain_len=chain)
        if timers is not None:
            timers[2].record()
        val, idx, sig = engine.scan_argmax(sol.lam.reshape(ns, na * nt))
        if world > 1:
            val, idx = scan.gather_surface_maxima(val, idx, ns * world)
        if timers is not None:
            timers[3].record()
        return sol, val, idx

    # the clock sampler starts BEFORE the warm-up (nvidia-smi's start-up stalls the device for a few ms) and keeps
    # sampling through the timed region
    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        sol, val, idx = step()
        flush.zero_()
    torch.cuda.synchronize()
    flags = (sol.info >> 16).cpu().numpy()
    nbad = int(np.count_nonzero(flags & 3))
    mean_iters = float((sol.info & 0xFFFF).double().mean().item())

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    timers = [[ev() for _ in range(4)] for _ in range(args.steps)]
    t_wall = time.perf_counter()
    for k in range(args.steps):
        step(timers[k])
        flush.zero_()                      # L2 flush between timed iterations (outside the event brackets)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if sampler else None
    t_step = np.array([t[0].elapsed_time(t[3]) for t in timers])          # ms
    t_geo = np.array([t[0].elapsed_time(t[1]) for t in timers])
    t_solve = np.array([t[1].elapsed_time(t[2]) for t in timers])
    total_ms = torch.tensor([t_step.sum()], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = total_ms.item() * 1e-3
    value = world * nsolve * args.steps / total_s

    # ---- end to end through the host-buffer C-ABI call (pinned host buffers, copies inside the timing)
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        import dataclasses
        st_p = dataclasses.replace(st, tab_mn=pin(st.tab_mn), tab_nyq=pin(st.tab_nyq), scal=pin(st.scal))
        out = dict(gamma=pin(np.empty((ns, na, nt))), val=pin(np.empty(ns)), sigma0=pin(np.empty(ns)),
                   idx=pin(np.empty(ns, dtype=np.int32)), xbest=pin(np.empty((ns, N))))
        a_p, t0_p, th_p = pin(alpha), pin(theta0), pin(theta)
        for _ in range(2):
            engine.scan_host(st_p, a_p, t0_p, th_p, want_xbest=True, out=out)
        if world > 1:
            dist.barrier()
        ke = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(ke):
            r = engine.scan_host(st_p, a_p, t0_p, th_p, want_xbest=True, out=out)
        dt_e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt_e, op=dist.ReduceOp.MAX)
        h2d = st.tab_mn.nbytes + st.tab_nyq.nbytes + st.scal.nbytes + alpha.nbytes + theta.nbytes + 8 * nsolve
        d2h = 8 * nsolve + 4 * nsolve + ns * (8 + 8 + 4) + 8 * ns * N
        e2e = {"value": world * nsolve * ke / dt_e.item(), "unit": "solves/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "bad_solves": int(r[4])}

    if world > 1 and os.environ.get("IBS_BENCH_VERBOSE"):
        print(f"[rank {rank}] step ms min/med/max {t_step.min():.3f}/{np.median(t_step):.3f}/{t_step.max():.3f} geo {t_geo.mean():.3f} "
              f"solve {t_solve.mean():.3f} iters {mean_iters:.2f}", file=sys.stderr, flush=True)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = load_peaks()
    alg_bytes = 32.0 * N * nsolve                       # g, c, f in + X out per solve
    solve_ms = float(t_solve.mean())
    achieved = alg_bytes / (solve_ms * 1e-3) / 1e9
    traffic = None
    rf = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.isfile(rf):
        try:
            per_solve = json.load(open(rf)).get(args.workload, {}).get("dram_bytes_per_solve")
            traffic = None if per_solve is None else float(per_solve) * nsolve
        except Exception:
            traffic = None
    line = {
        "metric": "field-line ballooning solves/sec (fp64, lambda_max+eigvec)",
        "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": float(total_ms.item() / args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": describe(args.workload, E), "solves_per_step_per_gpu": nsolve,
                   "field_lines_per_step_per_gpu": nlines, "l2": "flushed between timed steps (256 MB write)",
                   "eigvec": "X written to HBM for every solve",
                   "mean_solver_iterations": mean_iters,      # fine-grid-equivalent evaluations per solve (output passes not counted)
                   "solver": solver_kernel_name(nt, N), "theta0_chain": chain,
                   "bad_solves": nbad},
        "roofline": {"bound": "hbm", "kernel": solver_kernel_name(nt, N), "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": solve_ms,
                     "share_of_step": float(t_solve.sum() / t_step.sum()),
                     "note": "FP64-pipe bound, not HBM bound (DESIGN.md section 3): frac is the HBM fraction asked for; kernel_ms = CUDA-event time of the solve call (coefficient prep + solver kernel); traffic = ncu dram bytes per launch of the solver kernel"},
        "kernel_ms": {"geometry(K1 incl. pack+dPdrho)": float(t_geo.mean()), "solve(K2+K3)": solve_ms,
                      "step": float(t_step.mean())},
        "gpu_launches": KERNELS_PER_STEP * args.steps,
        "step_ms": {"min": float(t_step.min()), "median": float(np.median(t_step)), "max": float(t_step.max())},
        "clocks": clocks, "wall_s_timed_region": wall,
    }
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample, cw = cpu_reference_rate(args.workload)
        line["cpu_baseline"] = {"value": v, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
