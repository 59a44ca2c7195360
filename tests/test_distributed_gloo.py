"""CPU, world_size = 2 over gloo: the one exchange step of the scan (all-gather of the per-surface maxima,
replacing the three MPI.Gather of ball_scan.py:345-347) and the surface sharding arithmetic."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ns_total, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from ideal_ballooning_solver_b200 import scan
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        gamma = rng.standard_normal((ns_total, 6, 5)) * 1e-3          # the "global" growth-rate grids
        gamma[min(1, ns_total - 1)] = 0.0                             # all-zero guard on one surface
        lo, hi = scan.shard_range(ns_total, rank, world)
        loc = gamma[lo:hi].reshape(hi - lo, gamma.shape[1] * gamma.shape[2])
        # local per-surface arg-max with the reference's rules (what ibs_scan_argmax produces on the GPU)
        val = torch.tensor([g.max() if hi > lo else 0.0 for g in loc], dtype=torch.float64)
        idx = torch.tensor([-1 if g.max() == 0.0 else int(np.flatnonzero(g == g.max())[0]) for g in loc], dtype=torch.int32)
        v_all, i_all = scan.gather_surface_maxima(val, idx, ns_total)
        # the form the scan itself uses: the producer writes packed (max, index) pairs straight into the send slot and the
        # exchange is ONE all_gather_into_tensor into a preallocated buffer (no zeros / cat / cast on the way)
        g = scan.SurfaceGather(ns_total, torch.device("cpu"))
        assert g.send.shape == (hi - lo, 2) and g.recv.shape[0] == world
        g.send[:, 0] = val
        g.send[:, 1] = idx.to(torch.float64)
        buf = g.exchange()
        assert buf.data_ptr() == g.recv.data_ptr()
        v2, i2 = g.unpack()
        assert torch.equal(v2, v_all) and torch.equal(i2, i_all)
        # asynchronous form with two buffers used alternately (the bench's N > 1 path): step k's exchange is waited for in step k + 1
        gs = [scan.SurfaceGather(ns_total, torch.device("cpu")) for _ in range(2)]
        for k in range(5):
            ga, gb = gs[k % 2], gs[(k + 1) % 2]
            ga.wait()
            ga.send[:, 0] = val + k
            ga.send[:, 1] = idx.to(torch.float64)
            ga.exchange(async_op=True)
            gb.wait()
            if k >= 1:
                vb, ib = gb.unpack()
                assert torch.equal(vb, v_all + (k - 1)) and torch.equal(ib, i_all)
        gs[0].wait()
        v4, i4 = gs[0].unpack()
        assert torch.equal(v4, v_all + 4) and torch.equal(i4, i_all)
        q.put((rank, v_all.numpy().copy(), i_all.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ns_total", [7, 8, 1])      # 1: more ranks than surfaces (a rank with an empty block still gathers)
def test_gather_surface_maxima_world2(ns_total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ns_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(7)
    gamma = rng.standard_normal((ns_total, 6, 5)) * 1e-3
    gamma[min(1, ns_total - 1)] = 0.0
    flat = gamma.reshape(ns_total, -1)
    want_v = flat.max(axis=1)
    want_i = np.array([-1 if g.max() == 0.0 else int(np.flatnonzero(g == g.max())[0]) for g in flat])
    for rank, v, i in res:          # every rank ends up with the full per-surface result, bit-exact
        assert np.array_equal(v, want_v), rank
        assert np.array_equal(i, want_i), rank
