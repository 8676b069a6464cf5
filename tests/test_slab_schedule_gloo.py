"""The multi-GPU schedule (which rows are exchanged when, which are recomputed redundantly) checked
on CPU with world_size 2 and 3 over gloo.  Each rank keeps FULL-SIZE arrays poisoned with NaN outside
the rows the plan says it may trust, runs the oracle passes, and exchanges exactly the messages
`SlabPlan` / fluid.cu issue.  After every step the rows a rank owns must be NaN-free and bitwise
equal to the single-domain oracle: a halo that is one row too thin shows up as a NaN."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

W, H, WD, HD = 64, 96, 96, 192
ITERS, DT, HALO = 23, 0.016666, 14   # halo >= iterations+2 is added by the plan when it fits
NSTEPS = 3


def _exchange(arr, r0, r1, n, rank, world):
    """send my top/bottom n owned rows, receive the neighbours' into rows just outside [r0,r1)."""
    if n <= 0:
        return
    reqs, bufs = [], []
    for peer, send_rows, recv_rows in ((rank + 1, slice(r1 - n, r1), slice(r1, r1 + n)),
                                       (rank - 1, slice(r0, r0 + n), slice(r0 - n, r0))):
        if 0 <= peer < world:
            s = torch.from_numpy(np.ascontiguousarray(arr[send_rows]))
            r = torch.empty_like(s)
            reqs += [dist.isend(s, peer), dist.irecv(r, peer)]
            bufs.append((recv_rows, r))
    for q in reqs:
        q.wait()
    for rows_, r in bufs:
        arr[rows_] = r.numpy()


def _poison(arr, lo, hi):
    """keep rows [lo,hi), NaN everywhere else"""
    out = np.full_like(arr, np.nan)
    lo, hi = max(lo, 0), min(hi, arr.shape[0])
    out[lo:hi] = arr[lo:hi]
    return out


def _worker(rank, world, port, q, q_iters):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from webgl_fluid_simulation_b200.slab import SlabPlan, jacobi_launches
    plan = SlabPlan(H, HD, rank, world, halo=HALO, iterations=q_iters)
    r0, r1, d0, d1, G, Gd = plan.row0, plan.row1, plan.drow0, plan.drow1, plan.G, plan.Gd
    rng = np.random.default_rng(0)                      # same on every rank
    v = (rng.standard_normal((H, W, 2)) * 40).astype(np.float32)
    dye = rng.random((HD, WD, 4), dtype=np.float32)
    p = rng.standard_normal((H, W)).astype(np.float32)
    ref = O.OracleSim(W, H, WD, HD, PRESSURE_ITERATIONS=ITERS)
    ref.velocity, ref.dye, ref.pressure = v.copy(), dye.copy(), p.copy()
    # what this rank may trust at step start: v on owned +-3, p and dye on owned rows
    v = _poison(v, r0 - 3, r1 + 3); p = _poison(p, r0, r1); dye = _poison(dye, d0, d1)
    ok = True
    with np.errstate(invalid="ignore"):
        for _ in range(NSTEPS):
            c = O.curl(v)                                            # valid owned +-2
            v = O.vorticity(v, c, 30.0, DT)                          # valid owned +-1
            div = _poison(O.divergence(v), r0, r1)                   # valid owned
            ext_of = plan.launch_extents(ITERS)
            deep = plan.deep(ITERS)
            if deep:                                                 # one group: p (iters+1) + div (iters)
                _exchange(p, r0, r1, ITERS + 1, rank, world); p = _poison(p, r0 - ITERS - 1, r1 + ITERS + 1)
                _exchange(div, r0, r1, ITERS, rank, world)
            else:
                _exchange(div, r0, r1, max(k for k, _ in ext_of), rank, world)
            for i, (k, ext) in enumerate(ext_of):
                if not deep:
                    _exchange(p, r0, r1, k + ext, rank, world)
                    p = _poison(p, r0 - k - ext, r1 + k + ext)
                if i == 0:
                    p = O.clear(p, 0.8)
                p = O.jacobi(p, div, k)
                p = _poison(p, r0 - ext, r1 + ext)                   # what the launch writes
            v = _poison(O.gradient_subtract(p, v), r0, r1)           # valid owned
            _exchange(v, r0, r1, G, rank, world)
            v = _poison(O.advect(v, v, DT, 0.2), r0 - 3, r1 + 3)     # redundant ghost compute
            _exchange(dye, d0, d1, Gd, rank, world)
            dye = _poison(O.advect(v, dye, DT, 1.0), d0, d1)
            ref.step(DT)
            for mine, full, a, b in ((v, ref.velocity, r0, r1), (p, ref.pressure, r0, r1),
                                     (dye, ref.dye, d0, d1), (div, ref.divergence, r0, r1)):
                same = np.array_equal(mine[a:b].view(np.uint32), np.ascontiguousarray(full[a:b]).view(np.uint32))
                ok = ok and same and not np.isnan(mine[a:b]).any()
            # the +-3 velocity ghosts must be right too: the next step's curl reads them
            lo, hi = max(r0 - 3, 0), min(r1 + 3, H)
            ok = ok and np.array_equal(v[lo:hi].view(np.uint32), np.ascontiguousarray(ref.velocity[lo:hi]).view(np.uint32))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,plan_iters", [(2, ITERS), (3, ITERS), (2, 5)])
def test_slab_schedule_is_exact(world, plan_iters):
    """plan_iters = ITERS: ghost zone sized for the deep (one message per solve) form;
    plan_iters = 5: ghost zone too thin for 23 sweeps -> per-launch fallback."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    from conftest import free_port
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, plan_iters)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
    assert res == [(r, True) for r in range(world)], res


def test_plan_arithmetic():
    from webgl_fluid_simulation_b200.slab import SlabPlan, jacobi_launches, rows
    assert jacobi_launches(50, 10) == [10] * 5 and jacobi_launches(50, 8) == [8, 7, 7, 7, 7, 7, 7]
    assert jacobi_launches(20, 12) == [10, 10] and jacobi_launches(0) == [] and sum(jacobi_launches(37, 9)) == 37
    assert [rows(10, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]
    pl = SlabPlan(8192, 8192, 3, 8, iterations=40)
    assert (pl.row0, pl.row1, pl.G, pl.Gd) == (3072, 4096, 64, 64)
    assert pl.jacobi_messages(40) == [("pressure+divergence", 41)]
    assert pl.launch_extents(40) == [(10, 31), (10, 21), (10, 11), (10, 1)]
    thin = SlabPlan(8192, 8192, 3, 8, halo=32, iterations=5)        # iterations raised after creation
    assert thin.G == 32 and not thin.deep(40)
    assert thin.jacobi_messages(40) == [("divergence", 10)] + [("pressure", 10)] * 3 + [("pressure", 11)]
    assert SlabPlan(4096, 4096, 0, 1).jacobi_messages(50) == []
    with pytest.raises(ValueError):
        SlabPlan(64, 64, 0, 8)          # 8-row slabs cannot hold a 14-row halo
