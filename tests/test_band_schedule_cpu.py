"""The banded schedule of fluid_pressure_solve_host (fluid.cu solve_host_banded, mirrored by slab.host_bands)
checked on CPU in the style of test_slab_schedule_gloo.py: rows that have not been uploaded yet are NaN, every
launch keeps only the rows the plan says it writes, and the assembled result must be NaN-free and bitwise equal
to the one-piece solve.  An upload chunk or a launch extent that is one row short shows up as a NaN."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _keep(arr, lo, hi):
    out = np.full_like(arr, np.nan)
    out[max(lo, 0):min(hi, arr.shape[0])] = arr[max(lo, 0):min(hi, arr.shape[0])]
    return out


@pytest.mark.parametrize("H,iters,bands,block", [(1024, 20, 16, 10), (1000, 23, 3, 7), (2048, 50, 8, 10), (600, 9, 2, 12)])
def test_banded_host_solve_schedule_is_exact(H, iters, bands, block):
    from oracle import oracle as O
    from webgl_fluid_simulation_b200.slab import host_bands
    W = 16
    rng = np.random.default_rng(3)
    p = rng.standard_normal((H, W)).astype(np.float32)
    d = rng.uniform(-1, 1, (H, W)).astype(np.float32)
    want = O.jacobi(O.clear(p, 0.8), d, iters)
    plan = host_bands(H, iters, bands, block)
    assert len(plan) >= 2 and plan[0]["lo"] == 0 and plan[-1]["hi"] == H
    assert all(a["hi"] == b["lo"] for a, b in zip(plan, plan[1:]))                    # bands tile the grid
    assert plan[0]["up_lo"] == 0 and plan[-1]["up_hi"] == H
    assert all(a["up_hi"] == b["up_lo"] for a, b in zip(plan, plan[1:]))              # every row uploaded exactly once
    dev_p = np.full_like(p, np.nan); dev_d = np.full_like(d, np.nan)                   # device fields before any upload
    result = np.full_like(p, np.nan)
    with np.errstate(invalid="ignore"):
        for b in plan:
            dev_p[b["up_lo"]:b["up_hi"]] = p[b["up_lo"]:b["up_hi"]]
            dev_d[b["up_lo"]:b["up_hi"]] = d[b["up_lo"]:b["up_hi"]]
            cur = dev_p
            for k, (K, lo, hi) in enumerate(b["launches"]):
                src = O.clear(cur, 0.8) if k == 0 else cur                             # SCALE fused into the first launch
                cur = _keep(O.jacobi(src, dev_d, K), lo, hi)                           # band-private rows: what the launch writes
            assert (lo, hi) == (b["lo"], b["hi"])                                      # the last launch writes the owned rows only
            result[b["lo"]:b["hi"]] = cur[b["lo"]:b["hi"]]
    assert not np.isnan(result).any()
    assert np.array_equal(result.view(np.uint32), want.view(np.uint32))


def test_short_grids_take_the_one_piece_path():
    from webgl_fluid_simulation_b200.slab import host_bands
    assert host_bands(500, 20) == [] and host_bands(4096, 0) == [] and host_bands(300, 80) == []
    assert len(host_bands(4096, 50)) == 16 and len(host_bands(2048, 50)) == 8 and len(host_bands(1000, 20)) == 3
