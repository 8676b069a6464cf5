"""The temporally-blocked Jacobi kernel's algorithm (lane-level model of the CUDA code) must equal
K plain sweeps of the oracle BITWISE — including walls, mirrored x-halos, chunk seams."""
import numpy as np
import pytest

from conftest import bits_equal
from jacobi_tb_model import jacobi_tb_model


@pytest.mark.parametrize("W,H,K,R", [(128, 40, 3, 16), (256, 33, 8, 12), (16, 9, 2, 4),
                                     (132, 50, 5, 50), (240, 31, 1, 7), (112, 2, 4, 2),
                                     (360, 21, 10, 9)])
def test_model_equals_k_oracle_sweeps(oracle, W, H, K, R):
    rng = np.random.default_rng(W * 1000 + H * 10 + K)
    p = rng.standard_normal((H, W)).astype(np.float32)
    d = rng.standard_normal((H, W)).astype(np.float32)
    got = jacobi_tb_model(p, d, K, R)
    assert bits_equal(got, oracle.jacobi(p, d, K))


def test_model_with_fused_clear(oracle):
    rng = np.random.default_rng(5)
    p = rng.standard_normal((24, 64)).astype(np.float32)
    d = rng.standard_normal((24, 64)).astype(np.float32)
    got = jacobi_tb_model(p, d, 4, 8, scale=0.8)
    assert bits_equal(got, oracle.jacobi(oracle.clear(p, 0.8), d, 4))


def test_model_on_a_slab_with_ghost_rows(oracle):
    """Interior slab of a taller grid: local buffer = owned rows +- K ghost rows; no wall logic."""
    rng = np.random.default_rng(6)
    H, W, K = 64, 128, 4
    p = rng.standard_normal((H, W)).astype(np.float32)
    d = rng.standard_normal((H, W)).astype(np.float32)
    full = oracle.jacobi(p, d, K)
    for r0, r1 in [(0, 20), (20, 44), (44, 64)]:
        lo, hi = max(r0 - K, 0), min(r1 + K, H)
        got = jacobi_tb_model(p[lo:hi], d[lo:hi], K, 10, row_off=lo, out_lo=r0, out_hi=r1, H=H)
        assert bits_equal(got[r0 - lo:r1 - lo], full[r0:r1])
