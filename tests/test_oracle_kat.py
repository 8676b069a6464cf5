"""Hand-derived known-answer tests of the CPU oracle (SURVEY §8c list).  No GPU."""
import numpy as np

from conftest import bits_equal


def test_zero_is_a_fixed_point(oracle):
    O = oracle; H, W = 12, 20
    v = np.zeros((H, W, 2), np.float32); p = np.zeros((H, W), np.float32)
    assert not O.curl(v).any() and not O.divergence(v).any()
    assert not O.vorticity(v, p, 30.0, 0.016).any()
    assert not O.jacobi(p, p, 5).any() and not O.gradient_subtract(p, v).any()
    assert not O.advect(v, v, 0.016, 0.2).any()


def test_constant_pressure(oracle):
    O = oracle; H, W = 9, 13
    p = np.full((H, W), 3.0, np.float32)
    div = np.random.default_rng(0).standard_normal((H, W)).astype(np.float32)
    assert bits_equal(O.jacobi(p, div, 1), (np.float32(12.0) - div) * np.float32(0.25))
    v = np.random.default_rng(1).standard_normal((H, W, 2)).astype(np.float32)
    assert bits_equal(O.gradient_subtract(p, v), v)      # clamp cancels the gradient on the walls


def test_constant_velocity_curl_and_wall_divergence(oracle):
    O = oracle; H, W = 10, 14; a, b = 2.0, -3.0
    v = np.empty((H, W, 2), np.float32); v[..., 0] = a; v[..., 1] = b
    assert not O.curl(v).any()
    d = O.divergence(v)
    exp = np.zeros((H, W), np.float32)
    exp[:, 0] += a; exp[:, W - 1] -= a; exp[0, :] += b; exp[H - 1, :] -= b   # S:804-807
    assert np.array_equal(d, exp)


def test_single_impulse_jacobi(oracle):
    O = oracle; H, W = 8, 8
    p = np.zeros((H, W), np.float32); p[3, 4] = 1.0
    q = O.jacobi(p, np.zeros_like(p), 1)
    exp = np.zeros_like(p); exp[3, 3] = exp[3, 5] = exp[2, 4] = exp[4, 4] = 0.25
    assert np.array_equal(q, exp)
    p = np.zeros((H, W), np.float32); p[0, 0] = 1.0        # corner: clamp feeds itself twice
    q = O.jacobi(p, np.zeros_like(p), 1)
    assert q[0, 0] == 0.5 and q[0, 1] == 0.25 and q[1, 0] == 0.25 and q.sum() == 1.0


def test_linear_ramp_gradient(oracle):
    O = oracle; H, W = 6, 10
    p = np.tile(np.arange(W, dtype=np.float32), (H, 1))
    g = O.gradient_subtract(p, np.zeros((H, W, 2), np.float32))
    assert np.all(g[:, 1:-1, 0] == -2) and np.all(g[:, 0, 0] == -1) and np.all(g[:, -1, 0] == -1)
    assert not g[..., 1].any()


def test_advection_rest_and_integer_shift(oracle):
    O = oracle; H, W = 16, 16
    rng = np.random.default_rng(2)
    src = rng.random((H, W, 4), dtype=np.float32)
    v0 = np.zeros((H, W, 2), np.float32)
    dt = np.float32(0.5)
    out = O.advect(v0, src, dt, 1.0)
    assert bits_equal(out, src / (np.float32(1) + np.float32(1.0) * dt))
    k = 3
    v = np.zeros((H, W, 2), np.float32); v[..., 0] = 2 * k      # dt*v = k texels (S:777)
    out = O.advect(v, src, dt, 0.0)
    exp = src[:, np.clip(np.arange(W) - k, 0, W - 1)]
    assert bits_equal(out, exp)


def test_splat_is_centred_gaussian(oracle):
    O = oracle; H, W = 32, 32
    base = np.zeros((H, W, 4), np.float32)
    rad = O.correct_radius(0.25, 1.0)
    out = O.splat(base, 1.0, 0.5, 0.5, (1.0, 2.0, 3.0), rad)
    assert np.array_equal(out, out[::-1]) and np.array_equal(out, out[:, ::-1])
    r2 = 2 * (0.5 / 32) ** 2
    assert abs(out[16, 16, 0] - np.exp(-r2 / 0.0025)) < 1e-6
    assert np.all(out[..., 3] == 1.0)
    assert np.allclose(out[..., 1], 2 * out[..., 0]) and np.allclose(out[..., 2], 3 * out[..., 0])


def test_half_storage_rounds_to_fp16(oracle):
    O = oracle
    a = np.array([1.0, 1.0 + 2 ** -11, 1.0 + 3 * 2 ** -11, 65520.0, 1e-8], np.float32)
    assert np.array_equal(O.round_half(a), a.astype(np.float16).astype(np.float32))
