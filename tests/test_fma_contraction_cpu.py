"""The arithmetic identity the temporally blocked Jacobi kernel (csrc/jacobi.cuh, generation 7) relies on:

        fma(S, 0.25, -0.25*d)  ==  (S - d) * 0.25        bit for bit in fp32,

whenever d == 0 or |d| >= 2^-123 (and nothing overflows) -- and that the guard is NEEDED: for
0 < |d| < 2^-123 the two sides can differ (double rounding into the subnormal range), which is why
the producers of `divergence` keep the tiny-value map and flagged warps run the exact form.
Checked here with exact rational arithmetic (no GPU, no fma instruction needed)."""
from fractions import Fraction

import numpy as np

F32 = np.float32
TINY = 2.0 ** -123


def rn32(x: Fraction) -> np.float32:
    """Correctly rounded (nearest-even) fp32 of an exact rational, subnormals included."""
    if x == 0:
        return F32(0.0)
    s = -1 if x < 0 else 1
    a = abs(x)
    # exponent e with 2^e <= a < 2^(e+1)
    e = a.numerator.bit_length() - a.denominator.bit_length()
    if Fraction(2) ** e > a:
        e -= 1
    e = max(e, -126)                      # subnormals share the exponent of the smallest normal
    ulp = Fraction(2) ** (e - 23)
    q, r = divmod(a, ulp)
    q = int(q)
    if r * 2 > ulp or (r * 2 == ulp and (q & 1)):
        q += 1
    v = Fraction(q) * ulp
    if v >= Fraction(2) ** 128:
        return F32(s * np.inf)
    return F32(s * float(v))              # exactly representable: float() is exact


def ref(S, d):     # the reference tail: two roundings
    return rn32(Fraction(float(rn32(Fraction(float(S)) - Fraction(float(d))))) * Fraction(1, 4))


def fma(S, d):     # the contracted tail: d' = rn(-0.25 d), then one rounding of S*0.25 + d'
    dq = rn32(Fraction(float(d)) * Fraction(-1, 4))
    return rn32(Fraction(float(S)) * Fraction(1, 4) + Fraction(float(dq)))


def _bits(x):
    return np.array([x], F32).view(np.uint32)[0]


def _samples(rng, n):
    """(S, d) pairs stressing cancellation and the bottom of the exponent range."""
    out = []
    for _ in range(n):
        kind = rng.integers(0, 6)
        if kind == 0:       # ordinary magnitudes
            S, d = rng.standard_normal() * 10 ** rng.uniform(-3, 3), rng.uniform(-1, 1)
        elif kind == 1:     # heavy cancellation: S within a few ulps of d
            d = rng.standard_normal()
            S = float(F32(d)) * (1 + rng.integers(-8, 9) * 2.0 ** -23)
        elif kind == 2:     # both near the bottom of the normal range
            S, d = rng.uniform(-1, 1) * 2.0 ** rng.integers(-126, -110), rng.uniform(-1, 1) * 2.0 ** rng.integers(-123, -110)
        elif kind == 3:     # subnormal S, d just above the guard
            S, d = rng.uniform(-1, 1) * 2.0 ** -127, (1 + rng.uniform(0, 1)) * 2.0 ** -123 * rng.choice([-1, 1])
        elif kind == 4:     # d == 0 with tiny S
            S, d = rng.uniform(-1, 1) * 2.0 ** rng.integers(-149, -120), 0.0
        else:               # |S - d| lands in the double-rounding band [2^-125, 2^-124) with a safe d
            d = (1 + rng.uniform(0, 1)) * 2.0 ** -123
            S = d + rng.uniform(1, 2) * 2.0 ** -125 * rng.choice([-1, 1])
        out.append((F32(S), F32(d)))
    return out


def test_contracted_tail_equals_reference_outside_the_guard():
    rng = np.random.default_rng(123)
    checked = 0
    for S, d in _samples(rng, 6000):
        if d != 0 and abs(float(d)) < TINY:
            continue
        assert _bits(ref(S, d)) == _bits(fma(S, d)), (float(S), float(d))
        checked += 1
    assert checked > 5000


def test_guard_is_needed_for_tiny_nonzero_divergence():
    """A concrete pair with 0 < |d| < 2^-123 where the contraction is NOT bit-exact."""
    rng = np.random.default_rng(7)
    found = 0
    for _ in range(4000):
        d = F32(rng.integers(1, 2 ** 22) * 2.0 ** -149 * rng.choice([-1, 1]))          # subnormal d
        S = F32(rng.uniform(1, 2) * 2.0 ** -125 * rng.choice([-1, 1]))
        if _bits(ref(S, d)) != _bits(fma(S, d)):
            found += 1
    assert found > 0


def test_signed_zeros_agree():
    for S in (F32(0.0), F32(-0.0)):
        for d in (F32(0.0), F32(-0.0)):
            assert _bits(ref(S, d)) == _bits(fma(S, d))
