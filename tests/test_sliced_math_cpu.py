"""The arithmetic behind the int8-sliced f64 GEMM (easy-ml_b200/csrc/gemm_sliced_f64.cu), restated in numpy so that the
error analysis the guard relies on is pinned without a GPU:

  x in (-1, 1)  ->  digits d_p in [-64, 64] with  x = sum_{p<s} d_p 2^(-6-7p) + rho,  |rho| <= 2^(-7s) (1 + 2^-7);
  keeping only the first s' digits of an s-digit expansion obeys the same bound with s';
  |x y - sum_{p+q<s} dx_p dy_q 2^(-12-7(p+q))| <= (s + 2) 2^(-7s);
  the int32 accumulators cannot overflow for s K 2^12 < 2^31.
"""
from fractions import Fraction

import numpy as np
import pytest


def digits_two_halves(x, slices):
    """digits(): the 32-bit two-half fixed-point form used for 7 and 8 slices."""
    xs = x * 134217728.0
    h = int(np.rint(xs))
    low = int(np.rint((xs - h) * (268435456.0 if slices == 8 else 2097152.0)))
    d = [0] * slices
    for p in range(slices - 1, 4, -1):
        r = ((low + 64) & 127) - 64
        d[p] = r
        low = (low - r) >> 7
    d[4] = low
    for p in (3, 2, 1):
        r = ((h + 64) & 127) - 64
        d[p] = r
        h = (h - r) >> 7
    d[0] = h
    return d


def digits_fp(x, slices):
    """digits(): the FP64 form (round to nearest from the top) used for 6, 9 and 10 slices."""
    t = x * 64.0
    d = []
    for _ in range(slices):
        r = float(np.rint(t))
        d.append(int(r))
        t = (t - r) * 128.0
    return d


def value(d):
    return sum(Fraction(v) * Fraction(1, 2 ** (6 + 7 * p)) for p, v in enumerate(d))


def samples():
    r = np.random.default_rng(5)
    xs = list(r.uniform(-1, 1, 300)) + list(10.0 ** r.uniform(-12, 0, 200) * r.choice([-1, 1], 200))
    xs += [0.0, 0.5, -0.5, 1 - 2.0 ** -53, -(1 - 2.0 ** -53), 2.0 ** -6, 2.0 ** -7, 2.0 ** -55, 2.0 ** -56, 3 * 2.0 ** -57,
           63.5 / 64, -63.5 / 64, 1 / 128 + 2.0 ** -60, 2.0 ** -300]
    return xs


@pytest.mark.parametrize("slices,scheme", [(7, digits_two_halves), (8, digits_two_halves), (6, digits_fp), (8, digits_fp),
                                           (9, digits_fp), (10, digits_fp)])
def test_digit_expansion_bounds(slices, scheme):
    for x in samples():
        d = scheme(float(x), slices)
        assert all(-64 <= v <= 64 for v in d), (x, d)
        rest = abs(Fraction(float(x)) - value(d))
        assert rest <= Fraction(1, 2 ** (7 * slices)) * (1 + Fraction(1, 128)), (x, d)
        # the first s' digits of the expansion are an s'-digit expansion within the same bound
        for keep in range(6, slices):
            rest = abs(Fraction(float(x)) - value(d[:keep]))
            assert rest <= Fraction(1, 2 ** (7 * keep)) * (1 + Fraction(1, 128)), (x, keep)


@pytest.mark.parametrize("slices", [7, 8])
def test_truncated_product_error_bound(slices):
    r = np.random.default_rng(9)
    xs = r.uniform(-1, 1, 120)
    ys = r.uniform(-1, 1, 120)
    worst = Fraction(0)
    for x, y in zip(xs, ys):
        dx, dy = digits_two_halves(float(x), 8), digits_two_halves(float(y), 8)  # prepared with 8, used with `slices`
        partial = sum(Fraction(dx[p] * dy[q], 2 ** (12 + 7 * (p + q)))
                      for p in range(slices) for q in range(slices) if p + q < slices)
        worst = max(worst, abs(Fraction(float(x)) * Fraction(float(y)) - partial))
    assert worst <= (slices + 2) * Fraction(1, 2 ** (7 * slices))


def test_int32_accumulators_cannot_overflow():
    # sliced_shape_ok: s * K * 4096 < 2^31  with |d d'| <= 64 * 64 and at most s pairs of K products per diagonal
    for slices, k in ((7, 65536), (8, 65535), (8, 32768), (10, 52428)):
        assert slices * k * 4096 < 2 ** 31
        assert slices * k * 64 * 64 <= 2 ** 31 - 1
    assert 8 * 65536 * 4096 >= 2 ** 31       # the first K that is refused for 8 slices


def test_guard_threshold_implies_the_contract():
    """acc_ij = sum_k floor(64|x|) floor(64|y|) >= T  =>  K (s + 2) 2^(-7s) <= 0.9e-12 sum_k |x y|  (floor digits are lower
    bounds of 64 |x|), which with the product bound above gives |C - ref| <= 1e-12 sum_k |a b| after the exact scaling."""
    for slices in (7, 8, 10):
        for k in (256, 8192, 32768):
            t = int(np.ceil(k * (slices + 2) * 2.0 ** (12 - 7 * slices) / 0.9e-12))
            sum_xy_lower = Fraction(t, 4096)             # acc 2^-12 <= sum_k |x||y|
            assert Fraction(k * (slices + 2), 2 ** (7 * slices)) <= Fraction(9, 10 ** 13) * sum_xy_lower * (1 + Fraction(1, 10 ** 9))
