"""The numerical claim behind k_score_sorted's fast cell index, checked on the CPU with exact rational arithmetic.

The kernel replaces Java's  (int) ((x*c - y*s + px - posx) / res)  (GridMap.java:273-274, Transform.java:13-32) by two
FMAs on per-particle constants plus a magic-number add that leaves round(q * 2^k) in the low mantissa word, and accepts
the result only when the fraction is at least `margin` units of 2^-k away from an integer (otherwise the literal
expression decides).  This test re-creates that arithmetic bit for bit (FMA through fractions.Fraction, everything else
IEEE binary64 as in the kernel) and asserts, for random and for adversarial near-integer inputs and for every grid size
class, that (1) |q~ - q_java| stays far below the margin and (2) an accepted index always equals Java's.
No GPU, no library: pure arithmetic (csrc/kernels.cuh k_score_sorted, csrc/gms.cu Geometry set-up)."""
import math
import struct
from fractions import Fraction

import numpy as np
import pytest


def f32(x):
    return float(np.float32(x))


def fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))  # int/int true division: correctly rounded


def d2i(d):
    if d != d:
        return 0
    return max(-2147483648, min(2147483647, int(d)))


def geometry(W):
    bits = 1
    while (1 << bits) <= W:
        bits += 1
    k = min(20, 30 - bits)
    magic = math.ldexp(1.5, 52 - k)
    hi = struct.unpack("<Q", struct.pack("<d", magic))[0] >> 32
    return k, magic, hi, max(8, 1 << max(0, k - 14))


def fast_index(m, cinv, sinv, pq, k, magic, hi, marg, sign):
    """One coordinate of the kernel's fast path -> (accepted, cell)."""
    q = fma(m[0], cinv, fma(-m[1], sinv, pq)) if sign > 0 else fma(m[0], sinv, fma(m[1], cinv, pq))
    t = q + magic
    bits = struct.unpack("<Q", struct.pack("<d", t))[0]
    i, thi = bits & 0xFFFFFFFF, bits >> 32
    fr = i & ((1 << k) - 1)
    ok = thi == hi and marg <= fr < (1 << k) - marg
    return ok, i >> k, q


@pytest.mark.parametrize("W", [120, 400, 1024, 2048, 4096, 16384])
def test_accepted_fast_index_equals_java(W):
    rng = np.random.default_rng(W)
    res = f32(0.05)
    inv_res = 1.0 / res
    posx = f32(-W * 0.05 / 2)
    k, magic, hi, marg = geometry(W)
    worst, accepted, rejected = 0.0, 0, 0
    for n in range(3000):
        theta = f32(rng.uniform(-math.pi, math.pi))
        c, s = f32(math.cos(theta)), f32(math.sin(theta))
        px = f32(rng.uniform(-0.1, 0.1) * W * 0.05)
        reach = min(30.0, 0.6 * W * 0.05)  # some end points leave the map on purpose
        m = [float(rng.uniform(-reach, reach)), float(rng.uniform(-reach, reach))]
        for sign in (+1, -1):  # +1: the x row of the transform, -1: the y row
            def java_q(mm):
                t = mm[0] * c - mm[1] * s + px if sign > 0 else mm[0] * s + mm[1] * c + px
                return (t - posx) / res
            if n % 2 and abs(c if sign > 0 else s) > 0.2:  # adversarial: steer q next to an integer
                target = round(java_q(m)) + float(rng.choice([0.0, 1e-13, -1e-13, 1e-10, -1e-10, 1e-7, -1e-7, 5e-5, -5e-5,
                                                               7e-5, -7e-5]))
                m[0] += (target - java_q(m)) * res / (c if sign > 0 else s)
            qj = java_q(m)
            cinv, sinv, pq = c * inv_res, s * inv_res, (px - posx) * inv_res
            ok, cell, q = fast_index(m, cinv, sinv, pq, k, magic, hi, marg, sign)
            worst = max(worst, abs(q - qj))
            if ok:
                accepted += 1
                assert cell == d2i(qj), (W, theta, px, m, q, qj)
            else:
                rejected += 1
    # the error of the FMA form is rounding noise; the acceptance margin is >= 8 * 2^-k (6e-5 cells for W <= 8192)
    assert worst < 1e-9, worst
    assert worst * 1000 < marg / (1 << k)
    assert accepted > 2000 and rejected > 200, (accepted, rejected)  # both branches exercised


def fast_index_v2(m, cinv, sinv, pq, k, magic, marg, sign, lp):
    """One coordinate of k_score_sorted<G, 2>: the magic constant is folded into the per-particle offset, the
    fraction test runs on (i + margin) * 2^(32-k) mod 2^32, the range test is one compare of the low word."""
    pqm = pq + magic  # rounded to a multiple of 2^-k once per particle
    t = fma(m[0], cinv, fma(-m[1], sinv, pqm)) if sign > 0 else fma(m[0], sinv, fma(m[1], cinv, pqm))
    i = struct.unpack("<Q", struct.pack("<d", t))[0] & 0xFFFFFFFF
    u = (i * (1 << (32 - k)) + (marg << (32 - k))) & 0xFFFFFFFF
    ok = u >= ((2 * marg) << (32 - k)) and i < (1 << (k + lp))
    return ok, i >> k


@pytest.mark.parametrize("W", [120, 400, 1024, 2048, 4096, 16384])
def test_accepted_fast_index_v2_equals_java(W):
    """The ALU-lean variant: an accepted index equals Java's truncation, inside or outside the map (cells of the
    padded square beyond the map hold 1.0), for every pose the per-particle range guard lets through."""
    rng = np.random.default_rng(1000 + W)
    res = f32(0.05)
    inv_res = 1.0 / res
    posx = f32(-W * 0.05 / 2)
    k, magic, hi, marg = geometry(W)
    lp = max(1, (W - 1).bit_length())
    lim = float(1 << (31 - k)) - 2.0
    accepted = rejected = guarded = 0
    for n in range(3000):
        theta = f32(rng.uniform(-math.pi, math.pi))
        c, s = f32(math.cos(theta)), f32(math.sin(theta))
        px = f32(rng.uniform(-0.7, 0.7) * W * 0.05)  # poses well outside the map too
        reach = min(30.0, 0.6 * W * 0.05)
        m = [float(rng.uniform(-reach, reach)), float(rng.uniform(-reach, reach))]
        for sign in (+1, -1):
            def java_q(mm):
                t = mm[0] * c - mm[1] * s + px if sign > 0 else mm[0] * s + mm[1] * c + px
                return (t - posx) / res
            if n % 2 and abs(c if sign > 0 else s) > 0.2:
                target = round(java_q(m)) + float(rng.choice([0.0, 1e-13, -1e-13, 1e-10, -1e-10, 1e-7, -1e-7, 5e-5, -5e-5,
                                                               7e-5, -7e-5, 2e-5, -2e-5]))
                m[0] += (target - java_q(m)) * res / (c if sign > 0 else s)
            qj = java_q(m)
            cinv, sinv, pq = c * inv_res, s * inv_res, (px - posx) * inv_res
            R = math.sqrt(m[0] * m[0] + m[1] * m[1]) * inv_res * 1.000001 + 2.0
            if not (abs(pq) + R < lim):
                guarded += 1  # the kernel takes the exact path for every beam of such a particle
                continue
            ok, cell = fast_index_v2(m, cinv, sinv, pq, k, magic, marg, sign, lp)
            if ok:
                accepted += 1
                assert cell == d2i(qj) and 0 <= qj, (W, theta, px, m, qj)
                assert cell < (1 << lp)
            else:
                rejected += 1
    assert accepted > 1500 and rejected > 200, (accepted, rejected, guarded)


def test_guarded_reciprocal_cell_of():
    """cell_of() of csrc/device_math.cuh (k_score, per-particle maps): trunc(t * (1/res)) is used only when it lies
    more than 1e-5 from both neighbouring integers; then it must equal Java's (int) (t / res)."""
    rng = np.random.default_rng(77)
    for res in (f32(0.05), f32(0.025), f32(0.1), f32(0.03)):
        inv = 1.0 / res
        fast = slow = 0
        for n in range(20000):
            t = float(rng.uniform(-20, 250))
            if n % 2:  # next to a cell boundary, on either side, down to a few ulps
                cell = float(rng.integers(-50, 5000))
                t = cell * res + float(rng.choice([0.0, 1e-16, -1e-16, 1e-13, -1e-13, 1e-9, -1e-9, 4e-7, -4e-7, 6e-7, -6e-7]))
            q = t * inv
            nq = int(q)
            fr = abs(q - nq)
            if 1e-5 < fr < 1.0 - 1e-5:
                fast += 1
                assert nq == d2i(t / res), (res, t)
            else:
                slow += 1
        assert fast > 9000 and slow > 2000, (fast, slow)
