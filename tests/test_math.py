"""CPU: gen_b200/csrc/gsmc_math.h (the IEEE-only transcendentals shared by device and oracle)
against glibc, in ulps."""
import math

import numpy as np
import pytest


def ulps(ref, got):
    ref, got = np.asarray(ref, dtype=np.float64), np.asarray(got, dtype=np.float64)
    return np.abs(got - ref) / np.spacing(np.abs(ref))


def test_exp_log_accuracy(orc):
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-700, 700, 20000), rng.uniform(-2, 2, 20000), [0.0, 1.0, -1.0, 709.0, -708.0]])
    got = np.array([orc.L.orc_exp(float(x)) for x in xs])
    assert ulps(np.exp(xs), got).max() <= 2.0
    ys = np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), 1 + rng.uniform(-1e-3, 1e-3, 20000), [1.0, 2.0, 0.5, 1e-310]])
    got = np.array([orc.L.orc_log(float(y)) for y in ys])
    assert ulps(np.log(ys), got)[np.log(ys) != 0].max() <= 2.5
    assert orc.L.orc_log(1.0) == 0.0 and orc.L.orc_exp(0.0) == 1.0
    assert orc.L.orc_exp(-math.inf) == 0.0 and orc.L.orc_exp(-750.0) == 0.0 and orc.L.orc_exp(800.0) == math.inf
    assert orc.L.orc_log(0.0) == -math.inf and math.isnan(orc.L.orc_log(-1.0)) and math.isnan(orc.L.orc_exp(math.nan))


def test_sincospi_atan2_accuracy(orc):
    rng = np.random.default_rng(1)
    for t in rng.uniform(0, 2, 5000):
        s, c = orc.sincospi(float(t))
        assert abs(s - math.sin(math.pi * t)) < 1e-15 and abs(c - math.cos(math.pi * t)) < 1e-15   # the reference side rounds pi*t too
    assert orc.sincospi(0.0) == (0.0, 1.0)
    assert orc.sincospi(0.5)[0] == 1.0 and orc.sincospi(1.5)[0] == -1.0
    for _ in range(5000):
        y, x = rng.uniform(-30, 30, 2)
        assert ulps(math.atan2(y, x), orc.L.orc_atan2(float(y), float(x))) <= 4.0


def test_fp32_box_muller_pieces_against_glibc(orc):
    """gm_nlog_u32f / gm_sincos_u32f (the fp32 transform behind the per-particle normals, shared by device and oracle)
    against glibc in double: -ln((w + 1/2) 2^-32) to 1e-5 relative over the whole range (next to 1 through the
    complement), always positive; (sin, cos)(2 pi a 2^-32) to 1e-7 absolute."""
    import ctypes as C
    rng = np.random.default_rng(5)
    ws = np.concatenate([rng.integers(0, 2 ** 32, 60000, dtype=np.uint64), rng.integers(2 ** 32 - 2 ** 26, 2 ** 32, 20000, dtype=np.uint64),
                         np.arange(0, 2000, dtype=np.uint64), 2 ** 32 - 1 - np.arange(0, 2000, dtype=np.uint64),
                         np.uint64(1) << np.arange(0, 32, dtype=np.uint64), (np.uint64(1) << np.arange(1, 32, dtype=np.uint64)) - np.uint64(1),
                         np.array([0xfe000000 - 1, 0xfe000000, 0xfe000001], dtype=np.uint64)])
    worst = 0.0
    for w in ws:
        w = int(w)
        got, ref = float(orc.L.orc_nlog_u32f(w)), -math.log1p(-(2 ** 32 - w - 0.5) * 2.0 ** -32)
        assert got > 0.0
        worst = max(worst, abs(got - ref) / ref)
    assert worst < 1e-5, worst
    s, c = C.c_float(), C.c_float()
    worst = 0.0
    j = np.arange(0, 128, dtype=np.uint64) << np.uint64(25)
    for a in np.concatenate([rng.integers(0, 2 ** 32, 60000, dtype=np.uint64), np.arange(0, 1000, dtype=np.uint64), j, j + np.uint64((1 << 24) - 1), j + np.uint64(1 << 24)]):
        a = int(a) & 0xffffffff
        orc.L.orc_sincos_u32f(a, C.byref(s), C.byref(c))
        th = 2 * math.pi * a / 2 ** 32
        worst = max(worst, abs(s.value - math.sin(th)), abs(c.value - math.cos(th)))
    assert worst < 1e-7, worst
    orc.L.orc_sincos_u32f(0, C.byref(s), C.byref(c))
    assert (s.value, c.value) == (0.0, 1.0)
    orc.L.orc_sincos_u32f(1 << 30, C.byref(s), C.byref(c))
    assert (s.value, c.value) == (1.0, 0.0)


def test_box_muller_normals_are_standard(orc):
    z = orc.normals(12345, 7, 0, 400000)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs(np.mean(z ** 3)) < 0.03 and abs(np.mean(z ** 4) - 3) < 0.06
    # the normals are fp32 values widened to fp64; Kolmogorov-Smirnov against the normal CDF (1% critical value) and
    # no correlation between the cos and sin branches or between neighbouring calls
    from scipy import stats
    assert np.array_equal(z, z.astype(np.float32).astype(np.float64))
    assert stats.kstest(z, "norm").statistic < 1.63 / math.sqrt(len(z))
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 0.01 and abs(np.corrcoef(z[0::4], z[2::4])[0, 1]) < 0.01
    assert abs(np.mean(np.abs(z) > 3.0) - 0.0026998) < 0.0005
    assert np.array_equal(orc.normals(12345, 7, 1000, 10), z[1000:1010])          # counter-based: any slice
    u = orc.uniforms(1, 2, 1, 0, 100000)
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01


def test_gamma_gaps_and_grouped_order_statistics_are_exact_in_distribution(orc):
    """Marsaglia-Tsang gaps: Gamma(shape) moments for the shapes the resampler uses (1 = head, partial groups, 256);
    and the thresholds of a whole event are uniform order statistics (Kolmogorov-Smirnov against U(0,1))."""
    for shape, n in ((1, 40000), (7, 20000), (256, 20000)):
        g = np.array([orc.gap_variate(5, 3, j, shape) for j in range(n)], dtype=np.float64) / 2 ** 20
        se = math.sqrt(shape / n)
        assert abs(g.mean() - shape) < 5 * se, (shape, g.mean())
        assert abs(g.var() / shape - 1) < 0.08, (shape, g.var())
        assert g.min() >= 0
    # exponential head: P(g > 1) = e^-1
    g1 = np.array([orc.gap_variate(9, 0, j, 1) for j in range(40000)], dtype=np.float64) / 2 ** 20
    assert abs(np.mean(g1 > 1.0) - math.exp(-1)) < 0.01
    # an event with M draws against a CDF with 2^40 equal steps: anc / n is the uniform itself
    M, n = 50000, 1 << 20
    cdf = (np.arange(1, n + 1, dtype=np.uint64)) << np.uint64(20)
    for rho in (0, 1):
        anc = orc.search_sorted(cdf, 77, rho, M)
        u = np.sort((anc + 0.5) / n)
        d = np.max(np.abs(u - (np.arange(M) + 0.5) / M))
        assert d < 1.63 / math.sqrt(M), d                 # KS 1% critical value
    a0, a1 = orc.search_sorted(cdf, 77, 0, M), orc.search_sorted(cdf, 77, 1, M)
    assert not np.array_equal(a0, a1)


def test_division_by_invariant_is_correctly_rounded(orc):
    """gm_div_inv (Markstein sequence, used for -(diff^2)/(2 var) with a launch-invariant variance)
    must give the bits of the IEEE division the reference formula performs (normal.jl:59)."""
    assert orc.L.orc_div_inv_mismatches(1, 10_000_000) == 0
    assert orc.L.orc_div_inv_mismatches(987654321, 10_000_000) == 0
    for x, c in ((0.0, 2.0), (-0.0, 2.0), (1e-300, 3.0), (-1e300, 7.0), (math.inf, 2.0), (1.0, 1e-200), (1.0, 1e200)):
        got, want = orc.L.orc_div_inv(x, c), x / c
        assert (got == want and math.copysign(1, got) == math.copysign(1, want)) or (math.isnan(got) and math.isnan(want))
    assert math.isnan(orc.L.orc_div_inv(math.nan, 2.0))


def test_log_fast_path_has_the_same_bits(orc):
    rng = np.random.default_rng(3)
    for x in np.concatenate([rng.random(20000), np.exp(rng.uniform(-600, 600, 20000)), [2.0 ** -53 * 0.5, 1.0 - 2.0 ** -53]]):
        assert orc.L.orc_log_pos(float(x)) == orc.L.orc_log(float(x))


def test_exp_nonpos_has_the_same_bits(orc):
    rng = np.random.default_rng(4)
    xs = np.concatenate([-rng.random(20000) * 50, -np.exp(rng.uniform(-30, 6.56, 20000)), [0.0, -0.0, -708.39, -708.4, -745.0, -1e300, -math.inf]])
    for x in xs:
        assert orc.L.orc_exp_nonpos(float(x)) == orc.L.orc_exp(float(x)), x


def test_muldiv_floor_is_exact(orc):
    """gsmc_fixed.h: T_k = floor(S_k C_N / S_tot) against unsigned __int128 division."""
    assert orc.L.orc_muldiv_mismatches(1, 5_000_000) == 0
    assert orc.L.orc_muldiv_mismatches(2024, 5_000_000) == 0


def test_table_log_of_the_box_muller_radius(orc):
    """gm_log_unit on 53-bit uniforms in (0,1): absolute error at the level of one ulp of max(1, |log u|),
    result never positive (the radius sqrt(-2 log u) is always defined)."""
    rng = np.random.default_rng(6)
    us = np.concatenate([
        (rng.integers(0, 2 ** 53, 60000).astype(np.float64) + 0.5) * 2.0 ** -53,
        [0.5 * 2.0 ** -53, 1.5 * 2.0 ** -53, 1.0 - 2.0 ** -54, 1.0 - 3 * 2.0 ** -54, 0.5, 0.25, 0.75, 2.0 ** -30],
        1.0 - (rng.integers(0, 2 ** 20, 2000).astype(np.float64) + 0.5) * 2.0 ** -53,    # next to 1
        (rng.integers(0, 2 ** 20, 2000).astype(np.float64) + 0.5) * 2.0 ** -53])          # next to 0
    worst = 0.0
    for u in us:
        got, ref = orc.L.orc_log_unit(float(u)), math.log(float(u))
        assert got <= 0.0
        worst = max(worst, abs(got - ref) / max(1.0, abs(ref)))
    assert worst < 4e-16, worst


def test_lookup_tables_are_the_generated_ones():
    """gen_b200/csrc/gsmc_tables.h (exp2, sin/cos, and the fp32 log / sin-cos tables of the normal generator) is what
    scripts/gen_math_tables.py generates with mpmath: correctly rounded entries, no hand edits."""
    import os
    import subprocess
    import sys
    pytest.importorskip("mpmath")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "gen_math_tables.py")], capture_output=True, text=True, check=True).stdout
    assert out.strip() == open(os.path.join(root, "gen_b200", "csrc", "gsmc_tables.h")).read().strip()


def test_fp32_box_muller_independent_restatement(orc):
    """The definition of the per-particle normals, restated here with exact rational arithmetic and ONE correct rounding
    to float32 per operation (no C, no shared header, tables recomputed with mpmath), must give the bits of the
    shared gm_box_muller_u32: radius, angle, and whole draws taken from Philox words (element e -> call e >> 2, pair
    (e >> 1) & 1, cos / sin branch e & 1)."""
    import ctypes as C
    import struct
    from fractions import Fraction as Fr
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200

    def rn32(x):
        """exact rational (or mpf) -> nearest float32 (ties to even), as an exact Fraction; normal range only"""
        if not isinstance(x, Fr):                           # an mpmath value (200 bits): far finer than any float32 tie
            x = Fr(int(mp.nint(mp.mpf(x) * mp.mpf(2) ** 300)), 2 ** 300)
        if x == 0:
            return Fr(0)
        s, a = (1 if x > 0 else -1), abs(x)
        e = a.numerator.bit_length() - a.denominator.bit_length()
        if Fr(2) ** e > a:
            e -= 1                                          # 2^e <= a < 2^(e+1)
        assert e >= -126
        scaled = a / Fr(2) ** (e - 23)                     # in [2^23, 2^24)
        n, rem = divmod(scaled.numerator, scaled.denominator)
        twice = 2 * rem
        if twice > scaled.denominator or (twice == scaled.denominator and (n & 1)):
            n += 1
        return s * Fr(n) * Fr(2) ** (e - 23)

    def bits32(x):
        return struct.unpack("<I", struct.pack("<f", float(x)))[0]      # x is exactly a float32: float() is exact

    c = lambda v: Fr(np.float32(v).item())                              # the float32 constant the C source spells
    ln2f, kpi = rn32(mp.log(2)), rn32(mp.pi / 64 / 2 ** 25)
    ltab = []
    for i in range(64):
        ic = rn32(1 / (1 + (mp.mpf(i) + mp.mpf(1) / 2) / 64))
        ltab.append((ic, rn32(mp.log(2 * mp.mpf(ic.numerator) / mp.mpf(ic.denominator)))))
    sctab = [(rn32(mp.sin(mp.pi * j / 64)) if j % 64 else Fr(0), rn32(mp.cos(mp.pi * j / 64)) if (j - 32) % 64 else Fr(0)) for j in range(128)]

    def nlog(w):
        if w >= 0xfe000000:
            y = rn32(rn32(rn32(Fr((~w) & 0xffffffff)) + Fr(1, 2)) * c(2.32830644e-10))
            q = rn32(c(0.25) * y + c(0.333333343))
            q = rn32(q * y + Fr(1, 2))
            q = rn32(q * y + 1)
            return rn32(q * y)
        v = 2 * w + 1
        lz = 33 - v.bit_length()
        t = (v << lz) >> 1
        m = 1 + Fr((t >> 8) & 0x7fffff, 2 ** 23)
        ic, t2 = ltab[(t >> 25) & 63]
        r = rn32(m * ic - 1)
        q = rn32(c(0.333333343) * r - Fr(1, 2))
        q = rn32(q * r + 1)
        base = rn32(lz * ln2f + t2)
        l = rn32(-q * r + base)
        return l if l > 0 else Fr(0)

    def sincos(a):
        b = (a + 0x01000000) & 0xffffffff
        d = (b & 0x01ffffff) - 0x01000000
        x = rn32(Fr(d) * kpi)
        z = rn32(x * x)
        sx = rn32(rn32(x * z) * c(-0.166666672) + x)
        cm = rn32(z * rn32(z * c(0.0416666679) - Fr(1, 2)))
        S, Cc = sctab[b >> 25]
        sn = rn32(S + rn32(S * cm + rn32(Cc * sx)))
        cs = rn32(Cc + rn32(Cc * cm - rn32(S * sx)))
        return sn, cs

    def box(wr, wa):
        l = nlog(wr)
        l2 = rn32(l + l)
        r = rn32(mp.sqrt(mp.mpf(l2.numerator) / mp.mpf(l2.denominator))) if l2 > 0 else Fr(0)
        sn, cs = sincos(wa)
        return float(rn32(r * cs)), float(rn32(r * sn))

    rng = np.random.default_rng(11)
    words = [int(w) for w in rng.integers(0, 2 ** 32, 1500, dtype=np.uint64)] + [0, 1, 2, 2 ** 31, 2 ** 32 - 1, 0xfe000000 - 1, 0xfe000000,
                                                                                   0xffffff00, 2 ** 24, 2 ** 24 - 1, 0x80000001, 0x7fffffff]
    s, cc = C.c_float(), C.c_float()
    for w in words:
        assert bits32(nlog(w)) == bits32(Fr(float(orc.L.orc_nlog_u32f(w)))), hex(w)
        orc.L.orc_sincos_u32f(w, C.byref(s), C.byref(cc))
        sn, cs = sincos(w)
        assert (bits32(sn), bits32(cs)) == (bits32(Fr(s.value)), bits32(Fr(cc.value))), hex(w)
    seed, t = 0x0123456789abcdef, 9
    z = orc.normals(seed, t, 0, 400)
    for call in range(100):
        o = orc.philox([call, 0, t, 0], [seed & 0xffffffff, seed >> 32])
        exp = box(o[0], o[1]) + box(o[2], o[3])
        assert [float(v) for v in z[4 * call: 4 * call + 4]] == list(exp), call
