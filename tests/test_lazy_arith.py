"""The 64-bit headroom arguments of the lazy-reduction kernels, checked on the CPU with Python integers.

The GPU parity tests compare whole kernels with the oracle on random inputs, which essentially never produce limbs at the edges of the
lazy ranges (probability ~2^-58 per value).  These tests restate the lazy formulas of csrc/ntt.cu (dft8_lazy, lmul_w8, lrot30), csrc/field.cuh
(dot61_raw + a canonical addend: the fused fold of the sumcheck round kernel) and csrc/encode.cu (acc_mac4) operation by operation, run them
with every intermediate wrapped to 64 bits exactly as the hardware would, feed them limbs AT the proven bounds, and compare with exact
arithmetic mod p.  An overflow anywhere shows up as a wrong residue."""
import itertools
import random

import pytest

P = (1 << 61) - 1
M64 = (1 << 64) - 1


def w(x):                      # what a 64-bit register holds
    assert x >= 0, "a lazy subtraction went negative: K p was smaller than the subtrahend"
    return x & M64


def fold(x):
    return w((x & P) + (x >> 61))


def lsub(a, b, K):
    assert b <= K * P, "subtrahend above K p"
    return w(a + (K * P - b))


def lrot30(x):
    return w((x >> 31) + ((x & 0x7FFFFFFF) << 30))


class Z:                       # an element of F_p[i] held lazily: (re, im) as 64-bit words
    def __init__(self, re, im):
        self.re, self.im = re, im

    def val(self):
        return (self.re % P, self.im % P)


def ladd(a, b):
    return Z(w(a.re + b.re), w(a.im + b.im))


def lsubz(a, b, K):
    return Z(lsub(a.re, b.re, K), lsub(a.im, b.im, K))


def lsub_j(a, b, K, jn):       # (a - b) * (+i) or (a - b) * (-i)
    return Z(lsub(a.im, b.im, K), lsub(b.re, a.re, K)) if jn else Z(lsub(b.im, a.im, K), lsub(a.re, b.re, K))


def lfold(a):
    return Z(fold(a.re), fold(a.im))


def lmul_w8(x, f):             # x folded (limbs <= p + 7)
    assert x.re <= P + 7 and x.im <= P + 7
    apb, amb, bma = w(x.re + x.im), lsub(x.re, x.im, 2), lsub(x.im, x.re, 2)
    napb = lsub(0, apb, 4)
    re, im = {0: (amb, apb), 2: (apb, bma), 1: (napb, amb), 3: (bma, napb)}[f]
    return Z(lrot30(re), lrot30(im))


def dft8_lazy(x, f):
    jn = (f & 1) != ((f >> 1) & 1)
    a0, a1, a2, a3 = ladd(x[0], x[1]), lsubz(x[0], x[1], 2), ladd(x[2], x[3]), lsub_j(x[2], x[3], 2, jn)
    a4, a5, a6, a7 = ladd(x[4], x[5]), lsubz(x[4], x[5], 2), ladd(x[6], x[7]), lsub_j(x[6], x[7], 2, jn)
    b0, b2, b1, b3 = lfold(ladd(a0, a2)), lfold(lsubz(a0, a2, 4)), lfold(ladd(a1, a3)), lfold(lsubz(a1, a3, 4))
    b4, b6 = lfold(ladd(a4, a6)), lfold(lsub_j(a4, a6, 4, jn))
    b5, b7 = lmul_w8(lfold(ladd(a5, a7)), f), lmul_w8(lfold(lsub_j(a5, a7, 4, jn)), f)
    out = [None] * 8
    out[0], out[4] = lfold(ladd(b0, b4)), lfold(lsubz(b0, b4, 2))
    out[1], out[5] = lfold(ladd(b1, b5)), lfold(lsubz(b1, b5, 2))
    out[2], out[6] = lfold(ladd(b2, b6)), lfold(lsubz(b2, b6, 2))
    out[3], out[7] = lfold(ladd(b3, b7)), lfold(lsubz(b3, b7, 2))
    return out


# exact reference: the same butterfly in F_p[i]
def cmul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def cadd(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def csub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def dft8_exact(x, f):
    sr, si = (-1 if f & 1 else 1), (-1 if f & 2 else 1)
    w8 = ((sr << 30) % P, (si << 30) % P)
    j = cmul(w8, w8)
    assert j in ((0, 1), (0, P - 1))
    a0, a1, a2, a3 = cadd(x[0], x[1]), csub(x[0], x[1]), cadd(x[2], x[3]), cmul(csub(x[2], x[3]), j)
    a4, a5, a6, a7 = cadd(x[4], x[5]), csub(x[4], x[5]), cadd(x[6], x[7]), cmul(csub(x[6], x[7]), j)
    b0, b2, b1, b3 = cadd(a0, a2), csub(a0, a2), cadd(a1, a3), csub(a1, a3)
    b4, b6, b5, b7 = cadd(a4, a6), cmul(csub(a4, a6), j), cmul(cadd(a5, a7), w8), cmul(cmul(csub(a5, a7), j), w8)
    return [cadd(b0, b4), cadd(b1, b5), cadd(b2, b6), cadd(b3, b7), csub(b0, b4), csub(b1, b5), csub(b2, b6), csub(b3, b7)]


B0 = P + (1 << 34)             # bound of every butterfly input (folded value, lazy product or W8 output)
EDGE = [0, 1, P - 1, P, P + 7, (1 << 61) + (1 << 33), B0]


@pytest.mark.parametrize("f", [0, 1, 2, 3])
def test_dft8_lazy_at_the_bounds(f):
    rng = random.Random(f)
    cases = [[Z(B0, B0)] * 8, [Z(0, 0)] * 8, [Z(B0, 0), Z(0, B0)] * 4, [Z(0, B0), Z(B0, 0)] * 4]
    for _ in range(300):
        cases.append([Z(rng.choice(EDGE), rng.choice(EDGE)) for _ in range(8)])
    for _ in range(100):
        cases.append([Z(rng.randrange(B0 + 1), rng.randrange(B0 + 1)) for _ in range(8)])
    for x in cases:
        got = dft8_lazy(x, f)
        want = dft8_exact([z.val() for z in x], f)
        for g, e in zip(got, want):
            assert g.re <= P + 7 and g.im <= P + 7            # folded: what shared memory may hold
            assert g.val() == e


@pytest.mark.parametrize("jn", [False, True])
def test_radix4_and_radix2_passes_at_the_bounds(jn):
    """The radix-4 / radix-2 passes of ntt_tile_lazy_kernel: x0 folded (<= p + 7), X1..X3 lazy products (<= p + 7)."""
    rng = random.Random(7 + jn)
    j = (0, P - 1) if jn else (0, 1)
    edge = [0, 1, P - 1, P, P + 7]
    for _ in range(2000):
        x0, X1, X2, X3 = (Z(rng.choice(edge + [rng.randrange(P + 8)]), rng.choice(edge + [rng.randrange(P + 8)])) for _ in range(4))
        a0, a1, b, c = ladd(x0, X1), lsubz(x0, X1, 2), ladd(X2, X3), lsub_j(X2, X3, 2, jn)
        got = [lfold(ladd(a0, b)), lfold(ladd(a1, c)), lfold(lsubz(a0, b, 4)), lfold(lsubz(a1, c, 4))]
        e0, e1, eb, ec = cadd(x0.val(), X1.val()), csub(x0.val(), X1.val()), cadd(X2.val(), X3.val()), cmul(csub(X2.val(), X3.val()), j)
        want = [cadd(e0, eb), cadd(e1, ec), csub(e0, eb), csub(e1, ec)]
        for g, e in zip(got, want):
            assert g.re <= P + 7 and g.im <= P + 7 and g.val() == e
        u, v = x0, X1                                                                    # radix-2: s[p0] = u + v, s[p1] = u - v
        assert lfold(ladd(u, v)).val() == cadd(u.val(), v.val()) and lfold(lsubz(u, v, 2)).val() == csub(u.val(), v.val())


def test_first_pass_operands_stay_folded():
    """Stages 1-4 in registers (zero-extended rows): canonical inputs through lmul_j (p - limb) and lmul_w8 must be valid butterfly inputs."""
    rng = random.Random(9)
    for f in range(4):
        jn = (f & 1) != ((f >> 1) & 1)
        sr, si = (-1 if f & 1 else 1), (-1 if f & 2 else 1)
        w8 = ((sr << 30) % P, (si << 30) % P)
        j = cmul(w8, w8)
        for _ in range(500):
            x = Z(rng.choice([0, 1, P - 1, rng.randrange(P)]), rng.choice([0, 1, P - 1, rng.randrange(P)]))
            mj = Z(x.im, P - x.re) if jn else Z(P - x.im, x.re)                           # lmul_j of a canonical value: limbs in [0, p]
            assert mj.re <= P and mj.im <= P and mj.val() == cmul(x.val(), j)
            for y in (x, mj):
                r = lmul_w8(y, f)
                assert r.re <= B0 and r.im <= B0 and r.val() == cmul(y.val(), w8)


def test_lrot30_any_64_bit_value():
    rng = random.Random(1)
    for x in [0, 1, P, P + 1, M64, M64 - 1, 1 << 63, (1 << 31) - 1, 1 << 31] + [rng.getrandbits(64) for _ in range(2000)]:
        r = lrot30(x)
        assert r <= (1 << 61) + (1 << 33) and r % P == (x << 30) % P


# ---- field.cuh: dot61_raw and the fused fold of the sumcheck round kernel ---------------------------------------------------------------
def split(x):
    return x & 0x7FFFFFFF, x >> 31


def dot61_raw(a, e, c, d):      # a c + e d with a, e lazy (<= p + 7) and c, d canonical; every IMAD.WIDE result wrapped to 64 bits
    a0, a1 = split(a); e0, e1 = split(e); c0, c1 = split(c); d0, d1 = split(d)
    assert max(a0, a1, e0, e1, c0, c1, d0, d1, 2 * c1, 2 * d1) < (1 << 32)
    mid = w(a1 * c0); mid = w(a0 * c1 + mid); mid = w(e1 * d0 + mid); mid = w(e0 * d1 + mid)
    assert mid == a1 * c0 + a0 * c1 + e1 * d0 + e0 * d1
    u = w((mid & 0x3FFFFFFF) * 0x80000000 + (mid >> 30))
    t = w(a0 * c0 + u); t = w(e0 * d0 + t); t = w(a1 * (2 * c1) + t); t = w(e1 * (2 * d1) + t)
    assert t == a0 * c0 + e0 * d0 + 2 * a1 * c1 + 2 * e1 * d1 + (mid & 0x3FFFFFFF) * 0x80000000 + (mid >> 30), "a chain overflowed 64 bits"
    return t


def test_dot61_raw_leaves_room_for_one_canonical_addend():
    rng = random.Random(2)
    lazy_edge, canon_edge = [0, 1, P - 1, P, P + 7, (1 << 61) - 2, (1 << 31) - 1, 1 << 31], [0, 1, P - 1, P - 2, (1 << 31) - 1, 1 << 31, (1 << 60)]
    cases = list(itertools.product([P + 7, P, 0], [P + 7, 0], [P - 1, 0], [P - 1, 0]))
    cases += [(rng.choice(lazy_edge), rng.choice(lazy_edge), rng.choice(canon_edge), rng.choice(canon_edge)) for _ in range(2000)]
    cases += [(rng.randrange(P + 8), rng.randrange(P + 8), rng.randrange(P), rng.randrange(P)) for _ in range(2000)]
    for a, e, c, d in cases:
        t = dot61_raw(a, e, c, d)
        assert t % P == (a * c + e * d) % P
        x = P - 1                                              # the table entry the fold adds BEFORE the single reduction
        assert t + x <= M64, "dot61_raw + canonical addend overflows"
        v = fold(w(t + x))
        assert v <= P + 7 and v % P == (a * c + e * d + x) % P


# ---- encode.cu: acc_mac4 --------------------------------------------------------------------------------------------------------------
def test_acc_mac4_partial_sums_fit_64_bits():
    wmax, xmax = (1 << 31) - 1, P - 1                          # weights < 2^31 (w31), canonical entries
    lo, hi = xmax & 0xFFFFFFFF, xmax >> 32
    assert 2 * (0xFFFFFFFF * wmax) <= M64                      # a pair of low-half products on one IMAD.WIDE addend
    assert 4 * (hi * wmax) <= M64 and hi < (1 << 29)           # all four high-half products of a quad
    rng = random.Random(3)
    for _ in range(500):
        xs = [rng.choice([0, 1, P - 1, rng.randrange(P)]) for _ in range(4)]
        ws = [rng.choice([0, 1, wmax, rng.randrange(1 << 31)]) for _ in range(4)]
        ua = w((xs[1] & 0xFFFFFFFF) * ws[1] + w((xs[0] & 0xFFFFFFFF) * ws[0]))
        ub = w((xs[3] & 0xFFFFFFFF) * ws[3] + w((xs[2] & 0xFFFFFFFF) * ws[2]))
        v = 0
        for x, wt in zip(xs, ws):
            v = w((x >> 32) * wt + v)
        assert (ua + ub + (v << 32)) % P == sum(x * wt for x, wt in zip(xs, ws)) % P


def test_row_accumulators_reduced_once():
    """encode.cu acc_reduce: U + V 2^32 (two 96-bit accumulators) assembled as one 128-bit number and reduced once by red128, whose upper half
    must stay below 2^58 — true for any in-degree below 2^29 (hb_expander_set rejects larger ones)."""
    M32 = (1 << 32) - 1

    def red128(lo, hi):
        assert hi < (1 << 58)
        t = w(w((hi << 3) & M64 | (lo >> 61)) + (lo & P))
        assert t == (hi << 3) + (lo >> 61) + (lo & P)
        f = (t & P) + (t >> 61)
        return f - P if f >= P else f

    rng = random.Random(5)
    for _ in range(20000):
        terms = rng.choice([1, 4, 64, 4096, 1 << 20, (1 << 29) - 1])
        U = rng.choice([terms * M64 >> 0, rng.randrange(terms << 64)]) % (1 << 96)      # <= terms low-half sums of < 2^64 each
        V = rng.choice([terms * ((1 << 61) - 1), rng.randrange(terms << 61)]) % (1 << 96)  # <= terms high-half sums of < 2^61 each
        u, v = [U & M32, (U >> 32) & M32, U >> 64], [V & M32, (V >> 32) & M32, V >> 64]
        t = u[1] + v[0]; s1, c = t & M32, t >> 32
        t = u[2] + v[1] + c; s2, c = t & M32, t >> 32
        s3 = v[2] + c
        assert s3 <= M32
        assert red128((s1 << 32) | u[0], (s3 << 32) | s2) == (U + (V << 32)) % P
