// F_{p^2}, p = 2^61 - 1, i^2 = -1 — device arithmetic for sm_100a.
// Replaces virgo::fieldElement (reference src/fieldElement.hpp:15-16,96-97; fieldElement.cpp:34-104,336-360).
// Every result is canonical (both limbs in [0,p)), exactly like the reference's operators, so equality is a
// raw limb compare and any exact formula reproduces the reference bits.
//
// B200 notes: there is no 64-bit integer multiplier; mul.lo/mul.hi.u64 lower to IMAD.WIDE.U32 chains on the
// FMA-heavy pipe, reductions (shift/and/add/select) go to the ALU pipe.  See the note above fmul.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hb {

typedef unsigned long long u64;
struct __align__(16) F { u64 re, im; };

static constexpr u64 P61 = 2305843009213693951ULL;

__host__ __device__ __forceinline__ F mkF(u64 re, u64 im = 0) { F r; r.re = re; r.im = im; return r; }
// read-only (non-coherent) 16-byte load
__device__ __forceinline__ F ldgF(const F *p) { ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p)); return mkF(v.x, v.y); }

// x < 2^64 -> x mod p in [0, p]  then canonicalised
__host__ __device__ __forceinline__ u64 fold61(u64 x) { return (x & P61) + (x >> 61); }
__host__ __device__ __forceinline__ u64 canon61(u64 x) { return x >= P61 ? x - P61 : x; }

__host__ __device__ __forceinline__ u64 add61(u64 a, u64 b) { return canon61(a + b); }
__host__ __device__ __forceinline__ u64 sub61(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    // a - b, plus p when it borrowed: the borrow becomes an all-ones mask, p & mask is added with one carry chain (6 instructions)
    uint32_t lo, hi, m;
    asm("{\n\t.reg .u32 al, ah, bl, bh;\n\tmov.b64 {al, ah}, %3;\n\tmov.b64 {bl, bh}, %4;\n\t"
        "sub.cc.u32 %0, al, bl;\n\tsubc.cc.u32 %1, ah, bh;\n\tsubc.u32 %2, 0, 0;\n\t}"
        : "=r"(lo), "=r"(hi), "=r"(m) : "l"(a), "l"(b));
    uint32_t rl, rh;
    asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;" : "=r"(rl), "=r"(rh) : "r"(lo), "r"(hi), "r"(m), "r"(m & 0x1fffffffu));
    return ((u64)rh << 32) | rl;
#else
    u64 d = a - b; return (a < b) ? d + P61 : d;
#endif
}

// 64x64 -> 128 product, folded with 2^61 == 1: returns a value congruent to a*b, < 2^63 + 2^61 for a,b < 2^62.
__device__ __forceinline__ u64 mul61_lazy(u64 a, u64 b) {
    u64 lo = a * b, hi = __umul64hi(a, b);
    return ((hi << 3) | (lo >> 61)) + (lo & P61);
}
__device__ __forceinline__ u64 mul61(u64 a, u64 b) { return canon61(fold61(mul61_lazy(a, b))); }

// 61-bit x 32-bit product (expander weights are 31-bit reals): a*w < 2^93
__device__ __forceinline__ void mul61x32_wide(u64 a, uint32_t w, u64 &lo, u64 &hi) {
    u64 p0 = (u64)(uint32_t)a * w;            // < 2^64
    u64 p1 = (a >> 32) * (u64)w;              // < 2^61
    lo = p0 + (p1 << 32);
    hi = (p1 >> 32) + (lo < p0);
}
// 128-bit (hi:lo) -> canonical residue; requires hi < 2^58 so that hi<<3 does not overflow
__host__ __device__ __forceinline__ u64 red128(u64 lo, u64 hi) {
    u64 t = ((hi << 3) | (lo >> 61)) + (lo & P61);      // < 2^62
    return canon61(fold61(t));
}

__host__ __device__ __forceinline__ F fadd(F a, F b) { return mkF(add61(a.re, b.re), add61(a.im, b.im)); }
__host__ __device__ __forceinline__ F fsub(F a, F b) { return mkF(sub61(a.re, b.re), sub61(a.im, b.im)); }
__host__ __device__ __forceinline__ F fneg(F a) { return mkF(a.re ? P61 - a.re : 0, a.im ? P61 - a.im : 0); }
__host__ __device__ __forceinline__ bool feq(F a, F b) { return a.re == b.re && a.im == b.im; }
__host__ __device__ __forceinline__ bool fzero(F a) { return (a.re | a.im) == 0; }

// ---- F_{p^2} multiply for sm_100a ------------------------------------------------------------------------------------------------
// B200 has no 64-bit multiplier: the only wide product is IMAD.WIDE.U32 (32x32 + 64-bit addend, FMA-heavy pipe, ~0.8-1 warp-inst/clk/SM),
// everything else (shifts, masks, carries, selects) issues on the ALU pipe at 2 warp-inst/clk/SM.  The multiply is therefore written so
// that nearly all additions ride on the free 64-bit addend of IMAD.WIDE: operands are split as x = x1*2^31 + x0 (x0 < 2^31, x1 <= 2^30),
// and with 2^61 == 1 (mod p), i.e. 2^62 == 2,
//     a*c + e*d  ==  (a0 c0 + e0 d0 + a1 (2 c1) + e1 (2 d1))  +  2^31 * (a1 c0 + a0 c1 + e1 d0 + e0 d1)          (mod p)
// The middle sum M < 2^63 is one chain of four IMAD.WIDE; 2^31 M == (M mod 2^30) 2^31 + (M >> 30) is one more IMAD.WIDE whose addend
// is the shifted-out part; the outer sum is a second chain of four that starts from it and stays below 2^64.  One fold (4 ALU ops)
// gives a value <= p + 7.  Both limbs of a product are such dot products: re = a.re b.re + a.im (p - b.im), im = a.re b.im + a.im b.re:
// 18 IMAD.WIDE + ~37 ALU instructions per canonical product, against 12 + ~80 for Karatsuba on 64x64 products (tools/ubench_fmul.cu).
__device__ __forceinline__ u64 madwide(uint32_t a, uint32_t b, u64 c) { u64 d; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mulwide(uint32_t a, uint32_t b) { u64 d; asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b)); return d; }

// a right-hand operand prepared once and reused (a sumcheck challenge, a twiddle): limbs of re, im and p - im, high limbs also doubled
struct FN { uint32_t r0, r1, r1d, i0, i1, i1d, n0, n1, n1d; };
__device__ __forceinline__ FN fprep(F b) {                       // b canonical
    FN q;
    q.r0 = (uint32_t)b.re & 0x7fffffffu; q.r1 = (uint32_t)(b.re >> 31); q.r1d = q.r1 << 1;
    q.i0 = (uint32_t)b.im & 0x7fffffffu; q.i1 = (uint32_t)(b.im >> 31); q.i1d = q.i1 << 1;
    // p - b.im == -b.im (mod p), in [1, p]: the limbs of p are all ones (2^30 - 1 and 2^31 - 1), so the subtraction never borrows
    q.n0 = q.i0 ^ 0x7fffffffu; q.n1 = q.i1 ^ 0x3fffffffu; q.n1d = q.i1d ^ 0x7ffffffeu;
    return q;
}
// (a*c + e*d) mod p, result <= p + 7.  a, e given as (x0, x1) limbs of values <= p + 7; c, d as (y0, y1, 2 y1) of values <= p.
// the unreduced sum: < 2^63 + 2^62 + 2^61 (two products < (2^31-1)^2, two <= 2^30 (2^31-2), u < 2^61 + 2^33), so one more canonical
// value can be added to it without leaving 64 bits
__device__ __forceinline__ u64 dot61_raw(uint32_t a0, uint32_t a1, uint32_t e0, uint32_t e1,
                                         uint32_t c0, uint32_t c1, uint32_t c1d, uint32_t d0, uint32_t d1, uint32_t d1d) {
    const u64 mid = madwide(e0, d1, madwide(e1, d0, madwide(a0, c1, mulwide(a1, c0))));                 // < 2^63
    const u64 u = madwide((uint32_t)mid & 0x3fffffffu, 0x80000000u, mid >> 30);                         // == 2^31 mid, < 2^61 + 2^33
    return madwide(e1, d1d, madwide(a1, c1d, madwide(e0, d0, madwide(a0, c0, u))));                     // < 2^64 - 2^61
}
__device__ __forceinline__ u64 dot61_lazy(uint32_t a0, uint32_t a1, uint32_t e0, uint32_t e1,
                                          uint32_t c0, uint32_t c1, uint32_t c1d, uint32_t d0, uint32_t d1, uint32_t d1d) {
    const u64 t = dot61_raw(a0, a1, e0, e1, c0, c1, c1d, d0, d1, d1d);
    return (t & P61) + (t >> 61);
}
// a (limbs <= p + 7, e.g. a previous lazy product) times a prepared operand; limbs of the result <= p + 7, NOT canonical
__device__ __forceinline__ F fmul_n_lazy(F a, const FN &b) {
    const uint32_t x0 = (uint32_t)a.re & 0x7fffffffu, x1 = (uint32_t)(a.re >> 31), y0 = (uint32_t)a.im & 0x7fffffffu, y1 = (uint32_t)(a.im >> 31);
    return mkF(dot61_lazy(x0, x1, y0, y1, b.r0, b.r1, b.r1d, b.n0, b.n1, b.n1d),
               dot61_lazy(x0, x1, y0, y1, b.i0, b.i1, b.i1d, b.r0, b.r1, b.r1d));
}
// the same product with both limbs left as raw 64-bit sums (< 2^64 - 2^61): for callers that add something before the one fold
__device__ __forceinline__ F fmul_n_raw(F a, const FN &b) {
    const uint32_t x0 = (uint32_t)a.re & 0x7fffffffu, x1 = (uint32_t)(a.re >> 31), y0 = (uint32_t)a.im & 0x7fffffffu, y1 = (uint32_t)(a.im >> 31);
    return mkF(dot61_raw(x0, x1, y0, y1, b.r0, b.r1, b.r1d, b.n0, b.n1, b.n1d),
               dot61_raw(x0, x1, y0, y1, b.i0, b.i1, b.i1d, b.r0, b.r1, b.r1d));
}
__device__ __forceinline__ F fcanon(F a) { return mkF(canon61(a.re), canon61(a.im)); }
__device__ __forceinline__ F fmul_n(F a, const FN &b) { return fcanon(fmul_n_lazy(a, b)); }
// canonical in, canonical out (reference fieldElement.cpp:49-78 computes the same value)
__device__ __forceinline__ F fmul(F a, F b) { return fmul_n(a, fprep(b)); }
// lazy helpers for accumulation loops: limbs stay below 2^64, one fold per few additions
__device__ __forceinline__ F lfold(F a) { return mkF(fold61(a.re), fold61(a.im)); }                       // any u64 limbs -> limbs <= p + 7
__device__ __forceinline__ void lacc(F &acc, F v) { acc.re += v.re; acc.im += v.im; }                       // caller keeps the sum below 2^64
__device__ __forceinline__ F fcanon2(F a) { return fcanon(lfold(a)); }                                      // limbs < 2^62 -> canonical
// a * real scalar s (s canonical)
__device__ __forceinline__ F fmul_real(F a, u64 s) { return mkF(mul61(a.re, s), mul61(a.im, s)); }

// host twins (host-side control code: MiMC chain, twiddle tables, a handful of scalar combinations per round).
// MiMC is 322 dependent multiplications per hash and sits on the critical path between sumcheck rounds, so the host
// multiply does one 128-bit reduction per output limb: re = ac + (p^2 - bd), im = ad + bc.
static inline u64 h_red128(unsigned __int128 x) {          // x < 2^124 -> canonical
    u64 s = ((u64)x & P61) + ((u64)(x >> 61) & P61) + (u64)(x >> 122);
    return canon61(fold61(s));
}
static inline u64 h_mul61(u64 a, u64 b) { return h_red128((unsigned __int128)a * b); }
static inline F h_fmul(F a, F b) {
    typedef unsigned __int128 u128;
    const u128 p2 = (u128)P61 * P61;
    u128 ac = (u128)a.re * b.re, bd = (u128)a.im * b.im, ad = (u128)a.re * b.im, bc = (u128)a.im * b.re;
    return mkF(h_red128(ac + (p2 - bd)), h_red128(ad + bc));
}
}  // namespace hb
