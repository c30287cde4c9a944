// F_{p^2}, p = 2^61 - 1, i^2 = -1 — device arithmetic for sm_100a.
// Replaces virgo::fieldElement (reference src/fieldElement.hpp:15-16,96-97; fieldElement.cpp:34-104,336-360).
// Every result is canonical (both limbs in [0,p)), exactly like the reference's operators, so equality is a
// raw limb compare and any exact formula reproduces the reference bits.
//
// B200 notes: there is no 64-bit integer multiplier; mul.lo/mul.hi.u64 lower to IMAD.WIDE.U32 chains on the
// FMA-heavy pipe, reductions (shift/and/add/select) go to the ALU pipe.  The code below keeps the number of
// 64x64 products minimal (3 per F mul, Karatsuba) and reduces lazily (one fold per output limb).
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
__host__ __device__ __forceinline__ u64 sub61(u64 a, u64 b) { u64 d = a - b; return (a < b) ? d + P61 : d; }

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

// (a.re + i a.im)(b.re + i b.im): Karatsuba, 3 wide products (reference fieldElement.cpp:49-78)
__device__ __forceinline__ F fmul(F a, F b) {
    u64 ac = fold61(mul61_lazy(a.re, b.re));                       // <= p + 4
    u64 bd = fold61(mul61_lazy(a.im, b.im));
    u64 all = fold61(mul61_lazy(a.re + a.im, b.re + b.im));
    // re = ac - bd ; im = all - ac - bd   (add multiples of p to stay non-negative)
    u64 re = ac + (2 * P61 - bd);                                  // < 3p + 8
    u64 im = all + (4 * P61 - ac - bd);                            // < 5p + 8
    return mkF(canon61(fold61(re)), canon61(fold61(im)));
}
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
