// N1 — batched NTT over F_{p^2}, bit-identical to the reference's `_fft(arr, logn, false)`
// (src/utils.cpp:605-673): bit-reversal permutation, twiddles w[k] = omega^k with
// omega = getRootOfUnity(logn) (utils.cpp:452-463), radix-2 DIT stages
//     u = a[j+k]; v = a[j+k+i/2] * w[len/i*k]; a[j+k] = u+v; a[j+k+i/2] = u-v.
// Field arithmetic is exact and canonical, so the stage order/grouping below is free to differ.
//
// B200 mapping: one CTA owns up to 4096 consecutive (bit-reversed) positions of one row in shared memory
// (64 KB, 64 registers -> 2 CTAs/SM) and runs the first 12 stages there (ntt_tile_lazy_kernel: lazy reduction, per-pass twiddle tables); longer transforms finish with register-resident
// radix-2^k passes over global memory (coalesced: consecutive threads own consecutive low index bits).
// The zero-extension of the message rows (the RS encoding evaluates a degree < len/2 polynomial on len points)
// is fused into the load, so the tensor's upper half is never memset nor read.
#include "common.cuh"

namespace hb {

static constexpr int kLogTile = 12;             // 4096 elements * 16 B = 64 KB of shared memory

// ---------------------------------------------------------------------------------------------------------
// twiddles: the FULL period w[k] = omega^k, k < len (the radix-4 passes index up to 3*len/4), computed once per length
// on the host with the same repeated multiplication as the reference
int get_twiddles(hb_ctx *ctx, int logn, const F **out) {
    if (logn < 1 || logn > 28) HB_FAIL(ctx, "get_twiddles: logn out of range");
    if (!ctx->tw[logn]) {
        size_t len = (size_t)1 << logn;
        std::vector<F> w(len);
        hb_F rou; hb_root_of_unity(logn, &rou);
        F w1 = mkF(rou.real, rou.img);
        w[0] = mkF(1, 0);
        for (size_t i = 1; i < len; i++) w[i] = h_fmul(w[i - 1], w1);
        HB_CHECK(ctx, cudaMalloc(&ctx->tw[logn], len * sizeof(F)));
        HB_CHECK(ctx, cudaMemcpyAsync(ctx->tw[logn], w.data(), len * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
        HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // w is a local
        // omega^(len/4) is a primitive 4th root of unity, i.e. +i or -i in F_p[i]: multiplying by it is a limb swap
        if (logn >= 2) {
            F j = w[len / 4];
            if (j.re != 0 || (j.im != 1 && j.im != P61 - 1)) HB_FAIL(ctx, "get_twiddles: omega^(len/4) is not +-i");
            ctx->tw_j_neg[logn] = (j.im != 1);
        }
        // omega^(len/8) is a primitive 8th root of unity: (1 + i)^2 = 2i and 2^61 = 1 mod p, so the four of them are 2^30 (+-1 +- i) and
        // multiplying by one is two additions and a 30-bit rotation per limb (no wide multiply)
        if (logn >= 3) {
            F w8 = w[len / 8];
            const u64 r30 = (u64)1 << 30;
            if ((w8.re != r30 && w8.re != P61 - r30) || (w8.im != r30 && w8.im != P61 - r30)) HB_FAIL(ctx, "get_twiddles: omega^(len/8) is not 2^30 (+-1 +- i)");
            ctx->tw_w8[logn] = (uint8_t)((w8.re != r30 ? 1 : 0) | (w8.im != r30 ? 2 : 0));
        }
    }
    *out = ctx->tw[logn];
    return 0;
}

__device__ __forceinline__ F mul_j(F x, bool neg) {            // x * (+i) = (-im, re) ; x * (-i) = (im, -re)
    return neg ? mkF(x.im, x.re ? P61 - x.re : 0) : mkF(x.im ? P61 - x.im : 0, x.re);
}

// ---------------------------------------------------------------------------------------------------------
// x * omega^(len/8) with omega^(len/8) = 2^30 (sr + si i), sr, si = +-1 (flags: bit 0 = sr negative, bit 1 = si negative):
//   2^30 ((sr a - si b) + (si a + sr b) i); the factor 2^30 is a rotation by 30 inside the 61-bit limb
__device__ __forceinline__ u64 rot30(u64 x) { return ((x << 30) & P61) | (x >> 31); }
__device__ __forceinline__ u64 neg61(u64 x) { return x ? P61 - x : 0; }
__device__ __forceinline__ F mul_w8(F x, unsigned flags) {
    const u64 apb = add61(x.re, x.im), amb = sub61(x.re, x.im);
    u64 re, im;
    if (flags == 0) { re = amb; im = apb; }
    else if (flags == 2) { re = apb; im = neg61(amb); }
    else if (flags == 1) { re = neg61(apb); im = amb; }
    else { re = neg61(amb); im = neg61(apb); }
    return mkF(rot30(re), rot30(im));
}
// 8-point butterfly (three radix-2 stages) on inputs that already carry their stage twiddles: x[0], X1 .. X7 -> c_0 .. c_7 in place.
// Inside, only powers of omega^(len/8) appear: J = omega^(len/4) = +-i (a limb swap) and W8 (above): no wide multiplies.
__device__ __forceinline__ void dft8(F *x, bool j_neg, unsigned w8) {
    const F a0 = fadd(x[0], x[1]), a1 = fsub(x[0], x[1]), a2 = fadd(x[2], x[3]), a3 = mul_j(fsub(x[2], x[3]), j_neg);
    const F a4 = fadd(x[4], x[5]), a5 = fsub(x[4], x[5]), a6 = fadd(x[6], x[7]), a7 = mul_j(fsub(x[6], x[7]), j_neg);
    const F b0 = fadd(a0, a2), b2 = fsub(a0, a2), b1 = fadd(a1, a3), b3 = fsub(a1, a3);
    const F b4 = fadd(a4, a6), b6 = mul_j(fsub(a4, a6), j_neg), b5 = mul_w8(fadd(a5, a7), w8), b7 = mul_w8(mul_j(fsub(a5, a7), j_neg), w8);
    x[0] = fadd(b0, b4); x[4] = fsub(b0, b4); x[1] = fadd(b1, b5); x[5] = fsub(b1, b5);
    x[2] = fadd(b2, b6); x[6] = fsub(b2, b6); x[3] = fadd(b3, b7); x[7] = fsub(b3, b7);
}

// Tile kernel: stages 1..lb of a length-2^logn transform on positions [tile*2^lb, (tile+1)*2^lb) of one row.
// Loads src[rev(p)] (zero if rev(p) >= in_len), writes dst[p].  src may alias dst only when lb == logn
// (then the CTA reads its whole row before it writes anything).
//  * rows are grouped in chunks: row r -> chunk r / rows_per_chunk; chunk strides are given separately so that all
//    chunks of a commit go through ONE launch (no per-chunk grid tail).
//  * stages are taken three at a time (radix-8): with e the twiddle exponent of the LAST of the three stages, the inputs are multiplied
//    by w^(4e), w^(2e), w^(6e), w^e, w^(5e), w^(3e), w^(7e) (7 multiplications per 8 points per 3 stages instead of 12) and everything
//    inside the 8-point butterfly is a power of omega^(len/8), which costs additions and rotations only.  A remainder of two stages is a
//    radix-4 pass (3 multiplications per 4 points), of one stage a radix-2 pass.
//  * zero-extended input (in_len == len/2, the RS encoding of a message row): every odd bit-reversed position is zero, so stage 1 is a
//    plain duplication and stages 1..4 are done in registers while loading: the 16 positions of a group see the same 8 inputs; the
//    even ones are their plain 8-point butterfly, the odd ones need the 16th roots: J, W8, W8 J and four real multiplications.
// 64 registers: two CTAs per SM.  Forcing three (40 registers) spills the butterfly and is slower (13.68 vs 12.91 ms per 2^26 commit).
__global__ void __launch_bounds__(512, 2)
ntt_tile_kernel(const F *__restrict__ src, size_t src_stride, size_t src_chunk_stride, size_t in_len,
                F *__restrict__ dst, size_t dst_stride, size_t dst_chunk_stride, unsigned rows_per_chunk,
                int logn, int lb, const F *__restrict__ tw, bool j_neg, unsigned w8) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F *s = reinterpret_cast<F *>(smem_raw);
    const unsigned tiles_per_row = 1u << (logn - lb);
    const size_t row = blockIdx.x / tiles_per_row;
    const unsigned tile = blockIdx.x % tiles_per_row;
    const unsigned tlen = 1u << lb, base = tile << lb;
    const size_t chunk = row / rows_per_chunk, rr = row % rows_per_chunk;
    const F *in = src + chunk * src_chunk_stride + rr * src_stride;
    F *out = dst + chunk * dst_chunk_stride + rr * dst_stride;
    const unsigned len = 1u << logn;

    int st = 1;
    if (in_len * 2 == len && lb >= 4) {
        const unsigned ngroups = tlen >> 4, e = len >> 4;
        for (unsigned u = threadIdx.x; u < 2 * ngroups; u += blockDim.x) {     // k = u / ngroups: uniform per warp for tiles of >= 512 positions
            const unsigned g = u % ngroups, k = u / ngroups;
            F x[8];
#pragma unroll
            for (int m = 0; m < 8; m++) x[m] = in[__brev(base + 16 * g + 2 * m) >> (32 - logn)];
            if (k) {
                x[1] = mul_j(x[1], j_neg); x[2] = mul_w8(x[2], w8); x[3] = mul_w8(mul_j(x[3], j_neg), w8);
                x[4] = fmul(x[4], ldgF(&tw[e])); x[5] = fmul(x[5], ldgF(&tw[5 * e])); x[6] = fmul(x[6], ldgF(&tw[3 * e])); x[7] = fmul(x[7], ldgF(&tw[7 * e]));
            }
            dft8(x, j_neg, w8);
#pragma unroll
            for (int m = 0; m < 8; m++) s[16 * g + 2 * m + k] = x[m];
        }
        st = 5;
    } else if (in_len * 2 == len && lb >= 1) {
        for (unsigned i = threadIdx.x; i < (tlen >> 1); i += blockDim.x) {
            unsigned p = base + 2 * i;
            F v = in[__brev(p) >> (32 - logn)];
            s[2 * i] = v; s[2 * i + 1] = v;
        }
        st = 2;
    } else {
        for (unsigned i = threadIdx.x; i < tlen; i += blockDim.x) {
            unsigned q = __brev(base + i) >> (32 - logn);
            s[i] = (q < in_len) ? in[q] : mkF(0, 0);
        }
    }
    __syncthreads();
    for (; st + 2 <= lb; st += 3) {                       // radix-8 pass: stages st, st+1, st+2
        const unsigned h = 1u << (st - 1);
        const unsigned tws3 = len >> (st + 2);            // twiddle stride of stage st+2
        for (unsigned q = threadIdx.x; q < (tlen >> 3); q += blockDim.x) {
            const unsigned k = q & (h - 1);
            const unsigned p0 = ((q >> (st - 1)) << (st + 2)) + k;
            const size_t e = (size_t)tws3 * k;
            F x[8];
#pragma unroll
            for (int m = 0; m < 8; m++) x[m] = s[p0 + m * h];
            x[1] = fmul(x[1], ldgF(&tw[4 * e])); x[2] = fmul(x[2], ldgF(&tw[2 * e])); x[3] = fmul(x[3], ldgF(&tw[6 * e]));
            x[4] = fmul(x[4], ldgF(&tw[e])); x[5] = fmul(x[5], ldgF(&tw[5 * e])); x[6] = fmul(x[6], ldgF(&tw[3 * e])); x[7] = fmul(x[7], ldgF(&tw[7 * e]));
            dft8(x, j_neg, w8);
#pragma unroll
            for (int m = 0; m < 8; m++) s[p0 + m * h] = x[m];
        }
        __syncthreads();
    }
    if (st + 1 <= lb) {                                   // radix-4 pass: stages st and st+1
        const unsigned h = 1u << (st - 1);
        const unsigned tws2 = len >> (st + 1);            // twiddle stride of stage st+1; stage st uses 2*tws2
        for (unsigned q = threadIdx.x; q < (tlen >> 2); q += blockDim.x) {
            unsigned k = q & (h - 1);
            unsigned p0 = ((q >> (st - 1)) << (st + 1)) + k;
            F x0 = s[p0], x1 = s[p0 + h], x2 = s[p0 + 2 * h], x3 = s[p0 + 3 * h];
            size_t e = (size_t)tws2 * k;
            F X1 = fmul(x1, ldgF(&tw[2 * e])), X2 = fmul(x2, ldgF(&tw[e])), X3 = fmul(x3, ldgF(&tw[3 * e]));
            F a0 = fadd(x0, X1), a1 = fsub(x0, X1), b = fadd(X2, X3), c = mul_j(fsub(X2, X3), j_neg);
            s[p0] = fadd(a0, b); s[p0 + 2 * h] = fsub(a0, b);
            s[p0 + h] = fadd(a1, c); s[p0 + 3 * h] = fsub(a1, c);
        }
        __syncthreads();
        st += 2;
    }
    if (st <= lb) {                                       // one radix-2 stage left
        const unsigned half = 1u << (st - 1), tws = len >> st;
        for (unsigned b = threadIdx.x; b < (tlen >> 1); b += blockDim.x) {
            unsigned k = b & (half - 1);
            unsigned p0 = ((b >> (st - 1)) << st) + k, p1 = p0 + half;
            F u = s[p0], v = fmul(s[p1], ldgF(&tw[(size_t)tws * k]));
            s[p0] = fadd(u, v); s[p1] = fsub(u, v);
        }
        __syncthreads();
    }
    for (unsigned i = threadIdx.x; i < tlen; i += blockDim.x) out[base + i] = s[i];
}

// ---------------------------------------------------------------------------------------------------------
// Lazy-reduction variant of the tile kernel (HB_NTT_LAZY, the default).  The tile kernel is bound by the ALU pipe, and more than half of its
// ALU work was canonicalisation: every fadd / fsub / fmul ended in a compare-subtract-select.  Here values inside a pass are only kept
// CONGRUENT mod p and below 2^64; the bounds (p = 2^61 - 1, so 8p = 2^64 - 8):
//   * shared memory holds "folded" limbs, <= p + 7 (fold61 of any 64-bit value); global input is canonical; the final store canonicalises;
//   * products are fmul_n_lazy (left limbs <= p + 7, canonical twiddle) -> <= p + 7;  x * 2^30 of any 64-bit x is (x >> 31) + ((x & (2^31-1)) << 30)
//     <= 2^61 + 2^33;  so every input of an 8-point butterfly is <= B0 := p + 2^34;
//   * a + b is a plain 64-bit addition; a - b is a + (K p - b) with K p >= the bound of b; multiplying by +-i swaps the limbs and turns one
//     of the two subtractions around (no negation);
//   * layer 1 of the butterfly: sums <= 2 B0, differences (K = 2) <= B0 + 2p, both <= U1 := 3p + 2^35;
//     layer 2: sums <= 2 U1, differences (K = 4) <= U1 + 4p < 8p; the two values that go through W8 are folded first (W8 adds both limbs);
//     before layer 3 everything is folded (<= p + 7, W8 outputs <= 2^61 + 2^33): sums <= 2 B0, differences (K = 2) <= 3p + 2^35; one more
//     fold per output.  16 folds (8 instructions each) + 24 additions (4 each) per 8 points instead of 24 canonical additions (12 - 16 each)
//     and 7 canonical products.
// The arithmetic is exact, so the canonical result is bit-identical to the canonical path (and to the reference's _fft).
__device__ __forceinline__ F ladd(F a, F b) { return mkF(a.re + b.re, a.im + b.im); }
template <int K> __device__ __forceinline__ u64 lsub1(u64 a, u64 b) { return a + ((u64)K * P61 - b); }             // requires b <= K p
template <int K> __device__ __forceinline__ F lsub(F a, F b) { return mkF(lsub1<K>(a.re, b.re), lsub1<K>(a.im, b.im)); }
// (a - b) * (+i) = (b.im - a.im, a.re - b.re);  (a - b) * (-i) = (a.im - b.im, b.re - a.re)
template <int K, bool JNEG> __device__ __forceinline__ F lsub_j(F a, F b) {
    return JNEG ? mkF(lsub1<K>(a.im, b.im), lsub1<K>(b.re, a.re)) : mkF(lsub1<K>(b.im, a.im), lsub1<K>(a.re, b.re));
}
// x * (+-i) for a CANONICAL x: p - limb is in [1, p] (congruent to the negated limb, 0 -> p)
template <bool JNEG> __device__ __forceinline__ F lmul_j(F x) { return JNEG ? mkF(x.im, P61 - x.re) : mkF(P61 - x.im, x.re); }
__device__ __forceinline__ u64 lrot30(u64 x) { return (x >> 31) + ((x & 0x7fffffffull) << 30); }                 // == 2^30 x (mod p), <= 2^61 + 2^33
// x * omega^(len/8) = 2^30 (sr + si i) x for FOLDED x (limbs <= p + 7); W8F bit 0: sr negative, bit 1: si negative
template <unsigned W8F> __device__ __forceinline__ F lmul_w8(F x) {
    const u64 apb = x.re + x.im, amb = lsub1<2>(x.re, x.im), bma = lsub1<2>(x.im, x.re), napb = 4 * P61 - apb;     // apb <= 2p + 14 <= 4p
    u64 re, im;
    if (W8F == 0) { re = amb; im = apb; }
    else if (W8F == 2) { re = apb; im = bma; }
    else if (W8F == 1) { re = napb; im = amb; }
    else { re = bma; im = napb; }
    return mkF(lrot30(re), lrot30(im));
}
// 8-point butterfly, inputs <= B0 per limb, outputs folded (<= p + 7)
template <unsigned W8F> __device__ __forceinline__ void dft8_lazy(F *x) {
    constexpr bool JN = ((W8F & 1) != ((W8F >> 1) & 1));          // omega^(len/4) = (omega^(len/8))^2 = sr si i
    const F a0 = ladd(x[0], x[1]), a1 = lsub<2>(x[0], x[1]), a2 = ladd(x[2], x[3]), a3 = lsub_j<2, JN>(x[2], x[3]);
    const F a4 = ladd(x[4], x[5]), a5 = lsub<2>(x[4], x[5]), a6 = ladd(x[6], x[7]), a7 = lsub_j<2, JN>(x[6], x[7]);
    const F b0 = lfold(ladd(a0, a2)), b2 = lfold(lsub<4>(a0, a2)), b1 = lfold(ladd(a1, a3)), b3 = lfold(lsub<4>(a1, a3));
    const F b4 = lfold(ladd(a4, a6)), b6 = lfold(lsub_j<4, JN>(a4, a6));
    const F b5 = lmul_w8<W8F>(lfold(ladd(a5, a7))), b7 = lmul_w8<W8F>(lfold(lsub_j<4, JN>(a5, a7)));
    x[0] = lfold(ladd(b0, b4)); x[4] = lfold(lsub<2>(b0, b4)); x[1] = lfold(ladd(b1, b5)); x[5] = lfold(lsub<2>(b1, b5));
    x[2] = lfold(ladd(b2, b6)); x[6] = lfold(lsub<2>(b2, b6)); x[3] = lfold(ladd(b3, b7)); x[7] = lfold(lsub<2>(b3, b7));
}
__device__ __forceinline__ F lmul_tw(F x, const F *tw) { return fmul_n_lazy(x, fprep(ldgF(tw))); }                 // x folded, twiddle canonical
// (Measured and not kept: reading the twiddles as PREPARED operands — limbs already split, doubled and negated, 32 bytes each — saves fprep's nine
// instructions per product but doubles the twiddle loads: step 11.29 vs 10.52 ms from the strided full-period table, 9.72 vs 9.72 ms as two
// lane-contiguous planes next to the per-pass tables.)

// Per-pass twiddle tables.  A pass at stage st (h = 2^(st-1) butterflies apart) multiplies input j of butterfly k by omega^(mult_j * stride * k):
// read from the full-period table that is a different stride per input and 32 scattered sectors per warp request (ncu, round 2: the L1
// data pipe of the tile kernel was as busy as the ALU pipe, 61 %, two thirds of it these loads and the bit-reversed input gather).  The
// tables below hold, for every stage st <= 12 of one transform length, the factors of a radix-8 pass (7 per butterfly), of a radix-4 pass (3)
// and of a radix-2 pass (1) as [input j][k], so that the lanes of a warp (consecutive k) read consecutive 16-byte entries.
//   layout: [radix-8: 7 (2^12 - 1)] [radix-4: 3 (2^12 - 1)] [radix-2: 2^12 - 1], the block of stage st at offset R (2^(st-1) - 1)
static constexpr size_t kPassR8 = 0, kPassR4 = 7 * (((size_t)1 << kLogTile) - 1), kPassR2 = kPassR4 + 3 * (((size_t)1 << kLogTile) - 1),
                        kPassTotal = kPassR2 + (((size_t)1 << kLogTile) - 1);
__global__ void __launch_bounds__(256) pass_table_kernel(const F *__restrict__ tw, F *__restrict__ out, int logn, int lb) {
    const unsigned len = 1u << logn;
    for (int st = 1; st <= lb; st++) {
        const unsigned h = 1u << (st - 1);
        for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < h; k += gridDim.x * blockDim.x) {
            if (st + 2 <= lb) {
                const size_t e = (size_t)(len >> (st + 2)) * k;
                const unsigned mult[7] = {4, 2, 6, 1, 5, 3, 7};
                for (int j = 0; j < 7; j++) out[kPassR8 + 7 * (size_t)(h - 1) + (size_t)j * h + k] = tw[mult[j] * e];
            }
            if (st + 1 <= lb) {
                const size_t e = (size_t)(len >> (st + 1)) * k;
                const unsigned mult[3] = {2, 1, 3};
                for (int j = 0; j < 3; j++) out[kPassR4 + 3 * (size_t)(h - 1) + (size_t)j * h + k] = tw[mult[j] * e];
            }
            out[kPassR2 + (size_t)(h - 1) + k] = tw[(size_t)(len >> st) * k];
        }
    }
}
static int get_pass_tables(hb_ctx *ctx, int logn, int lb, const F *tw, const F **out) {
    if (!ctx->tw_pass[logn]) {
        HB_CHECK(ctx, cudaMalloc(&ctx->tw_pass[logn], kPassTotal * sizeof(F)));
        HB_CHECK(ctx, cudaMemsetAsync(ctx->tw_pass[logn], 0, kPassTotal * sizeof(F), ctx->stream));
        HB_LAUNCH(ctx, pass_table_kernel, 8, 256, 0, tw, ctx->tw_pass[logn], logn, lb);
    }
    *out = ctx->tw_pass[logn];
    return 0;
}

template <unsigned W8F>
__global__ void __launch_bounds__(512, 2)
ntt_tile_lazy_kernel(const F *__restrict__ src, size_t src_stride, size_t src_chunk_stride, size_t in_len,
                     F *__restrict__ dst, size_t dst_stride, size_t dst_chunk_stride, unsigned rows_per_chunk,
                     int logn, int lb, const F *__restrict__ tw, const F *__restrict__ tp) {
    constexpr bool JN = ((W8F & 1) != ((W8F >> 1) & 1));
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F *s = reinterpret_cast<F *>(smem_raw);
    // Shared-memory index padding: with the bit-reversed thread order of the first pass the 8 lanes of a store phase differ in the TOP three
    // bits of their position, all in one bank group; one extra 16-byte slot per 2^(lb-3) positions (7 slots in all) moves them to 8 different
    // groups, and every later pass still reads aligned runs of 8 consecutive slots.  (Round 1 measured the padding alone at +1 %: the kernel was
    // ALU-bound then.)
    const int psh = lb >= 7 ? lb - 3 : 31;
    auto P = [psh](unsigned i) { return i + (i >> psh); };
    const unsigned tiles_per_row = 1u << (logn - lb);
    const size_t row = blockIdx.x / tiles_per_row;
    const unsigned tile = blockIdx.x % tiles_per_row;
    const unsigned tlen = 1u << lb, base = tile << lb;
    const size_t chunk = row / rows_per_chunk, rr = row % rows_per_chunk;
    const F *in = src + chunk * src_chunk_stride + rr * src_stride;
    F *out = dst + chunk * dst_chunk_stride + rr * dst_stride;
    const unsigned len = 1u << logn;

    int st = 1;
    if (in_len * 2 == len && lb >= 4) {
        const unsigned ngroups = tlen >> 4, e = len >> 4;
        const int lg = lb - 4;                                                  // log2(ngroups)
        for (unsigned u = threadIdx.x; u < 2 * ngroups; u += blockDim.x) {     // k = u / ngroups: uniform per warp for tiles of >= 512 positions
            // thread -> group in bit-reversed order: the input index of (group g, input m) is brev(m) * 2^.. + brev(g) (+ the tile's bits), so
            // consecutive threads gather consecutive input elements (4 wavefronts per warp request instead of 32)
            const unsigned gu = u % ngroups, k = u / ngroups, g = lg ? (__brev(gu) >> (32 - lg)) : 0u;
            F x[8];
#pragma unroll
            for (int m = 0; m < 8; m++) x[m] = in[__brev(base + 16 * g + 2 * m) >> (32 - logn)];      // canonical
            if (k) {
                x[1] = lmul_j<JN>(x[1]); x[2] = lmul_w8<W8F>(x[2]); x[3] = lmul_w8<W8F>(lmul_j<JN>(x[3]));       // canonical inputs: limbs stay <= p
                x[4] = lmul_tw(x[4], &tw[e]); x[5] = lmul_tw(x[5], &tw[5 * e]); x[6] = lmul_tw(x[6], &tw[3 * e]); x[7] = lmul_tw(x[7], &tw[7 * e]);
            }
            dft8_lazy<W8F>(x);
#pragma unroll
            for (int m = 0; m < 8; m++) s[P(16 * g + 2 * m + k)] = x[m];
        }
        st = 5;
    } else if (in_len * 2 == len && lb >= 1) {
        for (unsigned i = threadIdx.x; i < (tlen >> 1); i += blockDim.x) {
            unsigned p = base + 2 * i;
            F v = in[__brev(p) >> (32 - logn)];
            s[P(2 * i)] = v; s[P(2 * i + 1)] = v;
        }
        st = 2;
    } else {
        for (unsigned i = threadIdx.x; i < tlen; i += blockDim.x) {
            unsigned q = __brev(base + i) >> (32 - logn);
            s[P(i)] = (q < in_len) ? in[q] : mkF(0, 0);
        }
    }
    __syncthreads();
    for (; st + 2 <= lb; st += 3) {                       // radix-8 pass: stages st, st+1, st+2
        const unsigned h = 1u << (st - 1);
        for (unsigned q = threadIdx.x; q < (tlen >> 3); q += blockDim.x) {
            const unsigned k = q & (h - 1);
            const unsigned p0 = ((q >> (st - 1)) << (st + 2)) + k;
            const F *t8 = tp + kPassR8 + 7 * (size_t)(h - 1) + k;            // [input j][k]: omega^({4,2,6,1,5,3,7}[j] * tws3 * k)
            F x[8];
#pragma unroll
            for (int m = 0; m < 8; m++) x[m] = s[P(p0 + m * h)];
#pragma unroll
            for (int j = 0; j < 7; j++) x[j + 1] = lmul_tw(x[j + 1], &t8[(size_t)j * h]);
            dft8_lazy<W8F>(x);
#pragma unroll
            for (int m = 0; m < 8; m++) s[P(p0 + m * h)] = x[m];
        }
        __syncthreads();
    }
    if (st + 1 <= lb) {                                   // radix-4 pass: stages st and st+1
        const unsigned h = 1u << (st - 1);
        for (unsigned q = threadIdx.x; q < (tlen >> 2); q += blockDim.x) {
            unsigned k = q & (h - 1);
            unsigned p0 = ((q >> (st - 1)) << (st + 1)) + k;
            const F x0 = s[P(p0)], x1 = s[P(p0 + h)], x2 = s[P(p0 + 2 * h)], x3 = s[P(p0 + 3 * h)];
            const F *t4 = tp + kPassR4 + 3 * (size_t)(h - 1) + k;            // [j][k]: omega^({2,1,3}[j] * tws2 * k)
            const F X1 = lmul_tw(x1, &t4[0]), X2 = lmul_tw(x2, &t4[h]), X3 = lmul_tw(x3, &t4[2 * (size_t)h]);      // <= p + 7
            const F a0 = ladd(x0, X1), a1 = lsub<2>(x0, X1), b = ladd(X2, X3), c = lsub_j<2, JN>(X2, X3);       // <= 3p + 14
            s[P(p0)] = lfold(ladd(a0, b)); s[P(p0 + 2 * h)] = lfold(lsub<4>(a0, b));
            s[P(p0 + h)] = lfold(ladd(a1, c)); s[P(p0 + 3 * h)] = lfold(lsub<4>(a1, c));
        }
        __syncthreads();
        st += 2;
    }
    if (st <= lb) {                                       // one radix-2 stage left
        const unsigned half = 1u << (st - 1);
        for (unsigned b = threadIdx.x; b < (tlen >> 1); b += blockDim.x) {
            unsigned k = b & (half - 1);
            unsigned p0 = ((b >> (st - 1)) << st) + k, p1 = p0 + half;
            const F u = s[P(p0)], v = lmul_tw(s[P(p1)], &tp[kPassR2 + (size_t)(half - 1) + k]);
            s[P(p0)] = lfold(ladd(u, v)); s[P(p1)] = lfold(lsub<2>(u, v));
        }
        __syncthreads();
    }
    for (unsigned i = threadIdx.x; i < tlen; i += blockDim.x) out[base + i] = fcanon(s[P(i)]);       // folded (<= p + 7) -> canonical
}

// Register-resident radix-2^CNT pass over global memory: stages s_lo+1 .. s_lo+CNT of a length-2^logn transform.
template <int CNT>
__global__ void __launch_bounds__(256)
ntt_global_kernel(F *__restrict__ data, size_t stride, int logn, int s_lo, const F *__restrict__ tw, size_t total_groups, bool j_neg, unsigned w8) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total_groups) return;
    const unsigned groups_per_row = 1u << (logn - CNT);
    const size_t row = gid / groups_per_row;
    const unsigned g = (unsigned)(gid % groups_per_row);
    const unsigned low = g & ((1u << s_lo) - 1), high = g >> s_lo;
    const unsigned base = (high << (s_lo + CNT)) | low;
    F *a = data + row * stride;
    F v[1 << CNT];
#pragma unroll
    for (int q = 0; q < (1 << CNT); q++) v[q] = a[base + ((unsigned)q << s_lo)];
    if constexpr (CNT == 3) {
        // the same radix-8 butterfly as in the tile kernel: stages s_lo+1 .. s_lo+3, h = 2^s_lo, k = low
        const size_t e = (size_t)((1u << logn) >> (s_lo + 3)) * low;
        v[1] = fmul(v[1], ldgF(&tw[4 * e])); v[2] = fmul(v[2], ldgF(&tw[2 * e])); v[3] = fmul(v[3], ldgF(&tw[6 * e]));
        v[4] = fmul(v[4], ldgF(&tw[e])); v[5] = fmul(v[5], ldgF(&tw[5 * e])); v[6] = fmul(v[6], ldgF(&tw[3 * e])); v[7] = fmul(v[7], ldgF(&tw[7 * e]));
        dft8(v, j_neg, w8);
    } else {
#pragma unroll
    for (int t = 0; t < CNT; t++) {
        const int st = s_lo + t + 1;
        const unsigned tws = (1u << logn) >> st;
#pragma unroll
        for (int q = 0; q < (1 << CNT); q++) {
            if (q & (1 << t)) continue;
            // k = position inside the half-block of this stage
            unsigned k = (((unsigned)q & ((1u << t) - 1)) << s_lo) | low;
            F u = v[q];
            F x = fmul(v[q | (1 << t)], ldgF(&tw[(size_t)tws * k]));
            v[q] = fadd(u, x);
            v[q | (1 << t)] = fsub(u, x);
        }
    }
    }
#pragma unroll
    for (int q = 0; q < (1 << CNT); q++) a[base + ((unsigned)q << s_lo)] = v[q];
}

static int ntt_rows_impl(hb_ctx *ctx, const F *src, size_t src_stride, size_t in_len, F *dst, size_t dst_stride,
                         int logn, size_t batch, size_t rows_per_chunk = 0, size_t src_chunk_stride = 0, size_t dst_chunk_stride = 0) {
    if (rows_per_chunk == 0) { rows_per_chunk = batch ? batch : 1; }
    if (batch == 0) return 0;
    if (logn == 0) {
        if (src != dst) HB_CHECK(ctx, cudaMemcpy2DAsync(dst, dst_stride * sizeof(F), src, src_stride * sizeof(F), sizeof(F), batch,
                                                         cudaMemcpyDeviceToDevice, ctx->stream));
        return 0;
    }
    const F *tw; HB_TRY(get_twiddles(ctx, logn, &tw));
    const int lb = logn < kLogTile ? logn : kLogTile;
    if (src == dst && lb != logn) HB_FAIL(ctx, "ntt: in-place transform longer than one tile needs a distinct source");
    const size_t smem = (sizeof(F) << lb) + 8 * sizeof(F);         // + the 7 padding slots of the lazy tile kernel
    static bool attr_set_dev[64] = {};                        // the attribute is per device
    bool &attr_set = attr_set_dev[ctx->device & 63];
    if (!attr_set) {
        HB_CHECK(ctx, cudaFuncSetAttribute(ntt_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((sizeof(F) << kLogTile) + 8 * sizeof(F))));
        attr_set = true;
    }
    unsigned threads = 1u << (lb > 3 ? lb - 3 : 0);          // one radix-8 butterfly per thread up to 512 threads
    if (threads > 512) threads = 512;
    if (threads < 32) threads = 32;
    size_t grid = batch << (logn - lb);
    if (rows_per_chunk != batch && lb != logn) HB_FAIL(ctx, "ntt: chunked launch supports transforms up to one tile");
#ifndef HB_NTT_LAZY
#define HB_NTT_LAZY 1
#endif
    static const int lazy_ok = getenv("HB_NTT_LAZY") ? atoi(getenv("HB_NTT_LAZY")) : HB_NTT_LAZY;      // experiment switch
    // the lazy kernel is specialised on the 8th root (and derives +-i from it: (omega^(len/8))^2 = omega^(len/4)); lengths below 8 have no 8th root
    const unsigned w8f = logn >= 3 ? (unsigned)ctx->tw_w8[logn] : (ctx->tw_j_neg[logn] ? 1u : 0u);
    const bool jn_derived = ((w8f & 1) != ((w8f >> 1) & 1));
    if (lazy_ok && (logn < 2 || jn_derived == ctx->tw_j_neg[logn])) {
        static bool lazy_attr_dev[64] = {};
        if (!lazy_attr_dev[ctx->device & 63]) {
            HB_CHECK(ctx, cudaFuncSetAttribute(ntt_tile_lazy_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((sizeof(F) << kLogTile) + 8 * sizeof(F))));
            HB_CHECK(ctx, cudaFuncSetAttribute(ntt_tile_lazy_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((sizeof(F) << kLogTile) + 8 * sizeof(F))));
            HB_CHECK(ctx, cudaFuncSetAttribute(ntt_tile_lazy_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((sizeof(F) << kLogTile) + 8 * sizeof(F))));
            HB_CHECK(ctx, cudaFuncSetAttribute(ntt_tile_lazy_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((sizeof(F) << kLogTile) + 8 * sizeof(F))));
            lazy_attr_dev[ctx->device & 63] = true;
        }
        const F *tp; HB_TRY(get_pass_tables(ctx, logn, lb, tw, &tp));
#define HB_NTT_LAZY_LAUNCH(W) HB_LAUNCH(ctx, ntt_tile_lazy_kernel<W>, (unsigned)grid, threads, smem, src, src_stride, src_chunk_stride, in_len, dst, dst_stride, \
                                       dst_chunk_stride, (unsigned)rows_per_chunk, logn, lb, tw, tp)
        if (w8f == 0) { HB_NTT_LAZY_LAUNCH(0); } else if (w8f == 1) { HB_NTT_LAZY_LAUNCH(1); } else if (w8f == 2) { HB_NTT_LAZY_LAUNCH(2); } else { HB_NTT_LAZY_LAUNCH(3); }
#undef HB_NTT_LAZY_LAUNCH
    } else
    HB_LAUNCH(ctx, ntt_tile_kernel, (unsigned)grid, threads, smem, src, src_stride, src_chunk_stride, in_len, dst, dst_stride, dst_chunk_stride,
              (unsigned)rows_per_chunk, logn, lb, tw, ctx->tw_j_neg[logn], (unsigned)ctx->tw_w8[logn]);
    int s_lo = lb;
    while (s_lo < logn) {
        int cnt = logn - s_lo; if (cnt > 3) cnt = 3;
        size_t groups = batch << (logn - cnt);
        unsigned g = (unsigned)((groups + 255) / 256);
        if (cnt == 3) HB_LAUNCH(ctx, ntt_global_kernel<3>, g, 256, 0, dst, dst_stride, logn, s_lo, tw, groups, ctx->tw_j_neg[logn], (unsigned)ctx->tw_w8[logn]);
        else if (cnt == 2) HB_LAUNCH(ctx, ntt_global_kernel<2>, g, 256, 0, dst, dst_stride, logn, s_lo, tw, groups, ctx->tw_j_neg[logn], (unsigned)ctx->tw_w8[logn]);
        else HB_LAUNCH(ctx, ntt_global_kernel<1>, g, 256, 0, dst, dst_stride, logn, s_lo, tw, groups, ctx->tw_j_neg[logn], (unsigned)ctx->tw_w8[logn]);
        s_lo += cnt;
    }
    return 0;
}

int ntt_rows_dev(hb_ctx *ctx, F *data, int logn, size_t batch, size_t stride) {
    if (logn <= kLogTile) return ntt_rows_impl(ctx, data, stride, (size_t)1 << logn, data, stride, logn, batch);
    // long in-place transform: stage the input in a temporary (the tile kernel gathers bit-reversed across tiles)
    F *tmp; size_t len = (size_t)1 << logn;
    HB_CHECK(ctx, cudaMallocAsync(&tmp, batch * len * sizeof(F), ctx->stream));
    HB_CHECK(ctx, cudaMemcpy2DAsync(tmp, len * sizeof(F), data, stride * sizeof(F), len * sizeof(F), batch, cudaMemcpyDeviceToDevice, ctx->stream));
    int r = ntt_rows_impl(ctx, tmp, len, len, data, stride, logn, batch);
    cudaFreeAsync(tmp, ctx->stream);
    return r;
}

int ntt_rows_padded_dev(hb_ctx *ctx, const F *src, size_t in_len, F *dst, size_t dst_stride, int logn, size_t rows_per_chunk,
                        size_t nchunks, size_t src_chunk_stride, size_t dst_chunk_stride) {
    if (nchunks > 1 && logn > kLogTile) {       // long rows: one chunk per launch (the global passes address one matrix)
        for (size_t c = 0; c < nchunks; c++)
            HB_TRY(ntt_rows_impl(ctx, src + c * src_chunk_stride, in_len, in_len, dst + c * dst_chunk_stride, dst_stride, logn, rows_per_chunk));
        return 0;
    }
    return ntt_rows_impl(ctx, src, in_len, in_len, dst, dst_stride, logn, rows_per_chunk * nchunks, rows_per_chunk, src_chunk_stride, dst_chunk_stride);
}

// ---------------------------------------------------------------------------------------------------------
// Column NTT (RS x RS tensor code, PC_utils.cpp:28-39): a CTA owns CB adjacent columns of all 2^logn rows in
// shared memory as s[row][CB]; a quarter-warp reads one 16*CB-byte row segment -> conflict-free LDS.128 and
// full-sector global accesses.  Rows >= nz_rows are zero on input and are not read.
template <int CB>
__global__ void __launch_bounds__(512, 2)
ntt_cols_kernel(F *__restrict__ mat, size_t cols, int logn, unsigned nz_rows, const F *__restrict__ tw, bool j_neg, unsigned w8) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F *s = reinterpret_cast<F *>(smem_raw);
    const unsigned rows = 1u << logn;
    const size_t col0 = (size_t)blockIdx.x * CB;
    const unsigned c = threadIdx.x % CB, t0 = threadIdx.x / CB, tstep = blockDim.x / CB;
    for (unsigned p = t0; p < rows; p += tstep) {
        unsigned q = __brev(p) >> (32 - logn);
        s[p * CB + c] = (q < nz_rows) ? mat[(size_t)q * cols + col0 + c] : mkF(0, 0);
    }
    __syncthreads();
    int st = 1;
    for (; st + 2 <= logn; st += 3) {                     // radix-8 pass (see ntt_tile_kernel): stages st, st+1, st+2
        const unsigned h = 1u << (st - 1), tws3 = rows >> (st + 2);
        for (unsigned q = t0; q < (rows >> 3); q += tstep) {
            const unsigned k = q & (h - 1), p0 = ((q >> (st - 1)) << (st + 2)) + k;
            F x[8];
#pragma unroll
            for (int m = 0; m < 8; m++) x[m] = s[(p0 + m * h) * CB + c];
            if (k) {
                const size_t e = (size_t)tws3 * k;
                x[1] = fmul(x[1], ldgF(&tw[4 * e])); x[2] = fmul(x[2], ldgF(&tw[2 * e])); x[3] = fmul(x[3], ldgF(&tw[6 * e]));
                x[4] = fmul(x[4], ldgF(&tw[e])); x[5] = fmul(x[5], ldgF(&tw[5 * e])); x[6] = fmul(x[6], ldgF(&tw[3 * e])); x[7] = fmul(x[7], ldgF(&tw[7 * e]));
            }
            dft8(x, j_neg, w8);
#pragma unroll
            for (int m = 0; m < 8; m++) s[(p0 + m * h) * CB + c] = x[m];
        }
        __syncthreads();
    }
    if (st + 1 <= logn) {                                 // radix-4 pass: stages st, st+1
        const unsigned h = 1u << (st - 1), tws2 = rows >> (st + 1);
        for (unsigned q = t0; q < (rows >> 2); q += tstep) {
            const unsigned k = q & (h - 1), p0 = ((q >> (st - 1)) << (st + 1)) + k;
            F x0 = s[p0 * CB + c], x1 = s[(p0 + h) * CB + c], x2 = s[(p0 + 2 * h) * CB + c], x3 = s[(p0 + 3 * h) * CB + c];
            if (k) {
                const size_t e = (size_t)tws2 * k;
                x1 = fmul(x1, ldgF(&tw[2 * e])); x2 = fmul(x2, ldgF(&tw[e])); x3 = fmul(x3, ldgF(&tw[3 * e]));
            }
            const F a0 = fadd(x0, x1), a1 = fsub(x0, x1), b = fadd(x2, x3), d = mul_j(fsub(x2, x3), j_neg);
            s[p0 * CB + c] = fadd(a0, b); s[(p0 + 2 * h) * CB + c] = fsub(a0, b);
            s[(p0 + h) * CB + c] = fadd(a1, d); s[(p0 + 3 * h) * CB + c] = fsub(a1, d);
        }
        __syncthreads();
        st += 2;
    }
    if (st <= logn) {                                     // one radix-2 stage left
        const unsigned half = 1u << (st - 1), tws = rows >> st;
        for (unsigned b = t0; b < (rows >> 1); b += tstep) {
            unsigned k = b & (half - 1), j = (b >> (st - 1)) << st;
            unsigned p0 = j + k, p1 = p0 + half;
            F u = s[p0 * CB + c], x = s[p1 * CB + c];
            F v = (k == 0) ? x : fmul(x, ldgF(&tw[(size_t)tws * k]));
            s[p0 * CB + c] = fadd(u, v);
            s[p1 * CB + c] = fsub(u, v);
        }
        __syncthreads();
    }
    for (unsigned p = t0; p < rows; p += tstep) mat[(size_t)p * cols + col0 + c] = s[p * CB + c];
}

template <int CB>
static int launch_cols(hb_ctx *ctx, F *mat, int logn, size_t cols, size_t nz_rows, const F *tw) {
    size_t smem = (sizeof(F) * CB) << logn;
    HB_CHECK(ctx, cudaFuncSetAttribute(ntt_cols_kernel<CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HB_LAUNCH(ctx, ntt_cols_kernel<CB>, (unsigned)(cols / CB), 512, smem, mat, cols, logn, (unsigned)nz_rows, tw, ctx->tw_j_neg[logn], (unsigned)ctx->tw_w8[logn]);
    return 0;
}

int ntt_cols_dev(hb_ctx *ctx, F *mat, int logn, size_t cols, size_t nz_rows) {
    if (logn == 0 || cols == 0) return 0;
    const F *tw; HB_TRY(get_twiddles(ctx, logn, &tw));
    // pick the widest column block whose tile fits comfortably (<= 64 KB keeps 3 CTAs per SM)
    size_t rows = (size_t)1 << logn;
    if (rows * 16 * 16 <= 65536 && cols % 16 == 0) return launch_cols<16>(ctx, mat, logn, cols, nz_rows, tw);
    if (rows * 8 * 16 <= 131072 && cols % 8 == 0) return launch_cols<8>(ctx, mat, logn, cols, nz_rows, tw);
    if (rows * 4 * 16 <= 196608 && cols % 4 == 0) return launch_cols<4>(ctx, mat, logn, cols, nz_rows, tw);
    if (rows * 1 * 16 <= 196608) return launch_cols<1>(ctx, mat, logn, cols, nz_rows, tw);
    HB_FAIL(ctx, "ntt_cols: column length too large for the shared-memory kernel");
}

}  // namespace hb

// ---------------------------------------------------------------------------------------------------------
extern "C" int hb_ntt_batch(hb_ctx *ctx, hb_F *data, int logn, size_t batch, size_t stride) { HB_DEV(ctx);
    using namespace hb;
    if (logn < 0 || logn > 30) HB_FAIL(ctx, "hb_ntt_batch: logn out of range");
    size_t len = (size_t)1 << logn;
    if (stride < len) HB_FAIL(ctx, "hb_ntt_batch: stride < 2^logn");
    if (batch == 0) return 0;
    Staged d(ctx);
    HB_TRY(d.outbuf(data, ((batch - 1) * stride + len) * sizeof(F), true));
    HB_TRY(ntt_rows_dev(ctx, d.as<F>(), logn, batch, stride));
    HB_TRY(d.finish());
    HB_TRY(end_call(ctx));
    return 0;
}
