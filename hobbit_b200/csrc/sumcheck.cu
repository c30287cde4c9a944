// S1/S2/S3/S5/S9 — sumcheck provers and the product-tree GKR on the GPU.
// Reference: generate_2product_sumcheck_proof (src/sumcheck.cpp:2391-2460), _generate_3product_sumcheck_proof
// (:1974-2058), batch_3product_sumcheck (:275-372), prove_multiplication_tree_new (:35-257),
// precompute_beta (src/utils.cpp:251-296), evaluate_vector (:789-802).
//
// One kernel per round fuses everything that touches the bookkeeping tables: it folds the previous round's tables
// with that round's challenge ON THE FLY (reading each table exactly once), writes the half-size tables, and
// accumulates the new round polynomial with a warp-shuffle + shared-memory + last-CTA (ticket) reduction.  So a round
// moves 16*NT*(2L) bytes in and 16*NT*L bytes out for L output pairs — the algorithmic minimum of SURVEY §8d — and
// the only thing that leaves the GPU per round is the 3–4 coefficients (48–64 B) the Fiat–Shamir MiMC chain needs.
// MiMC (161 dependent cubings) is latency-bound and runs on the host between rounds, as does the ≤4-element
// coefficient algebra; no table entry is ever touched by the CPU.
//
// Variable order: every fold pairs (2j, 2j+1) — variable 0 is the LSB.  Challenge/fold ordering differs
// per prover (N7): S1/S3 derive the challenge after the round polynomial and fold with it (here: folded lazily by the
// NEXT round's kernel); S2 folds with the incoming challenge in the same pass that accumulates the polynomial.
#include "common.cuh"
#include "reduce.cuh"
#include <algorithm>

namespace hb {

enum { POLY_ONLY = 0, FOLD_THEN_POLY = 1, POLY_AND_FOLD = 2, FOLD_ONLY = 3 };

template <int NT> struct Tabs { const F *in[NT]; F *out[NT]; };

__device__ __forceinline__ F fold1(F x, F y, F r) { return fadd(x, fmul(r, fsub(y, x))); }
// the same against a prepared challenge: x + r (y - x), canonical; d = y - x is handed back for the round polynomial
#ifndef HB_SC_FUSEFOLD
#define HB_SC_FUSEFOLD 1
#endif
__device__ __forceinline__ F fold1n(F x, F d, const FN &rn) {
#if HB_SC_FUSEFOLD
    const F t = fmul_n_raw(d, rn);                                   // raw sums < 2^64 - 2^61: x (canonical) is added BEFORE the one fold
    return fcanon(lfold(mkF(x.re + t.re, x.im + t.im)));
#else
    const F t = fmul_n_lazy(d, rn);                                  // limbs <= p + 7
    return fcanon2(mkF(x.re + t.re, x.im + t.im));                   // < 2p + 7 < 2^62
#endif
}

// Round polynomial prod_t (x_t + X (y_t - x_t)) accumulated in EVALUATION form — fewer multiplications than its coefficients:
//   NT = 2: acc = { P(inf) = d1 d2, P(1) = y1 y2, P(0) = x1 x2 }                                  (3 multiplications instead of 4)
//   NT = 3: acc = { P(inf) = d1 d2 d3, P(1) = y1 y2 y3, P(-1) = (2x1-y1)(2x2-y2)(2x3-y3), P(0) }    (7 instead of 10: the pair products
//           A = x1 x2, B = y1 y2, C = d1 d2 give (2x1-y1)(2x2-y2) = 2A + 2C - B for free)
// The sums over all pairs are linear in these, so the coefficients (a, b, c[, d]) the reference accumulates directly are recovered
// exactly on the host from the reduced sums (coeffs_from_evals); all arithmetic is exact in F_p^2, so the bits are the same.
// Products are LAZY (limbs <= p + 7, fmul_n_lazy) and so are the accumulators: the two integer pipes are the binding units of this
// kernel and canonicalising intermediates would only add ALU work.  Accumulator flavours: WideAcc (the default: the UNFOLDED 64-bit dot
// products of the last multiplication are added into 96-bit sums, no fold per product), RegAcc (64-bit sums of folded products, re-folded
// every 4th pair; HB_SC_ACC3=0) and SmemAcc (the same in a thread-private shared-memory column, an experiment that freed registers for a
// third resident CTA and was slower; HB_SC_SACC=1).
template <int NC> struct RegAcc {
    F a[NC];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int c = 0; c < NC; c++) a[c] = mkF(0, 0);
    }
    __device__ __forceinline__ void add(int c, F v) { lacc(a[c], v); }
    __device__ __forceinline__ void fold() {
#pragma unroll
        for (int c = 0; c < NC; c++) a[c] = lfold(a[c]);
    }
    __device__ __forceinline__ F get(int c) { return a[c]; }
};
// raw product sums (< 2^64 each) added into 96-bit accumulators: 3 instructions per limb and product instead of a fold (4) and an add (2),
// and no periodic re-fold; 2^64 == 8 (mod p) brings the top word back at the end.  add_raw takes UNFOLDED limbs (fmul_n_raw).
template <int NC> struct WideAcc {
    uint32_t w[NC][2][3];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
            for (int l = 0; l < 2; l++) w[c][l][0] = w[c][l][1] = w[c][l][2] = 0;
    }
    __device__ __forceinline__ void add1(uint32_t (&a)[3], u64 v) {
        asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;"
            : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]) : "r"((uint32_t)v), "r"((uint32_t)(v >> 32)));
    }
    __device__ __forceinline__ void add_raw(int c, F v) { add1(w[c][0], v.re); add1(w[c][1], v.im); }
    __device__ __forceinline__ void fold() {}
    __device__ __forceinline__ u64 get1(const uint32_t (&a)[3]) { return fold61(((u64)a[1] << 32) | a[0]) + 8ull * a[2]; }   // < 2^62
    __device__ __forceinline__ F get(int c) { return mkF(get1(w[c][0]), get1(w[c][1])); }
};
template <int NC> struct SmemAcc {
    F *col;                                             // col[c * blockDim.x]: consecutive threads -> consecutive 16-byte slots, conflict-free
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int c = 0; c < NC; c++) col[c * blockDim.x] = mkF(0, 0);
    }
    __device__ __forceinline__ void add(int c, F v) { F t = col[c * blockDim.x]; lacc(t, v); col[c * blockDim.x] = t; }
    __device__ __forceinline__ void fold() {
#pragma unroll
        for (int c = 0; c < NC; c++) col[c * blockDim.x] = lfold(col[c * blockDim.x]);
    }
    __device__ __forceinline__ F get(int c) { return col[c * blockDim.x]; }
};
#ifndef HB_SC_ACC3
#define HB_SC_ACC3 1
#endif
template <int NT, class Acc> __device__ __forceinline__ void poly_acc(Acc &acc, const F (&x)[NT], const F (&y)[NT], const F (&d)[NT]) {
    if (NT == 2) {
#if HB_SC_ACC3
        acc.add_raw(0, fmul_n_raw(d[0], fprep(d[1])));
        acc.add_raw(1, fmul_n_raw(y[0], fprep(y[1])));
        acc.add_raw(2, fmul_n_raw(x[0], fprep(x[1])));
#else
        acc.add(0, fmul_n_lazy(d[0], fprep(d[1])));
        acc.add(1, fmul_n_lazy(y[0], fprep(y[1])));
        acc.add(2, fmul_n_lazy(x[0], fprep(x[1])));
#endif
    } else if (NT == 3) {
        const F A = fmul_n_lazy(x[0], fprep(x[1])), B = fmul_n_lazy(y[0], fprep(y[1])), C = fmul_n_lazy(d[0], fprep(d[1]));
        // M = 2A + 2C - B: 2 (A + C) <= 4p + 28, + (2p - B) stays below 2^64; one fold -> limbs <= p + 7
        const F M = lfold(mkF(2 * (A.re + C.re) + (2 * P61 - B.re), 2 * (A.im + C.im) + (2 * P61 - B.im)));
        // m3 = 2 x3 - y3 = x3 - d3 (canonical: it is a right-hand operand; one borrow-chain subtraction)
        const F m3 = fsub(x[2], d[2]);
#if HB_SC_ACC3
        acc.add_raw(0, fmul_n_raw(C, fprep(d[2])));
        acc.add_raw(1, fmul_n_raw(B, fprep(y[2])));
        acc.add_raw(2, fmul_n_raw(M, fprep(m3)));
        acc.add_raw(3, fmul_n_raw(A, fprep(x[2])));
#else
        acc.add(0, fmul_n_lazy(C, fprep(d[2])));
        acc.add(1, fmul_n_lazy(B, fprep(y[2])));
        acc.add(2, fmul_n_lazy(M, fprep(m3)));
        acc.add(3, fmul_n_lazy(A, fprep(x[2])));
#endif
    }
}
// host side: evaluation sums -> coefficients, highest degree first (1/2 = 2^60 mod p)
template <int NT> static inline void coeffs_from_evals(F *co) {
    if (NT == 2) co[1] = fsub(fsub(co[1], co[0]), co[2]);
    else if (NT == 3) {
        const F inv2 = mkF((u64)1 << 60, 0);
        const F a = co[0], p1 = co[1], pm1 = co[2], d = co[3];
        co[1] = fsub(h_fmul(fadd(p1, pm1), inv2), d);
        co[2] = fsub(h_fmul(fsub(p1, pm1), inv2), a);
    }
}

// INTERLEAVED: tables 0 and 1 are the even/odd entries of one array t.in[0] (product-tree layer: in1[j]=prev[2j],
// in2[j]=prev[2j+1], sumcheck.cpp:84-101), so the layer is consumed in place without materialising in1/in2.
#ifndef HB_SC_MINB
#define HB_SC_MINB 2
#endif
#ifndef HB_SC_THREADS
#define HB_SC_THREADS 256
#endif
// the tables are streamed exactly once per round: HB_SC_STREAM=1 marks the loads evict-first (ld.global.cs)
#if defined(HB_SC_STREAM) && HB_SC_STREAM
__device__ __forceinline__ F ld_stream(const F *p) { ulonglong2 v; asm volatile("ld.global.cs.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p)); return mkF(v.x, v.y); }
#define HB_SC_LD(p) ld_stream(p)
#else
#define HB_SC_LD(p) (*(p))
#endif
// (A variant that kept the next iterations' table entries in flight with cp.async into thread-private shared-memory slots was measured
// and was 4 % slower; what did pay is the register ping-pong below — the next pair's entries loaded into a second register set —
// which took long_scoreboard from 1.62 to 0.09 stall cycles per issue, see profiles/r02_summary.md section 0.)
template <int NT, int MODE, bool INTERLEAVED>
__global__ void __launch_bounds__(HB_SC_THREADS, HB_SC_MINB)
sc_round_kernel(Tabs<NT> t, size_t L, F r, RedArgs ra) {
    constexpr int NC = NT + 1;
#ifndef HB_SC_SACC
#define HB_SC_SACC 0
#endif
#if HB_SC_SACC
    __shared__ F sacc[NC * HB_SC_THREADS];
    SmemAcc<NC> acc; acc.col = sacc + threadIdx.x;
#elif HB_SC_ACC3
    WideAcc<NC> acc;
#else
    RegAcc<NC> acc;
#endif
    acc.init();
    const FN rn = fprep(r);
    unsigned it = 0;

#ifndef HB_SC_PF
#define HB_SC_PF 2
#endif
    const size_t stride = (size_t)gridDim.x * blockDim.x;
#ifndef HB_SC_RPF
#define HB_SC_RPF 2
#endif
    // HB_SC_RPF: the table entries of this thread's NEXT pair are loaded into registers before the current pair is multiplied out (software
    // pipelining one iteration deep), so that the DRAM round trip overlaps ~800 instructions of arithmetic instead of stalling the warp at the
    // top of every iteration (ncu, round 2: long_scoreboard was the largest stall with 4 warps per scheduler).
    constexpr int RPF = (MODE != FOLD_THEN_POLY && NT <= 3) ? HB_SC_RPF : 0;       // wider table sets (the fold-only passes of S7/S8) would spill
    auto load_pair = [&](size_t jj, F (&xx)[NT], F (&yy)[NT]) {
        if (INTERLEAVED) {
            const F *p = t.in[0] + 4 * jj;
            xx[0] = p[0]; xx[1] = p[1]; yy[0] = p[2]; yy[1] = p[3];
#pragma unroll
            for (int k = 2; k < NT; k++) { xx[k] = t.in[k][2 * jj]; yy[k] = t.in[k][2 * jj + 1]; }
        } else {
#pragma unroll
            for (int k = 0; k < NT; k++) { xx[k] = HB_SC_LD(&t.in[k][2 * jj]); yy[k] = HB_SC_LD(&t.in[k][2 * jj + 1]); }
        }
    };
    // The tables are streamed at ~3 TB/s: the loads of a pair would wait a loaded-DRAM round trip (ncu: long_scoreboard is the top stall
    // with only 16 warps per SM).  The sectors of the pair this thread handles HB_SC_PF iterations later are pulled into L2 now — one
    // prefetch instruction per 32-byte sector, no registers held.
    auto prefetch_l2 = [&](size_t j) {
        if (HB_SC_PF > 0 && j + HB_SC_PF * stride < L) {
            const size_t jp = j + HB_SC_PF * stride;
#pragma unroll
            for (int k = 0; k < NT; k++) {
                if (MODE == FOLD_THEN_POLY) { asm volatile("prefetch.global.L2 [%0];" ::"l"(t.in[k] + 4 * jp)); asm volatile("prefetch.global.L2 [%0];" ::"l"(t.in[k] + 4 * jp + 2)); }
                else if (INTERLEAVED && k == 0) { asm volatile("prefetch.global.L2 [%0];" ::"l"(t.in[0] + 4 * jp)); asm volatile("prefetch.global.L2 [%0];" ::"l"(t.in[0] + 4 * jp + 2)); }
                else if (!(INTERLEAVED && k == 1)) asm volatile("prefetch.global.L2 [%0];" ::"l"(t.in[k] + 2 * jp));
            }
        }
    };
    // one pair whose entries are in registers: round polynomial terms and/or the folded entry
    auto process = [&](size_t j, const F (&x)[NT], const F (&y)[NT]) {
        F d[NT];
#pragma unroll
        for (int k = 0; k < NT; k++) d[k] = fsub(y[k], x[k]);
        if (MODE != FOLD_ONLY) {
            poly_acc<NT>(acc, x, y, d);
            if ((++it & 3) == 0) acc.fold();                         // raw 64-bit sums: at most 4 lazy terms on top of a folded value
        }
        if (MODE == POLY_AND_FOLD || MODE == FOLD_ONLY) {
#pragma unroll
            for (int k = 0; k < NT; k++) t.out[k][j] = fold1n(x[k], d[k], rn);
        }
    };
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (RPF == 2) {
        // two register sets in ping-pong (the loop is unrolled twice so that no set is ever copied)
        F ax[NT], ay[NT], bx[NT], by[NT];
        if (j < L) load_pair(j, ax, ay);
        while (j < L) {
            prefetch_l2(j);
            if (j + stride < L) load_pair(j + stride, bx, by);
            process(j, ax, ay);
            j += stride;
            if (j >= L) break;
            prefetch_l2(j);
            if (j + stride < L) load_pair(j + stride, ax, ay);
            process(j, bx, by);
            j += stride;
        }
    } else if (RPF == 1) {
        F nx[NT], ny[NT];
        if (j < L) load_pair(j, nx, ny);
        for (; j < L; j += stride) {
            F x[NT], y[NT];
            prefetch_l2(j);
#pragma unroll
            for (int k = 0; k < NT; k++) { x[k] = nx[k]; y[k] = ny[k]; }
            if (j + stride < L) load_pair(j + stride, nx, ny);
            process(j, x, y);
        }
    } else {
        for (; j < L; j += stride) {
            F x[NT], y[NT];
            prefetch_l2(j);
            if (MODE == FOLD_THEN_POLY) {
#pragma unroll
                for (int k = 0; k < NT; k++) {
                    const F *p = t.in[k] + 4 * j;
                    const F a = p[0], b = p[1], c = p[2], e = p[3];
                    x[k] = fold1n(a, fsub(b, a), rn); y[k] = fold1n(c, fsub(e, c), rn);
                    t.out[k][2 * j] = x[k]; t.out[k][2 * j + 1] = y[k];
                }
            } else load_pair(j, x, y);
            process(j, x, y);
        }
    }
    if (MODE == FOLD_ONLY) return;
    F accv[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) accv[c] = fcanon2(lfold(acc.get(c)));          // get(): limbs < 2^64 (RegAcc) or < 2^62 (WideAcc)

    grid_reduce<NC>(accv, ra);
}

// ---- S4: streaming folding sumcheck over one product-tree layer (sumcheck.cpp:1093-1136, 1150-1392) -----------------------
// A block of the layer array holds B interleaved pairs (b1, b2) = (blk[2k], blk[2k+1]); b3 is the eq table of the low variables.
// first block: fold tables := block, K = sum f1 f2 f3
__global__ void __launch_bounds__(256)
stream_init_kernel(const F *__restrict__ blk, const F *__restrict__ eq_low, F *__restrict__ f1, F *__restrict__ f2, F *__restrict__ f3, size_t B,
                   RedArgs ra) {
    F acc[1] = {mkF(0, 0)};
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < B; k += (size_t)gridDim.x * blockDim.x) {
        F b1 = blk[2 * k], b2 = blk[2 * k + 1], b3 = eq_low[k];
        f1[k] = b1; f2[k] = b2; f3[k] = b3;
        acc[0] = fadd(acc[0], fmul(fmul(b1, b2), b3));
    }
    grid_reduce<1>(acc, ra);
}
// error terms of folding one more block into (f1,f2,f3)  (batch_prod, :1103-1111):
//   K1 = sum f3 (b1 f2 + b2 f1) + b3 f1 f2 ;  K2 = sum b3 (b1 f2 + b2 f1) + f3 b1 b2 ;  K3 = sum b1 b2 b3
__global__ void __launch_bounds__(256)
stream_err_kernel(const F *__restrict__ f1, const F *__restrict__ f2, const F *__restrict__ f3, const F *__restrict__ blk,
                  const F *__restrict__ eq_low, size_t B, RedArgs ra) {
    F acc[3] = {mkF(0, 0), mkF(0, 0), mkF(0, 0)};
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < B; k += (size_t)gridDim.x * blockDim.x) {
        F b1 = blk[2 * k], b2 = blk[2 * k + 1], b3 = eq_low[k], x1 = f1[k], x2 = f2[k], x3 = f3[k];
        F t1 = fadd(fmul(b1, x2), fmul(b2, x1)), t2 = fmul(b1, b2);
        acc[0] = fadd(acc[0], fadd(fmul(x3, t1), fmul(fmul(b3, x1), x2)));
        acc[1] = fadd(acc[1], fadd(fmul(b3, t1), fmul(x3, t2)));
        acc[2] = fadd(acc[2], fmul(t2, b3));
    }
    grid_reduce<3>(acc, ra);
}
// f += rho * block  (:1130-1135)
__global__ void __launch_bounds__(256)
stream_fold_kernel(F *__restrict__ f1, F *__restrict__ f2, F *__restrict__ f3, const F *__restrict__ blk, const F *__restrict__ eq_low, F rho, size_t B) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < B; k += (size_t)gridDim.x * blockDim.x) {
        f1[k] = fadd(f1[k], fmul(rho, blk[2 * k]));
        f2[k] = fadd(f2[k], fmul(rho, blk[2 * k + 1]));
        f3[k] = fadd(f3[k], fmul(rho, eq_low[k]));
    }
}
// pass B (:1327-1340): PE0[g] = sum_j beta[j] A[2(gB+j)], PE1[g] = sum_j beta[j] A[2(gB+j)+1]; grid (parts, nb), out[(g*parts+part)*2 + {0,1}]
__global__ void __launch_bounds__(256)
partial_evals_kernel(const F *__restrict__ A, const F *__restrict__ beta, size_t B, size_t j0, size_t jn, F *__restrict__ out) {
    __shared__ F sred[8][2];
    const F *blk = A + 2 * (size_t)blockIdx.y * B;
    F a0 = mkF(0, 0), a1 = mkF(0, 0);
    for (size_t jj = (size_t)blockIdx.x * blockDim.x + threadIdx.x; jj < jn; jj += (size_t)gridDim.x * blockDim.x) {
        const size_t j = j0 + jj;
        F bj = beta[j];
        a0 = fadd(a0, fmul(bj, blk[2 * j])); a1 = fadd(a1, fmul(bj, blk[2 * j + 1]));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { a0 = fadd(a0, shfl_down_F(a0, d)); a1 = fadd(a1, shfl_down_F(a1, d)); }
    if (lane == 0) { sred[warp][0] = a0; sred[warp][1] = a1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        F v = sred[0][threadIdx.x];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) v = fadd(v, sred[w][threadIdx.x]);
        out[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 + threadIdx.x] = v;
    }
}

// generate_claims_opt (:1014-1055): per half-chunk g the sum over k of eq_low[k] A[2(gB+k)] A[2(gB+k)+1]; grid (parts, nb), out[g*parts+part]
__global__ void __launch_bounds__(256)
layer_claim_kernel(const F *__restrict__ A, const F *__restrict__ eq_low, size_t B, size_t j0, size_t jn, F *__restrict__ out) {
    __shared__ F sred[8];
    const F *blk = A + 2 * (size_t)blockIdx.y * B;
    F a = mkF(0, 0);
    for (size_t jj = (size_t)blockIdx.x * blockDim.x + threadIdx.x; jj < jn; jj += (size_t)gridDim.x * blockDim.x) {
        const size_t j = j0 + jj;
        a = fadd(a, fmul(fmul(eq_low[j], blk[2 * j]), blk[2 * j + 1]));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a = fadd(a, shfl_down_F(a, d));
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) a = fadd(a, sred[w]);
        out[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = a;
    }
}

// ---- gate consistency (sumcheck.cpp:434-501 and the final phase of :796-981): degree-4 round polynomial of
//      beta(X) * ( mul(X) L(X) R(X) + add(X) (L(X) + R(X)) - O(X) ),  mul = 1 - add  (so mul needs no table of its own).
// Tables: 0 add, 1 beta, 2 L, 3 R, 4 O.  Same fused fold-then-accumulate structure as sc_round_kernel.
// General form (streaming variant, :877-935): beta(X) * ( w.a2 * mul(X) L(X) R(X) + add(X) (w.a0 L(X) + w.a1 R(X)) + w.a3 O(X) ) with
// mul = w.c - add (the folded selector tables satisfy fold_mul = (sum of the fold challenges) - fold_add).  The in-memory prover is
// a0 = a1 = a2 = 1, a3 = -1, c = 1.
struct GateW { F a0, a1, a2, a3, c; };
template <int MODE>
__global__ void __launch_bounds__(256)
gate_round_kernel(Tabs<5> t, size_t L, F r, GateW w, RedArgs ra) {
    F acc[5];
#pragma unroll
    for (int c = 0; c < 5; c++) acc[c] = mkF(0, 0);
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < L; j += (size_t)gridDim.x * blockDim.x) {
        F x[5], y[5];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            if (MODE == FOLD_THEN_POLY) {
                const F *p = t.in[k] + 4 * j;
                F a = p[0], b = p[1], c = p[2], d = p[3];
                x[k] = fold1(a, b, r); y[k] = fold1(c, d, r);
                t.out[k][2 * j] = x[k]; t.out[k][2 * j + 1] = y[k];
            } else { x[k] = t.in[k][2 * j]; y[k] = t.in[k][2 * j + 1]; }
        }
        F a0 = x[0], a1 = fsub(y[0], x[0]), m0 = fmul(w.a2, fsub(w.c, a0)), m1 = fmul(w.a2, fneg(a1));     // a2 * mul(X)
        F b0 = x[1], b1 = fsub(y[1], x[1]), l0 = x[2], l1 = fsub(y[2], x[2]), r0 = x[3], r1 = fsub(y[3], x[3]), o0 = x[4], o1 = fsub(y[4], x[4]);
        F ml2 = fmul(m1, l1), ml1 = fadd(fmul(m1, l0), fmul(m0, l1)), ml0 = fmul(m0, l0);
        F q3 = fmul(ml2, r1);
        F q2 = fadd(fmul(ml2, r0), fmul(ml1, r1));
        F q1 = fadd(fmul(ml1, r0), fmul(ml0, r1));
        F q0 = fmul(ml0, r0);
        F s0 = fadd(fmul(w.a0, l0), fmul(w.a1, r0)), s1 = fadd(fmul(w.a0, l1), fmul(w.a1, r1));
        q2 = fadd(q2, fmul(a1, s1));
        q1 = fadd(fadd(q1, fadd(fmul(a1, s0), fmul(a0, s1))), fmul(w.a3, o1));
        q0 = fadd(fadd(q0, fmul(a0, s0)), fmul(w.a3, o0));
        acc[0] = fadd(acc[0], fmul(b1, q3));
        acc[1] = fadd(acc[1], fadd(fmul(b1, q2), fmul(b0, q3)));
        acc[2] = fadd(acc[2], fadd(fmul(b1, q1), fmul(b0, q2)));
        acc[3] = fadd(acc[3], fadd(fmul(b1, q0), fmul(b0, q1)));
        acc[4] = fadd(acc[4], fmul(b0, q0));
    }
    grid_reduce<5>(acc, ra);
}

// ---- S7 streaming pass (sumcheck.cpp:796-870): fold tables 0 add(S), 1 beta, 2 L, 3 R, 4 O; fold_mul = csum - fold_add ---------
// chunk 0: folds := chunk, Kf_O, Kf_L, Kf_R, Kf_M (:815-824)
__global__ void __launch_bounds__(256)
gs_init_kernel(const F *__restrict__ L, const F *__restrict__ R, const F *__restrict__ O, const F *__restrict__ S, const F *__restrict__ beta,
               F *__restrict__ fA, F *__restrict__ fB, F *__restrict__ fL, F *__restrict__ fR, F *__restrict__ fO, size_t B,
               RedArgs ra) {
    F acc[4] = {mkF(0, 0), mkF(0, 0), mkF(0, 0), mkF(0, 0)};
    const F one = mkF(1, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (size_t)gridDim.x * blockDim.x) {
        F l = L[i], r = R[i], o = O[i], s = S[i], b = beta[i];
        fA[i] = s; fB[i] = b; fL[i] = l; fR[i] = r; fO[i] = o;
        F bl = fmul(b, l), br = fmul(b, r);
        acc[0] = fadd(acc[0], fmul(b, o));
        acc[1] = fadd(acc[1], fmul(bl, s));
        acc[2] = fadd(acc[2], fmul(br, s));
        acc[3] = fadd(acc[3], fmul(fmul(br, l), fsub(one, s)));
    }
    grid_reduce<4>(acc, ra);
}
// the 12 error terms of folding one more chunk (compute2p/3p/4p_error_terms, :374-432): K1_O K2_O | K1_L K2_L K3_L | K1_R K2_R K3_R | K1_M..K4_M
__global__ void __launch_bounds__(256)
gs_err_kernel(const F *__restrict__ bL, const F *__restrict__ bR, const F *__restrict__ bO, const F *__restrict__ bS, const F *__restrict__ beta,
              const F *__restrict__ fA, const F *__restrict__ fB, const F *__restrict__ fL, const F *__restrict__ fR, const F *__restrict__ fO,
              F csum, size_t B, RedArgs ra) {
    F acc[12];
#pragma unroll
    for (int c = 0; c < 12; c++) acc[c] = mkF(0, 0);
    const F one = mkF(1, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (size_t)gridDim.x * blockDim.x) {
        F xl = bL[i], xr = bR[i], xo = bO[i], g = bS[i], ng = fsub(one, g), be = beta[i];
        F a = fA[i], fb = fB[i], l = fL[i], r = fR[i], o = fO[i], m = fsub(csum, a);
        acc[0] = fadd(acc[0], fadd(fmul(xo, fb), fmul(be, o)));
        acc[1] = fadd(acc[1], fmul(xo, be));
        F t1 = fadd(fmul(xl, a), fmul(g, l)), t2 = fmul(xl, g);
        acc[2] = fadd(acc[2], fadd(fmul(fb, t1), fmul(fmul(be, l), a)));
        acc[3] = fadd(acc[3], fadd(fmul(be, t1), fmul(fb, t2)));
        acc[4] = fadd(acc[4], fmul(t2, be));
        t1 = fadd(fmul(xr, a), fmul(g, r)); t2 = fmul(xr, g);
        acc[5] = fadd(acc[5], fadd(fmul(fb, t1), fmul(fmul(be, r), a)));
        acc[6] = fadd(acc[6], fadd(fmul(be, t1), fmul(fb, t2)));
        acc[7] = fadd(acc[7], fmul(t2, be));
        F u1 = fadd(fmul(l, xr), fmul(r, xl)), u2 = fadd(fmul(fb, ng), fmul(m, be));
        F u3 = fmul(xl, xr), u4 = fmul(ng, be), u5 = fmul(l, r), u6 = fmul(fb, m);
        acc[8] = fadd(acc[8], fadd(fmul(u1, u6), fmul(u2, u5)));
        acc[9] = fadd(acc[9], fadd(fadd(fmul(u1, u2), fmul(u3, u6)), fmul(u4, u5)));
        acc[10] = fadd(acc[10], fadd(fmul(u1, u4), fmul(u2, u3)));
        acc[11] = fadd(acc[11], fmul(u3, u4));
    }
    grid_reduce<12>(acc, ra);
}
__global__ void __launch_bounds__(256)
gs_fold_kernel(const F *__restrict__ bL, const F *__restrict__ bR, const F *__restrict__ bO, const F *__restrict__ bS, const F *__restrict__ beta,
               F *__restrict__ fA, F *__restrict__ fB, F *__restrict__ fL, F *__restrict__ fR, F *__restrict__ fO, F rho, size_t B) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (size_t)gridDim.x * blockDim.x) {
        fA[i] = fadd(fA[i], fmul(rho, bS[i])); fB[i] = fadd(fB[i], fmul(rho, beta[i]));
        fL[i] = fadd(fL[i], fmul(rho, bL[i])); fR[i] = fadd(fR[i], fmul(rho, bR[i])); fO[i] = fadd(fO[i], fmul(rho, bO[i]));
    }
}
// pass B (:943-955): per chunk c the dot products of eq(sumcheck_rand) with L, R, O, S; grid (parts, nch), out[(c*parts+part)*4 + k];
// chunk-independent sums (sum beta1, sum beta1*beta) are appended by part 0.. of an extra pseudo-chunk (blockIdx.y == nch).
__global__ void __launch_bounds__(256)
gs_peval_kernel(const F *__restrict__ L, const F *__restrict__ R, const F *__restrict__ O, const F *__restrict__ S, const F *__restrict__ beta,
                const F *__restrict__ beta1, size_t B, size_t j0, size_t jn, unsigned nch, F *__restrict__ out) {
    __shared__ F sred[8][4];
    const size_t base = (size_t)blockIdx.y * B;
    const bool extra = blockIdx.y == nch;
    F a[4] = {mkF(0, 0), mkF(0, 0), mkF(0, 0), mkF(0, 0)};
    for (size_t jj = (size_t)blockIdx.x * blockDim.x + threadIdx.x; jj < jn; jj += (size_t)gridDim.x * blockDim.x) {
        const size_t j = j0 + jj;
        F b1 = beta1[j];
        if (extra) { a[0] = fadd(a[0], b1); a[1] = fadd(a[1], fmul(b1, beta[j])); }
        else {
            a[0] = fadd(a[0], fmul(b1, L[base + j])); a[1] = fadd(a[1], fmul(b1, R[base + j]));
            a[2] = fadd(a[2], fmul(b1, O[base + j])); a[3] = fadd(a[3], fmul(b1, S[base + j]));
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a[k] = fadd(a[k], shfl_down_F(a[k], d));
        if (lane == 0) sred[warp][k] = a[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        F v = sred[0][threadIdx.x];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) v = fadd(v, sred[w][threadIdx.x]);
        out[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = v;
    }
}

// ---- S8: prove_gate_consistency_lookups (sumcheck.cpp:503-794) ---------------------------------------------------------------------
// Nine fold tables: 0 add_L, 1 add_R, 2 mul, 3 lkp, 4 L, 5 R, 6 O, 7 lkp_O, 8 beta.  Everything a chunk contributes is derived on the fly
// from (L, R, O, S) — S = 0 add / 1 mul / 2 lookup — and lookup_rand[0..1]:
//   gate_L = 1 / 0 / lr0, gate_R = 1 / 0 / lr1, gate_mul = 0 / 1 / 0, gate_lkp = 0 / 0 / 1, lkp_O = lr0 L + lr1 R - O on lookup rows
// (what the reference's in-place rewrites of buff_S to 3, -1, 4 select through compute3p/4p_error_terms, :568-586, :382-432).
struct S8Row { F l, r, o, gL, gR, gM, gK, bK; };
__device__ __forceinline__ S8Row s8_row(const F *L, const F *R, const F *O, const F *S, size_t i, F lr0, F lr1) {
    S8Row x; x.l = L[i]; x.r = R[i]; x.o = O[i];
    const u64 s = S[i].re;
    const F one = mkF(1, 0), zero = mkF(0, 0);
    x.gL = s == 0 ? one : s == 1 ? zero : lr0;
    x.gR = s == 0 ? one : s == 1 ? zero : lr1;
    x.gM = s == 1 ? one : zero;
    x.gK = s >= 2 ? one : zero;
    x.bK = s >= 2 ? fsub(fadd(fmul(lr0, x.l), fmul(lr1, x.r)), x.o) : zero;
    return x;
}
struct S8Folds { F *t[9]; };
// chunk 0: folds := chunk; Kf_O, Kf_L, Kf_R, Kf_M, Kf_lkp (:538-546)
__global__ void __launch_bounds__(256)
gl_init_kernel(const F *__restrict__ L, const F *__restrict__ R, const F *__restrict__ O, const F *__restrict__ S, const F *__restrict__ beta,
               S8Folds f, F lr0, F lr1, size_t B, RedArgs ra) {
    F acc[5] = {mkF(0, 0), mkF(0, 0), mkF(0, 0), mkF(0, 0), mkF(0, 0)};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (size_t)gridDim.x * blockDim.x) {
        S8Row x = s8_row(L, R, O, S, i, lr0, lr1); F b = beta[i];
        f.t[0][i] = x.gL; f.t[1][i] = x.gR; f.t[2][i] = x.gM; f.t[3][i] = x.gK; f.t[4][i] = x.l; f.t[5][i] = x.r; f.t[6][i] = x.o; f.t[7][i] = x.bK; f.t[8][i] = b;
        F bl = fmul(b, x.l), br = fmul(b, x.r);
        acc[0] = fadd(acc[0], fmul(b, x.o));
        acc[1] = fadd(acc[1], fmul(bl, x.gL));
        acc[2] = fadd(acc[2], fmul(br, x.gR));
        acc[3] = fadd(acc[3], fmul(fmul(br, x.l), x.gM));
        acc[4] = fadd(acc[4], fmul(fmul(b, x.bK), x.gK));
    }
    grid_reduce<5>(acc, ra);
}
__device__ __forceinline__ void s8_err3(F b1, F gate, F f1, F f2, F fb, F be, F &k1, F &k2, F &k3) {
    F t1 = fadd(fmul(b1, f2), fmul(gate, f1)), t2 = fmul(b1, gate);
    k1 = fadd(k1, fadd(fmul(fb, t1), fmul(fmul(be, f1), f2)));
    k2 = fadd(k2, fadd(fmul(be, t1), fmul(fb, t2)));
    k3 = fadd(k3, fmul(t2, be));
}
// the 15 error terms of folding one more chunk: K1_O K2_O | K1..3_L | K1..3_R | K1..3_lkp | K1..4_M  (:560-588)
__global__ void __launch_bounds__(256)
gl_err_kernel(const F *__restrict__ L, const F *__restrict__ R, const F *__restrict__ O, const F *__restrict__ S, const F *__restrict__ beta,
              S8Folds f, F lr0, F lr1, size_t B, RedArgs ra) {
    F acc[15];
#pragma unroll
    for (int c = 0; c < 15; c++) acc[c] = mkF(0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (size_t)gridDim.x * blockDim.x) {
        S8Row x = s8_row(L, R, O, S, i, lr0, lr1); const F be = beta[i];
        const F fAL = f.t[0][i], fAR = f.t[1][i], fM = f.t[2][i], fK = f.t[3][i], fL = f.t[4][i], fR = f.t[5][i], fO = f.t[6][i], fKO = f.t[7][i], fB = f.t[8][i];
        acc[0] = fadd(acc[0], fadd(fmul(x.o, fB), fmul(be, fO)));
        acc[1] = fadd(acc[1], fmul(x.o, be));
        s8_err3(x.l, x.gL, fL, fAL, fB, be, acc[2], acc[3], acc[4]);
        s8_err3(x.r, x.gR, fR, fAR, fB, be, acc[5], acc[6], acc[7]);
        s8_err3(x.bK, x.gK, fKO, fK, fB, be, acc[8], acc[9], acc[10]);
        F u1 = fadd(fmul(fL, x.r), fmul(fR, x.l)), u2 = fadd(fmul(fB, x.gM), fmul(fM, be));
        F u3 = fmul(x.l, x.r), u4 = fmul(x.gM, be), u5 = fmul(fL, fR), u6 = fmul(fB, fM);
        acc[11] = fadd(acc[11], fadd(fmul(u1, u6), fmul(u2, u5)));
        acc[12] = fadd(acc[12], fadd(fadd(fmul(u1, u2), fmul(u3, u6)), fmul(u4, u5)));
        acc[13] = fadd(acc[13], fadd(fmul(u1, u4), fmul(u2, u3)));
        acc[14] = fadd(acc[14], fmul(u3, u4));
    }
    grid_reduce<15>(acc, ra);
}
__global__ void __launch_bounds__(256)
gl_fold_kernel(const F *__restrict__ L, const F *__restrict__ R, const F *__restrict__ O, const F *__restrict__ S, const F *__restrict__ beta,
               S8Folds f, F lr0, F lr1, F rho, size_t B) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (size_t)gridDim.x * blockDim.x) {
        S8Row x = s8_row(L, R, O, S, i, lr0, lr1);
        const F v[9] = {x.gL, x.gR, x.gM, x.gK, x.l, x.r, x.o, x.bK, beta[i]};
#pragma unroll
        for (int t = 0; t < 9; t++) f.t[t][i] = fadd(f.t[t][i], fmul(rho, v[t]));
    }
}
// degree-4 round polynomial of beta(X) * ( a2 mul L R + a0 add_L L + a1 add_R R + a4 lkp lkp_O + a3 O )   (:646-712)
struct S8W { F a[5]; };
template <int MODE>
__global__ void __launch_bounds__(256)
gatel_round_kernel(Tabs<9> t, size_t L, F r, S8W w, RedArgs ra) {
    F acc[5];
#pragma unroll
    for (int c = 0; c < 5; c++) acc[c] = mkF(0, 0);
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < L; j += (size_t)gridDim.x * blockDim.x) {
        F x[9], d[9];
#pragma unroll
        for (int k = 0; k < 9; k++) {
            F y;
            if (MODE == FOLD_THEN_POLY) {
                const F *p = t.in[k] + 4 * j;
                F a = p[0], b = p[1], c = p[2], e = p[3];
                x[k] = fold1(a, b, r); y = fold1(c, e, r);
                t.out[k][2 * j] = x[k]; t.out[k][2 * j + 1] = y;
            } else { x[k] = t.in[k][2 * j]; y = t.in[k][2 * j + 1]; }
            d[k] = fsub(y, x[k]);
        }
        F m0 = fmul(w.a[2], x[2]), m1 = fmul(w.a[2], d[2]);
        F ml2 = fmul(m1, d[4]), ml1 = fadd(fmul(m1, x[4]), fmul(m0, d[4])), ml0 = fmul(m0, x[4]);
        F q3 = fmul(ml2, d[5]);
        F q2 = fadd(fmul(ml2, x[5]), fmul(ml1, d[5]));
        F q1 = fadd(fmul(ml1, x[5]), fmul(ml0, d[5]));
        F q0 = fmul(ml0, x[5]);
#define HB_S8_TERM(W, SEL, VAL) { F s0 = fmul(w.a[W], x[SEL]), s1 = fmul(w.a[W], d[SEL]);                                     \
            q2 = fadd(q2, fmul(s1, d[VAL])); q1 = fadd(q1, fadd(fmul(s1, x[VAL]), fmul(s0, d[VAL]))); q0 = fadd(q0, fmul(s0, x[VAL])); }
        HB_S8_TERM(0, 0, 4) HB_S8_TERM(1, 1, 5) HB_S8_TERM(4, 3, 7)
#undef HB_S8_TERM
        q1 = fadd(q1, fmul(w.a[3], d[6])); q0 = fadd(q0, fmul(w.a[3], x[6]));
        acc[0] = fadd(acc[0], fmul(d[8], q3));
        acc[1] = fadd(acc[1], fadd(fmul(d[8], q2), fmul(x[8], q3)));
        acc[2] = fadd(acc[2], fadd(fmul(d[8], q1), fmul(x[8], q2)));
        acc[3] = fadd(acc[3], fadd(fmul(d[8], q0), fmul(x[8], q1)));
        acc[4] = fadd(acc[4], fmul(x[8], q0));
    }
    grid_reduce<5>(acc, ra);
}
// pass B (:737-768): per chunk the dot products of eq(sumcheck_rand) with L, R, O, gate_L, gate_R, gate_mul, gate_lkp, lkp_O;
// grid (parts, nch), out[(c*parts+part)*8 + k]
__global__ void __launch_bounds__(256)
gl_peval_kernel(const F *__restrict__ L, const F *__restrict__ R, const F *__restrict__ O, const F *__restrict__ S, const F *__restrict__ beta1,
                F lr0, F lr1, size_t B, size_t j0, size_t jn, F *__restrict__ out) {
    __shared__ F sred[8][8];
    const size_t base = (size_t)blockIdx.y * B;
    F a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = mkF(0, 0);
    for (size_t jj = (size_t)blockIdx.x * blockDim.x + threadIdx.x; jj < jn; jj += (size_t)gridDim.x * blockDim.x) {
        const size_t j = j0 + jj;
        S8Row x = s8_row(L, R, O, S, base + j, lr0, lr1); const F b1 = beta1[j];
        const F v[8] = {x.l, x.r, x.o, x.gL, x.gR, x.gM, x.gK, x.bK};
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = fadd(a[k], fmul(b1, v[k]));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) a[k] = fadd(a[k], shfl_down_F(a[k], dd));
        if (lane == 0) sred[warp][k] = a[k];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        F v = sred[0][threadIdx.x];
        for (int wq = 1; wq < (int)(blockDim.x >> 5); wq++) v = fadd(v, sred[wq][threadIdx.x]);
        out[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + threadIdx.x] = v;
    }
}

// product-tree level: out[j] = in[2j] * in[2j+1]   (sumcheck.cpp:84-101)
__global__ void __launch_bounds__(256) prod_level_kernel(const F *__restrict__ in, F *__restrict__ out, size_t n_out) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += (size_t)gridDim.x * blockDim.x)
        out[j] = fmul(in[2 * j], in[2 * j + 1]);
}

// eq table of <= 2^12 entries by doubling inside one CTA (utils.cpp:251-296: step i uses r[nr-1-i])
__global__ void __launch_bounds__(1024) beta_small_kernel(const F *__restrict__ r, int nr, F *__restrict__ out, F *__restrict__ tmp) {
    F *cur = (nr & 1) ? tmp : out, *nxt = (nr & 1) ? out : tmp;      // so that the last write lands in `out`
    if (threadIdx.x == 0) cur[0] = mkF(1, 0);
    __syncthreads();
    for (int i = 0; i < nr; i++) {
        F ri = r[nr - 1 - i];
        for (unsigned j = threadIdx.x; j < (1u << i); j += blockDim.x) {
            F b = cur[j], tq = fmul(ri, b);
            nxt[2 * j] = fsub(b, tq); nxt[2 * j + 1] = tq;
        }
        __syncthreads();
        F *s = cur; cur = nxt; nxt = s;
    }
}
// out[idx] = hi[idx >> h] * lo[idx & (2^h - 1)]
__global__ void __launch_bounds__(256) beta_combine_kernel(const F *__restrict__ lo, const F *__restrict__ hi, int h, F *__restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = fmul(hi[i >> h], lo[i & (((size_t)1 << h) - 1)]);
}

// ---------------------------------------------------------------------------------------------------------
static inline unsigned grid_for(hb_ctx *ctx, size_t L) {
    size_t g = (L + 255) / 256;
    size_t cap = std::min<size_t>((size_t)ctx->sm_count * 4, kMaxRedBlocks);
    return (unsigned)std::max<size_t>(1, std::min(g, cap));
}

template <int NT, int MODE, bool IL>
static int launch_round(hb_ctx *ctx, const Tabs<NT> &t, size_t L, F r, F *coeffs_host /* NT+1, may be null for FOLD_ONLY */) {
    HB_TRY(ensure_scratch(ctx));
#ifndef HB_SC_WAVES
#define HB_SC_WAVES (1024 / HB_SC_THREADS)
#endif
    const size_t want = (L + HB_SC_THREADS - 1) / HB_SC_THREADS, cap = std::min<size_t>((size_t)ctx->sm_count * HB_SC_WAVES, kMaxRedBlocks);
    HB_LAUNCH(ctx, (sc_round_kernel<NT, MODE, IL>), (unsigned)std::max<size_t>(1, std::min(want, cap)), HB_SC_THREADS, 0, t, L, r, red_args(ctx));
    if (MODE != FOLD_ONLY) { HB_TRY(read_result(ctx, NT + 1, coeffs_host)); coeffs_from_evals<NT>(coeffs_host); }
    return 0;
}

static inline F h_mimc(F in, F k) { hb_F a{in.re, in.im}, b{k.re, k.im}, o; hb_mimc_hash(&a, &b, &o); return mkF(o.real, o.img); }
static inline F h_fold(F x, F y, F r) { return fadd(x, h_fmul(r, fsub(y, x))); }
static inline hb_F toabi(F x) { hb_F o{x.re, x.im}; return o; }

// download k scalars that sit at the head of k device tables
static int fetch_heads(hb_ctx *ctx, F *const *tabs, int k, F *out) {
    for (int i = 0; i < k; i++) HB_CHECK(ctx, cudaMemcpyAsync(ctx->mailbox + kMailHeads + i, tabs[i], sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < k; i++) out[i] = ctx->mailbox[kMailHeads + i];
    transcript_absorb(ctx, out, k);
    return 0;
}

// ---- multi-GPU slicing of a set of tables inside a prover (sumcheck sharded by hypercube prefix) ----------------------------
// While `on`, cur[k] points at this rank's contiguous part of table k and the round sums are added across ranks inside the kernel
// (ctx->dist.reduce_on).  Adjacent-pair folding keeps pairs rank-local until a round would have fewer than one pair per rank; then every
// rank's remaining 1-2 entries per table are all-gathered (rank order == index order) into `small` and the last log2(world) rounds run
// redundantly on every rank.  gsz = global entries per table before this round, L = global pairs of this round.
struct DistRedGuard { hb_ctx *c; DistRedGuard(hb_ctx *ctx, bool on) : c(ctx) { c->dist.reduce_on = on; } ~DistRedGuard() { c->dist.reduce_on = false; } };
static int slice_step(hb_ctx *ctx, bool &on, const F **cur, int nt, size_t gsz, size_t L, F *small, size_t *Lloc) {
    const size_t G = (size_t)ctx->dist.world;
    if (on && L < G) {
        const int cnt = (int)(gsz / G);
        HB_TRY(dist_gather_small(ctx, cur, nt, cnt, small));
        for (int k = 0; k < nt; k++) cur[k] = small + (size_t)k * gsz;
        on = false;
    }
    *Lloc = on ? L / G : L;
    return 0;
}

// eq table on the device (out: 2^nr entries)
int beta_dev(hb_ctx *ctx, const F *r_dev, int nr, F *out) {
    if (nr <= 12) {
        F *tmp; HB_CHECK(ctx, cudaMallocAsync(&tmp, sizeof(F) << nr, ctx->stream));
        HB_LAUNCH(ctx, beta_small_kernel, 1, 1024, 0, r_dev, nr, out, tmp);
        cudaFreeAsync(tmp, ctx->stream);
        return 0;
    }
    int h = nr / 2;                       // bits [0,h) <-> r[0..h) ; bits [h,nr) <-> r[h..nr)
    F *buf; HB_CHECK(ctx, cudaMallocAsync(&buf, (sizeof(F) << h) * 2 + (sizeof(F) << (nr - h)) * 2, ctx->stream));
    F *lo = buf, *lot = lo + ((size_t)1 << h), *hi = lot + ((size_t)1 << h), *hit = hi + ((size_t)1 << (nr - h));
    if (nr - h > 12) { cudaFreeAsync(buf, ctx->stream); HB_FAIL(ctx, "beta: table larger than 2^24 entries is not supported"); }
    HB_LAUNCH(ctx, beta_small_kernel, 1, 1024, 0, r_dev, h, lo, lot);
    HB_LAUNCH(ctx, beta_small_kernel, 1, 1024, 0, r_dev + h, nr - h, hi, hit);
    size_t n = (size_t)1 << nr;
    HB_LAUNCH(ctx, beta_combine_kernel, (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8), 256, 0, lo, hi, h, out, n);
    cudaFreeAsync(buf, ctx->stream);
    return 0;
}

// MLE evaluation by adjacent-pair folding; v_dev is not modified.  Result downloaded into *out.
static int evaluate_dev(hb_ctx *ctx, const F *v_dev, size_t n, const F *r_host, F *out) {
    int nr = ilog2(n);
    if (nr == 0) { F *p = const_cast<F *>(v_dev); return fetch_heads(ctx, &p, 1, out); }
    F *buf; HB_CHECK(ctx, cudaMallocAsync(&buf, (n / 2 + n / 4 + 1) * sizeof(F), ctx->stream));
    F *a = buf, *b = buf + n / 2;
    const F *cur = v_dev;
    for (int i = 0; i < nr; i++) {
        size_t L = n >> (i + 1);
        Tabs<1> t; t.in[0] = cur; t.out[0] = a;
        int rc = launch_round<1, FOLD_ONLY, false>(ctx, t, L, r_host[i], nullptr);
        if (rc) { cudaFreeAsync(buf, ctx->stream); return rc; }
        cur = a; std::swap(a, b);
    }
    F *p = const_cast<F *>(cur);
    int rc = fetch_heads(ctx, &p, 1, out);
    cudaFreeAsync(buf, ctx->stream);
    return rc;
}

// S2 on device tables (tables 0/1 optionally interleaved in `il`): writes the flat proof, returns ps increment.
// v[k] are read-only; scratch holds NT * (n/2 + n/4) elements.
static int sumcheck3_dev(hb_ctx *ctx, const F *v1, const F *v2, const F *v3, const F *il, size_t n, F prev_r,
                         F *scratch, hb_F *proof, double *ps) {
    int rounds = ilog2(n);
    F rand = prev_r;
    Slice sl = dist_slice(ctx, n);
    bool on = sl.on;
    const F *cur[3] = {v1 ? v1 + sl.off : nullptr, v2 ? v2 + sl.off : nullptr, v3 + sl.off};
    if (il) il += 2 * sl.off;
    F *bufA[3], *bufB[3];
    for (int k = 0; k < 3; k++) { bufA[k] = scratch + (size_t)k * (n / 2 + n / 4); bufB[k] = bufA[k] + n / 2; }
    F *small = nullptr;
    if (on) HB_CHECK(ctx, cudaMallocAsync(&small, 3 * 2 * kMaxRanks * sizeof(F), ctx->stream));
    auto done = [&](int rc) { if (small) cudaFreeAsync(small, ctx->stream); return rc; };
    hb_F *rs = proof + 4 * rounds;
    for (int i = 0; i < rounds; i++) {
        size_t L = n >> (i + 1), Ll;
        F co[4];
        int rc = slice_step(ctx, on, cur, 3, 2 * L, L, small, &Ll);
        if (rc) return done(rc);
        Tabs<3> t;
        for (int k = 0; k < 3; k++) { t.in[k] = cur[k]; t.out[k] = bufA[k]; }
        {
            DistRedGuard g(ctx, on);
            if (i == 0 && il) { t.in[0] = il; t.in[1] = nullptr; rc = launch_round<3, POLY_AND_FOLD, true>(ctx, t, Ll, rand, co); }
            else rc = launch_round<3, POLY_AND_FOLD, false>(ctx, t, Ll, rand, co);
        }
        if (rc) return done(rc);
        rs[i] = toabi(rand);
        for (int c = 0; c < 4; c++) { proof[4 * i + c] = toabi(co[c]); rand = h_mimc(rand, co[c]); }
        *ps += 5 * 16 / 1024.0;
        for (int k = 0; k < 3; k++) { cur[k] = bufA[k]; std::swap(bufA[k], bufB[k]); }
    }
    F vr[3];
    if (rounds == 0 && il) { done(0); HB_FAIL(ctx, "sumcheck3: interleaved input needs n >= 2"); }
    F *heads[3] = {const_cast<F *>(cur[0]), const_cast<F *>(cur[1]), const_cast<F *>(cur[2])};
    { int rc = fetch_heads(ctx, heads, 3, vr); if (rc) return done(rc); }
    done(0);
    rand = h_mimc(rand, vr[0]); rand = h_mimc(rand, vr[1]);
    *ps += 3 * 16 / 1024.0;
    hb_F *o = proof + 5 * rounds;
    o[0] = toabi(vr[0]); o[1] = toabi(vr[1]); o[2] = toabi(vr[2]); o[3] = toabi(rand);
    return 0;
}

}  // namespace hb

using namespace hb;

// =========================================================================================================
extern "C" int hb_precompute_beta(hb_ctx *ctx, const hb_F *r, int nr, hb_F *out) { HB_DEV(ctx);
    if (nr < 0 || nr > 24) HB_FAIL(ctx, "hb_precompute_beta: nr out of range");
    Staged sr(ctx), so(ctx);
    HB_TRY(sr.in(r, (size_t)std::max(nr, 1) * sizeof(F)));
    HB_TRY(so.outbuf(out, sizeof(F) << nr));
    HB_TRY(beta_dev(ctx, sr.as<F>(), nr, so.as<F>()));
    HB_TRY(so.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_evaluate_vector(hb_ctx *ctx, const hb_F *v, size_t n, const hb_F *r, hb_F *out) { HB_DEV(ctx);
    if (n == 0 || (n & (n - 1))) HB_FAIL(ctx, "hb_evaluate_vector: n must be a power of two");
    HB_TRY(ensure_scratch(ctx));
    Staged sv(ctx);
    HB_TRY(sv.in(v, n * sizeof(F)));
    int nr = ilog2(n);
    std::vector<F> rh(std::max(nr, 1));
    if (nr) HB_CHECK(ctx, cudaMemcpy(rh.data(), r, nr * sizeof(F), cudaMemcpyDefault));
    F e; HB_TRY(evaluate_dev(ctx, sv.as<F>(), n, rh.data(), &e));
    hb_F eo = toabi(e);
    HB_CHECK(ctx, cudaMemcpy(out, &eo, sizeof(F), cudaMemcpyDefault));
    return 0;
}

extern "C" int hb_sumcheck2(hb_ctx *ctx, const hb_F *v1, const hb_F *v2, size_t n, const hb_F *prev_r, hb_F *proof, double *ps) { HB_DEV(ctx);
    if (n == 0 || (n & (n - 1))) HB_FAIL(ctx, "hb_sumcheck2: n must be a power of two");
    HB_TRY(ensure_scratch(ctx));
    int rounds = ilog2(n);
    Staged a(ctx), b(ctx);
    HB_TRY(a.in(v1, n * sizeof(F))); HB_TRY(b.in(v2, n * sizeof(F)));
    F *scratch; HB_CHECK(ctx, cudaMallocAsync(&scratch, 2 * (n / 2 + n / 4 + 2) * sizeof(F), ctx->stream));
    F *bufA[2] = {scratch, scratch + (n / 2 + n / 4 + 2)}, *bufB[2] = {bufA[0] + n / 2 + 1, bufA[1] + n / 2 + 1};
    const F *cur[2] = {a.as<F>(), b.as<F>()};
    F rand = mkF(prev_r->real, prev_r->img);
    hb_F *rs = proof + 3 * rounds;
    int rc = 0;
    for (int i = 0; i < rounds && !rc; i++) {
        size_t L = n >> (i + 1);
        F co[3];
        Tabs<2> t;
        for (int k = 0; k < 2; k++) { t.in[k] = cur[k]; t.out[k] = bufA[k]; }
        if (i == 0) rc = launch_round<2, POLY_ONLY, false>(ctx, t, L, rand, co);
        else {
            rc = launch_round<2, FOLD_THEN_POLY, false>(ctx, t, L, rand, co);
            for (int k = 0; k < 2; k++) { cur[k] = bufA[k]; std::swap(bufA[k], bufB[k]); }
        }
        for (int c = 0; c < 3; c++) { proof[3 * i + c] = toabi(co[c]); rand = h_mimc(rand, co[c]); }
        rs[i] = toabi(rand);
        *ps += 3 * 16 / 1024.0;
    }
    F vr[2];
    if (!rc && rounds > 0) {          // the last challenge folds the final pair
        Tabs<2> t;
        for (int k = 0; k < 2; k++) { t.in[k] = cur[k]; t.out[k] = bufA[k]; }
        rc = launch_round<2, FOLD_ONLY, false>(ctx, t, 1, rand, nullptr);
        for (int k = 0; k < 2; k++) cur[k] = bufA[k];
    }
    if (!rc) { F *heads[2] = {const_cast<F *>(cur[0]), const_cast<F *>(cur[1])}; rc = fetch_heads(ctx, heads, 2, vr); }
    cudaFreeAsync(scratch, ctx->stream);
    if (rc) return rc;
    rand = h_mimc(rand, vr[0]); rand = h_mimc(rand, vr[1]);
    *ps += 2 * 16 / 1024.0;
    hb_F *o = proof + 4 * rounds;
    o[0] = toabi(vr[0]); o[1] = toabi(vr[1]); o[2] = toabi(rand);
    return 0;
}

extern "C" int hb_sumcheck3(hb_ctx *ctx, const hb_F *v1, const hb_F *v2, const hb_F *v3, size_t n, const hb_F *prev_r,
                            hb_F *proof, double *ps) { HB_DEV(ctx);
    if (n == 0 || (n & (n - 1))) HB_FAIL(ctx, "hb_sumcheck3: n must be a power of two");
    HB_TRY(ensure_scratch(ctx));
    Staged a(ctx), b(ctx), c(ctx);
    HB_TRY(a.in(v1, n * sizeof(F))); HB_TRY(b.in(v2, n * sizeof(F))); HB_TRY(c.in(v3, n * sizeof(F)));
    F *scratch; HB_CHECK(ctx, cudaMallocAsync(&scratch, (3 * (n / 2 + n / 4) + 4) * sizeof(F), ctx->stream));
    int rc = sumcheck3_dev(ctx, a.as<F>(), b.as<F>(), c.as<F>(), nullptr, n, mkF(prev_r->real, prev_r->img), scratch, proof, ps);
    cudaFreeAsync(scratch, ctx->stream);
    return rc;
}

// S3 on DEVICE tables (batch b at d1/d2/d3 + sum of earlier sizes); a: host.  Inputs are not modified.
static int batch_sumcheck3_dev(hb_ctx *ctx, const F *d1, const F *d2, const F *d3, const size_t *sizes, int batches,
                               const F *a_host, hb_F *proof, double *ps) {
    HB_TRY(ensure_scratch(ctx));
    size_t tot = 0, Lmax = 0;
    for (int b = 0; b < batches; b++) {
        if (sizes[b] == 0 || (sizes[b] & (sizes[b] - 1))) HB_FAIL(ctx, "hb_batch_sumcheck3: batch sizes must be powers of two");
        tot += sizes[b]; Lmax = std::max(Lmax, sizes[b]);
    }
    int rounds = ilog2(Lmax);
    std::vector<F> a(a_host, a_host + batches);
    // per batch: ping-pong scratch of size/2 + size/4 per table
    F *scratch; HB_CHECK(ctx, cudaMallocAsync(&scratch, (3 * tot + 8) * sizeof(F), ctx->stream));
    struct Bt { const F *cur[3]; F *A[3], *B[3]; size_t size; bool exhausted; F x[3]; bool on; F *small; };
    std::vector<Bt> bt(batches);
    size_t off = 0, soff = 0;
    F *small_all = nullptr;
    for (int b = 0; b < batches; b++) {
        Bt &q = bt[b]; q.size = sizes[b]; q.exhausted = false;
        const Slice sl = dist_slice(ctx, sizes[b]);                   // every batch is sliced on its own (its tables may hold only this rank's part)
        q.on = sl.on; q.small = nullptr;
        const F *base[3] = {d1 + off + sl.off, d2 + off + sl.off, d3 + off + sl.off};
        for (int k = 0; k < 3; k++) { q.cur[k] = base[k]; q.A[k] = scratch + soff; q.B[k] = q.A[k] + sizes[b] / 2; soff += sizes[b] / 2 + sizes[b] / 4 + 1; }
        off += sizes[b];
        if (q.on && !small_all) HB_CHECK(ctx, cudaMallocAsync(&small_all, (size_t)batches * 3 * 2 * kMaxRanks * sizeof(F), ctx->stream));
        if (q.on) q.small = small_all + (size_t)b * 3 * 2 * kMaxRanks;
    }
    const F unset = mkF(P61 - 1, 0);                  // F(-1), the reference's "not yet set" marker (sumcheck.cpp:289)
    std::vector<F> vr(3 * batches, unset);
    F rand = mkF(312, 0);
    hb_F *rs = proof + 4 * rounds;
    int rc = 0;
    auto cubic_host = [](F *p, const F *x) {          // (1-t)^3 * x1 x2 x3 : linear factors (-x t + x)
        F m = h_fmul(h_fmul(x[0], x[1]), x[2]);
        F m3 = fadd(fadd(m, m), m);
        p[0] = fneg(m); p[1] = m3; p[2] = fneg(m3); p[3] = m;
    };
    for (int i = 0; i < rounds && !rc; i++) {
        F poly[4] = {mkF(0, 0), mkF(0, 0), mkF(0, 0), mkF(0, 0)};
        for (int b = 0; b < batches && !rc; b++) {
            Bt &q = bt[b];
            int lg = ilog2(q.size) - 1 - i;
            F p[4];
            if (lg >= 0) {
                size_t L = (size_t)1 << lg, Ll;
                rc = slice_step(ctx, q.on, q.cur, 3, i == 0 ? 2 * L : 4 * L, L, q.small, &Ll);
                if (rc) break;
                Tabs<3> t;
                for (int k = 0; k < 3; k++) { t.in[k] = q.cur[k]; t.out[k] = q.A[k]; }
                DistRedGuard g(ctx, q.on);
                if (i == 0) rc = launch_round<3, POLY_ONLY, false>(ctx, t, Ll, rand, p);
                else {
                    rc = launch_round<3, FOLD_THEN_POLY, false>(ctx, t, Ll, rand, p);
                    for (int k = 0; k < 3; k++) { q.cur[k] = q.A[k]; std::swap(q.A[k], q.B[k]); }
                }
            } else {
                if (!q.exhausted) {
                    // first exhausted round: materialise the single remaining value of each table
                    if (q.size >= 2) {            // fold the last pair with the previous round's challenge
                        Tabs<3> t;
                        for (int k = 0; k < 3; k++) { t.in[k] = q.cur[k]; t.out[k] = q.A[k]; }
                        rc = launch_round<3, FOLD_ONLY, false>(ctx, t, 1, rand, nullptr);
                        for (int k = 0; k < 3; k++) q.cur[k] = q.A[k];
                    }
                    F *heads[3] = {const_cast<F *>(q.cur[0]), const_cast<F *>(q.cur[1]), const_cast<F *>(q.cur[2])};
                    if (!rc) rc = fetch_heads(ctx, heads, 3, q.x);
                    q.exhausted = true;
                    // every later round scales by (1 - rand); rounds already passed since exhaustion: none (this is the first)
                }
                if (feq(vr[3 * b], unset)) { vr[3 * b] = q.x[0]; vr[3 * b + 1] = q.x[1]; vr[3 * b + 2] = q.x[2]; }
                cubic_host(p, q.x);
            }
            for (int c = 0; c < 4; c++) poly[c] = fadd(poly[c], h_fmul(a[b], p[c]));
        }
        if (rc) break;
        for (int c = 0; c < 4; c++) { proof[4 * i + c] = toabi(poly[c]); rand = h_mimc(rand, poly[c]); }
        rs[i] = toabi(rand);
        *ps += 4 * 16 / 1024.0;
        F om = fsub(mkF(1, 0), rand);
        for (int b = 0; b < batches; b++) {
            Bt &q = bt[b];
            if (q.exhausted) for (int k = 0; k < 3; k++) q.x[k] = h_fmul(om, q.x[k]);
        }
    }
    // batches that never ran out: the last challenge folds their final pair
    for (int b = 0; b < batches && !rc; b++) {
        Bt &q = bt[b];
        if (!feq(vr[3 * b], unset)) continue;
        if (!q.exhausted) {
            if (q.size >= 2) {
                Tabs<3> t;
                for (int k = 0; k < 3; k++) { t.in[k] = q.cur[k]; t.out[k] = q.A[k]; }
                rc = launch_round<3, FOLD_ONLY, false>(ctx, t, 1, rand, nullptr);
                for (int k = 0; k < 3; k++) q.cur[k] = q.A[k];
            }
            F *heads[3] = {const_cast<F *>(q.cur[0]), const_cast<F *>(q.cur[1]), const_cast<F *>(q.cur[2])};
            if (!rc) rc = fetch_heads(ctx, heads, 3, q.x);
        }
        vr[3 * b] = q.x[0]; vr[3 * b + 1] = q.x[1]; vr[3 * b + 2] = q.x[2];
    }
    if (small_all) cudaFreeAsync(small_all, ctx->stream);
    cudaFreeAsync(scratch, ctx->stream);
    if (rc) return rc;
    *ps += (3 * batches - batches) * 16 / 1024.0;
    for (int j = 0; j < 3 * batches; j++) proof[5 * rounds + j] = toabi(vr[j]);
    return 0;
}

extern "C" int hb_batch_sumcheck3(hb_ctx *ctx, const hb_F *t1, const hb_F *t2, const hb_F *t3, const size_t *sizes, int batches,
                                  const hb_F *a_in, hb_F *proof, double *ps) { HB_DEV(ctx);
    size_t tot = 0;
    for (int b = 0; b < batches; b++) tot += sizes[b];
    Staged s1(ctx), s2(ctx), s3(ctx);
    HB_TRY(s1.in(t1, tot * sizeof(F))); HB_TRY(s2.in(t2, tot * sizeof(F))); HB_TRY(s3.in(t3, tot * sizeof(F)));
    std::vector<F> a(batches);
    HB_CHECK(ctx, cudaMemcpy(a.data(), a_in, batches * sizeof(F), cudaMemcpyDefault));
    return batch_sumcheck3_dev(ctx, s1.as<F>(), s2.as<F>(), s3.as<F>(), sizes, batches, a.data(), proof, ps);
}

extern "C" int hb_mul_tree(hb_ctx *ctx, const hb_F *input, int vectors, size_t n, const hb_F *prev_r, const hb_F *x_rand,
                           hb_F *out, size_t *written, int *nfr, double *ps) { HB_DEV(ctx);
    if (n < 2 || (n & (n - 1)) || vectors < 1 || (vectors & (vectors - 1)))
        HB_FAIL(ctx, "hb_mul_tree: vectors and n must be powers of two (pad with F(1)/zero vectors as the reference does)");
    HB_TRY(ensure_scratch(ctx));
    const int depth = ilog2(n);
    const size_t total = (size_t)vectors * n;
    Staged in(ctx);
    HB_TRY(in.in(input, total * sizeof(F)));
    // transcript levels: lvl[0] = input, lvl[i+1][j] = lvl[i][2j]*lvl[i][2j+1]
    F *tree; HB_CHECK(ctx, cudaMallocAsync(&tree, (total + 8) * sizeof(F), ctx->stream));
    std::vector<const F *> lvl(depth + 1);
    lvl[0] = in.as<F>();
    {
        F *p = tree;
        for (int i = 0; i < depth; i++) {
            size_t sz = total >> (i + 1);
            HB_LAUNCH(ctx, prod_level_kernel, (unsigned)std::min<size_t>((sz + 255) / 256, (size_t)ctx->sm_count * 8), 256, 0, lvl[i], p, sz);
            lvl[i + 1] = p; p += sz;
        }
    }
    const int maxr = ilog2(total);
    F *beta, *scratch, *r_dev;
    HB_CHECK(ctx, cudaMallocAsync(&beta, (total / 2 + 1) * sizeof(F), ctx->stream));
    HB_CHECK(ctx, cudaMallocAsync(&scratch, (3 * (total / 4 + total / 8) + 8) * sizeof(F), ctx->stream));
    HB_CHECK(ctx, cudaMallocAsync(&r_dev, (maxr + 1) * sizeof(F), ctx->stream));
    auto cleanup = [&]() { cudaFreeAsync(tree, ctx->stream); cudaFreeAsync(beta, ctx->stream); cudaFreeAsync(scratch, ctx->stream); cudaFreeAsync(r_dev, ctx->stream); };

    std::vector<F> outputs(vectors);
    HB_CHECK(ctx, cudaMemcpyAsync(outputs.data(), lvl[depth], vectors * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    size_t k = 0;
    for (int v = 0; v < vectors; v++) out[k++] = toabi(outputs[v]);

    std::vector<F> r; F previous_r = mkF(prev_r->real, prev_r->img), sum, out_eval;
    std::vector<hb_F> proofs; std::vector<hb_F> pbuf(5 * (size_t)maxr + 8);
    int rc = 0;
    auto run_layer = [&](int i) -> int {
        size_t sz = total >> (i + 1); int rounds = ilog2(sz);
        HB_CHECK(ctx, cudaMemcpyAsync(r_dev, r.data(), r.size() * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
        HB_TRY(beta_dev(ctx, r_dev, (int)r.size(), beta));
        HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));        // r.data() is reused below
        HB_TRY(sumcheck3_dev(ctx, nullptr, nullptr, beta, lvl[i], sz, previous_r, scratch, pbuf.data(), ps));
        proofs.insert(proofs.end(), pbuf.begin(), pbuf.begin() + 4 * rounds);
        proofs.insert(proofs.end(), pbuf.begin() + 5 * rounds, pbuf.begin() + 5 * rounds + 4);
        F vr0 = mkF(pbuf[5 * rounds].real, pbuf[5 * rounds].img), vr1 = mkF(pbuf[5 * rounds + 1].real, pbuf[5 * rounds + 1].img);
        previous_r = mkF(pbuf[5 * rounds + 3].real, pbuf[5 * rounds + 3].img);
        sum = fadd(h_fmul(vr0, fsub(mkF(1, 0), previous_r)), h_fmul(vr1, previous_r));
        r.resize(rounds + 1);
        r[0] = previous_r;
        for (int q = 0; q < rounds; q++) r[q + 1] = mkF(pbuf[4 * rounds + q].real, pbuf[4 * rounds + q].img);
        return 0;
    };
    if (vectors == 1) {
        previous_r = h_mimc(previous_r, outputs[0]);
        sum = outputs[0]; out_eval = sum;
        for (int i = depth - 1; i >= 0 && !rc; i--) {
            if (r.empty()) {
                F pair[2];
                HB_CHECK(ctx, cudaMemcpy(pair, lvl[i], 2 * sizeof(F), cudaMemcpyDeviceToHost));   // in1[i][0], in2[i][0]
                F num = h_mimc(previous_r, pair[0]);
                previous_r = h_mimc(num, pair[1]);
                sum = fadd(h_fmul(fsub(mkF(1, 0), previous_r), pair[0]), h_fmul(previous_r, pair[1]));
                r.push_back(previous_r);
            } else rc = run_layer(i);
        }
    } else {
        int nr0 = ilog2((size_t)vectors);
        r.resize(nr0);
        for (int q = 0; q < nr0; q++) r[q] = mkF(x_rand[q].real, x_rand[q].img);
        // evaluate_vector on the `vectors` outputs — a handful of values already on the host
        std::vector<F> ev(outputs);
        for (int q = 0; q < nr0; q++) for (size_t j = 0; j < ((size_t)vectors >> (q + 1)); j++) ev[j] = h_fold(ev[2 * j], ev[2 * j + 1], r[q]);
        sum = ev[0]; out_eval = sum;
        previous_r = h_mimc(r[nr0 - 1], sum);
        for (int i = depth - 1; i >= 0 && !rc; i--) rc = run_layer(i);
        if (!rc) {      // the reference's closing self-check (sumcheck.cpp:213-216)
            F e; rc = evaluate_dev(ctx, lvl[0], total, r.data(), &e);
            if (!rc && !feq(e, sum)) { cleanup(); HB_FAIL(ctx, "Error in mul tree final"); }
        }
    }
    cleanup();
    if (rc) return rc;
    out[k++] = toabi(out_eval);
    for (auto &x : r) out[k++] = toabi(x);
    out[k++] = toabi(sum);
    for (auto &x : proofs) out[k++] = x;
    *written = k; *nfr = (int)r.size();
    return 0;
}

// =========================================================================================================
// S4 on a layer array A (device, S entries, natural [seg(X) | seg(Y)] order).  rnd4 = (a, b0, b1, pad) drawn by the host with
// libc in the reference's order.  r: log2(S/2) host points.  Outputs on the host.
static inline F fromabi(const hb_F &x) { return mkF(x.real, x.img); }
static int stream_layer_dev(hb_ctx *ctx, const F *A, size_t S, size_t B, const F *r, F old_claim, const F *rnd4,
                            F *new_claim, std::vector<F> &new_r, double *ps) {
    HB_TRY(ensure_scratch(ctx));
    if (S < 4 * B || (S & (S - 1)) || (B & (B - 1))) HB_FAIL(ctx, "stream layer: need a power-of-two layer of at least 4*BUFFER_SPACE entries");
    const size_t nb = S / (2 * B);
    const int lgB = ilog2(B), lgnb = ilog2(nb);
    F *buf; HB_CHECK(ctx, cudaMallocAsync(&buf, (5 * B + nb + 64) * sizeof(F), ctx->stream));
    F *eq_low = buf, *f1 = buf + B, *f2 = f1 + B, *f3 = f2 + B, *beta = f3 + B, *eqh_dev = beta + B, *r_dev = eqh_dev + nb;
    auto fail = [&](int rc) { cudaFreeAsync(buf, ctx->stream); return rc; };
    int rc;
    HB_CHECK(ctx, cudaMemcpyAsync(r_dev, r, (lgB + lgnb) * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = beta_dev(ctx, r_dev, lgB, eq_low))) return fail(rc);
    if ((rc = beta_dev(ctx, r_dev + lgB, lgnb, eqh_dev))) return fail(rc);
    std::vector<F> eq_high(nb);
    HB_CHECK(ctx, cudaMemcpyAsync(eq_high.data(), eqh_dev, nb * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
    // multi-GPU: every rank folds positions [so, so + sn) of each block; the error terms are summed across ranks inside the kernels
    const Slice sl = dist_slice(ctx, B);
    const size_t so = sl.off, sn = sl.len;
    DistRedGuard dist_guard(ctx, sl.on);
    const unsigned grid = grid_for(ctx, sn);
    F Kp;
    HB_LAUNCH(ctx, stream_init_kernel, grid, 256, 0, A + 2 * so, eq_low + so, f1 + so, f2 + so, f3 + so, sn, red_args(ctx));
    if ((rc = read_result(ctx, 1, &Kp))) return fail(rc);                 // also completes the eq_high download
    const F a = rnd4[0];
    F Kf = h_fmul(a, Kp);
    Kp = h_fmul(Kp, eq_high[0]);
    *ps += 2 * 16 / 1024.0;
    std::vector<F> R; R.push_back(mkF(1, 0));
    for (size_t step = 1; step < nb; step++) {
        // processing order of the reference: X0 | Y0, X1, Y1, ...   (natural block index g selects eq_high)
        const size_t g = (step % 2) ? nb / 2 + (step - 1) / 2 : step / 2;
        const F *blk = A + 2 * g * B + 2 * so;
        F K[3];
        HB_LAUNCH(ctx, stream_err_kernel, grid, 256, 0, f1 + so, f2 + so, f3 + so, blk, eq_low + so, sn, red_args(ctx));
        if ((rc = read_result(ctx, 3, K))) return fail(rc);
        F K1 = h_fmul(a, K[0]), K2 = h_fmul(a, K[1]), K3 = K[2];
        F rand = R.back();
        rand = h_mimc(K1, rand); rand = h_mimc(K2, rand); rand = h_mimc(K3, rand);     // argument order (value, rand): N7
        F x1 = rand, x2 = h_fmul(rand, x1), x3 = h_fmul(rand, x2);
        Kp = fadd(Kp, h_fmul(eq_high[g], K3));
        Kf = fadd(Kf, h_fmul(h_fmul(x3, a), K3));
        Kf = fadd(Kf, fadd(h_fmul(x2, K2), h_fmul(x1, K1)));
        R.push_back(rand);
        *ps += 2 * 16 / 1024.0;
        HB_LAUNCH(ctx, stream_fold_kernel, grid, 256, 0, f1 + so, f2 + so, f3 + so, blk, eq_low + so, rand, sn);
    }
    ctx->dist.reduce_on = false;                                              // the provers below slice (and reduce) on their own
    if (!feq(Kp, old_claim)) printf("Error in sumcheck 0 %d\n", 0);           // the reference only warns (sumcheck.cpp:1246-1251)
    std::vector<hb_F> p1(5 * (size_t)lgB + 8);
    size_t szB = B;
    if ((rc = batch_sumcheck3_dev(ctx, f1, f2, f3, &szB, 1, &a, p1.data(), ps))) return fail(rc);
    {
        F s = fadd(fadd(fadd(fromabi(p1[0]), fromabi(p1[1])), fadd(fromabi(p1[2]), fromabi(p1[3]))), fromabi(p1[3]));
        if (!feq(s, Kf)) { cudaFreeAsync(buf, ctx->stream); HB_FAIL(ctx, "Error in sumcheck 1"); }
    }
    // pass B: per-block partial evaluations against eq(P1.randomness[0 .. log2 B))
    std::vector<F> P1r(lgB);
    for (int j = 0; j < lgB; j++) P1r[j] = fromabi(p1[4 * lgB + j]);
    HB_CHECK(ctx, cudaMemcpyAsync(r_dev, P1r.data(), lgB * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = beta_dev(ctx, r_dev, lgB, beta))) return fail(rc);
    unsigned parts = (unsigned)std::max<size_t>(1, std::min<size_t>((sn + 2047) / 2048, (size_t)(2 * ctx->sm_count) / nb + 1));
    F *pe_dev; HB_CHECK(ctx, cudaMallocAsync(&pe_dev, nb * parts * 2 * sizeof(F), ctx->stream));
    HB_LAUNCH(ctx, partial_evals_kernel, dim3(parts, (unsigned)nb), 256, 0, A, beta, B, so, sn, pe_dev);
    if (sl.on) HB_TRY(dist_allreduce_vec(ctx, pe_dev, nb * parts * 2));
    std::vector<F> pe(nb * parts * 2);
    HB_CHECK(ctx, cudaMemcpyAsync(pe.data(), pe_dev, pe.size() * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(pe_dev, ctx->stream);
    cudaFreeAsync(buf, ctx->stream);
    std::vector<F> PE0(nb, mkF(0, 0)), PE1(nb, mkF(0, 0));
    for (size_t g = 0; g < nb; g++) for (unsigned q = 0; q < parts; q++) { PE0[g] = fadd(PE0[g], pe[(g * parts + q) * 2]); PE1[g] = fadd(PE1[g], pe[(g * parts + q) * 2 + 1]); }
    // de-interleave R (permute_partial_evals) -> natural block order
    std::vector<F> Rp; Rp.reserve(nb);
    for (size_t i = 0; i < nb / 2; i++) Rp.push_back(R[2 * i]);
    for (size_t i = 0; i < nb / 2; i++) Rp.push_back(R[2 * i + 1]);
    const F b0 = rnd4[1], b1 = rnd4[2], pad = rnd4[3];
    std::vector<F> aggr(nb);
    for (size_t j = 0; j < nb; j++) aggr[j] = fadd(h_fmul(b0, PE0[j]), h_fmul(b1, PE1[j]));
    std::vector<hb_F> p2(4 * (size_t)lgnb + 8);
    hb_F zero{0, 0};
    HB_TRY(hb_sumcheck2(ctx, (const hb_F *)Rp.data(), (const hb_F *)aggr.data(), nb, &zero, p2.data(), ps));
    {
        F sum = fadd(h_fmul(b0, fromabi(p1[5 * lgB])), h_fmul(b1, fromabi(p1[5 * lgB + 1])));
        F q = fadd(fadd(fromabi(p2[0]), fromabi(p2[1])), fadd(fromabi(p2[2]), fromabi(p2[2])));
        if (!feq(sum, q)) HB_FAIL(ctx, "Error in sumcheck 2");
    }
    new_r.clear(); new_r.push_back(pad);
    for (int j = 0; j < lgB; j++) new_r.push_back(P1r[j]);
    std::vector<F> P2r(lgnb);
    for (int j = 0; j < lgnb; j++) { P2r[j] = fromabi(p2[3 * lgnb + j]); new_r.push_back(P2r[j]); }
    auto eval_small = [&](std::vector<F> v) { for (int q = 0; q < lgnb; q++) for (size_t j = 0; j < (nb >> (q + 1)); j++) v[j] = h_fold(v[2 * j], v[2 * j + 1], P2r[q]); return v[0]; };
    *new_claim = fadd(h_fmul(fsub(mkF(1, 0), pad), eval_small(PE0)), h_fmul(pad, eval_small(PE1)));
    return 0;
}

// The batched form (generate_3product_sumcheck_beta_stream_batch_optimized with batches > 1, as called when layers > distance,
// sumcheck.cpp:1893-1900): batch j is the layer array A[j] (layer_id + j*distance) with block size B >> (j*distance); all batches have
// the same number nb of half-chunks and share the challenges.  r[j]: log2(B_j) + log2(nb) points.  rnd = a[batches] | b[2*batches] | pad.
static int stream_batch_dev(hb_ctx *ctx, const F *const *A, size_t S0, size_t B, int distance, int batches, const std::vector<std::vector<F>> &r,
                            const F *old_claims, const F *rnd, F *new_claims, std::vector<std::vector<F>> &new_r, double *ps) {
    HB_TRY(ensure_scratch(ctx));
    enum { MAXB = 8 };
    if (batches < 1 || batches > MAXB) HB_FAIL(ctx, "stream batch: 1..8 batches");
    if (S0 < 4 * B || (S0 & (S0 - 1)) || (B & (B - 1)) || (B >> ((batches - 1) * distance)) < 2) HB_FAIL(ctx, "stream batch: bad sizes");
    const size_t nb = S0 / (2 * B);
    const int lgnb = ilog2(nb);
    size_t Bj[MAXB], tot = 0; int lgBj[MAXB];
    for (int j = 0; j < batches; j++) { Bj[j] = B >> (j * distance); lgBj[j] = ilog2(Bj[j]); tot += Bj[j]; }
    // per batch: eq_low | f1 | f2 | f3 | beta (5 * B_j) ; folds of all batches contiguous per table for the batched sumcheck: f1[all] f2[all] f3[all]
    F *buf; HB_CHECK(ctx, cudaMallocAsync(&buf, (5 * tot + nb + 64) * sizeof(F), ctx->stream));
    F *eq_low[MAXB], *f1[MAXB], *f2[MAXB], *f3[MAXB], *beta[MAXB];
    { F *p = buf; size_t off = 0;
      for (int j = 0; j < batches; j++) { f1[j] = p + off; f2[j] = p + tot + off; f3[j] = p + 2 * tot + off; eq_low[j] = p + 3 * tot + off; beta[j] = p + 4 * tot + off; off += Bj[j]; } }
    F *eqh_dev = buf + 5 * tot, *r_dev = eqh_dev + nb;
    auto fail = [&](int rc) { cudaFreeAsync(buf, ctx->stream); return rc; };
    int rc;
    std::vector<std::vector<F>> eq_high(batches, std::vector<F>(nb));
    F Kp[MAXB];
    Slice sl[MAXB];                                                         // multi-GPU: positions [off, off + len) of every block of batch j
    for (int j = 0; j < batches; j++) sl[j] = dist_slice(ctx, Bj[j]);
    for (int j = 0; j < batches; j++) {
        if ((int)r[j].size() < lgBj[j] + lgnb) return fail((ctx->err = "stream batch: point too short", 2));
        HB_CHECK(ctx, cudaMemcpyAsync(r_dev, r[j].data(), (lgBj[j] + lgnb) * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = beta_dev(ctx, r_dev, lgBj[j], eq_low[j]))) return fail(rc);
        if ((rc = beta_dev(ctx, r_dev + lgBj[j], lgnb, eqh_dev))) return fail(rc);
        HB_CHECK(ctx, cudaMemcpyAsync(eq_high[j].data(), eqh_dev, nb * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
        {
            DistRedGuard g(ctx, sl[j].on);
            const size_t so = sl[j].off;
            HB_LAUNCH(ctx, stream_init_kernel, grid_for(ctx, sl[j].len), 256, 0, A[j] + 2 * so, eq_low[j] + so, f1[j] + so, f2[j] + so, f3[j] + so, sl[j].len, red_args(ctx));
        }
        if ((rc = read_result(ctx, 1, &Kp[j]))) return fail(rc);
        HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));                 // eq_high[j] is on the host; r_dev / eqh_dev are reused by the next batch
    }
    const F *a = rnd, *b = rnd + batches; const F pad = rnd[3 * batches];
    F Kf = mkF(0, 0);
    for (int j = 0; j < batches; j++) { Kf = fadd(Kf, h_fmul(a[j], Kp[j])); Kp[j] = h_fmul(Kp[j], eq_high[j][0]); }
    *ps += (1 + batches) * 16 / 1024.0;
    std::vector<F> R; R.push_back(mkF(1, 0));
    for (size_t step = 1; step < nb; step++) {
        const size_t g = (step % 2) ? nb / 2 + (step - 1) / 2 : step / 2;
        F K1 = mkF(0, 0), K2 = mkF(0, 0), K3[MAXB];
        for (int j = 0; j < batches; j++) {
            F K[3];
            {
                DistRedGuard gd(ctx, sl[j].on);
                const size_t so = sl[j].off;
                HB_LAUNCH(ctx, stream_err_kernel, grid_for(ctx, sl[j].len), 256, 0, f1[j] + so, f2[j] + so, f3[j] + so, A[j] + 2 * g * Bj[j] + 2 * so, eq_low[j] + so, sl[j].len, red_args(ctx));
            }
            if ((rc = read_result(ctx, 3, K))) return fail(rc);
            K1 = fadd(K1, h_fmul(a[j], K[0])); K2 = fadd(K2, h_fmul(a[j], K[1])); K3[j] = K[2];
        }
        F rand = R.back();
        rand = h_mimc(K1, rand); rand = h_mimc(K2, rand);
        for (int j = 0; j < batches; j++) rand = h_mimc(K3[j], rand);
        F x1 = rand, x2 = h_fmul(rand, x1), x3 = h_fmul(rand, x2);
        for (int j = 0; j < batches; j++) { Kp[j] = fadd(Kp[j], h_fmul(eq_high[j][g], K3[j])); Kf = fadd(Kf, h_fmul(h_fmul(x3, a[j]), K3[j])); }
        Kf = fadd(Kf, fadd(h_fmul(x2, K2), h_fmul(x1, K1)));
        R.push_back(rand);
        *ps += (1 + batches) * 16 / 1024.0;
        for (int j = 0; j < batches; j++) {
            const size_t so = sl[j].off;
            HB_LAUNCH(ctx, stream_fold_kernel, grid_for(ctx, sl[j].len), 256, 0, f1[j] + so, f2[j] + so, f3[j] + so, A[j] + 2 * g * Bj[j] + 2 * so, eq_low[j] + so, rand, sl[j].len);
        }
    }
    for (int j = 0; j < batches; j++) if (!feq(Kp[j], old_claims[j])) printf("Error in sumcheck 0 %d\n", j);   // the reference only warns (:1246-1251)
    const int lgB = lgBj[0];
    std::vector<hb_F> p1(5 * (size_t)lgB + 3 * batches + 8);
    if ((rc = batch_sumcheck3_dev(ctx, f1[0], f2[0], f3[0], Bj, batches, a, p1.data(), ps))) return fail(rc);
    {
        F s = fadd(fadd(fadd(fromabi(p1[0]), fromabi(p1[1])), fadd(fromabi(p1[2]), fromabi(p1[3]))), fromabi(p1[3]));
        if (!feq(s, Kf)) { cudaFreeAsync(buf, ctx->stream); HB_FAIL(ctx, "Error in sumcheck 1"); }
    }
    if (nb < 2) { cudaFreeAsync(buf, ctx->stream); HB_FAIL(ctx, "stream batch: single-chunk layers belong to the in-memory prover"); }
    std::vector<F> P1r(lgB);
    for (int q = 0; q < lgB; q++) P1r[q] = fromabi(p1[4 * lgB + q]);
    HB_CHECK(ctx, cudaMemcpyAsync(r_dev, P1r.data(), lgB * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<std::vector<F>> PE(2 * batches, std::vector<F>(nb, mkF(0, 0)));
    for (int j = 0; j < batches; j++) {
        if ((rc = beta_dev(ctx, r_dev, lgBj[j], beta[j]))) return fail(rc);
        unsigned parts = (unsigned)std::max<size_t>(1, std::min<size_t>((sl[j].len + 2047) / 2048, (size_t)(2 * ctx->sm_count) / nb + 1));
        F *pe_dev; HB_CHECK(ctx, cudaMallocAsync(&pe_dev, nb * parts * 2 * sizeof(F), ctx->stream));
        HB_LAUNCH(ctx, partial_evals_kernel, dim3(parts, (unsigned)nb), 256, 0, A[j], beta[j], Bj[j], sl[j].off, sl[j].len, pe_dev);
        if (sl[j].on) HB_TRY(dist_allreduce_vec(ctx, pe_dev, nb * parts * 2));
        std::vector<F> pe(nb * parts * 2);
        HB_CHECK(ctx, cudaMemcpyAsync(pe.data(), pe_dev, pe.size() * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
        HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFreeAsync(pe_dev, ctx->stream);
        for (size_t g = 0; g < nb; g++) for (unsigned q = 0; q < parts; q++) {
            PE[2 * j][g] = fadd(PE[2 * j][g], pe[(g * parts + q) * 2]); PE[2 * j + 1][g] = fadd(PE[2 * j + 1][g], pe[(g * parts + q) * 2 + 1]);
        }
    }
    cudaFreeAsync(buf, ctx->stream);
    std::vector<F> Rp; Rp.reserve(nb);
    for (size_t i = 0; i < nb / 2; i++) Rp.push_back(R[2 * i]);
    for (size_t i = 0; i < nb / 2; i++) Rp.push_back(R[2 * i + 1]);
    std::vector<F> aggr(nb, mkF(0, 0));
    for (int i = 0; i < 2 * batches; i++) for (size_t g = 0; g < nb; g++) aggr[g] = fadd(aggr[g], h_fmul(b[i], PE[i][g]));
    std::vector<hb_F> p2(4 * (size_t)lgnb + 8);
    hb_F zero{0, 0};
    HB_TRY(hb_sumcheck2(ctx, (const hb_F *)Rp.data(), (const hb_F *)aggr.data(), nb, &zero, p2.data(), ps));
    {
        F sum = mkF(0, 0);
        for (int i = 0; i < batches; i++) sum = fadd(sum, fadd(h_fmul(b[2 * i], fromabi(p1[5 * lgB + 3 * i])), h_fmul(b[2 * i + 1], fromabi(p1[5 * lgB + 3 * i + 1]))));
        F q = fadd(fadd(fromabi(p2[0]), fromabi(p2[1])), fadd(fromabi(p2[2]), fromabi(p2[2])));
        if (!feq(sum, q)) HB_FAIL(ctx, "Error in sumcheck 2");
    }
    std::vector<F> P2r(lgnb);
    for (int q = 0; q < lgnb; q++) P2r[q] = fromabi(p2[3 * lgnb + q]);
    auto eval_small = [&](std::vector<F> v) { for (int q = 0; q < lgnb; q++) for (size_t j = 0; j < (nb >> (q + 1)); j++) v[j] = h_fold(v[2 * j], v[2 * j + 1], P2r[q]); return v[0]; };
    new_r.assign(batches, std::vector<F>());
    for (int j = 0; j < batches; j++) {
        new_r[j].push_back(pad);
        for (int q = 0; q < lgBj[j]; q++) new_r[j].push_back(P1r[q]);
        for (int q = 0; q < lgnb; q++) new_r[j].push_back(P2r[q]);
        new_claims[j] = fadd(h_fmul(fsub(mkF(1, 0), pad), eval_small(PE[2 * j])), h_fmul(pad, eval_small(PE[2 * j + 1])));
    }
    return 0;
}

// generate_claims_opt (sumcheck.cpp:1014-1055) on the resident layer arrays
static int layer_claims_dev(hb_ctx *ctx, const F *const *A, size_t S0, size_t B, int distance, int batches, const std::vector<F> &r, F *claims) {
    const size_t nb = S0 / (2 * B); const int lgnb = ilog2(nb);
    for (int j = 0; j < batches; j++) {
        const size_t Bj = B >> (j * distance); const int lgBj = ilog2(Bj);
        F *buf; HB_CHECK(ctx, cudaMallocAsync(&buf, (Bj + 2 * nb + 64) * sizeof(F), ctx->stream));
        F *lo = buf, *hi = buf + Bj, *r_dev = hi + nb;
        int rc;
        HB_CHECK(ctx, cudaMemcpyAsync(r_dev, r.data(), (lgBj + lgnb) * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = beta_dev(ctx, r_dev, lgBj, lo)) || (rc = beta_dev(ctx, r_dev + lgBj, lgnb, hi))) { cudaFreeAsync(buf, ctx->stream); return rc; }
        const Slice sl = dist_slice(ctx, Bj);
        unsigned parts = (unsigned)std::max<size_t>(1, std::min<size_t>((sl.len + 2047) / 2048, (size_t)(2 * ctx->sm_count) / nb + 1));
        F *out_dev; HB_CHECK(ctx, cudaMallocAsync(&out_dev, nb * parts * sizeof(F), ctx->stream));
        HB_LAUNCH(ctx, layer_claim_kernel, dim3(parts, (unsigned)nb), 256, 0, A[j], lo, Bj, sl.off, sl.len, out_dev);
        if (sl.on) HB_TRY(dist_allreduce_vec(ctx, out_dev, nb * parts));
        std::vector<F> part(nb * parts), eh(nb);
        HB_CHECK(ctx, cudaMemcpyAsync(part.data(), out_dev, part.size() * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
        HB_CHECK(ctx, cudaMemcpyAsync(eh.data(), hi, nb * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
        HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFreeAsync(out_dev, ctx->stream); cudaFreeAsync(buf, ctx->stream);
        F c = mkF(0, 0);
        for (size_t g = 0; g < nb; g++) { F sg = mkF(0, 0); for (unsigned q = 0; q < parts; q++) sg = fadd(sg, part[g * parts + q]); c = fadd(c, h_fmul(eh[g], sg)); }
        claims[j] = c;
    }
    return 0;
}

// layer arrays of a resident stream: lv[0] = xy, lv[l+1][j] = lv[l][2j] * lv[l][2j+1]  (segments never straddle X|Y)
static int build_layers(hb_ctx *ctx, const F *xy, size_t total, int layers, std::vector<const F *> &lv, F **owned) {
    lv.assign(layers + 1, nullptr); lv[0] = xy; *owned = nullptr;
    if (layers == 0) return 0;
    HB_CHECK(ctx, cudaMallocAsync(owned, total * sizeof(F), ctx->stream));
    F *p = *owned;
    for (int l = 0; l < layers; l++) {
        size_t sz = total >> (l + 1);
        HB_LAUNCH(ctx, prod_level_kernel, (unsigned)std::min<size_t>((sz + 255) / 256, (size_t)ctx->sm_count * 8), 256, 0, lv[l], p, sz);
        lv[l + 1] = p; p += sz;
    }
    return 0;
}

extern "C" int hb_stream_sumcheck_layer(hb_ctx *ctx, const hb_F *xy, size_t total, size_t B, int layer_id, const hb_F *r, int nr,
                                        const hb_F *old_claim, const hb_F *rnd4, hb_F *new_claim, hb_F *new_r, int *n_new_r, double *ps) { HB_DEV(ctx);
    if (total == 0 || (total & (total - 1))) HB_FAIL(ctx, "hb_stream_sumcheck_layer: stream size must be a power of two");
    Staged sx(ctx);
    HB_TRY(sx.in(xy, total * sizeof(F)));
    std::vector<const F *> lv; F *owned;
    HB_TRY(build_layers(ctx, sx.as<F>(), total, layer_id, lv, &owned));
    size_t S = total >> layer_id;
    if (nr != ilog2(S / 2)) { if (owned) cudaFreeAsync(owned, ctx->stream); HB_FAIL(ctx, "hb_stream_sumcheck_layer: r must hold log2(S/2) points"); }
    std::vector<F> nrv; F nc;
    int rc = stream_layer_dev(ctx, lv[layer_id], S, B, (const F *)r, fromabi(*old_claim), (const F *)rnd4, &nc, nrv, ps);
    if (owned) cudaFreeAsync(owned, ctx->stream);
    if (rc) return rc;
    *new_claim = toabi(nc);
    for (size_t i = 0; i < nrv.size(); i++) new_r[i] = toabi(nrv[i]);
    *n_new_r = (int)nrv.size();
    return 0;
}

// S6: prove_multiplication_tree_stream_shallow (sumcheck.cpp:1746-1915) with the stream resident in HBM.
// x_rand: log2(vectors) points for the product tree; rnd: 4 values (a, b0, b1, pad) per streamed layer, top layer first.
extern "C" int hb_mul_tree_stream(hb_ctx *ctx, const hb_F *xy, size_t total, int vectors, size_t B, int distance, int naive,
                                  const hb_F *prev_r, const hb_F *x_rand, const hb_F *rnd, hb_F *out, int *layers_out, double *ps) { HB_DEV(ctx);
    if (total == 0 || (total & (total - 1)) || vectors < 2 || (vectors & (vectors - 1))) HB_FAIL(ctx, "hb_mul_tree_stream: sizes must be powers of two, vectors >= 2");
    Staged sx(ctx);
    HB_TRY(sx.in(xy, total * sizeof(F)));
    const int maxr = ilog2(total);
    std::vector<hb_F> buf(64 + vectors + 8 * (size_t)(maxr + 2) * (maxr + 2));
    size_t written; int nfr;
    if (total <= 2 * B) {                                            // the whole stream fits the buffer: plain product tree (:1756-1773)
        HB_TRY(hb_mul_tree(ctx, (const hb_F *)sx.as<F>(), vectors, total / vectors, prev_r, x_rand, buf.data(), &written, &nfr, ps));
        memcpy(out, buf.data(), vectors * sizeof(hb_F)); *layers_out = 0;
        return 0;
    }
    int layers = ilog2(total / (2 * B));
    if (layers % distance != 0 && layers > distance) layers = distance + layers - (layers % distance);
    std::vector<const F *> lv; F *owned;
    HB_TRY(build_layers(ctx, sx.as<F>(), total, layers, lv, &owned));
    size_t St = total >> layers;
    int rc = hb_mul_tree(ctx, (const hb_F *)lv[layers], vectors, St / vectors, prev_r, x_rand, buf.data(), &written, &nfr, ps);
    if (rc) { cudaFreeAsync(owned, ctx->stream); return rc; }
    memcpy(out, buf.data(), vectors * sizeof(hb_F));
    F claim = fromabi(buf[vectors + 1 + nfr]);
    std::vector<F> r(nfr), nrv;
    for (int i = 0; i < nfr; i++) r[i] = fromabi(buf[vectors + 1 + i]);
    if (layers <= distance || naive) {
        for (int i = layers - 1, q = 0; i >= 0 && !rc; i--, q++) {
            F nc;
            rc = stream_layer_dev(ctx, lv[i], total >> i, B, r.data(), claim, (const F *)rnd + 4 * q, &nc, nrv, ps);
            claim = nc; r = nrv;
        }
    } else {
        // layers > distance (:1871-1908): `batches` layers, `distance` apart, are proven together in `distance` batched passes.
        // rnd = generate_randomness(layers - distance) | per pass a[batches], b[2*batches], pad.  commit_layers / open_layers (the Elastic_PC
        // commitment to the intermediate layers, :983-1011) are separate calls made by the host mirror around this one.
        const int batches = layers / distance, extra = layers - distance;
        if (batches > 8) { cudaFreeAsync(owned, ctx->stream); HB_FAIL(ctx, "hb_mul_tree_stream: more than 8 batches"); }
        const F *rq = (const F *)rnd;
        for (int i = 0; i < extra; i++) r.push_back(rq[i]);
        rq += extra;
        std::vector<std::vector<F>> rb(batches, r), nrb;
        F claims[8], nclaims[8];
        const F *Aj[8];
        for (int j = 0; j < batches; j++) Aj[j] = lv[distance - 1 + j * distance];
        rc = layer_claims_dev(ctx, Aj, total >> (distance - 1), B, distance, batches, r, claims);
        for (int i = distance - 1; i >= 0 && !rc; i--) {
            for (int j = 0; j < batches; j++) Aj[j] = lv[i + j * distance];
            rc = stream_batch_dev(ctx, Aj, total >> i, B, distance, batches, rb, claims, rq, nclaims, nrb, ps);
            rq += 3 * batches + 1;
            for (int j = 0; j < batches; j++) claims[j] = nclaims[j];
            rb = nrb;
        }
    }
    cudaFreeAsync(owned, ctx->stream);
    *layers_out = layers;
    return rc;
}

// prove_gate_consistency_standard (sumcheck.cpp:434-501).  out: (a,b,c,d,e,rand) per round, then the final add, L, R, O, mul, beta.
extern "C" int hb_gate_consistency_standard(hb_ctx *ctx, const hb_F *L, const hb_F *R, const hb_F *O, const hb_F *add_gate, size_t n,
                                            const hb_F *r, hb_F *out) { HB_DEV(ctx);
    if (n < 2 || (n & (n - 1))) HB_FAIL(ctx, "hb_gate_consistency_standard: n must be a power of two >= 2");
    HB_TRY(ensure_scratch(ctx));
    const int rounds = ilog2(n);
    Staged sl(ctx), sr(ctx), so(ctx), sa(ctx), srr(ctx);
    HB_TRY(sl.in(L, n * sizeof(F))); HB_TRY(sr.in(R, n * sizeof(F))); HB_TRY(so.in(O, n * sizeof(F))); HB_TRY(sa.in(add_gate, n * sizeof(F)));
    HB_TRY(srr.in(r, rounds * sizeof(F)));
    F *buf; HB_CHECK(ctx, cudaMallocAsync(&buf, (n + 5 * (n / 2 + n / 4 + 2)) * sizeof(F), ctx->stream));
    F *beta = buf;
    int rc = beta_dev(ctx, srr.as<F>(), rounds, beta);
    const F *cur[5] = {sa.as<F>(), beta, sl.as<F>(), sr.as<F>(), so.as<F>()};
    F *A[5], *Bf[5];
    for (int k = 0; k < 5; k++) { A[k] = buf + n + (size_t)k * (n / 2 + n / 4 + 2); Bf[k] = A[k] + n / 2 + 1; }
    F rand = mkF(213, 0);
    for (int i = 0; i < rounds && !rc; i++) {
        size_t Lp = n >> (i + 1);
        F co[5];
        Tabs<5> t;
        for (int k = 0; k < 5; k++) { t.in[k] = cur[k]; t.out[k] = A[k]; }
        const GateW w = {mkF(1, 0), mkF(1, 0), mkF(1, 0), mkF(P61 - 1, 0), mkF(1, 0)};
        if (i == 0) { HB_LAUNCH(ctx, gate_round_kernel<POLY_ONLY>, grid_for(ctx, Lp), 256, 0, t, Lp, rand, w, red_args(ctx)); }
        else {
            HB_LAUNCH(ctx, gate_round_kernel<FOLD_THEN_POLY>, grid_for(ctx, Lp), 256, 0, t, Lp, rand, w, red_args(ctx));
            for (int k = 0; k < 5; k++) { cur[k] = A[k]; std::swap(A[k], Bf[k]); }
        }
        rc = read_result(ctx, 5, co);
        for (int c = 0; c < 5; c++) { rand = h_mimc(co[c], rand); out[6 * i + c] = toabi(co[c]); }     // mimc_hash(value, rand): N7
        out[6 * i + 5] = toabi(rand);
    }
    F fin[5];
    if (!rc) {
        Tabs<5> t;
        for (int k = 0; k < 5; k++) { t.in[k] = cur[k]; t.out[k] = A[k]; }
        rc = launch_round<5, FOLD_ONLY, false>(ctx, t, 1, rand, nullptr);
        if (!rc) rc = fetch_heads(ctx, A, 5, fin);
    }
    cudaFreeAsync(buf, ctx->stream);
    if (rc) return rc;
    hb_F *o = out + 6 * rounds;
    o[0] = toabi(fin[0]); o[1] = toabi(fin[2]); o[2] = toabi(fin[3]); o[3] = toabi(fin[4]);
    o[4] = toabi(fsub(mkF(1, 0), fin[0]));                   // mul = 1 - add folds to 1 - fold(add)
    o[5] = toabi(fin[1]);
    return 0;
}

// S7: prove_gate_consistency (sumcheck.cpp:796-981, no lookups) on a transcript resident in HBM.
// rnd10 = a[4] (generate_randomness(4)) | b[6] (generate_randomness(6)), drawn by the host in that order.
// out: R[nch] | (a,b,c,d,e,rand) x log2 B | final L,R,O,add,mul,beta | Peval[6][nch] | P2 flat proof (4*log2 nch + 3).
extern "C" int hb_gate_consistency_stream(hb_ctx *ctx, const hb_F *L, const hb_F *R, const hb_F *O, const hb_F *S, size_t cs, size_t B,
                                          const hb_F *r, const hb_F *rnd10, hb_F *out, double *ps) { HB_DEV(ctx);
    if (B < 2 || (B & (B - 1)) || cs < B || (cs & (cs - 1))) HB_FAIL(ctx, "hb_gate_consistency_stream: sizes must be powers of two, cs >= B >= 2");
    HB_TRY(ensure_scratch(ctx));
    const size_t nch = cs / B; const int lgB = ilog2(B), lgn = ilog2(nch);
    Staged sl(ctx), sr(ctx), so(ctx), ss(ctx), srr(ctx);
    HB_TRY(sl.in(L, cs * sizeof(F))); HB_TRY(sr.in(R, cs * sizeof(F))); HB_TRY(so.in(O, cs * sizeof(F))); HB_TRY(ss.in(S, cs * sizeof(F)));
    HB_TRY(srr.in(r, lgB * sizeof(F)));
    const F *dL = sl.as<F>(), *dR = sr.as<F>(), *dO = so.as<F>(), *dS = ss.as<F>();
    F *buf; HB_CHECK(ctx, cudaMallocAsync(&buf, (7 * B + 5 * (B / 2 + B / 4 + 2) + 64) * sizeof(F), ctx->stream));
    F *beta = buf, *fA = buf + B, *fB = fA + B, *fL = fB + B, *fR = fL + B, *fO = fR + B, *beta1 = fO + B, *pp = beta1 + B, *r_dev = pp + 5 * (B / 2 + B / 4 + 2);
    auto fail = [&](int rc) { cudaFreeAsync(buf, ctx->stream); return rc; };
    int rc;
    if ((rc = beta_dev(ctx, srr.as<F>(), lgB, beta))) return fail(rc);
    // multi-GPU: every rank folds positions [so, so + dsn) of each chunk; sums are added across ranks inside the kernels
    const Slice dsl = dist_slice(ctx, B);
    const size_t dso = dsl.off, dsn = dsl.len;
    DistRedGuard dist_guard(ctx, dsl.on);
    const unsigned grid = grid_for(ctx, dsn);
    F Kf[4];                                                    // Kf_O, Kf_L, Kf_R, Kf_M
    HB_LAUNCH(ctx, gs_init_kernel, grid, 256, 0, dL + dso, dR + dso, dO + dso, dS + dso, beta + dso, fA + dso, fB + dso, fL + dso, fR + dso, fO + dso, dsn, red_args(ctx));
    if ((rc = read_result(ctx, 4, Kf))) return fail(rc);
    *ps += 4 * 16 / 1024.0;
    std::vector<F> Rv; Rv.push_back(mkF(1, 0));
    F rand = mkF(0, 0), csum = mkF(1, 0);
    for (size_t c = 1; c < nch; c++) {
        const F *bL = dL + c * B + dso, *bR = dR + c * B + dso, *bO = dO + c * B + dso, *bS = dS + c * B + dso;
        F K[12];
        HB_LAUNCH(ctx, gs_err_kernel, grid, 256, 0, bL, bR, bO, bS, beta + dso, fA + dso, fB + dso, fL + dso, fR + dso, fO + dso, csum, dsn, red_args(ctx));
        if ((rc = read_result(ctx, 12, K))) return fail(rc);
        if (!fzero(fsub(fadd(fadd(K[11], K[4]), K[7]), K[1]))) { cudaFreeAsync(buf, ctx->stream); HB_FAIL(ctx, "Error in gate consistency 1"); }
        for (int q = 0; q < 8; q++) rand = h_mimc(K[q], rand);                          // K*_M are not hashed (:844-851)
        Rv.push_back(rand);
        F x1 = rand, x2 = h_fmul(rand, x1), x3 = h_fmul(rand, x2), x4 = h_fmul(rand, x3);
        Kf[0] = fadd(Kf[0], fadd(h_fmul(x1, K[0]), h_fmul(x2, K[1])));
        Kf[1] = fadd(Kf[1], fadd(fadd(h_fmul(x1, K[2]), h_fmul(x2, K[3])), h_fmul(x3, K[4])));
        Kf[2] = fadd(Kf[2], fadd(fadd(h_fmul(x1, K[5]), h_fmul(x2, K[6])), h_fmul(x3, K[7])));
        Kf[3] = fadd(Kf[3], fadd(fadd(fadd(h_fmul(x1, K[8]), h_fmul(x2, K[9])), h_fmul(x3, K[10])), h_fmul(x4, K[11])));
        *ps += 12 * 16 / 1024.0;
        HB_LAUNCH(ctx, gs_fold_kernel, grid, 256, 0, bL, bR, bO, bS, beta + dso, fA + dso, fB + dso, fL + dso, fR + dso, fO + dso, rand, dsn);
        csum = fadd(csum, rand);
    }
    size_t k = 0;
    for (auto &x : Rv) out[k++] = toabi(x);
    const F *a = (const F *)rnd10, *b = a + 4;
    F sum = fadd(fadd(h_fmul(a[0], Kf[1]), h_fmul(a[1], Kf[2])), fadd(h_fmul(a[2], Kf[3]), h_fmul(Kf[0], a[3])));
    // degree-4 sumcheck over the folds (:877-935)
    const F *cur[5] = {fA + dso, fB + dso, fL + dso, fR + dso, fO + dso};
    bool on = dsl.on;
    F *small = nullptr;
    if (on) HB_CHECK(ctx, cudaMallocAsync(&small, 5 * 2 * kMaxRanks * sizeof(F), ctx->stream));
    F *A5[5], *B5[5];
    for (int q = 0; q < 5; q++) { A5[q] = pp + (size_t)q * (B / 2 + B / 4 + 2); B5[q] = A5[q] + B / 2 + 1; }
    const GateW w = {a[0], a[1], a[2], a[3], csum};
    std::vector<F> srnd;
    for (int i = 0; i < lgB; i++) {
        size_t Lp = B >> (i + 1), Ll;
        F co[5];
        if ((rc = slice_step(ctx, on, cur, 5, i == 0 ? 2 * Lp : 4 * Lp, Lp, small, &Ll))) return fail(rc);
        ctx->dist.reduce_on = on;
        Tabs<5> t;
        for (int q = 0; q < 5; q++) { t.in[q] = cur[q]; t.out[q] = A5[q]; }
        if (i == 0) { HB_LAUNCH(ctx, gate_round_kernel<POLY_ONLY>, grid_for(ctx, Ll), 256, 0, t, Ll, rand, w, red_args(ctx)); }
        else {
            HB_LAUNCH(ctx, gate_round_kernel<FOLD_THEN_POLY>, grid_for(ctx, Ll), 256, 0, t, Ll, rand, w, red_args(ctx));
            for (int q = 0; q < 5; q++) { cur[q] = A5[q]; std::swap(A5[q], B5[q]); }
        }
        if ((rc = read_result(ctx, 5, co))) return fail(rc);
        for (int q = 0; q < 5; q++) { rand = h_mimc(co[q], rand); out[k++] = toabi(co[q]); }
        out[k++] = toabi(rand);
        F s01 = fadd(fadd(fadd(co[0], co[1]), fadd(co[2], co[3])), fadd(co[4], co[4]));
        if (!feq(s01, sum)) { cudaFreeAsync(buf, ctx->stream); HB_FAIL(ctx, "Error in gate consistency 2"); }
        sum = fadd(h_fmul(fadd(h_fmul(fadd(h_fmul(fadd(h_fmul(co[0], rand), co[1]), rand), co[2]), rand), co[3]), rand), co[4]);
        srnd.push_back(rand);
        *ps += 5 * 16 / 1024.0;
    }
    ctx->dist.reduce_on = false;
    F fin[5];
    {
        Tabs<5> t;
        for (int q = 0; q < 5; q++) { t.in[q] = cur[q]; t.out[q] = A5[q]; }
        if ((rc = launch_round<5, FOLD_ONLY, false>(ctx, t, 1, rand, nullptr))) return fail(rc);
        if ((rc = fetch_heads(ctx, A5, 5, fin))) return fail(rc);
    }
    const F f_add = fin[0], f_beta = fin[1], f_L = fin[2], f_R = fin[3], f_O = fin[4], f_mul = fsub(csum, fin[0]);
    out[k++] = toabi(f_L); out[k++] = toabi(f_R); out[k++] = toabi(f_O); out[k++] = toabi(f_add); out[k++] = toabi(f_mul); out[k++] = toabi(f_beta);
    // pass B: partial evaluations per chunk against eq(sumcheck_rand)
    HB_CHECK(ctx, cudaMemcpyAsync(r_dev, srnd.data(), lgB * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = beta_dev(ctx, r_dev, lgB, beta1))) return fail(rc);
    if (small) cudaFreeAsync(small, ctx->stream);
    unsigned parts = (unsigned)std::max<size_t>(1, std::min<size_t>((dsn + 2047) / 2048, (size_t)(2 * ctx->sm_count) / (nch + 1) + 1));
    F *pe_dev; HB_CHECK(ctx, cudaMallocAsync(&pe_dev, (nch + 1) * parts * 4 * sizeof(F), ctx->stream));
    HB_LAUNCH(ctx, gs_peval_kernel, dim3(parts, (unsigned)nch + 1), 256, 0, dL, dR, dO, dS, beta, beta1, B, dso, dsn, (unsigned)nch, pe_dev);
    if (dsl.on) HB_TRY(dist_allreduce_vec(ctx, pe_dev, (nch + 1) * parts * 4));
    std::vector<F> pe((nch + 1) * parts * 4);
    HB_CHECK(ctx, cudaMemcpyAsync(pe.data(), pe_dev, pe.size() * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(pe_dev, ctx->stream);
    cudaFreeAsync(buf, ctx->stream);
    std::vector<F> Pe(6 * nch, mkF(0, 0));
    F sb1 = mkF(0, 0), sbb = mkF(0, 0);
    for (unsigned q = 0; q < parts; q++) { sb1 = fadd(sb1, pe[(nch * parts + q) * 4]); sbb = fadd(sbb, pe[(nch * parts + q) * 4 + 1]); }
    for (size_t c = 0; c < nch; c++) {
        for (unsigned q = 0; q < parts; q++) for (int t = 0; t < 4; t++) Pe[t * nch + c] = fadd(Pe[t * nch + c], pe[((c * parts + q) * 4) + t]);
        Pe[4 * nch + c] = fsub(sb1, Pe[3 * nch + c]);              // sum beta1 (1 - S) = sum beta1 - sum beta1 S
        Pe[5 * nch + c] = sbb;                                     // sum beta1 beta is the same for every chunk
    }
    for (auto &x : Pe) out[k++] = toabi(x);
    std::vector<F> pv(nch, mkF(0, 0));
    for (size_t j = 0; j < nch; j++) for (int i = 0; i < 6; i++) pv[j] = fadd(pv[j], h_fmul(b[i], Pe[i * nch + j]));
    std::vector<hb_F> p2(4 * (size_t)lgn + 8);
    hb_F rabi = toabi(rand);
    HB_TRY(hb_sumcheck2(ctx, (const hb_F *)Rv.data(), (const hb_F *)pv.data(), nch, &rabi, p2.data(), ps));
    *ps += 5 * 16 / 1024.0;
    F s2 = fadd(fadd(fadd(h_fmul(f_L, b[0]), h_fmul(f_R, b[1])), fadd(h_fmul(f_O, b[2]), h_fmul(b[3], f_add))), fadd(h_fmul(b[4], f_mul), h_fmul(b[5], f_beta)));
    if (lgn) {
        F qv = fadd(fadd(fromabi(p2[0]), fromabi(p2[1])), fadd(fromabi(p2[2]), fromabi(p2[2])));
        if (!feq(qv, s2)) HB_FAIL(ctx, "Error in gate consistency 3");
    }
    for (int i = 0; i < 4 * lgn + 3; i++) out[k++] = p2[i];
    return 0;
}

// One round of the 3-product sumcheck on DEVICE tables, for callers that own the round loop (multi-GPU sharding by hypercube prefix,
// hobbit_b200/dist.py): accumulates the cubic coefficients of sum_j prod_t (in_t[2j] + X (in_t[2j+1] - in_t[2j])) over L pairs and
// writes out_t[j] = in_t[2j] + rand * (in_t[2j+1] - in_t[2j])  (the S2 schedule: fold with the incoming challenge, sumcheck.cpp:1987-2014).
extern "C" int hb_sc3_round(hb_ctx *ctx, const hb_F *in1, const hb_F *in2, const hb_F *in3, hb_F *out1, hb_F *out2, hb_F *out3, size_t L,
                            const hb_F *rand, hb_F *coeffs4) { HB_DEV(ctx);
    HB_TRY(ensure_scratch(ctx));
    if (!is_device_ptr(in1) || !is_device_ptr(out1)) HB_FAIL(ctx, "hb_sc3_round: tables must be device memory");
    Tabs<3> t;
    t.in[0] = (const F *)in1; t.in[1] = (const F *)in2; t.in[2] = (const F *)in3;
    t.out[0] = (F *)out1; t.out[1] = (F *)out2; t.out[2] = (F *)out3;
    F co[4];
    HB_TRY((launch_round<3, POLY_AND_FOLD, false>(ctx, t, L, fromabi(*rand), co)));
    for (int c = 0; c < 4; c++) coeffs4[c] = toabi(co[c]);
    return 0;
}

// S8: prove_gate_consistency_lookups (sumcheck.cpp:503-794) on a transcript resident in HBM.  rnd13 = a[5] | b[8], lookup_rand2 = lookup_rand[0..1].
// out: R[nch] | (a,b,c,d,e,rand) x log2 B | final L,R,O,add_L,add_R,mul,lkp,lkp_O,beta | Peval[8][nch] | P2 flat proof (4*log2 nch + 3).
extern "C" int hb_gate_consistency_lookups_stream(hb_ctx *ctx, const hb_F *L, const hb_F *R, const hb_F *O, const hb_F *S, size_t cs, size_t B,
                                                  const hb_F *r, const hb_F *lookup_rand2, const hb_F *rnd13, hb_F *out, double *ps) { HB_DEV(ctx);
    if (B < 2 || (B & (B - 1)) || cs < B || (cs & (cs - 1))) HB_FAIL(ctx, "hb_gate_consistency_lookups_stream: sizes must be powers of two, cs >= B >= 2");
    HB_TRY(ensure_scratch(ctx));
    const size_t nch = cs / B; const int lgB = ilog2(B), lgn = ilog2(nch);
    Staged sl(ctx), sr(ctx), so(ctx), ss(ctx), srr(ctx);
    HB_TRY(sl.in(L, cs * sizeof(F))); HB_TRY(sr.in(R, cs * sizeof(F))); HB_TRY(so.in(O, cs * sizeof(F))); HB_TRY(ss.in(S, cs * sizeof(F)));
    HB_TRY(srr.in(r, lgB * sizeof(F)));
    const F *dL = sl.as<F>(), *dR = sr.as<F>(), *dO = so.as<F>(), *dS = ss.as<F>();
    const F lr0 = fromabi(lookup_rand2[0]), lr1 = fromabi(lookup_rand2[1]);
    const size_t pps = B / 2 + B / 4 + 2;
    F *buf; HB_CHECK(ctx, cudaMallocAsync(&buf, (11 * B + 9 * pps + 64) * sizeof(F), ctx->stream));
    F *beta = buf, *beta1 = buf + B, *pp = buf + 11 * B, *r_dev = pp + 9 * pps;
    S8Folds f; for (int q = 0; q < 9; q++) f.t[q] = buf + (size_t)(2 + q) * B;
    auto fail = [&](int rc) { cudaFreeAsync(buf, ctx->stream); return rc; };
    int rc;
    if ((rc = beta_dev(ctx, srr.as<F>(), lgB, beta))) return fail(rc);
    // multi-GPU: every rank folds positions [so, so + dsn) of each chunk; sums are added across ranks inside the kernels
    const Slice dsl = dist_slice(ctx, B);
    const size_t dso = dsl.off, dsn = dsl.len;
    DistRedGuard dist_guard(ctx, dsl.on);
    S8Folds fs; for (int q = 0; q < 9; q++) fs.t[q] = f.t[q] + dso;
    const unsigned grid = grid_for(ctx, dsn);
    F Kf[5];                                                    // Kf_O, Kf_L, Kf_R, Kf_M, Kf_lkp
    HB_LAUNCH(ctx, gl_init_kernel, grid, 256, 0, dL + dso, dR + dso, dO + dso, dS + dso, beta + dso, fs, lr0, lr1, dsn, red_args(ctx));
    if ((rc = read_result(ctx, 5, Kf))) return fail(rc);
    if (!fzero(fsub(fsub(fadd(fadd(Kf[3], Kf[1]), Kf[2]), Kf[4]), Kf[0]))) { cudaFreeAsync(buf, ctx->stream); HB_FAIL(ctx, "Error (gate consistency with lookups, first chunk)"); }
    *ps += 5 * 16 / 1024.0;
    std::vector<F> Rv; Rv.push_back(mkF(1, 0));
    F rand = mkF(0, 0);
    for (size_t c = 1; c < nch; c++) {
        const F *bL = dL + c * B + dso, *bR = dR + c * B + dso, *bO = dO + c * B + dso, *bS = dS + c * B + dso;
        F K[15];
        HB_LAUNCH(ctx, gl_err_kernel, grid, 256, 0, bL, bR, bO, bS, beta + dso, fs, lr0, lr1, dsn, red_args(ctx));
        if ((rc = read_result(ctx, 15, K))) return fail(rc);
        if (!fzero(fsub(fsub(fadd(fadd(K[14], K[4]), K[7]), K[10]), K[1]))) { cudaFreeAsync(buf, ctx->stream); HB_FAIL(ctx, "Error in gate consistency 1"); }
        for (int q = 0; q < 8; q++) rand = h_mimc(K[q], rand);                          // only the O, L, R terms are hashed (:589-597)
        Rv.push_back(rand);
        F x1 = rand, x2 = h_fmul(rand, x1), x3 = h_fmul(rand, x2), x4 = h_fmul(rand, x3);
        Kf[0] = fadd(Kf[0], fadd(h_fmul(x1, K[0]), h_fmul(x2, K[1])));
        Kf[1] = fadd(Kf[1], fadd(fadd(h_fmul(x1, K[2]), h_fmul(x2, K[3])), h_fmul(x3, K[4])));
        Kf[2] = fadd(Kf[2], fadd(fadd(h_fmul(x1, K[5]), h_fmul(x2, K[6])), h_fmul(x3, K[7])));
        Kf[4] = fadd(Kf[4], fadd(fadd(h_fmul(x1, K[8]), h_fmul(x2, K[9])), h_fmul(x3, K[10])));
        Kf[3] = fadd(Kf[3], fadd(fadd(fadd(h_fmul(x1, K[11]), h_fmul(x2, K[12])), h_fmul(x3, K[13])), h_fmul(x4, K[14])));
        *ps += 15 * 16 / 1024.0;
        HB_LAUNCH(ctx, gl_fold_kernel, grid, 256, 0, bL, bR, bO, bS, beta + dso, fs, lr0, lr1, rand, dsn);
    }
    size_t k = 0;
    for (auto &x : Rv) out[k++] = toabi(x);
    const F *a = (const F *)rnd13, *b = a + 5;
    F sum = fadd(fadd(fadd(h_fmul(a[0], Kf[1]), h_fmul(a[1], Kf[2])), fadd(h_fmul(a[2], Kf[3]), h_fmul(Kf[0], a[3]))), h_fmul(Kf[4], a[4]));
    const F *cur[9]; F *A9[9], *B9[9];
    for (int q = 0; q < 9; q++) { cur[q] = f.t[q] + dso; A9[q] = pp + (size_t)q * pps; B9[q] = A9[q] + B / 2 + 1; }
    bool on = dsl.on;
    F *small = nullptr;
    if (on) HB_CHECK(ctx, cudaMallocAsync(&small, 9 * 2 * kMaxRanks * sizeof(F), ctx->stream));
    S8W w; for (int q = 0; q < 5; q++) w.a[q] = a[q];
    std::vector<F> srnd;
    for (int i = 0; i < lgB; i++) {
        size_t Lp = B >> (i + 1), Ll;
        F co[5];
        if ((rc = slice_step(ctx, on, cur, 9, i == 0 ? 2 * Lp : 4 * Lp, Lp, small, &Ll))) return fail(rc);
        ctx->dist.reduce_on = on;
        Tabs<9> t;
        for (int q = 0; q < 9; q++) { t.in[q] = cur[q]; t.out[q] = A9[q]; }
        if (i == 0) { HB_LAUNCH(ctx, gatel_round_kernel<POLY_ONLY>, grid_for(ctx, Ll), 256, 0, t, Ll, rand, w, red_args(ctx)); }
        else {
            HB_LAUNCH(ctx, gatel_round_kernel<FOLD_THEN_POLY>, grid_for(ctx, Ll), 256, 0, t, Ll, rand, w, red_args(ctx));
            for (int q = 0; q < 9; q++) { cur[q] = A9[q]; std::swap(A9[q], B9[q]); }
        }
        if ((rc = read_result(ctx, 5, co))) return fail(rc);
        for (int q = 0; q < 5; q++) { rand = h_mimc(co[q], rand); out[k++] = toabi(co[q]); }
        out[k++] = toabi(rand);
        F s01 = fadd(fadd(fadd(co[0], co[1]), fadd(co[2], co[3])), fadd(co[4], co[4]));
        if (!feq(s01, sum)) { cudaFreeAsync(buf, ctx->stream); HB_FAIL(ctx, "Error in gate consistency 2"); }
        sum = fadd(h_fmul(fadd(h_fmul(fadd(h_fmul(fadd(h_fmul(co[0], rand), co[1]), rand), co[2]), rand), co[3]), rand), co[4]);
        srnd.push_back(rand);
        *ps += 5 * 16 / 1024.0;
    }
    ctx->dist.reduce_on = false;
    F fin[9];
    {
        Tabs<9> t;
        for (int q = 0; q < 9; q++) { t.in[q] = cur[q]; t.out[q] = A9[q]; }
        if ((rc = launch_round<9, FOLD_ONLY, false>(ctx, t, 1, rand, nullptr))) return fail(rc);
        if ((rc = fetch_heads(ctx, A9, 9, fin))) return fail(rc);
    }
    // final L, R, O, add_L, add_R, mul, lkp, lkp_O, beta
    const F fo[9] = {fin[4], fin[5], fin[6], fin[0], fin[1], fin[2], fin[3], fin[7], fin[8]};
    for (int q = 0; q < 9; q++) out[k++] = toabi(fo[q]);
    HB_CHECK(ctx, cudaMemcpyAsync(r_dev, srnd.data(), lgB * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = beta_dev(ctx, r_dev, lgB, beta1))) return fail(rc);
    if (small) cudaFreeAsync(small, ctx->stream);
    unsigned parts = (unsigned)std::max<size_t>(1, std::min<size_t>((dsn + 2047) / 2048, (size_t)(2 * ctx->sm_count) / nch + 1));
    F *pe_dev; HB_CHECK(ctx, cudaMallocAsync(&pe_dev, nch * parts * 8 * sizeof(F), ctx->stream));
    HB_LAUNCH(ctx, gl_peval_kernel, dim3(parts, (unsigned)nch), 256, 0, dL, dR, dO, dS, beta1, lr0, lr1, B, dso, dsn, pe_dev);
    if (dsl.on) HB_TRY(dist_allreduce_vec(ctx, pe_dev, nch * parts * 8));
    std::vector<F> pe(nch * parts * 8);
    HB_CHECK(ctx, cudaMemcpyAsync(pe.data(), pe_dev, pe.size() * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(pe_dev, ctx->stream);
    cudaFreeAsync(buf, ctx->stream);
    std::vector<F> Pe(8 * nch, mkF(0, 0));
    for (size_t c = 0; c < nch; c++)
        for (unsigned q = 0; q < parts; q++) for (int t = 0; t < 8; t++) Pe[t * nch + c] = fadd(Pe[t * nch + c], pe[((c * parts + q) * 8) + t]);
    for (auto &x : Pe) out[k++] = toabi(x);
    std::vector<F> pv(nch, mkF(0, 0));
    for (size_t j = 0; j < nch; j++) for (int i = 0; i < 8; i++) pv[j] = fadd(pv[j], h_fmul(b[i], Pe[i * nch + j]));
    std::vector<hb_F> p2(4 * (size_t)lgn + 8);
    hb_F rabi = toabi(rand);
    HB_TRY(hb_sumcheck2(ctx, (const hb_F *)Rv.data(), (const hb_F *)pv.data(), nch, &rabi, p2.data(), ps));
    *ps += 5 * 16 / 1024.0;
    F s2 = mkF(0, 0);
    for (int q = 0; q < 8; q++) s2 = fadd(s2, h_fmul(fo[q], b[q]));
    if (lgn) {
        F qv = fadd(fadd(fromabi(p2[0]), fromabi(p2[1])), fadd(fromabi(p2[2]), fromabi(p2[2])));
        if (!feq(qv, s2)) HB_FAIL(ctx, "Error in gate consistency 3");
    }
    for (int i = 0; i < 4 * lgn + 3; i++) out[k++] = p2[i];
    return 0;
}
