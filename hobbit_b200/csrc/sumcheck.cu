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
#include <algorithm>

namespace hb {

enum { POLY_ONLY = 0, FOLD_THEN_POLY = 1, POLY_AND_FOLD = 2, FOLD_ONLY = 3 };
static constexpr int kMaxRedBlocks = 148 * 4;

template <int NT> struct Tabs { const F *in[NT]; F *out[NT]; };

__device__ __forceinline__ F fold1(F x, F y, F r) { return fadd(x, fmul(r, fsub(y, x))); }

// coefficients of prod_t (d_t * X + x_t), highest degree first, accumulated into acc[NT+1]
template <int NT> __device__ __forceinline__ void poly_acc(F (&acc)[NT + 1], const F (&x)[NT], const F (&y)[NT]) {
    if (NT == 2) {
        F d1 = fsub(y[0], x[0]), d2 = fsub(y[1], x[1]);
        acc[0] = fadd(acc[0], fmul(d1, d2));
        acc[1] = fadd(acc[1], fadd(fmul(d1, x[1]), fmul(d2, x[0])));
        acc[2] = fadd(acc[2], fmul(x[0], x[1]));
    } else if (NT == 3) {
        F d1 = fsub(y[0], x[0]), d2 = fsub(y[1], x[1]), d3 = fsub(y[2], x[2]);
        F qa = fmul(d1, d2), qb = fadd(fmul(d1, x[1]), fmul(d2, x[0])), qc = fmul(x[0], x[1]);
        acc[0] = fadd(acc[0], fmul(qa, d3));
        acc[1] = fadd(acc[1], fadd(fmul(qa, x[2]), fmul(qb, d3)));
        acc[2] = fadd(acc[2], fadd(fmul(qb, x[2]), fmul(qc, d3)));
        acc[3] = fadd(acc[3], fmul(qc, x[2]));
    }
}

__device__ __forceinline__ F shfl_down_F(F v, int d) {
    F r; r.re = __shfl_down_sync(0xffffffffu, v.re, d); r.im = __shfl_down_sync(0xffffffffu, v.im, d); return r;
}

// INTERLEAVED: tables 0 and 1 are the even/odd entries of one array t.in[0] (product-tree layer: in1[j]=prev[2j],
// in2[j]=prev[2j+1], sumcheck.cpp:84-101), so the layer is consumed in place without materialising in1/in2.
template <int NT, int MODE, bool INTERLEAVED>
__global__ void __launch_bounds__(256)
sc_round_kernel(Tabs<NT> t, size_t L, F r, F *__restrict__ partial, unsigned *__restrict__ ticket, F *__restrict__ result) {
    constexpr int NC = NT + 1;
    F acc[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) acc[c] = mkF(0, 0);

    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < L; j += (size_t)gridDim.x * blockDim.x) {
        F x[NT], y[NT];
        if (MODE == FOLD_THEN_POLY) {
#pragma unroll
            for (int k = 0; k < NT; k++) {
                const F *p = t.in[k] + 4 * j;
                F a = p[0], b = p[1], c = p[2], d = p[3];
                x[k] = fold1(a, b, r); y[k] = fold1(c, d, r);
                t.out[k][2 * j] = x[k]; t.out[k][2 * j + 1] = y[k];
            }
        } else if (INTERLEAVED) {
            const F *p = t.in[0] + 4 * j;
            x[0] = p[0]; x[1] = p[1]; y[0] = p[2]; y[1] = p[3];
#pragma unroll
            for (int k = 2; k < NT; k++) { x[k] = t.in[k][2 * j]; y[k] = t.in[k][2 * j + 1]; }
        } else {
#pragma unroll
            for (int k = 0; k < NT; k++) { x[k] = t.in[k][2 * j]; y[k] = t.in[k][2 * j + 1]; }
        }
        if (MODE != FOLD_ONLY) poly_acc<NT>(acc, x, y);
        if (MODE == POLY_AND_FOLD || MODE == FOLD_ONLY) {
#pragma unroll
            for (int k = 0; k < NT; k++) t.out[k][j] = fold1(x[k], y[k], r);
        }
    }
    if (MODE == FOLD_ONLY) return;

    // warp -> CTA -> grid reduction of the NC coefficients
    __shared__ F sred[8][NC];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        F v = acc[c];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = fadd(v, shfl_down_F(v, d));
        if (lane == 0) sred[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < NC) {
        F v = sred[0][threadIdx.x];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) v = fadd(v, sred[w][threadIdx.x]);
        partial[(size_t)blockIdx.x * NC + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last CTA: sum the per-CTA partials (volatile: written by other CTAs during this launch)
    const volatile u64 *pv = reinterpret_cast<const volatile u64 *>(partial);
#pragma unroll
    for (int c = 0; c < NC; c++) {
        F v = mkF(0, 0);
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            F q; q.re = pv[((size_t)b * NC + c) * 2]; q.im = pv[((size_t)b * NC + c) * 2 + 1];
            v = fadd(v, q);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = fadd(v, shfl_down_F(v, d));
        if (lane == 0) sred[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < NC) {
        F v = sred[0][threadIdx.x];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) v = fadd(v, sred[w][threadIdx.x]);
        result[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) *ticket = 0;
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
static int ensure_scratch(hb_ctx *ctx) {
    if (!ctx->red) {
        HB_CHECK(ctx, cudaMalloc(&ctx->red, (size_t)kMaxRedBlocks * 4 * sizeof(F) + 4 * sizeof(F)));
        HB_CHECK(ctx, cudaMalloc(&ctx->ticket, sizeof(unsigned)));
        HB_CHECK(ctx, cudaMemset(ctx->ticket, 0, sizeof(unsigned)));
        HB_CHECK(ctx, cudaMallocHost(&ctx->mailbox, 64 * sizeof(F)));
        ctx->mailbox_dev = ctx->red + (size_t)kMaxRedBlocks * 4;
    }
    return 0;
}
static inline unsigned grid_for(hb_ctx *ctx, size_t L) {
    size_t g = (L + 255) / 256;
    size_t cap = std::min<size_t>((size_t)ctx->sm_count * 4, kMaxRedBlocks);
    return (unsigned)std::max<size_t>(1, std::min(g, cap));
}

template <int NT, int MODE, bool IL>
static int launch_round(hb_ctx *ctx, const Tabs<NT> &t, size_t L, F r, F *coeffs_host /* NT+1, may be null for FOLD_ONLY */) {
    HB_TRY(ensure_scratch(ctx));
    HB_LAUNCH(ctx, (sc_round_kernel<NT, MODE, IL>), grid_for(ctx, L), 256, 0, t, L, r, ctx->red, ctx->ticket, ctx->mailbox_dev);
    if (MODE != FOLD_ONLY) {
        HB_CHECK(ctx, cudaMemcpyAsync(ctx->mailbox, ctx->mailbox_dev, (NT + 1) * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
        HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        for (int c = 0; c <= NT; c++) coeffs_host[c] = ctx->mailbox[c];
    }
    return 0;
}

static inline F h_mimc(F in, F k) { hb_F a{in.re, in.im}, b{k.re, k.im}, o; hb_mimc_hash(&a, &b, &o); return mkF(o.real, o.img); }
static inline F h_fold(F x, F y, F r) { return fadd(x, h_fmul(r, fsub(y, x))); }
static inline hb_F toabi(F x) { hb_F o{x.re, x.im}; return o; }

// download k scalars that sit at the head of k device tables
static int fetch_heads(hb_ctx *ctx, F *const *tabs, int k, F *out) {
    for (int i = 0; i < k; i++) HB_CHECK(ctx, cudaMemcpyAsync(ctx->mailbox + 8 + i, tabs[i], sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < k; i++) out[i] = ctx->mailbox[8 + i];
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
    const F *cur[3] = {v1, v2, v3};
    F *bufA[3], *bufB[3];
    for (int k = 0; k < 3; k++) { bufA[k] = scratch + (size_t)k * (n / 2 + n / 4); bufB[k] = bufA[k] + n / 2; }
    hb_F *rs = proof + 4 * rounds;
    for (int i = 0; i < rounds; i++) {
        size_t L = n >> (i + 1);
        F co[4];
        Tabs<3> t;
        for (int k = 0; k < 3; k++) { t.in[k] = cur[k]; t.out[k] = bufA[k]; }
        if (i == 0 && il) { t.in[0] = il; t.in[1] = nullptr; HB_TRY((launch_round<3, POLY_AND_FOLD, true>(ctx, t, L, rand, co))); }
        else HB_TRY((launch_round<3, POLY_AND_FOLD, false>(ctx, t, L, rand, co)));
        rs[i] = toabi(rand);
        for (int c = 0; c < 4; c++) { proof[4 * i + c] = toabi(co[c]); rand = h_mimc(rand, co[c]); }
        *ps += 5 * 16 / 1024.0;
        for (int k = 0; k < 3; k++) { cur[k] = bufA[k]; std::swap(bufA[k], bufB[k]); }
    }
    F vr[3];
    if (rounds == 0 && il) HB_FAIL(ctx, "sumcheck3: interleaved input needs n >= 2");
    F *heads[3] = {const_cast<F *>(cur[0]), const_cast<F *>(cur[1]), const_cast<F *>(cur[2])};
    HB_TRY(fetch_heads(ctx, heads, 3, vr));
    rand = h_mimc(rand, vr[0]); rand = h_mimc(rand, vr[1]);
    *ps += 3 * 16 / 1024.0;
    hb_F *o = proof + 5 * rounds;
    o[0] = toabi(vr[0]); o[1] = toabi(vr[1]); o[2] = toabi(vr[2]); o[3] = toabi(rand);
    return 0;
}

}  // namespace hb

using namespace hb;

// =========================================================================================================
extern "C" int hb_precompute_beta(hb_ctx *ctx, const hb_F *r, int nr, hb_F *out) {
    if (nr < 0 || nr > 24) HB_FAIL(ctx, "hb_precompute_beta: nr out of range");
    Staged sr(ctx), so(ctx);
    HB_TRY(sr.in(r, (size_t)std::max(nr, 1) * sizeof(F)));
    HB_TRY(so.outbuf(out, sizeof(F) << nr));
    HB_TRY(beta_dev(ctx, sr.as<F>(), nr, so.as<F>()));
    HB_TRY(so.finish());
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int hb_evaluate_vector(hb_ctx *ctx, const hb_F *v, size_t n, const hb_F *r, hb_F *out) {
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

extern "C" int hb_sumcheck2(hb_ctx *ctx, const hb_F *v1, const hb_F *v2, size_t n, const hb_F *prev_r, hb_F *proof, double *ps) {
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
                            hb_F *proof, double *ps) {
    if (n == 0 || (n & (n - 1))) HB_FAIL(ctx, "hb_sumcheck3: n must be a power of two");
    HB_TRY(ensure_scratch(ctx));
    Staged a(ctx), b(ctx), c(ctx);
    HB_TRY(a.in(v1, n * sizeof(F))); HB_TRY(b.in(v2, n * sizeof(F))); HB_TRY(c.in(v3, n * sizeof(F)));
    F *scratch; HB_CHECK(ctx, cudaMallocAsync(&scratch, (3 * (n / 2 + n / 4) + 4) * sizeof(F), ctx->stream));
    int rc = sumcheck3_dev(ctx, a.as<F>(), b.as<F>(), c.as<F>(), nullptr, n, mkF(prev_r->real, prev_r->img), scratch, proof, ps);
    cudaFreeAsync(scratch, ctx->stream);
    return rc;
}

extern "C" int hb_batch_sumcheck3(hb_ctx *ctx, const hb_F *t1, const hb_F *t2, const hb_F *t3, const size_t *sizes, int batches,
                                  const hb_F *a_in, hb_F *proof, double *ps) {
    HB_TRY(ensure_scratch(ctx));
    size_t tot = 0, Lmax = 0;
    for (int b = 0; b < batches; b++) {
        if (sizes[b] == 0 || (sizes[b] & (sizes[b] - 1))) HB_FAIL(ctx, "hb_batch_sumcheck3: batch sizes must be powers of two");
        tot += sizes[b]; Lmax = std::max(Lmax, sizes[b]);
    }
    int rounds = ilog2(Lmax);
    Staged s1(ctx), s2(ctx), s3(ctx);
    HB_TRY(s1.in(t1, tot * sizeof(F))); HB_TRY(s2.in(t2, tot * sizeof(F))); HB_TRY(s3.in(t3, tot * sizeof(F)));
    std::vector<F> a(batches);
    HB_CHECK(ctx, cudaMemcpy(a.data(), a_in, batches * sizeof(F), cudaMemcpyDefault));
    // per batch: ping-pong scratch of size/2 + size/4 per table
    F *scratch; HB_CHECK(ctx, cudaMallocAsync(&scratch, (3 * tot + 8) * sizeof(F), ctx->stream));
    struct Bt { const F *cur[3]; F *A[3], *B[3]; size_t size; bool exhausted; F x[3]; };
    std::vector<Bt> bt(batches);
    size_t off = 0, soff = 0;
    for (int b = 0; b < batches; b++) {
        Bt &q = bt[b]; q.size = sizes[b]; q.exhausted = false;
        const F *base[3] = {s1.as<F>() + off, s2.as<F>() + off, s3.as<F>() + off};
        for (int k = 0; k < 3; k++) { q.cur[k] = base[k]; q.A[k] = scratch + soff; q.B[k] = q.A[k] + sizes[b] / 2; soff += sizes[b] / 2 + sizes[b] / 4 + 1; }
        off += sizes[b];
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
                size_t L = (size_t)1 << lg;
                Tabs<3> t;
                for (int k = 0; k < 3; k++) { t.in[k] = q.cur[k]; t.out[k] = q.A[k]; }
                if (i == 0) rc = launch_round<3, POLY_ONLY, false>(ctx, t, L, rand, p);
                else {
                    rc = launch_round<3, FOLD_THEN_POLY, false>(ctx, t, L, rand, p);
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
    cudaFreeAsync(scratch, ctx->stream);
    if (rc) return rc;
    *ps += (3 * batches - batches) * 16 / 1024.0;
    for (int j = 0; j < 3 * batches; j++) proof[5 * rounds + j] = toabi(vr[j]);
    return 0;
}

extern "C" int hb_mul_tree(hb_ctx *ctx, const hb_F *input, int vectors, size_t n, const hb_F *prev_r, const hb_F *x_rand,
                           hb_F *out, size_t *written, int *nfr, double *ps) {
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
