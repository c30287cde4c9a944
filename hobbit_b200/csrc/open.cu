// 8f.1 — data-parallel building blocks of the opening recursion behind open_standard / Elastic_PC open.
// Reference: shockwave_commit / shockwave_prove (src/Virgo.cpp:120-157, 435-517), whir_commit / _whir_prove (:160-178, 519-686),
// change_form (:104-118), recursive_prover_Spielman / _RS (src/PC_utils.cpp:290-512), prove_fft / prove_fft_matrix
// (src/sumcheck.cpp:2975-3027), phiGInit / prepare_matrix (src/utils.cpp:694-775).
//
// The recursion is a chain of small-to-medium dense steps (<= 4*BUFFER_SPACE elements each) glued by libc-drawn challenges, so the
// orchestration (RNG call order, Fiat–Shamir scalars, proof-size accounting) lives in the C++ host mirror and every step that touches a
// table is one of the kernels below, all operating on tables that stay resident in HBM between steps.  Every kernel is HBM-bound
// (one or two multiplications per 16-byte element); the tables are far smaller than the 126 MB L2 in every configured case.
// All results are exact field elements in canonical form, so any evaluation order reproduces the reference's bits.
#include "common.cuh"
#include "reduce.cuh"
#include "blake3.cuh"
#include <algorithm>

namespace hb {

static inline unsigned grid_1d(hb_ctx *ctx, size_t n, unsigned block = 256) {
    return (unsigned)std::max<size_t>(1, std::min<size_t>((n + block - 1) / block, (size_t)ctx->sm_count * 16));
}

// ---- dense helpers ---------------------------------------------------------------------------------------------------------------
// out[j] = sum_i w[i] * M[i*stride + j].  CTA = 32 columns x 8 row slices; grid.y splits the rows further (partials summed by a second pass).
__global__ void __launch_bounds__(256)
matvec_cols_kernel(const F *__restrict__ M, size_t rows, size_t cols, size_t stride, const F *__restrict__ w, F *__restrict__ out) {
    __shared__ F s[8][32];
    const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t j = (size_t)blockIdx.x * 32 + tx;
    F acc = mkF(0, 0);
    if (j < cols)
        for (size_t i = (size_t)blockIdx.y * 8 + ty; i < rows; i += (size_t)gridDim.y * 8) acc = fadd(acc, fmul(ldgF(&w[i]), M[i * stride + j]));
    s[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && j < cols) {
#pragma unroll
        for (int q = 1; q < 8; q++) acc = fadd(acc, s[q][tx]);
        out[(size_t)blockIdx.y * cols + j] = acc;
    }
}
__global__ void __launch_bounds__(256) sum_partials_kernel(const F *__restrict__ part, size_t n, unsigned S, F *__restrict__ out) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    F acc = part[j];
    for (unsigned q = 1; q < S; q++) acc = fadd(acc, part[(size_t)q * n + j]);
    out[j] = acc;
}
// out[i] = sum_j s[j] * M[i*stride + j]; one CTA per row
__global__ void __launch_bounds__(256)
matvec_rows_kernel(const F *__restrict__ M, size_t cols, size_t stride, const F *__restrict__ s, F *__restrict__ out) {
    __shared__ F sred[8];
    const F *row = M + (size_t)blockIdx.x * stride;
    F acc = mkF(0, 0);
    for (size_t j = threadIdx.x; j < cols; j += blockDim.x) acc = fadd(acc, fmul(ldgF(&s[j]), row[j]));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc = fadd(acc, shfl_down_F(acc, d));
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; q++) acc = fadd(acc, sred[q]);
        out[blockIdx.x] = acc;
    }
}
__global__ void __launch_bounds__(256) axpy_vec_kernel(F *__restrict__ y, const F *__restrict__ x, F a, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = fadd(y[i], fmul(a, x[i]));
}
__global__ void __launch_bounds__(256) scatter_kernel(F *__restrict__ out, const unsigned long long *__restrict__ idx, const F *__restrict__ val, size_t m) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) out[idx[k]] = val[k];
}
// out[q*rows + j] = M[j*stride + col[q]]
__global__ void __launch_bounds__(256)
gather_cols_kernel(const F *__restrict__ M, size_t rows, size_t stride, const unsigned long long *__restrict__ col, size_t m, F *__restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * rows) return;
    size_t q = t / rows, j = t - q * rows;
    out[t] = M[j * stride + col[q]];
}

// out[j*m + q] = M[j*stride + col[q]]: the selected columns as a rows x m matrix (message q = column q, the layout hb_encode_batch takes)
__global__ void __launch_bounds__(256)
select_cols_kernel(const F *__restrict__ M, size_t rows, size_t stride, const unsigned long long *__restrict__ col, size_t m, F *__restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * rows) return;
    size_t j = t / m, q = t - j * m;
    out[t] = M[j * stride + col[q]];
}
// out (cols x rows) = in (rows x cols)^T, 32 x 32 tiles through shared memory (both sides coalesced)
__global__ void __launch_bounds__(256) transpose_kernel(const F *__restrict__ in, size_t rows, size_t cols, F *__restrict__ out) {
    __shared__ F tile[32][33];
    const size_t c0 = (size_t)blockIdx.x * 32, r0 = (size_t)blockIdx.y * 32;
    const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (unsigned k = ty; k < 32; k += 8) if (r0 + k < rows && c0 + tx < cols) tile[k][tx] = in[(r0 + k) * cols + c0 + tx];
    __syncthreads();
    for (unsigned k = ty; k < 32; k += 8) if (c0 + k < cols && r0 + tx < rows) out[(c0 + k) * rows + r0 + tx] = tile[tx][k];
}
__global__ void __launch_bounds__(256) vec_nonzero_kernel(const F *__restrict__ v, size_t n, int *__restrict__ flag) {
    bool nz = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) nz |= !fzero(v[i]);
    if (__syncthreads_or(nz) && threadIdx.x == 0) *flag = 1;
}

// ---- phiGInit (utils.cpp:694-755, forward transform): the multilinear extension of the NTT matrix row selected by r ---------------
// phi_g[0] = 1; level i = 1..n-1 doubles the filled prefix: with m = n-i, t1 = 1 - r[m], t2 = r[m] * w^(b << m):
//   phi_g[b + 2^(i-1)] = phi_g[b] (t1 - t2),  phi_g[b] = phi_g[b] (t1 + t2),   b < 2^(i-1);
// final pass: phi_g[b] *= (1 - r[0]) + r[0] w^b, b < 2^(n-1).  Entries >= 2^(n-1) stay zero (the reference never writes them).
// tw[k] = w^k, k < 2^(n-1), is the NTT twiddle table of length 2^n.
__global__ void __launch_bounds__(1024)
phi_g_head_kernel(const F *__restrict__ r, int n, int last_level, const F *__restrict__ tw, F *__restrict__ g) {
    if (threadIdx.x == 0) g[0] = mkF(1, 0);
    __syncthreads();
    for (int i = 1; i <= last_level; i++) {
        const unsigned half = 1u << (i - 1);
        const int m = n - i;
        if (threadIdx.x < half) {
            const unsigned b = threadIdx.x;
            F rm = r[m], t1 = fsub(mkF(1, 0), rm), t2 = fmul(rm, tw[(size_t)b << m]), v = g[b];
            g[b + half] = fmul(v, fsub(t1, t2));
            g[b] = fmul(v, fadd(t1, t2));
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) phi_g_level_kernel(const F *__restrict__ r, int n, int i, const F *__restrict__ tw, F *__restrict__ g) {
    const size_t half = (size_t)1 << (i - 1);
    const int m = n - i;
    const F rm = r[m], t1 = fsub(mkF(1, 0), rm);
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < half; b += (size_t)gridDim.x * blockDim.x) {
        F t2 = fmul(rm, tw[b << m]), v = g[b];
        g[b + half] = fmul(v, fsub(t1, t2));
        g[b] = fmul(v, fadd(t1, t2));
    }
}
__global__ void __launch_bounds__(256) phi_g_final_kernel(const F *__restrict__ r, int n, const F *__restrict__ tw, F *__restrict__ g) {
    const size_t half = (size_t)1 << (n - 1);
    const F r0 = r[0], t1 = fsub(mkF(1, 0), r0);
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < half; b += (size_t)gridDim.x * blockDim.x)
        g[b] = fmul(g[b], fadd(t1, fmul(r0, tw[b])));
}

// ---- shockwave_commit column digests (Virgo.cpp:140-152): MT[0][c] = root of MT_commit_Blake over the k cells of column c.
// With the reference's parent rule H1(left || left) (merkle_tree.cpp:275-280) that root is H1^(log2(k/4)) applied to the first leaf
// H1(col[0..3]); the other k/4 - 1 leaves of the throw-away per-column tree never reach it.
__device__ __forceinline__ void o_cell_words(F x, uint32_t *m) {
    m[0] = (uint32_t)x.re; m[1] = (uint32_t)(x.re >> 32); m[2] = (uint32_t)x.im; m[3] = (uint32_t)(x.im >> 32);
}
__global__ void __launch_bounds__(256) shockwave_leaves_kernel(const F *__restrict__ enc, int lifts, size_t cols, uint8_t *__restrict__ leaves) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    uint32_t m[16], h[8];
#pragma unroll
    for (int q = 0; q < 4; q++) o_cell_words(enc[(size_t)q * cols + c], m + 4 * q);
    blake3_compress64(m, h);
    for (int l = 0; l < lifts; l++) {
#pragma unroll
        for (int i = 0; i < 8; i++) { m[i] = h[i]; m[8 + i] = h[i]; }
        blake3_compress64(m, h);
    }
    uint4 *q = reinterpret_cast<uint4 *>(leaves + c * 32);
    q[0] = make_uint4(h[0], h[1], h[2], h[3]);
    q[1] = make_uint4(h[4], h[5], h[6], h[7]);
}

// ---- WHIR (Virgo.cpp:104-118, 160-178, 519-686) -------------------------------------------------------------------------------------
// change_form, one recursion level: blocks of S = 2^(logn-l) elements, out[pos+i] = in[pos+2i], out[pos+S/2+i] = in[pos+2i+1] - in[pos+2i]
__global__ void __launch_bounds__(256) change_form_level_kernel(const F *__restrict__ in, F *__restrict__ out, size_t n, size_t S) {
    const size_t half = S >> 1;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n / 2; g += (size_t)gridDim.x * blockDim.x) {
        size_t pos = (g / half) * S, i = g % half;
        F a = in[pos + 2 * i], b = in[pos + 2 * i + 1];
        out[pos + i] = a;
        out[pos + half + i] = fsub(b, a);
    }
}
// out[j*K + t] = in[j + t*(n/K)]
__global__ void __launch_bounds__(256) regroup_kernel(const F *__restrict__ in, F *__restrict__ out, size_t n, int k) {
    const size_t K = (size_t)1 << k, q = n >> k;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (size_t)gridDim.x * blockDim.x) {
        size_t j = o >> k, t = o & (K - 1);
        out[o] = in[j + t * q];
    }
}
// round polynomial over the pairs (j, j+L) of (poly, beta): sum_j (p[j] + X dp)(b[j] + X db) -> (a, b, c)
__global__ void __launch_bounds__(256)
whir_poly_kernel(const F *__restrict__ p, const F *__restrict__ b, size_t L, RedArgs ra) {
    F acc[3] = {mkF(0, 0), mkF(0, 0), mkF(0, 0)};
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < L; j += (size_t)gridDim.x * blockDim.x) {
        F x1 = p[j], x2 = b[j], d1 = fsub(p[j + L], x1), d2 = fsub(b[j + L], x2);
        acc[0] = fadd(acc[0], fmul(d1, d2));
        acc[1] = fadd(acc[1], fadd(fmul(d1, x2), fmul(d2, x1)));
        acc[2] = fadd(acc[2], fmul(x1, x2));
    }
    grid_reduce<3>(acc, ra);
}
__global__ void __launch_bounds__(256) whir_fold_kernel(F *__restrict__ p, F *__restrict__ b, size_t L, F a) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < L; j += (size_t)gridDim.x * blockDim.x) {
        F x1 = p[j], x2 = b[j];
        p[j] = fadd(x1, fmul(a, fsub(p[j + L], x1)));
        b[j] = fadd(x2, fmul(a, fsub(b[j + L], x2)));
    }
}
// eq tables of the `repeats` points z_i (v coordinates each), split into a low table (hlo bits) and a high table (v - hlo bits):
// eq(z_i)[j] = lo_i[j & (2^hlo - 1)] * hi_i[j >> hlo];  coordinate k <-> bit k (precompute_beta, utils.cpp:251-296)
__global__ void __launch_bounds__(256) whir_eq_tables_kernel(const F *__restrict__ z, int v, int hlo, F *__restrict__ lo, F *__restrict__ hi) {
    const F *zi = z + (size_t)blockIdx.x * v;
    const unsigned nlo = 1u << hlo, nhi = 1u << (v - hlo);
    for (unsigned t = threadIdx.x; t < nlo + nhi; t += blockDim.x) {
        const bool high = t >= nlo;
        const unsigned idx = high ? t - nlo : t;
        const int k0 = high ? hlo : 0, nk = high ? v - hlo : hlo;
        F e = mkF(1, 0);
        for (int k = 0; k < nk; k++) { F zk = zi[k0 + k]; e = fmul(e, ((idx >> k) & 1) ? zk : fsub(mkF(1, 0), zk)); }
        (high ? hi + (size_t)blockIdx.x * nhi : lo + (size_t)blockIdx.x * nlo)[idx] = e;
    }
}
// y[i] = sum_j eq(z_i)[j] poly[j]; one CTA per point
__global__ void __launch_bounds__(256)
whir_zeta_y_kernel(const F *__restrict__ poly, size_t n, int hlo, int v, const F *__restrict__ lo, const F *__restrict__ hi, F *__restrict__ y) {
    __shared__ F sred[8];
    const F *l = lo + ((size_t)blockIdx.x << hlo), *h = hi + ((size_t)blockIdx.x << (v - hlo));
    const size_t mask = ((size_t)1 << hlo) - 1;
    F acc = mkF(0, 0);
    for (size_t j = threadIdx.x; j < n; j += blockDim.x) acc = fadd(acc, fmul(fmul(l[j & mask], h[j >> hlo]), poly[j]));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc = fadd(acc, shfl_down_F(acc, d));
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { for (int q = 1; q < 8; q++) acc = fadd(acc, sred[q]); y[blockIdx.x] = acc; }
}
// beta[j] += sum_i pows[i] eq(z_i)[j]
__global__ void __launch_bounds__(256)
whir_zeta_beta_kernel(F *__restrict__ beta, size_t n, int repeats, int hlo, int v, const F *__restrict__ lo, const F *__restrict__ hi, const F *__restrict__ pows) {
    const size_t mask = ((size_t)1 << hlo) - 1;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        F acc = beta[j];
        for (int i = 0; i < repeats; i++)
            acc = fadd(acc, fmul(ldgF(&pows[i]), fmul(lo[((size_t)i << hlo) + (j & mask)], hi[((size_t)i << (v - hlo)) + (j >> hlo)])));
        beta[j] = acc;
    }
}

}  // namespace hb

using namespace hb;

extern "C" int hb_vec_zero(hb_ctx *ctx, hb_F *v, size_t n) { HB_DEV(ctx);
    if (n == 0) return 0;
    if (is_device_ptr(v)) HB_CHECK(ctx, cudaMemsetAsync(v, 0, n * sizeof(F), ctx->stream));
    else { HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream)); memset(v, 0, n * sizeof(F)); }
    return 0;
}

extern "C" int hb_rs_encode_rows(hb_ctx *ctx, const hb_F *src, size_t in_len, size_t rows, hb_F *dst, int logn) { HB_DEV(ctx);
    const size_t len = (size_t)1 << logn;
    if (logn < 0 || logn > 28 || in_len > len) HB_FAIL(ctx, "hb_rs_encode_rows: need in_len <= 2^logn <= 2^28");
    if (rows == 0) return 0;
    Staged s(ctx), d(ctx);
    HB_TRY(s.in(src, rows * in_len * sizeof(F)));
    HB_TRY(d.outbuf(dst, rows * len * sizeof(F)));
    HB_TRY(ntt_rows_padded_dev(ctx, s.as<F>(), in_len, d.as<F>(), len, logn, rows, 1, 0, 0));
    HB_TRY(d.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_matvec_cols(hb_ctx *ctx, const hb_F *M, size_t rows, size_t cols, size_t stride, const hb_F *w, hb_F *out) { HB_DEV(ctx);
    if (rows == 0 || cols == 0) return 0;
    if (stride < cols) HB_FAIL(ctx, "hb_matvec_cols: stride < cols");
    Staged m(ctx), sw(ctx), o(ctx);
    HB_TRY(m.in(M, ((rows - 1) * stride + cols) * sizeof(F)));
    HB_TRY(sw.in(w, rows * sizeof(F)));
    HB_TRY(o.outbuf(out, cols * sizeof(F)));
    const unsigned gx = (unsigned)((cols + 31) / 32);
    unsigned S = 1;
    if (gx < (unsigned)ctx->sm_count * 4) S = (unsigned)std::min<size_t>(std::max<size_t>(1, rows / 8), ((size_t)ctx->sm_count * 4 + gx - 1) / gx);
    S = std::min(S, 64u);
    if (S == 1) {
        HB_LAUNCH(ctx, matvec_cols_kernel, dim3(gx, 1), 256, 0, m.as<F>(), rows, cols, stride, sw.as<F>(), o.as<F>());
    } else {
        F *part; HB_CHECK(ctx, cudaMallocAsync(&part, (size_t)S * cols * sizeof(F), ctx->stream));
        HB_LAUNCH(ctx, matvec_cols_kernel, dim3(gx, S), 256, 0, m.as<F>(), rows, cols, stride, sw.as<F>(), part);
        HB_LAUNCH(ctx, sum_partials_kernel, (unsigned)((cols + 255) / 256), 256, 0, part, cols, S, o.as<F>());
        cudaFreeAsync(part, ctx->stream);
    }
    HB_TRY(o.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_matvec_rows(hb_ctx *ctx, const hb_F *M, size_t rows, size_t cols, size_t stride, const hb_F *s, hb_F *out) { HB_DEV(ctx);
    if (rows == 0) return 0;
    if (stride < cols) HB_FAIL(ctx, "hb_matvec_rows: stride < cols");
    Staged m(ctx), ss(ctx), o(ctx);
    HB_TRY(m.in(M, ((rows - 1) * stride + cols) * sizeof(F)));
    HB_TRY(ss.in(s, cols * sizeof(F)));
    HB_TRY(o.outbuf(out, rows * sizeof(F)));
    HB_LAUNCH(ctx, matvec_rows_kernel, (unsigned)rows, 256, 0, m.as<F>(), cols, stride, ss.as<F>(), o.as<F>());
    HB_TRY(o.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_axpy(hb_ctx *ctx, hb_F *y, const hb_F *x, const hb_F *a, size_t n) { HB_DEV(ctx);
    if (n == 0) return 0;
    Staged sy(ctx), sx(ctx);
    HB_TRY(sy.outbuf(y, n * sizeof(F), true));
    HB_TRY(sx.in(x, n * sizeof(F)));
    hb_F ah; HB_CHECK(ctx, cudaMemcpy(&ah, a, sizeof(F), cudaMemcpyDefault));
    HB_LAUNCH(ctx, axpy_vec_kernel, grid_1d(ctx, n), 256, 0, sy.as<F>(), sx.as<F>(), mkF(ah.real, ah.img), n);
    HB_TRY(sy.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_scatter(hb_ctx *ctx, hb_F *out, size_t n, const uint64_t *idx, const hb_F *val, size_t m) { HB_DEV(ctx);
    if (n == 0) return 0;
    Staged so(ctx), si(ctx), sv(ctx);
    HB_TRY(so.outbuf(out, n * sizeof(F)));
    HB_CHECK(ctx, cudaMemsetAsync(so.dev, 0, n * sizeof(F), ctx->stream));
    if (m) {
        HB_TRY(si.in(idx, m * sizeof(uint64_t)));
        HB_TRY(sv.in(val, m * sizeof(F)));
        HB_LAUNCH(ctx, scatter_kernel, (unsigned)((m + 255) / 256), 256, 0, so.as<F>(), si.as<unsigned long long>(), sv.as<F>(), m);
    }
    HB_TRY(so.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_gather_cols(hb_ctx *ctx, const hb_F *M, size_t rows, size_t cols, size_t stride, const uint64_t *col, size_t m, hb_F *out) { HB_DEV(ctx);
    if (m == 0 || rows == 0) return 0;
    Staged sm(ctx), sc(ctx), so(ctx);
    HB_TRY(sm.in(M, ((rows - 1) * stride + cols) * sizeof(F)));
    HB_TRY(sc.in(col, m * sizeof(uint64_t)));
    HB_TRY(so.outbuf(out, m * rows * sizeof(F)));
    HB_LAUNCH(ctx, gather_cols_kernel, (unsigned)((m * rows + 255) / 256), 256, 0, sm.as<F>(), rows, stride, sc.as<unsigned long long>(), m, so.as<F>());
    HB_TRY(so.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_select_cols(hb_ctx *ctx, const hb_F *M, size_t rows, size_t cols, size_t stride, const uint64_t *col, size_t m, hb_F *out) { HB_DEV(ctx);
    if (m == 0 || rows == 0) return 0;
    Staged sm(ctx), sc(ctx), so(ctx);
    HB_TRY(sm.in(M, ((rows - 1) * stride + cols) * sizeof(F)));
    HB_TRY(sc.in(col, m * sizeof(uint64_t)));
    HB_TRY(so.outbuf(out, m * rows * sizeof(F)));
    HB_LAUNCH(ctx, select_cols_kernel, (unsigned)((m * rows + 255) / 256), 256, 0, sm.as<F>(), rows, stride, sc.as<unsigned long long>(), m, so.as<F>());
    HB_TRY(so.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_transpose(hb_ctx *ctx, const hb_F *in, size_t rows, size_t cols, hb_F *out) { HB_DEV(ctx);
    if (rows == 0 || cols == 0) return 0;
    Staged si(ctx), so(ctx);
    HB_TRY(si.in(in, rows * cols * sizeof(F)));
    HB_TRY(so.outbuf(out, rows * cols * sizeof(F)));
    HB_LAUNCH(ctx, transpose_kernel, dim3((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32)), 256, 0, si.as<F>(), rows, cols, so.as<F>());
    HB_TRY(so.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_any_nonzero(hb_ctx *ctx, const hb_F *v, size_t n, int *out) { HB_DEV(ctx);
    *out = 0;
    if (n == 0) return 0;
    if (!is_device_ptr(v)) {                       // host chunk: a linear scan on the host beats a round trip
        HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        const uint64_t *w = reinterpret_cast<const uint64_t *>(v);
        for (size_t i = 0; i < 2 * n; i++) if (w[i]) { *out = 1; break; }
        return 0;
    }
    int *flag; HB_CHECK(ctx, cudaMallocAsync(&flag, sizeof(int), ctx->stream));
    HB_CHECK(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    HB_LAUNCH(ctx, vec_nonzero_kernel, grid_1d(ctx, n), 256, 0, reinterpret_cast<const F *>(v), n, flag);
    HB_CHECK(ctx, cudaMemcpyAsync(out, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(flag, ctx->stream);
    return 0;
}

extern "C" int hb_phi_g_init(hb_ctx *ctx, const hb_F *r, int n, hb_F *out) { HB_DEV(ctx);
    if (n < 1 || n > 28) HB_FAIL(ctx, "hb_phi_g_init: n out of range");
    const size_t N = (size_t)1 << n;
    Staged sr(ctx), so(ctx);
    HB_TRY(sr.in(r, (size_t)n * sizeof(F)));
    HB_TRY(so.outbuf(out, N * sizeof(F)));
    HB_CHECK(ctx, cudaMemsetAsync(so.dev, 0, N * sizeof(F), ctx->stream));
    const F *tw; HB_TRY(get_twiddles(ctx, n, &tw));
    const int head = std::min(n - 1, 11);
    HB_LAUNCH(ctx, phi_g_head_kernel, 1, 1024, 0, sr.as<F>(), n, head, tw, so.as<F>());
    for (int i = head + 1; i <= n - 1; i++)
        HB_LAUNCH(ctx, phi_g_level_kernel, grid_1d(ctx, (size_t)1 << (i - 1)), 256, 0, sr.as<F>(), n, i, tw, so.as<F>());
    HB_LAUNCH(ctx, phi_g_final_kernel, grid_1d(ctx, N / 2), 256, 0, sr.as<F>(), n, tw, so.as<F>());
    HB_TRY(so.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_shockwave_leaves(hb_ctx *ctx, const hb_F *enc, int k, size_t cols, uint8_t *leaves) { HB_DEV(ctx);
    if (k < 4 || (k & (k - 1))) HB_FAIL(ctx, "hb_shockwave_leaves: k must be a power of two >= 4");
    if (cols == 0) return 0;
    Staged se(ctx), sl(ctx);
    HB_TRY(se.in(enc, (size_t)k * cols * sizeof(F)));
    HB_TRY(sl.outbuf(leaves, cols * 32));
    HB_LAUNCH(ctx, shockwave_leaves_kernel, (unsigned)((cols + 255) / 256), 256, 0, se.as<F>(), ilog2((size_t)k / 4), cols, sl.as<uint8_t>());
    HB_TRY(sl.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_change_form(hb_ctx *ctx, hb_F *poly, int logn) { HB_DEV(ctx);
    if (logn < 1 || logn > 30) HB_FAIL(ctx, "hb_change_form: logn out of range");
    const size_t n = (size_t)1 << logn;
    Staged sp(ctx);
    HB_TRY(sp.outbuf(poly, n * sizeof(F), true));
    F *tmp; HB_CHECK(ctx, cudaMallocAsync(&tmp, n * sizeof(F), ctx->stream));
    F *a = sp.as<F>(), *b = tmp;
    for (int l = 0; l < logn; l++) {
        HB_LAUNCH(ctx, change_form_level_kernel, grid_1d(ctx, n / 2), 256, 0, a, b, n, n >> l);
        std::swap(a, b);
    }
    if (a != sp.as<F>()) HB_CHECK(ctx, cudaMemcpyAsync(sp.dev, a, n * sizeof(F), cudaMemcpyDeviceToDevice, ctx->stream));
    cudaFreeAsync(tmp, ctx->stream);
    HB_TRY(sp.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_regroup(hb_ctx *ctx, const hb_F *in, size_t n, int k, hb_F *out) { HB_DEV(ctx);
    if (n == 0 || (n & (n - 1)) || ((size_t)1 << k) > n) HB_FAIL(ctx, "hb_regroup: n must be a power of two >= 2^k");
    Staged si(ctx), so(ctx);
    HB_TRY(si.in(in, n * sizeof(F)));
    HB_TRY(so.outbuf(out, n * sizeof(F)));
    HB_LAUNCH(ctx, regroup_kernel, grid_1d(ctx, n), 256, 0, si.as<F>(), so.as<F>(), n, k);
    HB_TRY(so.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_whir_poly(hb_ctx *ctx, const hb_F *poly, const hb_F *beta, size_t L, hb_F *coeffs3) { HB_DEV(ctx);
    if (L == 0) HB_FAIL(ctx, "hb_whir_poly: L must be positive");
    HB_TRY(ensure_scratch(ctx));
    Staged sp(ctx), sb(ctx);
    HB_TRY(sp.in(poly, 2 * L * sizeof(F))); HB_TRY(sb.in(beta, 2 * L * sizeof(F)));
    HB_LAUNCH(ctx, whir_poly_kernel, red_grid_for(ctx, L), 256, 0, sp.as<F>(), sb.as<F>(), L, red_args(ctx));
    F co[3]; HB_TRY(read_result(ctx, 3, co));
    for (int c = 0; c < 3; c++) { coeffs3[c].real = co[c].re; coeffs3[c].img = co[c].im; }
    return 0;
}

extern "C" int hb_whir_fold(hb_ctx *ctx, hb_F *poly, hb_F *beta, size_t L, const hb_F *a) { HB_DEV(ctx);
    if (L == 0) return 0;
    Staged sp(ctx), sb(ctx);
    HB_TRY(sp.outbuf(poly, 2 * L * sizeof(F), true)); HB_TRY(sb.outbuf(beta, 2 * L * sizeof(F), true));
    HB_LAUNCH(ctx, whir_fold_kernel, grid_1d(ctx, L), 256, 0, sp.as<F>(), sb.as<F>(), L, mkF(a->real, a->img));
    HB_TRY(sp.finish()); HB_TRY(sb.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_whir_zeta(hb_ctx *ctx, const hb_F *poly, hb_F *beta, int v, const hb_F *zetas, int repeats, const hb_F *pows, hb_F *y) { HB_DEV(ctx);
    if (v < 1 || v > 20 || repeats < 1) HB_FAIL(ctx, "hb_whir_zeta: need 1 <= v <= 20 and repeats >= 1");
    const size_t n = (size_t)1 << v;
    const int hlo = v / 2;
    Staged sp(ctx), sb(ctx), sz(ctx), spw(ctx), sy(ctx);
    HB_TRY(sp.in(poly, n * sizeof(F)));
    HB_TRY(sb.outbuf(beta, n * sizeof(F), true));
    HB_TRY(sz.in(zetas, (size_t)repeats * v * sizeof(F)));
    HB_TRY(spw.in(pows, (size_t)repeats * sizeof(F)));
    HB_TRY(sy.outbuf(y, (size_t)repeats * sizeof(F)));
    const size_t nlo = (size_t)1 << hlo, nhi = (size_t)1 << (v - hlo);
    F *tab; HB_CHECK(ctx, cudaMallocAsync(&tab, (size_t)repeats * (nlo + nhi) * sizeof(F), ctx->stream));
    F *lo = tab, *hi = tab + (size_t)repeats * nlo;
    HB_LAUNCH(ctx, whir_eq_tables_kernel, (unsigned)repeats, 256, 0, sz.as<F>(), v, hlo, lo, hi);
    HB_LAUNCH(ctx, whir_zeta_y_kernel, (unsigned)repeats, 256, 0, sp.as<F>(), n, hlo, v, lo, hi, sy.as<F>());
    HB_LAUNCH(ctx, whir_zeta_beta_kernel, grid_1d(ctx, n), 256, 0, sb.as<F>(), n, repeats, hlo, v, lo, hi, spw.as<F>());
    cudaFreeAsync(tab, ctx->stream);
    HB_TRY(sy.finish()); HB_TRY(sb.finish());
    HB_TRY(end_call(ctx));
    return 0;
}
