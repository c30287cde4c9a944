// E1/E2 — Orion/Spielman expander linear code, batched over all columns of the tensor.
// Reference: encode_monolithic (src/linear_code_encode.h:62-119) over graphs built by expander_init_store
// (src/expanders.h:20-47,78-92).  The recursion
//     enc_d(x) = x | enc_{d+1}(C_d x) | D_d * enc_{d+1}(C_d x)
// is flattened into a list of sparse mat-vec STAGES over one codeword buffer (C_0, C_1, .., C_last, D_last, .., D_0);
// each stage reads an already-final segment of the codeword and writes a disjoint later segment.
//
// B200 mapping: all columns share one graph, so a CTA takes CB adjacent columns and keeps their whole codewords
// in shared memory as cw[row][CB] (n=1024: 1761 rows x 8 columns x 16 B = 220 KB, one CTA per SM).  A thread owns
// (target row, column); each stage is a GATHER over the target's in-edges (CSR by target, built on the host from
// the reference's scatter lists) so there are no atomics, and the 61x32-bit products are accumulated in 128 bits
// and reduced once per target.  A quarter-warp reads one contiguous 16*CB-byte row -> conflict-free LDS.128.
// The weights are 31-bit reals (expanders.h:37 `F weight = random()`), so F x weight is two 61x32 products.
// The commit_standard leaf hashing (Our_PC.cpp:160-166) can be fused here: the CTA already holds every row of its
// columns, so the 4-row quads are hashed straight out of shared memory and the tensor is never re-read.
#include "common.cuh"
#include "blake3.cuh"
#include <algorithm>

namespace hb {

struct Acc128 { u64 lo, hi; };
__device__ __forceinline__ void acc_mul(Acc128 &a, u64 x, uint32_t w) {
    u64 lo, hi; mul61x32_wide(x, w, lo, hi);
    a.lo += lo; a.hi += hi + (a.lo < lo);
}

template <int CB, bool FUSE_LEAVES>
__global__ void __launch_bounds__(512)
encode_cols_kernel(F *__restrict__ T, size_t cols, int n, int cwlen,
                   const EncStage *__restrict__ stages, int nstages,
                   const int *__restrict__ rowptr, const uint2 *__restrict__ edges,
                   uint8_t *__restrict__ leaves) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F *cw = reinterpret_cast<F *>(smem_raw);                   // cw[row * CB + c]
    const size_t col0 = (size_t)blockIdx.x * CB;
    const unsigned c = threadIdx.x % CB, t0 = threadIdx.x / CB, tstep = blockDim.x / CB;

    for (unsigned r = t0; r < (unsigned)n; r += tstep) cw[r * CB + c] = T[(size_t)r * cols + col0 + c];
    __syncthreads();

    for (int s = 0; s < nstages; s++) {
        const EncStage st = stages[s];
        const int *rp = rowptr + st.rowptr_base;
        for (unsigned t = t0; t < (unsigned)st.R; t += tstep) {
            int e0 = __ldg(&rp[t]), e1 = __ldg(&rp[t + 1]);
            Acc128 are = {0, 0}, aim = {0, 0};
#pragma unroll 4
            for (int e = e0; e < e1; e++) {
                uint2 ed = __ldg(&edges[e]);
                F x = cw[ed.x * CB + c];
                acc_mul(are, x.re, ed.y);
                acc_mul(aim, x.im, ed.y);
            }
            cw[(st.out_off + t) * CB + c] = mkF(red128(are.lo, are.hi), red128(aim.lo, aim.hi));
        }
        __syncthreads();
    }

    // rows [n, cwlen) are new; rows [cwlen, 2n) are the zero tail of the reference's 2n-sized buffer
    for (unsigned r = n + t0; r < 2u * n; r += tstep)
        T[(size_t)r * cols + col0 + c] = (r < (unsigned)cwlen) ? cw[r * CB + c] : mkF(0, 0);

    if (FUSE_LEAVES) {
        // leaf (j, k) <- H1( H1(T[4j][k] | T[4j+1][k] | T[4j+2][k] | T[4j+3][k]) | leaf(j, k) ), j < n/2
        for (unsigned j = t0; j < (unsigned)n / 2; j += tstep) {
            uint32_t m[16], prev[8], out[8];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                unsigned r = 4 * j + q;
                F x = (r < (unsigned)cwlen) ? cw[r * CB + c] : mkF(0, 0);
                m[4 * q] = (uint32_t)x.re; m[4 * q + 1] = (uint32_t)(x.re >> 32);
                m[4 * q + 2] = (uint32_t)x.im; m[4 * q + 3] = (uint32_t)(x.im >> 32);
            }
            uint4 *lp = reinterpret_cast<uint4 *>(leaves + ((size_t)j * cols + col0 + c) * 32);
            uint4 p0 = lp[0], p1 = lp[1];
            prev[0] = p0.x; prev[1] = p0.y; prev[2] = p0.z; prev[3] = p0.w;
            prev[4] = p1.x; prev[5] = p1.y; prev[6] = p1.z; prev[7] = p1.w;
            md_leaf(m, prev, out);
            lp[0] = make_uint4(out[0], out[1], out[2], out[3]);
            lp[1] = make_uint4(out[4], out[5], out[6], out[7]);
        }
    }
}

template <int CB>
static int launch_encode(hb_ctx *ctx, F *T, long long n, size_t cols, uint8_t *leaves) {
    const ExpanderDev &ex = ctx->exp;
    size_t smem = (size_t)ex.cwlen * CB * sizeof(F);
    unsigned grid = (unsigned)(cols / CB);
    if (leaves) {
        HB_CHECK(ctx, cudaFuncSetAttribute(encode_cols_kernel<CB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HB_LAUNCH(ctx, (encode_cols_kernel<CB, true>), grid, 512, smem, T, cols, (int)n, ex.cwlen, ex.d_stages, (int)ex.stages.size(),
                  ex.d_rowptr, ex.d_edges, leaves);
    } else {
        HB_CHECK(ctx, cudaFuncSetAttribute(encode_cols_kernel<CB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HB_LAUNCH(ctx, (encode_cols_kernel<CB, false>), grid, 512, smem, T, cols, (int)n, ex.cwlen, ex.d_stages, (int)ex.stages.size(),
                  ex.d_rowptr, ex.d_edges, leaves);
    }
    return 0;
}

int encode_cols_dev(hb_ctx *ctx, F *T, long long n, size_t cols, uint8_t *leaves) {
    const ExpanderDev &ex = ctx->exp;
    if (ex.n != n) HB_FAIL(ctx, "encode: no expander installed for this message length (call hb_expander_set / expander_init_store first)");
    const size_t kMaxSmem = 227 * 1024;
    size_t per_col = (size_t)ex.cwlen * sizeof(F);
    // widest column block that fits; prefer <= ~100 KB tiles when the code is small so several CTAs share an SM
    if (cols % 32 == 0 && per_col * 32 <= 100 * 1024) return launch_encode<32>(ctx, T, n, cols, leaves);
    if (cols % 16 == 0 && per_col * 16 <= 100 * 1024) return launch_encode<16>(ctx, T, n, cols, leaves);
    if (cols % 8 == 0 && per_col * 8 <= kMaxSmem) return launch_encode<8>(ctx, T, n, cols, leaves);
    if (cols % 4 == 0 && per_col * 4 <= kMaxSmem) return launch_encode<4>(ctx, T, n, cols, leaves);
    if (cols % 2 == 0 && per_col * 2 <= kMaxSmem) return launch_encode<2>(ctx, T, n, cols, leaves);
    if (per_col <= kMaxSmem) return launch_encode<1>(ctx, T, n, cols, leaves);
    HB_FAIL(ctx, "encode: codeword does not fit in shared memory (message length too large for the column kernel)");
}

}  // namespace hb

// ---------------------------------------------------------------------------------------------------------
extern "C" int hb_expander_set(hb_ctx *ctx, long long n, int levels, int deg_C, int deg_D,
                               const long long *L_C, const long long *R_C, const uint32_t *const *nbr_C, const uint64_t *const *w_C,
                               const long long *L_D, const long long *R_D, const uint32_t *const *nbr_D, const uint64_t *const *w_D) {
    using namespace hb;
    ExpanderDev &ex = ctx->exp;
    if (ex.d_stages) { cudaFree(ex.d_stages); cudaFree(ex.d_rowptr); cudaFree(ex.d_edges); }
    ex = ExpanderDev();
    ex.n = n;
    // Codeword layout of encode_monolithic: level d occupies [off_d, off_d + len_d) with
    //   x_d at off_d (n_d entries), enc_{d+1} at off_d + n_d (Lenc_d entries), z_d after it (R_D[d] entries).
    std::vector<long long> nd(levels + 1), off(levels + 1), lenc(levels + 1);
    nd[0] = n; off[0] = 0;
    for (int d = 0; d < levels; d++) {
        if (L_C[d] != nd[d]) HB_FAIL(ctx, "hb_expander_set: C graph left size does not match the recursion");
        nd[d + 1] = R_C[d];
        off[d + 1] = off[d] + nd[d];
    }
    lenc[levels] = nd[levels];                       // base case copies its input (n <= distance_threshold)
    for (int d = levels - 1; d >= 0; d--) {
        if (L_D[d] != lenc[d + 1]) HB_FAIL(ctx, "hb_expander_set: D graph left size does not match the recursion");
        lenc[d] = nd[d] + lenc[d + 1] + R_D[d];
    }
    ex.cwlen = (int)lenc[0];
    if (levels == 0) return 0;

    std::vector<int> rowptr; std::vector<uint2> edges;
    auto add_stage = [&](long long in_off, long long L, long long out_off, long long R, int deg, const uint32_t *nbr, const uint64_t *w) -> int {
        EncStage st; st.in_off = (int)in_off; st.out_off = (int)out_off; st.L = (int)L; st.R = (int)R; st.rowptr_base = (int)rowptr.size();
        std::vector<int> cnt(R + 1, 0);
        for (long long i = 0; i < L * deg; i++) { if (nbr[i] >= (uint32_t)R) return 1; if (w[i] >> 32) return 1; cnt[nbr[i] + 1]++; }
        for (long long t = 0; t < R; t++) { ex.max_indeg = std::max(ex.max_indeg, cnt[t + 1]); cnt[t + 1] += cnt[t]; }
        size_t ebase = edges.size();
        edges.resize(ebase + (size_t)L * deg);
        std::vector<int> fill(cnt.begin(), cnt.end() - 1);
        for (long long i = 0; i < L; i++)
            for (int j = 0; j < deg; j++) {
                uint32_t t = nbr[i * deg + j];
                edges[ebase + fill[t]++] = make_uint2((unsigned)(in_off + i), (unsigned)w[i * deg + j]);
            }
        for (long long t = 0; t <= R; t++) rowptr.push_back((int)ebase + cnt[t]);
        ex.stages.push_back(st);
        return 0;
    };
    for (int d = 0; d < levels; d++)
        if (add_stage(off[d], nd[d], off[d + 1], R_C[d], deg_C, nbr_C[d], w_C[d])) HB_FAIL(ctx, "hb_expander_set: bad C graph (target >= R or weight >= 2^32)");
    for (int d = levels - 1; d >= 0; d--)
        if (add_stage(off[d + 1], lenc[d + 1], off[d + 1] + lenc[d + 1], R_D[d], deg_D, nbr_D[d], w_D[d])) HB_FAIL(ctx, "hb_expander_set: bad D graph (target >= R or weight >= 2^32)");
    ex.n_edges = edges.size();
    HB_CHECK(ctx, cudaMalloc(&ex.d_stages, ex.stages.size() * sizeof(EncStage)));
    HB_CHECK(ctx, cudaMalloc(&ex.d_rowptr, rowptr.size() * sizeof(int)));
    HB_CHECK(ctx, cudaMalloc(&ex.d_edges, edges.size() * sizeof(uint2)));
    HB_CHECK(ctx, cudaMemcpy(ex.d_stages, ex.stages.data(), ex.stages.size() * sizeof(EncStage), cudaMemcpyHostToDevice));
    HB_CHECK(ctx, cudaMemcpy(ex.d_rowptr, rowptr.data(), rowptr.size() * sizeof(int), cudaMemcpyHostToDevice));
    HB_CHECK(ctx, cudaMemcpy(ex.d_edges, edges.data(), edges.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" long long hb_expander_codeword_len(hb_ctx *ctx) { return ctx->exp.cwlen; }

extern "C" int hb_encode_batch(hb_ctx *ctx, const hb_F *src, hb_F *dst, long long n, size_t ncols) {
    using namespace hb;
    if (ncols == 0) return 0;
    Staged d(ctx);
    HB_TRY(d.outbuf(dst, 2 * (size_t)n * ncols * sizeof(F)));
    HB_CHECK(ctx, cudaMemcpyAsync(d.dev, src, (size_t)n * ncols * sizeof(F), cudaMemcpyDefault, ctx->stream));
    if (n <= 13 && ctx->exp.n != n) {
        // base case of the recursion: the codeword is the message (linear_code_encode.h:73-78)
        HB_CHECK(ctx, cudaMemsetAsync(d.as<F>() + (size_t)n * ncols, 0, (size_t)n * ncols * sizeof(F), ctx->stream));
    } else {
        HB_TRY(encode_cols_dev(ctx, d.as<F>(), n, ncols, nullptr));
    }
    HB_TRY(d.finish());
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
