// N1 — batched NTT over F_{p^2}, bit-identical to the reference's `_fft(arr, logn, false)`
// (src/utils.cpp:605-673): bit-reversal permutation, twiddles w[k] = omega^k with
// omega = getRootOfUnity(logn) (utils.cpp:452-463), radix-2 DIT stages
//     u = a[j+k]; v = a[j+k+i/2] * w[len/i*k]; a[j+k] = u+v; a[j+k+i/2] = u-v.
// Field arithmetic is exact and canonical, so the stage order/grouping below is free to differ.
//
// B200 mapping: one CTA owns up to 4096 consecutive (bit-reversed) positions of one row in shared memory
// (64 KB -> 3 CTAs/SM) and runs the first 12 stages there; longer transforms finish with register-resident
// radix-2^k passes over global memory (coalesced: consecutive threads own consecutive low index bits).
// The zero-extension of the message rows (the RS encoding evaluates a degree < len/2 polynomial on len points)
// is fused into the load, so the tensor's upper half is never memset nor read.
#include "common.cuh"

namespace hb {

static constexpr int kLogTile = 12;             // 4096 elements * 16 B = 64 KB of shared memory

// ---------------------------------------------------------------------------------------------------------
// twiddles: computed once per length on the host with the same repeated multiplication as the reference
int get_twiddles(hb_ctx *ctx, int logn, const F **out) {
    if (logn < 1 || logn > 31) HB_FAIL(ctx, "get_twiddles: logn out of range");
    if (!ctx->tw[logn]) {
        size_t half = (size_t)1 << (logn - 1);
        std::vector<F> w(half);
        hb_F rou; hb_root_of_unity(logn, &rou);
        F w1 = mkF(rou.real, rou.img);
        w[0] = mkF(1, 0);
        for (size_t i = 1; i < half; i++) w[i] = h_fmul(w[i - 1], w1);
        HB_CHECK(ctx, cudaMalloc(&ctx->tw[logn], half * sizeof(F)));
        HB_CHECK(ctx, cudaMemcpyAsync(ctx->tw[logn], w.data(), half * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
        HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // w is a local
    }
    *out = ctx->tw[logn];
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Tile kernel: stages 1..lb of a length-2^logn transform on positions [tile*2^lb, (tile+1)*2^lb) of row `r`.
// Loads src[rev(p)] (zero if rev(p) >= in_len), writes dst[p].  src may alias dst only when lb == logn
// (then the CTA reads its whole row before it writes anything).
__global__ void __launch_bounds__(512)
ntt_tile_kernel(const F *__restrict__ src, size_t src_stride, size_t in_len, F *__restrict__ dst, size_t dst_stride,
                int logn, int lb, const F *__restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F *s = reinterpret_cast<F *>(smem_raw);
    const unsigned tiles_per_row = 1u << (logn - lb);
    const size_t row = blockIdx.x / tiles_per_row;
    const unsigned tile = blockIdx.x % tiles_per_row;
    const unsigned tlen = 1u << lb, base = tile << lb;
    const F *in = src + row * src_stride;
    F *out = dst + row * dst_stride;

    for (unsigned i = threadIdx.x; i < tlen; i += blockDim.x) {
        unsigned p = base + i;
        unsigned q = __brev(p) >> (32 - logn);
        s[i] = (q < in_len) ? in[q] : mkF(0, 0);
    }
    __syncthreads();
    const unsigned len_half_stride = 1u << logn;      // twiddle index = (len >> st) * k
    for (int st = 1; st <= lb; st++) {
        const unsigned half = 1u << (st - 1);
        const unsigned tws = len_half_stride >> st;
        for (unsigned b = threadIdx.x; b < (tlen >> 1); b += blockDim.x) {
            unsigned k = b & (half - 1);
            unsigned j = (b >> (st - 1)) << st;
            unsigned p0 = j + k, p1 = p0 + half;
            F u = s[p0];
            F x = s[p1];
            F v = (k == 0) ? x : fmul(x, ldgF(&tw[(size_t)tws * k]));
            s[p0] = fadd(u, v);
            s[p1] = fsub(u, v);
        }
        __syncthreads();
    }
    for (unsigned i = threadIdx.x; i < tlen; i += blockDim.x) out[base + i] = s[i];
}

// Register-resident radix-2^CNT pass over global memory: stages s_lo+1 .. s_lo+CNT of a length-2^logn transform.
template <int CNT>
__global__ void __launch_bounds__(256)
ntt_global_kernel(F *__restrict__ data, size_t stride, int logn, int s_lo, const F *__restrict__ tw, size_t total_groups) {
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
#pragma unroll
    for (int q = 0; q < (1 << CNT); q++) a[base + ((unsigned)q << s_lo)] = v[q];
}

static int ntt_rows_impl(hb_ctx *ctx, const F *src, size_t src_stride, size_t in_len, F *dst, size_t dst_stride,
                         int logn, size_t batch) {
    if (batch == 0) return 0;
    if (logn == 0) {
        if (src != dst) HB_CHECK(ctx, cudaMemcpy2DAsync(dst, dst_stride * sizeof(F), src, src_stride * sizeof(F), sizeof(F), batch,
                                                         cudaMemcpyDeviceToDevice, ctx->stream));
        return 0;
    }
    const F *tw; HB_TRY(get_twiddles(ctx, logn, &tw));
    const int lb = logn < kLogTile ? logn : kLogTile;
    if (src == dst && lb != logn) HB_FAIL(ctx, "ntt: in-place transform longer than one tile needs a distinct source");
    const size_t smem = sizeof(F) << lb;
    static bool attr_set = false;
    if (!attr_set) {
        HB_CHECK(ctx, cudaFuncSetAttribute(ntt_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(F) << kLogTile)));
        attr_set = true;
    }
    unsigned threads = 1u << (lb > 1 ? lb - 1 : 0);
    if (threads > 512) threads = 512;
    if (threads < 32) threads = 32;
    size_t grid = batch << (logn - lb);
    HB_LAUNCH(ctx, ntt_tile_kernel, (unsigned)grid, threads, smem, src, src_stride, in_len, dst, dst_stride, logn, lb, tw);
    int s_lo = lb;
    while (s_lo < logn) {
        int cnt = logn - s_lo; if (cnt > 3) cnt = 3;
        size_t groups = batch << (logn - cnt);
        unsigned g = (unsigned)((groups + 255) / 256);
        if (cnt == 3) HB_LAUNCH(ctx, ntt_global_kernel<3>, g, 256, 0, dst, dst_stride, logn, s_lo, tw, groups);
        else if (cnt == 2) HB_LAUNCH(ctx, ntt_global_kernel<2>, g, 256, 0, dst, dst_stride, logn, s_lo, tw, groups);
        else HB_LAUNCH(ctx, ntt_global_kernel<1>, g, 256, 0, dst, dst_stride, logn, s_lo, tw, groups);
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

int ntt_rows_padded_dev(hb_ctx *ctx, const F *src, size_t in_len, F *dst, size_t dst_stride, int logn, size_t batch) {
    return ntt_rows_impl(ctx, src, in_len, in_len, dst, dst_stride, logn, batch);
}

// ---------------------------------------------------------------------------------------------------------
// Column NTT (RS x RS tensor code, PC_utils.cpp:28-39): a CTA owns CB adjacent columns of all 2^logn rows in
// shared memory as s[row][CB]; a quarter-warp reads one 16*CB-byte row segment -> conflict-free LDS.128 and
// full-sector global accesses.  Rows >= nz_rows are zero on input and are not read.
template <int CB>
__global__ void __launch_bounds__(512)
ntt_cols_kernel(F *__restrict__ mat, size_t cols, int logn, unsigned nz_rows, const F *__restrict__ tw) {
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
    for (int st = 1; st <= logn; st++) {
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
    HB_LAUNCH(ctx, ntt_cols_kernel<CB>, (unsigned)(cols / CB), 512, smem, mat, cols, logn, (unsigned)nz_rows, tw);
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
extern "C" int hb_ntt_batch(hb_ctx *ctx, hb_F *data, int logn, size_t batch, size_t stride) {
    using namespace hb;
    if (logn < 0 || logn > 30) HB_FAIL(ctx, "hb_ntt_batch: logn out of range");
    size_t len = (size_t)1 << logn;
    if (stride < len) HB_FAIL(ctx, "hb_ntt_batch: stride < 2^logn");
    if (batch == 0) return 0;
    Staged d(ctx);
    HB_TRY(d.outbuf(data, ((batch - 1) * stride + len) * sizeof(F), true));
    HB_TRY(ntt_rows_dev(ctx, d.as<F>(), logn, batch, stride));
    HB_TRY(d.finish());
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
