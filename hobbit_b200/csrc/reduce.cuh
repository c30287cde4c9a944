// Grid-wide reduction of a few field accumulators into a host-mapped mailbox, shared by the sumcheck kernels (sumcheck.cu) and the
// opening-recursion kernels (open.cu).  Internal, not part of the C ABI.
#pragma once
#include "common.cuh"
#include <algorithm>

namespace hb {

static constexpr int kMaxRedBlocks = 148 * 4;

__device__ __forceinline__ F shfl_down_F(F v, int d) {
    F r; r.re = __shfl_down_sync(0xffffffffu, v.re, d); r.im = __shfl_down_sync(0xffffffffu, v.im, d); return r;
}

// warp -> CTA -> grid reduction of NC field accumulators: warp shuffles, shared memory across warps, per-CTA partials in
// global memory, and the last CTA to take a ticket sums the partials and writes `result[0..NC)` (then re-arms the ticket).
// `result` is a host-mapped (pinned) mailbox: the coefficients are written straight into host memory, followed by a sequence
// number the host spins on — no D2H copy and no stream synchronisation on the round-to-round critical path.
template <int NC>
__device__ __forceinline__ void grid_reduce(F (&acc)[NC], F *__restrict__ partial, unsigned *__restrict__ ticket, F *__restrict__ result) {
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
        volatile u64 *rv = reinterpret_cast<volatile u64 *>(result + threadIdx.x);
        rv[0] = v.re; rv[1] = v.im;
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile u64 *seq = reinterpret_cast<volatile u64 *>(result + 15);       // mailbox slot 15 = sequence number
        ticket[0] = 0;
        unsigned next = ticket[1] + 1;                                            // launch counter kept in device memory
        ticket[1] = next;
        __threadfence_system();
        seq[0] = next;
    }
}

inline int ensure_scratch(hb_ctx *ctx) {
    if (!ctx->red) {
        HB_CHECK(ctx, cudaMalloc(&ctx->red, (size_t)kMaxRedBlocks * 16 * sizeof(F)));
        HB_CHECK(ctx, cudaMalloc(&ctx->ticket, 2 * sizeof(unsigned)));
        HB_CHECK(ctx, cudaMemset(ctx->ticket, 0, 2 * sizeof(unsigned)));
        HB_CHECK(ctx, cudaHostAlloc(&ctx->mailbox, 64 * sizeof(F), cudaHostAllocMapped));
        memset(ctx->mailbox, 0, 64 * sizeof(F));
        HB_CHECK(ctx, cudaHostGetDevicePointer(&ctx->mailbox_dev, ctx->mailbox, 0));
        ctx->seq = 0;
    }
    return 0;
}
// Wait for the reduction kernel launched last: its final CTA writes the coefficients into the host-mapped mailbox and then bumps the
// sequence number (grid_reduce).  Spinning on host memory avoids a D2H copy + stream synchronisation per sumcheck round.
inline int read_result(hb_ctx *ctx, int nc, F *out) {
    const u64 expect = ++ctx->seq;
    volatile u64 *flag = reinterpret_cast<volatile u64 *>(ctx->mailbox + 15);
    unsigned long spins = 0;
    while (*flag != expect) {
        if ((++spins & 0x3fff) == 0) {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q == cudaSuccess) { if (*flag == expect) break; HB_FAIL(ctx, "sumcheck: reduction kernel finished without publishing its result"); }
            if (q != cudaErrorNotReady) HB_CHECK(ctx, q);
        }
    }
    __sync_synchronize();
    const volatile u64 *m = reinterpret_cast<const volatile u64 *>(ctx->mailbox);
    for (int c = 0; c < nc; c++) { out[c].re = m[2 * c]; out[c].im = m[2 * c + 1]; }
    return 0;
}
inline unsigned red_grid_for(hb_ctx *ctx, size_t L) {
    size_t g = (L + 255) / 256;
    size_t cap = std::min<size_t>((size_t)ctx->sm_count * 4, kMaxRedBlocks);
    return (unsigned)std::max<size_t>(1, std::min(g, cap));
}

}  // namespace hb
