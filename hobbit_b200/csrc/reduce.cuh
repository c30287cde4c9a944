// Grid-wide reduction of a few field accumulators into a host-mapped mailbox, shared by the sumcheck kernels (sumcheck.cu) and the
// opening-recursion kernels (open.cu).  Internal, not part of the C ABI.
#pragma once
#include "common.cuh"
#include <algorithm>

namespace hb {

static constexpr int kMaxRedBlocks = 1024;            // capacity of the per-CTA partial array (grids are sm_count x waves, capped here)

__device__ __forceinline__ F shfl_down_F(F v, int d) {
    F r; r.re = __shfl_down_sync(0xffffffffu, v.re, d); r.im = __shfl_down_sync(0xffffffffu, v.im, d); return r;
}

// Mailbox layout (host-mapped, 64 F): [0,16) reduced values, [16,48) table heads fetched by the provers, [48] sequence number, [49] error.
static constexpr int kMailHeads = 16, kMailSeq = 48, kMailErr = 49;
static constexpr int kMailSlotF = 32;                 // F per (parity, rank) slot of the cross-GPU mail region

// Everything a reduction kernel needs, passed by value.  `seq` is chosen by the HOST at launch (no device-side counter that could drift).
struct RedArgs {
    F *partial; unsigned *ticket; F *result; unsigned long long seq;
    // world > 1: the sums of all ranks are added before they are published (sumcheck sharded by hypercube prefix):
    // the last CTA stores its sums into slot[rank] of EVERY rank's window (NVLink peer stores), then waits for the world's slots in its own.
    int world, rank; unsigned long long dseq; u64 *peer_mail[kMaxRanks];
};

// spin until *flag == want; false after ~30 s (a peer died): the caller reports it instead of hanging the GPU
__device__ __forceinline__ bool wait_flag(const volatile u64 *flag, u64 want) {
    const long long t0 = clock64();
    while (*flag != want) { if (clock64() - t0 > 60000000000LL) return false; }
    return true;
}

// warp -> CTA -> grid reduction of NC field accumulators: warp shuffles, shared memory across warps, per-CTA partials in
// global memory, and the last CTA to take a ticket sums the partials and writes `result[0..NC)` (then re-arms the ticket).
// `result` is a host-mapped (pinned) mailbox: the coefficients are written straight into host memory, followed by a sequence
// number the host spins on — no D2H copy and no stream synchronisation on the round-to-round critical path.
template <int NC>
__device__ __forceinline__ void grid_reduce(F (&acc)[NC], const RedArgs &ra) {
    __shared__ F sred[8][NC];
    __shared__ bool is_last;
    F *__restrict__ partial = ra.partial; unsigned *__restrict__ ticket = ra.ticket; F *__restrict__ result = ra.result;
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
    F v = mkF(0, 0);
    if (threadIdx.x < NC) {
        v = sred[0][threadIdx.x];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) v = fadd(v, sred[w][threadIdx.x]);
    }
    bool ok = true;
    if (ra.world > 1) {                                   // all-reduce over the ranks through the peer-mapped mail slots
        const size_t slot = ((ra.dseq & 1) * kMaxRanks + ra.rank) * (size_t)kMailSlotF * 2;          // u64 units
        if (threadIdx.x < NC) {
            for (int h = 0; h < ra.world; h++) {
                volatile u64 *dst = ra.peer_mail[h] + slot;
                dst[2 + 2 * threadIdx.x] = v.re; dst[3 + 2 * threadIdx.x] = v.im;
            }
            __threadfence_system();
        }
        __syncthreads();
        if ((int)threadIdx.x < ra.world) {
            __threadfence_system();
            volatile u64 *dst = ra.peer_mail[threadIdx.x] + slot;
            dst[0] = ra.dseq;                                                                        // flag after the data
            const volatile u64 *mine = ra.peer_mail[ra.rank] + ((ra.dseq & 1) * kMaxRanks + threadIdx.x) * (size_t)kMailSlotF * 2;
            ok = wait_flag(mine, ra.dseq);
            __threadfence_system();
        }
        ok = __syncthreads_and(ok);
        if (threadIdx.x < NC) {
            v = mkF(0, 0);
            for (int g = 0; g < ra.world; g++) {
                const volatile u64 *src = ra.peer_mail[ra.rank] + ((ra.dseq & 1) * kMaxRanks + g) * (size_t)kMailSlotF * 2;
                v = fadd(v, mkF(src[2 + 2 * threadIdx.x], src[3 + 2 * threadIdx.x]));
            }
        }
    }
    if (threadIdx.x < NC) {
        volatile u64 *rv = reinterpret_cast<volatile u64 *>(result + threadIdx.x);
        rv[0] = v.re; rv[1] = v.im;
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile u64 *seq = reinterpret_cast<volatile u64 *>(result + kMailSeq);
        ticket[0] = 0;
        if (!ok) reinterpret_cast<volatile u64 *>(result + kMailErr)[0] = 1;
        __threadfence_system();
        seq[0] = ra.seq;
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
// Arguments of the NEXT reduction launch: the sequence number its last CTA will publish, and — inside a prover's sliced phase on a
// multi-GPU context — the peer mail slots of the cross-rank sum.
inline RedArgs red_args(hb_ctx *ctx) {
    RedArgs ra;
    ra.partial = ctx->red; ra.ticket = ctx->ticket; ra.result = ctx->mailbox_dev; ra.seq = ++ctx->seq;
    ra.world = 1; ra.rank = 0; ra.dseq = 0;
    for (int h = 0; h < kMaxRanks; h++) ra.peer_mail[h] = nullptr;
    if (ctx->dist.world > 1 && ctx->dist.reduce_on) {
        ra.world = ctx->dist.world; ra.rank = ctx->dist.rank; ra.dseq = ++ctx->dist.rseq;
        for (int h = 0; h < ctx->dist.world; h++) ra.peer_mail[h] = reinterpret_cast<u64 *>(ctx->dist.peer[h] + kDistMailOff);
    }
    return ra;
}
// Wait for the reduction kernel launched last: its final CTA writes the coefficients into the host-mapped mailbox and then publishes the
// sequence number it was launched with (grid_reduce).  Spinning on host memory avoids a D2H copy + stream synchronisation per sumcheck round.
inline int read_result(hb_ctx *ctx, int nc, F *out) {
    const u64 expect = ctx->seq;
    volatile u64 *flag = reinterpret_cast<volatile u64 *>(ctx->mailbox + kMailSeq);
    unsigned long spins = 0;
    while (*flag != expect) {
        if ((++spins & 0x3fff) == 0) {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q == cudaSuccess) { if (*flag == expect) break; HB_FAIL(ctx, "sumcheck: reduction kernel finished without publishing its result"); }
            if (q != cudaErrorNotReady) HB_CHECK(ctx, q);
        }
    }
    __sync_synchronize();
    if (reinterpret_cast<volatile u64 *>(ctx->mailbox + kMailErr)[0]) HB_FAIL(ctx, "multi-GPU reduction: a peer rank did not answer within 30 s");
    const volatile u64 *m = reinterpret_cast<const volatile u64 *>(ctx->mailbox);
    for (int c = 0; c < nc; c++) { out[c].re = m[2 * c]; out[c].im = m[2 * c + 1]; }
    transcript_absorb(ctx, out, nc);
    return 0;
}
inline unsigned red_grid_for(hb_ctx *ctx, size_t L) {
    size_t g = (L + 255) / 256;
    size_t cap = std::min<size_t>((size_t)ctx->sm_count * 4, kMaxRedBlocks);
    return (unsigned)std::max<size_t>(1, std::min(g, cap));
}

}  // namespace hb
