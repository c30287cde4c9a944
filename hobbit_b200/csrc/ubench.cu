// On-box measurement of the integer-pipe roofs the kernels of this library are bound by (bench.py reports them next to the HBM roofline):
// warp-instructions per clock per SM for IMAD.WIDE.U32 (FMA-heavy pipe: the only wide multiplier of sm_100a), for 32-bit ALU work
// (IADD3 / LOP3 / SHF) and for a 1:3 mix.  Cycles are counted with clock64 inside the kernel, so the result does not depend on the clock.
#include "common.cuh"
#include <algorithm>

namespace hb {

template <int MODE> __global__ void __launch_bounds__(256) pipe_rate_kernel(unsigned long long *cycles, uint32_t *sink, uint32_t seed, int iters) {
    uint32_t a[8]; u64 w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + threadIdx.x * 8 + i; w[i] = a[i]; }
    const uint32_t b = seed | 1;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 2) w[i] = madwide((uint32_t)w[i], b, w[i]);
            // an ALU step is separate SASS instructions that can only issue on the ALU pipe (SHF, LOP3; ptxas may turn integer adds into
            // IMAD.IADD on the FMA pipe): volatile asm keeps them apart
            if (MODE == 1) {
                asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i]));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            }
            if (MODE == 2) {
                asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i]));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                asm volatile("shf.r.wrap.b32 %0, %0, %0, 3;" : "+r"(a[i]));
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

}  // namespace hb

using namespace hb;

// out[0] IMAD.WIDE.U32, out[1] ALU pipe only (SHF, LOP3: 2 instructions per step), out[2] mix of 1 IMAD.WIDE + 3 ALU: warp-instructions per
// SECOND, chip-wide, timed with CUDA events on the context's stream (the caller divides by the SM clock it samples and the SM count)
extern "C" int hb_ubench_pipes(hb_ctx *ctx, double *out3) {
    cudaSetDevice(ctx->device);
    const int per_sm = 4, blocks = ctx->sm_count * per_sm, iters = 4096;
    unsigned long long *cyc; uint32_t *sink;
    HB_CHECK(ctx, cudaMalloc(&cyc, blocks * sizeof(unsigned long long)));
    HB_CHECK(ctx, cudaMalloc(&sink, (size_t)blocks * 256 * sizeof(uint32_t)));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double ops[3] = {1, 2, 4};
    for (int mode = 0; mode < 3; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0, ctx->stream);
            if (mode == 0) { HB_LAUNCH(ctx, pipe_rate_kernel<0>, blocks, 256, 0, cyc, sink, 12345u, iters); }
            else if (mode == 1) { HB_LAUNCH(ctx, pipe_rate_kernel<1>, blocks, 256, 0, cyc, sink, 12345u, iters); }
            else { HB_LAUNCH(ctx, pipe_rate_kernel<2>, blocks, 256, 0, cyc, sink, 12345u, iters); }
            cudaEventRecord(e1, ctx->stream);
            HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) best = std::min(best, ms);
        }
        out3[mode] = (double)blocks * 8.0 * iters * 8.0 * ops[mode] / (best * 1e-3);          // CTAs x warps x iterations x chains x instructions / s
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(cyc); cudaFree(sink);
    return 0;
}
