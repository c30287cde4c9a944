// Development microbenchmark (not part of the product): issue rate of IMAD.WIDE.U32 on sm_100a with and without a 64-bit register addend.
// ptxas splits the mad.wide chains of field.cuh's dot products into independent `IMAD.WIDE R, a, b, RZ` plus three-input IADD3/IADD3.X sums;
// this measures whether the form without addend really issues faster (which would justify the extra ALU adds).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_wide tools/ubench_wide.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;

// MODE 0: mul.wide (addend RZ)   1: mad.wide with 64-bit addend   2: two mul.wide + one 3-input 64-bit add (IADD3 + IADD3.X) per two products
// 3: mad.wide chain of two (one RZ, one with addend)              4: the xor alone (baseline of the helper instruction)
template <int MODE> __global__ void __launch_bounds__(256) k(uint32_t *sink, uint32_t seed, int iters, unsigned long long *cyc) {
    u64 w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = ((u64)(seed + threadIdx.x * 8 + i) << 32) | (seed * 2654435761u + i);
    const uint32_t b = seed | 1;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t lo = (uint32_t)w[i], hi = (uint32_t)(w[i] >> 32), x;
            asm volatile("xor.b32 %0, %1, %2;" : "=r"(x) : "r"(lo), "r"(hi));
            if (MODE == 0) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(x), "r"(b));
            if (MODE == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(b));
            if (MODE == 2) {
                u64 p, q;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x), "r"(b));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(q) : "r"(hi), "r"(b));
                uint32_t rl, rh;
                asm volatile("{\n\t.reg .u32 pl, ph, ql, qh, wl, wh;\n\tmov.b64 {pl, ph}, %2;\n\tmov.b64 {ql, qh}, %3;\n\tmov.b64 {wl, wh}, %4;\n\t"
                             "add.cc.u32 %0, pl, ql;\n\taddc.u32 %1, ph, qh;\n\tadd.cc.u32 %0, %0, wl;\n\taddc.u32 %1, %1, wh;\n\t}"
                             : "=r"(rl), "=r"(rh) : "l"(p), "l"(q), "l"(w[i]));
                w[i] = ((u64)rh << 32) | rl;
            }
            if (MODE == 3) {
                u64 p;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w[i]) : "r"(hi), "r"(b), "l"(p));
            }
            if (MODE == 4) w[i] = ((u64)hi << 32) | x;
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE> static void run(const char *name, double wide_per_step, int warps_per_sm) {
    int dev = 0, sms = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 256, blocks = sms * warps_per_sm / 8, iters = 2048;
    uint32_t *sink; unsigned long long *cyc, h[4096];
    cudaMalloc(&sink, (size_t)blocks * threads * 4); cudaMalloc(&cyc, blocks * sizeof(unsigned long long));
    for (int rep = 0; rep < 2; rep++) k<MODE><<<blocks, threads>>>(sink, 12345u, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += (double)h[i]; avg /= blocks;
    // per SM: warps_per_sm warps x iters x 8 steps
    const double steps = (double)warps_per_sm * iters * 8;
    printf("{\"mode\": \"%s\", \"warps_per_sm\": %d, \"cycles_per_step_per_sm\": %.3f, \"imad_wide_per_clk_per_sm\": %.3f}\n", name, warps_per_sm, avg / steps,
           wide_per_step * steps / avg);
    cudaFree(sink); cudaFree(cyc);
}

int main() {
    for (int w : {16, 32}) {
        run<4>("xor only", 0, w);
        run<0>("mul.wide (RZ addend) + xor", 1, w);
        run<1>("mad.wide (64-bit addend) + xor", 1, w);
        run<2>("2 mul.wide + 3-input 64-bit add + xor", 2, w);
        run<3>("mul.wide then mad.wide + xor", 2, w);
    }
    return 0;
}
