// Micro-benchmarks for round 2 (development tool, not part of the product): candidate F_{p^2} multiplications for sm_100a, the
// sumcheck round body built from them, and the latency of a MiMC hash on ONE GPU thread (the reason Fiat–Shamir stays on the host).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench_fmul tools/ubench_fmul.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../hobbit_b200/csrc/field.cuh"
using namespace hb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// the pre-round-2 multiply (Karatsuba on 64x64 products), kept here as the baseline
__device__ __forceinline__ u64 mul61_lazy_old(u64 a, u64 b) { u64 lo = a * b, hi = __umul64hi(a, b); return ((hi << 3) | (lo >> 61)) + (lo & P61); }
__device__ __forceinline__ F fmul_old(F a, F b) {
    u64 ac = fold61(mul61_lazy_old(a.re, b.re)), bd = fold61(mul61_lazy_old(a.im, b.im));
    u64 all = fold61(mul61_lazy_old(a.re + a.im, b.re + b.im));
    u64 re = ac + (2 * P61 - bd), im = all + (4 * P61 - ac - bd);
    return mkF(canon61(fold61(re)), canon61(fold61(im)));
}

template <int V> __global__ void __launch_bounds__(256) fmul_kernel(F *out, F seed, int iters) {
    F x[4];
    for (int i = 0; i < 4; i++) x[i] = mkF((seed.re + threadIdx.x * 77 + i * 1234567) % P61, (seed.im + blockIdx.x * 31 + i) % P61);
    F m = seed;
    if (V == 2) {
        FN mn = fprep(m);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 4; i++) x[i] = fmul_n(x[i], mn);
        }
    } else if (V == 3) {                 // lazy chain: no canonicalisation between products
        FN mn = fprep(m);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 4; i++) x[i] = fmul_n_lazy(x[i], mn);
        }
        for (int i = 0; i < 4; i++) x[i] = fcanon(x[i]);
    } else if (V == 4 || V == 5) {       // both operands vary: nothing can be hoisted
        for (int it = 0; it < iters; it++) {
            F y[4];
#pragma unroll
            for (int i = 0; i < 4; i++) y[i] = (V == 4) ? fmul_old(x[i], x[(i + 1) & 3]) : fmul(x[i], x[(i + 1) & 3]);
#pragma unroll
            for (int i = 0; i < 4; i++) x[i] = y[i];
        }
    } else {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 4; i++) x[i] = (V == 0) ? fmul_old(x[i], m) : fmul(x[i], m);
        }
    }
    F s = fadd(fadd(x[0], x[1]), fadd(x[2], x[3]));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// one MiMC hash (161 cubings) per thread: latency with 1 thread, throughput with a full grid
__global__ void mimc_kernel(F *out, F in, F key, int reps) {
    F h = in;
    for (int r = 0; r < reps; r++) {
        F x = h, acc = mkF(0, 0);
        for (int i = 0; i < 161; i++) {
            F t = (i == 0) ? fadd(x, key) : fadd(fadd(acc, key), mkF((u64)(i - 1), 0));
            acc = fmul(fmul(t, t), t);
        }
        h = fadd(acc, key);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = h;
}

template <class Fn> static float timeit(Fn fn) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    fn(); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); fn(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    const int blocks = 148 * 8, threads = 256, iters = 2048;
    F *oF; CK(cudaMalloc(&oF, blocks * threads * 16));
    double nthr = (double)blocks * threads;
    F seed = mkF(1234567890123456789ULL % P61, 987654321987654321ULL % P61);
    std::vector<F> h[6];
    for (int v = 0; v < 6; v++) {
        h[v].resize(blocks * threads);
        if (v == 0) fmul_kernel<0><<<blocks, threads>>>(oF, seed, 16);
        if (v == 1) fmul_kernel<1><<<blocks, threads>>>(oF, seed, 16);
        if (v == 2) fmul_kernel<2><<<blocks, threads>>>(oF, seed, 16);
        if (v == 3) fmul_kernel<3><<<blocks, threads>>>(oF, seed, 16);
        if (v == 4) fmul_kernel<4><<<blocks, threads>>>(oF, seed, 16);
        if (v == 5) fmul_kernel<5><<<blocks, threads>>>(oF, seed, 16);
        CK(cudaMemcpy(h[v].data(), oF, blocks * threads * 16, cudaMemcpyDeviceToHost));
    }
    for (int v = 1; v < 6; v++) {
        if (v == 4) continue;
        const int ref = v == 5 ? 4 : 0;
        size_t bad = 0; for (int i = 0; i < blocks * threads; i++) bad += !(h[ref][i].re == h[v][i].re && h[ref][i].im == h[v][i].im);
        printf("fmul variant %d vs old: %zu mismatches\n", v, bad);
    }
    const char *nm[] = {"old (Karatsuba, umul64hi)", "fmul (field.cuh, canonical)", "fmul_n (prepared operand)", "fmul_n_lazy (no canon)", "old, both operands vary", "fmul, both operands vary"};
    float ms;
    ms = timeit([&] { fmul_kernel<0><<<blocks, threads>>>(oF, seed, iters); }); printf("%-32s %.3f ms -> %.1f G fmul/s\n", nm[0], ms, nthr * iters * 4 / ms / 1e6);
    ms = timeit([&] { fmul_kernel<1><<<blocks, threads>>>(oF, seed, iters); }); printf("%-32s %.3f ms -> %.1f G fmul/s\n", nm[1], ms, nthr * iters * 4 / ms / 1e6);
    ms = timeit([&] { fmul_kernel<2><<<blocks, threads>>>(oF, seed, iters); }); printf("%-32s %.3f ms -> %.1f G fmul/s\n", nm[2], ms, nthr * iters * 4 / ms / 1e6);
    ms = timeit([&] { fmul_kernel<3><<<blocks, threads>>>(oF, seed, iters); }); printf("%-32s %.3f ms -> %.1f G fmul/s\n", nm[3], ms, nthr * iters * 4 / ms / 1e6);

    ms = timeit([&] { fmul_kernel<4><<<blocks, threads>>>(oF, seed, iters); }); printf("%-32s %.3f ms -> %.1f G fmul/s\n", nm[4], ms, nthr * iters * 4 / ms / 1e6);
    ms = timeit([&] { fmul_kernel<5><<<blocks, threads>>>(oF, seed, iters); }); printf("%-32s %.3f ms -> %.1f G fmul/s\n", nm[5], ms, nthr * iters * 4 / ms / 1e6);
    ms = timeit([&] { mimc_kernel<<<1, 1>>>(oF, seed, seed, 64); });
    printf("MiMC on ONE GPU thread: %.2f us per hash (the host takes about 3 us: Fiat-Shamir stays on the host)\n", ms * 1e3 / 64);
    ms = timeit([&] { mimc_kernel<<<blocks, threads>>>(oF, seed, seed, 4); });
    printf("MiMC throughput, full grid: %.1f M hashes/s\n", nthr * 4 / ms / 1e3);
    return 0;
}
