// Micro-benchmarks (development tool, not part of the product): instruction-pipe rates on B200 and candidate
// formulations of the F_{p^2} multiply, the 61x31-bit multiply-accumulate and the BLAKE3 compression.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../hobbit_b200/csrc/field.cuh"
#include "../hobbit_b200/csrc/blake3.cuh"
using namespace hb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ u64 madwide(uint32_t a, uint32_t b, u64 c) { u64 d; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c)); return d; }

// ---------------- pipe rate kernels: ITER iterations of 8 independent ops per thread -------------------------
template <int MODE> __global__ void pipe_kernel(uint32_t *out, uint32_t seed, int iters) {
    uint32_t a[8]; u64 w[8];
    for (int i = 0; i < 8; i++) { a[i] = seed + threadIdx.x * 8 + i; w[i] = a[i]; }
    uint32_t b = seed | 1;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) w[i] = madwide((uint32_t)w[i], b, w[i]);                 // IMAD.WIDE.U32 dependent per chain
            if (MODE == 1) a[i] = a[i] * b + a[i];                                   // IMAD (32-bit)
            if (MODE == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); // IADD
            if (MODE == 3) a[i] = __funnelshift_r(a[i], a[i], 7) ^ b;                // SHF + LOP3
            if (MODE == 4) { w[i] = madwide((uint32_t)w[i], b, w[i]); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); }  // 1 WIDE + 1 IADD
            if (MODE == 5) { w[i] = madwide((uint32_t)w[i], b, w[i]); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); a[i] ^= (a[i] >> 3); } // 1 WIDE + 3 ALU
        }
    }
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------- fmul variants ------------------------------------------------------------------------------
// B: schoolbook on 31/30-bit limbs, products accumulated inside IMAD.WIDE, one fold per output limb
struct Lm { uint32_t l0, l1; };
__device__ __forceinline__ Lm split31(u64 x) { Lm r; r.l0 = (uint32_t)x & 0x7fffffffu; r.l1 = (uint32_t)(x >> 31); return r; }
__device__ __forceinline__ u64 dot2_61(Lm a, Lm c, Lm e, Lm d) {         // (a*c + e*d) mod p, canonical
    u64 P0 = madwide(e.l0, d.l0, (u64)a.l0 * c.l0);
    u64 P1 = madwide(e.l1, d.l0, madwide(e.l0, d.l1, madwide(a.l1, c.l0, (u64)a.l0 * c.l1)));
    u64 P2 = madwide(e.l1, d.l1, (u64)a.l1 * c.l1);
    u64 t = madwide((uint32_t)P1 & 0x3fffffffu, 0x80000000u, P0) + (P1 >> 30) + 2 * P2;
    return canon61(fold61(t));
}
__device__ __forceinline__ F fmulB(F a, F b) {
    Lm ar = split31(a.re), ai = split31(a.im), br = split31(b.re), bi = split31(b.im), nai = split31(P61 - a.im);
    return mkF(dot2_61(ar, br, nai, bi), dot2_61(ar, bi, ai, br));
}

template <int V> __global__ void fmul_kernel(F *out, F seed, int iters) {
    F x[4];
    for (int i = 0; i < 4; i++) x[i] = mkF((seed.re + threadIdx.x * 77 + i * 1234567) & P61 - 1, (seed.im + blockIdx.x * 31 + i) & P61 - 1);
    F m = seed;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = (V == 0) ? fmul(x[i], m) : fmulB(x[i], m);
    }
    F s = fadd(fadd(x[0], x[1]), fadd(x[2], x[3]));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------- 61 x 31-bit multiply-accumulate variants ---------------------------------------------------
// V0: current (128-bit accumulate with compares); V1: weight split 16/16, 4 IMAD.WIDE accumulators per limb;
// V2: 32-bit columns accumulated with IMAD.WIDE(x,1,acc)
template <int V> __global__ void mac_kernel(u64 *out, const u64 *xs, const uint32_t *ws, int n, int iters) {
    u64 res = 0;
    for (int it = 0; it < iters; it++) {
        if (V == 0) {
            u64 lo = 0, hi = 0;
            for (int e = 0; e < n; e++) { u64 l, h; mul61x32_wide(xs[e] + threadIdx.x, ws[e], l, h); lo += l; hi += h + (lo < l); }
            res += red128(lo, hi);
        } else if (V == 1) {
            u64 a00 = 0, a01 = 0, a10 = 0, a11 = 0;
#pragma unroll 4
            for (int e = 0; e < n; e++) {
                u64 x = xs[e] + threadIdx.x; uint32_t w = ws[e], w0 = w & 0xffffu, w1 = w >> 16;
                uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32);
                a00 = madwide(x0, w0, a00); a01 = madwide(x0, w1, a01); a10 = madwide(x1, w0, a10); a11 = madwide(x1, w1, a11);
            }
            // value = a00 + (a01 << 16) + (a10 << 32) + (a11 << 48)
            unsigned __int128 v = (unsigned __int128)a00 + ((unsigned __int128)a01 << 16) + ((unsigned __int128)a10 << 32) + ((unsigned __int128)a11 << 48);
            res += red128((u64)v, (u64)(v >> 64));
        } else {
            u64 c0 = 0, c1 = 0, c2 = 0;
#pragma unroll 4
            for (int e = 0; e < n; e++) {
                u64 x = xs[e] + threadIdx.x; uint32_t w = ws[e];
                u64 p0 = (u64)(uint32_t)x * w, p1 = (u64)(uint32_t)(x >> 32) * w;
                c0 = madwide((uint32_t)p0, 1u, c0); c1 = madwide((uint32_t)(p0 >> 32), 1u, c1);
                c1 = madwide((uint32_t)p1, 1u, c1); c2 = madwide((uint32_t)(p1 >> 32), 1u, c2);
            }
            unsigned __int128 v = (unsigned __int128)c0 + ((unsigned __int128)c1 << 32) + ((unsigned __int128)c2 << 64);
            res += red128((u64)v, (u64)(v >> 64));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = res;
}

// ---------------- BLAKE3 --------------------------------------------------------------------------------------
__global__ void blake_kernel(uint32_t *out, uint32_t seed, int iters) {
    uint32_t m[16], o[8];
    for (int i = 0; i < 16; i++) m[i] = seed + i * 977 + threadIdx.x;
    for (int it = 0; it < iters; it++) {
        blake3_compress64(m, o);
#pragma unroll
        for (int i = 0; i < 8; i++) { m[i] = o[i]; m[8 + i] ^= o[i]; }
    }
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s ^= o[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class Fn> static float timeit(Fn fn) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    fn(); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); fn(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    uint32_t *o32; u64 *o64; F *oF;
    CK(cudaMalloc(&o32, blocks * threads * 4)); CK(cudaMalloc(&o64, blocks * threads * 8)); CK(cudaMalloc(&oF, blocks * threads * 16));
    double nthr = (double)blocks * threads;
    const char *names[] = {"IMAD.WIDE.U32", "IMAD", "IADD", "SHF+LOP3 (2 ops)", "WIDE+IADD (2 ops)", "WIDE+3ALU (4 ops)"};
    int opsper[] = {1, 1, 1, 2, 2, 4};
    float ms;
#define RUNP(M) ms = timeit([&] { pipe_kernel<M><<<blocks, threads>>>(o32, 12345, iters); }); \
    printf("pipe %-20s: %.3f ms  -> %.1f Gop/s (thread-ops), %.2f warp-inst/clk/SM @1.9GHz\n", names[M], ms, nthr * iters * 8 * opsper[M] / ms / 1e6, nthr * iters * 8 * opsper[M] / 32 / (ms * 1e-3) / 148 / 1.9e9);
    RUNP(0) RUNP(1) RUNP(2) RUNP(3) RUNP(4) RUNP(5)

    // fmul correctness (A vs B) then rate
    {
        F seed = mkF(1234567890123456789ULL % P61, 987654321987654321ULL % P61);
        F *ha = new F[blocks * threads], *hb2 = new F[blocks * threads];
        fmul_kernel<0><<<blocks, threads>>>(oF, seed, 16); CK(cudaMemcpy(ha, oF, blocks * threads * 16, cudaMemcpyDeviceToHost));
        fmul_kernel<1><<<blocks, threads>>>(oF, seed, 16); CK(cudaMemcpy(hb2, oF, blocks * threads * 16, cudaMemcpyDeviceToHost));
        size_t bad = 0; for (int i = 0; i < blocks * threads; i++) bad += !(ha[i].re == hb2[i].re && ha[i].im == hb2[i].im);
        printf("fmul A vs B mismatches: %zu\n", bad);
        ms = timeit([&] { fmul_kernel<0><<<blocks, threads>>>(oF, seed, iters); });
        printf("fmul A (Karatsuba, umul64hi): %.3f ms -> %.1f G fmul/s\n", ms, nthr * iters * 4 / ms / 1e6);
        ms = timeit([&] { fmul_kernel<1><<<blocks, threads>>>(oF, seed, iters); });
        printf("fmul B (31/30 limbs, mad.wide): %.3f ms -> %.1f G fmul/s\n", ms, nthr * iters * 4 / ms / 1e6);
    }
    {
        const int n = 48;
        std::vector<u64> xs(n); std::vector<uint32_t> ws(n);
        for (int i = 0; i < n; i++) { xs[i] = (0x123456789abcdefULL * (i + 1)) % (P61 - 1000); ws[i] = (uint32_t)(2654435761u * (i + 1)) >> 1; }
        u64 *dx; uint32_t *dw; CK(cudaMalloc(&dx, n * 8)); CK(cudaMalloc(&dw, n * 4));
        CK(cudaMemcpy(dx, xs.data(), n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dw, ws.data(), n * 4, cudaMemcpyHostToDevice));
        u64 h[3][256];
        mac_kernel<0><<<1, 256>>>(o64, dx, dw, n, 1); CK(cudaMemcpy(h[0], o64, 256 * 8, cudaMemcpyDeviceToHost));
        mac_kernel<1><<<1, 256>>>(o64, dx, dw, n, 1); CK(cudaMemcpy(h[1], o64, 256 * 8, cudaMemcpyDeviceToHost));
        mac_kernel<2><<<1, 256>>>(o64, dx, dw, n, 1); CK(cudaMemcpy(h[2], o64, 256 * 8, cudaMemcpyDeviceToHost));
        size_t bad = 0; for (int i = 0; i < 256; i++) bad += (h[0][i] != h[1][i]) + (h[0][i] != h[2][i]);
        printf("mac variants mismatches: %zu\n", bad);
        const int it2 = 256;
        ms = timeit([&] { mac_kernel<0><<<blocks, threads>>>(o64, dx, dw, n, it2); }); printf("mac V0 (128-bit acc, compares): %.3f ms -> %.1f G mac/s\n", ms, nthr * it2 * n / ms / 1e6);
        ms = timeit([&] { mac_kernel<1><<<blocks, threads>>>(o64, dx, dw, n, it2); }); printf("mac V1 (w split 16/16, 4 WIDE): %.3f ms -> %.1f G mac/s\n", ms, nthr * it2 * n / ms / 1e6);
        ms = timeit([&] { mac_kernel<2><<<blocks, threads>>>(o64, dx, dw, n, it2); }); printf("mac V2 (32-bit columns, 6 WIDE): %.3f ms -> %.1f G mac/s\n", ms, nthr * it2 * n / ms / 1e6);
    }
    ms = timeit([&] { blake_kernel<<<blocks, threads>>>(o32, 99, 1024); });
    printf("blake3 compress64: %.3f ms -> %.2f G compress/s\n", ms, nthr * 1024 / ms / 1e6);
    return 0;
}
