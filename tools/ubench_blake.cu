// BLAKE3 add-placement variants (development tool): compile with -DHB_BLAKE_VARIANT=n
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../hobbit_b200/csrc/blake3.cuh"
using namespace hb;
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
int main() {
    const int blocks = 148 * 8, threads = 256;
    uint32_t *o; cudaMalloc(&o, blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    blake_kernel<<<blocks, threads>>>(o, 99, 1024); cudaDeviceSynchronize();
    cudaEventRecord(e0); blake_kernel<<<blocks, threads>>>(o, 99, 1024); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    uint32_t h; cudaMemcpy(&h, o, 4, cudaMemcpyDeviceToHost);
    printf("variant %d: %.3f ms -> %.2f G compress/s (check %08x)\n", HB_BLAKE_VARIANT, ms, (double)blocks * threads * 1024 / ms / 1e6, h);
    return 0;
}
