// Development probe (not part of the product): CUDA IPC peer windows between one-process-per-GPU ranks on one box —
// does cudaIpcOpenMemHandle work here, what does a flag round trip over NVLink cost, what bandwidth do plain peer stores reach.
// Usage: ipc_probe <rank> <world> <rendezvous-dir>     (start one process per rank)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ipc_probe tools/ipc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("rank %d: CUDA error %s at line %d\n", g_rank, cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
static int g_rank = 0;
typedef unsigned long long u64;

struct Peers { u64 *win[8]; };

// every rank: publish `seq` + 8 payload words into slot[my] of every peer, then wait until all slots of my own window carry seq
__global__ void exchange_kernel(Peers p, int rank, int world, u64 seq, u64 *out) {
    if (threadIdx.x < world) {
        volatile u64 *dst = p.win[threadIdx.x] + (seq & 1) * 8 * 16 + rank * 16;
        for (int i = 0; i < 8; i++) dst[1 + i] = seq * 1000 + rank * 10 + i;
        __threadfence_system();
        dst[0] = seq;
    }
    __syncthreads();
    if (threadIdx.x < world) {
        volatile u64 *src = p.win[rank] + (seq & 1) * 8 * 16 + threadIdx.x * 16;
        long long t0 = clock64();
        while (src[0] != seq) { if (clock64() - t0 > 4000000000LL) { out[1] = 0xdead; break; } }
        __threadfence_system();
        atomicAdd(out, src[1]);
    }
}
__global__ void bw_kernel(ulonglong2 *dst, const ulonglong2 *src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

static void wait_file(const std::string &f) { for (int i = 0; i < 60000 && access(f.c_str(), F_OK) != 0; i++) usleep(1000); }

int main(int argc, char **argv) {
    if (argc < 4) { printf("usage\n"); return 2; }
    int rank = atoi(argv[1]), world = atoi(argv[2]); std::string dir = argv[3];
    g_rank = rank;
    CK(cudaSetDevice(rank));
    const size_t bytes = (size_t)256 << 20;
    u64 *win; CK(cudaMalloc(&win, bytes)); CK(cudaMemset(win, 0, bytes));
    cudaIpcMemHandle_t h; CK(cudaIpcGetMemHandle(&h, win));
    { std::string f = dir + "/h" + std::to_string(rank) + ".tmp", g = dir + "/h" + std::to_string(rank);
      FILE *fp = fopen(f.c_str(), "wb"); fwrite(&h, sizeof(h), 1, fp); fclose(fp); rename(f.c_str(), g.c_str()); }
    Peers p; memset(&p, 0, sizeof(p));
    for (int r = 0; r < world; r++) {
        if (r == rank) { p.win[r] = win; continue; }
        std::string g = dir + "/h" + std::to_string(r); wait_file(g);
        cudaIpcMemHandle_t hr; FILE *fp = fopen(g.c_str(), "rb"); if (fread(&hr, sizeof(hr), 1, fp) != 1) return 3; fclose(fp);
        void *q; CK(cudaIpcOpenMemHandle(&q, hr, cudaIpcMemLazyEnablePeerAccess)); p.win[r] = (u64 *)q;
    }
    printf("rank %d: opened %d peer windows\n", rank, world - 1);
    // host-level barrier through files so that nobody writes before everybody mapped
    { std::string g = dir + "/m" + std::to_string(rank); FILE *fp = fopen(g.c_str(), "wb"); fclose(fp);
      for (int r = 0; r < world; r++) wait_file(dir + "/m" + std::to_string(r)); }
    u64 *out; CK(cudaMalloc(&out, 16)); CK(cudaMemset(out, 0, 16));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        for (int i = 0; i < iters; i++) exchange_kernel<<<1, 32>>>(p, rank, world, (u64)(rep * iters + i + 1), out);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        u64 ho[2]; CK(cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost));
        printf("rank %d: %d exchanges (64 B to every peer + flag, wait for all) %.3f ms -> %.2f us each, timeout flag %llx\n", rank, iters, ms, ms * 1e3 / iters, ho[1]);
    }
    if (world > 1) {
        int peer = (rank + 1) % world;
        ulonglong2 *src; CK(cudaMalloc(&src, bytes / 2)); CK(cudaMemset(src, 1, bytes / 2));
        ulonglong2 *dst = (ulonglong2 *)((char *)p.win[peer] + bytes / 2);
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0); bw_kernel<<<148 * 8, 256>>>(dst, src, bytes / 2 / 16); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("rank %d: peer store of %zu MiB to rank %d: %.3f ms -> %.1f GB/s\n", rank, bytes / 2 >> 20, peer, ms, bytes / 2 / ms / 1e6);
        }
    }
    { std::string g = dir + "/d" + std::to_string(rank); FILE *fp = fopen(g.c_str(), "wb"); fclose(fp);
      for (int r = 0; r < world; r++) wait_file(dir + "/d" + std::to_string(r)); }
    return 0;
}
