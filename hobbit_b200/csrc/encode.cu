// E1/E2 — Orion/Spielman expander linear code, batched over all columns of the tensor.
// Reference: encode_monolithic (src/linear_code_encode.h:62-119) over graphs built by expander_init_store
// (src/expanders.h:20-47,78-92).  The recursion
//     enc_d(x) = x | enc_{d+1}(C_d x) | D_d * enc_{d+1}(C_d x)
// is flattened into a list of sparse mat-vec STAGES over one codeword buffer (C_0, C_1, .., C_last, D_last, .., D_0);
// each stage reads an already-final segment of the codeword and writes a disjoint later segment.
//
// B200 mapping: all columns share one graph, so a CTA takes CB adjacent columns and keeps their whole codewords
// in shared memory as cw[row][CB] (n=1024: 1761 rows x 4 columns x 16 B = 110 KB, two CTAs of 512 threads per SM; 8 columns /
// 220 KB / one CTA where two do not fit).  A thread owns (row, column); each stage is a GATHER over the row's in-edges (CSR by
// target, built on the host from the reference's scatter lists, rows in order of in-degree) so there are no atomics, and the
// 61x32-bit products are accumulated in 128 bits and reduced once per row.  The lanes of a row read one contiguous 16*CB-byte
// segment (LDS.128); with 4 columns two rows share a wavefront and the host orders the edges so that their sources differ in parity.
// The weights are 31-bit reals (expanders.h:37 `F weight = random()`), so F x weight is two 61x32 products.
// The inner half of the commit_standard leaf hashing (Our_PC.cpp:160-166; H1 of each 4-row quad of a column) is fused
// here: the CTA already holds every row of its columns, so the quads are hashed straight out of shared memory and the
// tensor is never re-read.  The chunk-order-dependent half (H1(inner | previous leaf)) runs afterwards in
// md_chain_kernel, which lets every chunk of a commit be encoded by ONE launch (grid.y = chunk) with no grid tail.
#include "common.cuh"
#include "blake3.cuh"
#include <algorithm>
#include <cuda.h>

namespace hb {

// Sum of 61-bit x 32-bit products without carry chains through compares: the two 64-bit partial products
// p0 = x.lo32 * w and p1 = x.hi32 * w (IMAD.WIDE, FMA pipe) are accumulated in two independent 96-bit accumulators
// with add.cc/addc (3 IADD3 each, ALU pipe); value = U + V * 2^32.  Headroom: 2^29 terms (acc_reduce).  This is the one-edge form
// (tail of a row, graphs with weights >= 2^31); acc_mac4 below is the hot one.
struct Acc { uint32_t u0, u1, u2, v0, v1, v2; };
__device__ __forceinline__ void acc_mac(Acc &a, u64 x, uint32_t w) {
    u64 p0 = (u64)(uint32_t)x * w, p1 = (u64)(uint32_t)(x >> 32) * w;
    asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;"
        : "+r"(a.u0), "+r"(a.u1), "+r"(a.u2) : "r"((uint32_t)p0), "r"((uint32_t)(p0 >> 32)));
    asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;"
        : "+r"(a.v0), "+r"(a.v1), "+r"(a.v2) : "r"((uint32_t)p1), "r"((uint32_t)(p1 >> 32)));
}
// Four edges at once (HB_ENC_MAC4): with weights < 2^31 and canonical entries (< 2^61) a low-half product is < 2^63 and a high-half product
// < 2^60, so two low products and all four high products are summed on the 64-bit addend of IMAD.WIDE before anything touches the 96-bit
// accumulators: 9 additions per limb and four edges instead of 24.
__device__ __forceinline__ void add96(uint32_t &a0, uint32_t &a1, uint32_t &a2, u64 v) {
    asm("{\n\t.reg .u32 vl, vh;\n\tmov.b64 {vl, vh}, %3;\n\tadd.cc.u32 %0, %0, vl;\n\taddc.cc.u32 %1, %1, vh;\n\taddc.u32 %2, %2, 0;\n\t}"
        : "+r"(a0), "+r"(a1), "+r"(a2) : "l"(v));
}
__device__ __forceinline__ void acc_mac4(Acc &a, u64 x0, u64 x1, u64 x2, u64 x3, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    const u64 ua = madwide((uint32_t)x1, w1, mulwide((uint32_t)x0, w0));
    const u64 ub = madwide((uint32_t)x3, w3, mulwide((uint32_t)x2, w2));
    const u64 v = madwide((uint32_t)(x3 >> 32), w3, madwide((uint32_t)(x2 >> 32), w2, madwide((uint32_t)(x1 >> 32), w1, mulwide((uint32_t)(x0 >> 32), w0))));
    add96(a.u0, a.u1, a.u2, ua); add96(a.u0, a.u1, a.u2, ub); add96(a.v0, a.v1, a.v2, v);
}
#ifndef HB_ENC_MAC4
#define HB_ENC_MAC4 1
#endif
// value = U + V 2^32 is put together as ONE 128-bit number (three additions with carry) and reduced once: its upper half is below
// 2^29 x (number of terms), inside red128's 2^58 for any in-degree below 2^29 (hb_expander_set checks it).  Two reductions, a rotation and a
// modular addition before: ~100 instructions per row and both limbs, a third of the work of a 12-edge row of the D stages; now ~40.
__device__ __forceinline__ u64 acc_reduce(const Acc &a) {
    uint32_t s1, s2, s3;
    asm("add.cc.u32 %0, %3, %4;\n\taddc.cc.u32 %1, %5, %6;\n\taddc.u32 %2, %7, 0;"
        : "=r"(s1), "=r"(s2), "=r"(s3) : "r"(a.u1), "r"(a.v0), "r"(a.u2), "r"(a.v1), "r"(a.v2));
    return red128(((u64)s1 << 32) | a.u0, ((u64)s3 << 32) | s2);
}

// Edge and row-pointer loads: read-only path, L1 evict-last.  One SM runs dozens of CTAs per launch and every CTA walks the same graph;
// measured (bench.py, ms per launch of 16 chunks): evict-last on every edge 3.36, default policy 3.41, evict-first on the two big
// stages (so that the small stages' lists stay resident) 3.48-3.66 — the big stages lose their same-line hits.
__device__ __forceinline__ uint2 ld_edge(const uint2 *p) {
    uint2 v;
    asm("ld.global.nc.L1::evict_last.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_edge2(const uint2 *p) {            // two consecutive edge records, p 16-byte aligned
    uint4 v;
    asm("ld.global.nc.L1::evict_last.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_rowptr(const uint2 *p) { int v; asm("ld.global.nc.L1::evict_last.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }   // .x only

// grid: (cols / CB, nchunks).  inner != nullptr: also emit the inner leaf digests H1(T[4j][k] | .. | T[4j+3][k])
// of this chunk (commit_standard hashes 4-row quads of every column, Our_PC.cpp:160-166); the Merkle–Damgård chaining
// over chunks is done afterwards by md_chain_kernel so that all chunks can be encoded in one launch.
#ifndef HB_ENC_HELP_UNITS
#define HB_ENC_HELP_UNITS 2
#endif
constexpr int kHelpUnits = HB_ENC_HELP_UNITS;
// TMA (cp.async.bulk.tensor): the CTA's tile — CB adjacent columns x all message rows, a dense [rows][CB] block of 16-byte elements — is
// fetched by ONE thread as 2-D boxes of up to 256 rows that land in shared memory in exactly the cw[row][CB] layout the stages use
// (no swizzle: a quarter-warp reads one contiguous 16*CB-byte row), completion on an mbarrier; the parity rows go back the same way.
// The tensor map views all chunks of the launch as one (2*cols u64) x (2n*nchunks rows) matrix.
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, int x, int y, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(tm), "r"(x), "r"(y), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, int x, int y, const void *smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tm), "r"(x), "r"(y), "r"((unsigned)__cvta_generic_to_shared(smem_src)) : "memory");
}
template <int CB, bool INNER>
__global__ void __launch_bounds__(1024)
encode_cols_kernel(F *__restrict__ Tbase, size_t chunk_stride, size_t cols, int n, int cwlen,
                   const EncStage *__restrict__ stages, int nstages,
                   const uint2 *__restrict__ rowptr, const uint2 *__restrict__ edges,
                   uint8_t *__restrict__ inner_base, Digest zero_quad, InnerLayout lay, unsigned long long *__restrict__ prof, int split_ok, int help_ok,
                   const __grid_constant__ CUtensorMap tmap, int box_rows, int mac4_ok) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    F *cw = reinterpret_cast<F *>(smem_raw);                   // cw[row * CB + c]
    __shared__ __align__(8) unsigned long long tma_bar;
    long long tprev = prof ? clock64() : 0;                    // development aid (HB_ENCODE_PROF): cycles per phase, summed over CTAs by thread 0
    auto mark = [&](int slot) { if (prof && threadIdx.x == 0) { long long t = clock64(); atomicAdd(&prof[slot], (unsigned long long)(t - tprev)); tprev = t; } };
    F *T = Tbase + (size_t)blockIdx.y * chunk_stride;
    const size_t col0 = (size_t)blockIdx.x * CB;
    const unsigned c = threadIdx.x % CB, t0 = threadIdx.x / CB, tstep = blockDim.x / CB;
    // inner leaf digests: pair index idx = j * CB + c <-> quad j (rows 4j .. 4j+3) of column c
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const unsigned msg_pairs = (unsigned)(n / 4) * CB, msg_units = (msg_pairs + 31) / 32;    // quads made of message rows only, in units of one warp
    __shared__ unsigned hash_next;                                                            // next message unit nobody has taken yet
    auto hash_pair = [&](unsigned idx) {
        const unsigned j = idx / CB, cc = idx % CB;
        if (j >= (unsigned)n / 2) return;
        uint32_t out[8];
        if (4 * j >= (unsigned)cwlen) {
#pragma unroll
            for (int q = 0; q < 8; q++) out[q] = zero_quad.w[q];           // H1(64 zero bytes), precomputed
        } else {
            uint32_t m[16];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                unsigned r = 4 * j + q;
                F x = (r < (unsigned)cwlen) ? cw[r * CB + cc] : mkF(0, 0);
                m[4 * q] = (uint32_t)x.re; m[4 * q + 1] = (uint32_t)(x.re >> 32);
                m[4 * q + 2] = (uint32_t)x.im; m[4 * q + 3] = (uint32_t)(x.im >> 32);
            }
            blake3_compress64(m, out);
        }
        uint4 *lp = reinterpret_cast<uint4 *>(lay.addr(inner_base, blockIdx.y, (size_t)j * cols + col0 + cc));
        lp[0] = make_uint4(out[0], out[1], out[2], out[3]);
        lp[1] = make_uint4(out[4], out[5], out[6], out[7]);
    };
    if (INNER && threadIdx.x == 0) hash_next = 0;
    // edges e, e + step, .. below e1 in groups of four.  (Measured and not kept, profiles/r02_summary.md: fetching the NEXT group's edge
    // records early — into a second register set, or into L1 with prefetch.global.L1 — to hide their L2 round trip: ptxas sinks the loads
    // to the end of the group under the 64-register cap, 3.58 vs 3.32 ms; the L1 prefetch changes nothing.  Two CTAs per SM hide it instead.)
#ifndef HB_ENC_MAC4_UNROLL
#define HB_ENC_MAC4_UNROLL 1
#endif
    constexpr int kMac4Unroll = HB_ENC_MAC4_UNROLL;
    auto mac4_run = [&](Acc &are, Acc &aim, int &e, int e1, int step) {
        if (step == 1) {
            // whole rows: lists start at even indices (hb_expander_set), two records per 16-byte load
#pragma unroll kMac4Unroll
            for (; e + 3 < e1; e += 4) {
                const uint4 q0 = ld_edge2(&edges[e]), q1 = ld_edge2(&edges[e + 2]);
                const F x0 = cw[q0.x * CB + c], x1 = cw[q0.z * CB + c], x2 = cw[q1.x * CB + c], x3 = cw[q1.z * CB + c];
                acc_mac4(are, x0.re, x1.re, x2.re, x3.re, q0.y, q0.w, q1.y, q1.w);
                acc_mac4(aim, x0.im, x1.im, x2.im, x3.im, q0.y, q0.w, q1.y, q1.w);
            }
            return;
        }
#pragma unroll kMac4Unroll
        for (; e + 3 * step < e1; e += 4 * step) {
            const uint2 a0 = ld_edge(&edges[e]), a1 = ld_edge(&edges[e + step]), a2 = ld_edge(&edges[e + 2 * step]), a3 = ld_edge(&edges[e + 3 * step]);
            const F x0 = cw[a0.x * CB + c], x1 = cw[a1.x * CB + c], x2 = cw[a2.x * CB + c], x3 = cw[a3.x * CB + c];
            acc_mac4(are, x0.re, x1.re, x2.re, x3.re, a0.y, a1.y, a2.y, a3.y);
            acc_mac4(aim, x0.im, x1.im, x2.im, x3.im, a0.y, a1.y, a2.y, a3.y);
        }
    };

    if (box_rows) {
        // message rows -> shared memory by TMA: one thread, n / box_rows boxes in flight, one mbarrier
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&tma_bar)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&tma_bar)), "r"((unsigned)(n * CB * sizeof(F))) : "memory");
            for (int r = 0; r < n; r += box_rows) tma_load_2d(&cw[r * CB], &tmap, (int)(2 * col0), (int)(blockIdx.y * 2 * n + r), &tma_bar);
        }
        __syncthreads();                                         // the barrier is initialised before anybody polls it
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"((unsigned)__cvta_generic_to_shared(&tma_bar)) : "memory");
    } else {
        // shapes the tensor map cannot describe (rows not a multiple of the box, strided chunks): asynchronous 16-byte copies, every thread
        // has all of its loads in flight at once
        for (unsigned r = t0; r < (unsigned)n; r += tstep) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&cw[r * CB + c]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(&T[(size_t)r * cols + col0 + c]) : "memory");
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
    mark(0);

    for (int s = 0; s < nstages; s++) {
        const EncStage st = stages[s];
        const uint2 *rp = rowptr + st.rowptr_base;        // rows in processing order (sorted by in-degree on the host): {first edge, target row}
        // Small stages have fewer rows than the CTA has row slots (blockDim / CB): their rows are split over P = 2 or 4 adjacent slots
        // (lanes 8 or 8 and 16 apart in the same warp when CB == 8), every slot sums every P-th edge and the partial sums are added
        // with shuffles — exact in the field, so the result is the same canonical value.
        // (CB == 4: the slots of a row are lanes 4, 8 and 16 apart and a row may be split eight ways.)
        constexpr unsigned SPW = 32 / CB;                                     // row slots per warp
        const unsigned P = (!split_ok || (CB != 8 && CB != 4)) ? 1u : (CB == 4 && (unsigned)st.R * 8 <= tstep) ? 8u : ((unsigned)st.R * 4 <= tstep) ? 4u
                         : ((unsigned)st.R * 2 <= tstep) ? 2u : 1u;
        if (P == 1) {
            for (unsigned t = t0; t < (unsigned)st.R; t += tstep) {
                const uint2 row = ld_edge(&rp[t]);
                const int e0 = (int)row.x, e1 = ld_rowptr(&rp[t + 1]);
                Acc are = {0, 0, 0, 0, 0, 0}, aim = {0, 0, 0, 0, 0, 0};
                int e = e0;
                if (HB_ENC_MAC4 && mac4_ok) mac4_run(are, aim, e, e1, 1);
#pragma unroll 4
                for (; e < e1; e++) {
                    uint2 ed = ld_edge(&edges[e]);
                    F x = cw[ed.x * CB + c];
                    acc_mac(are, x.re, ed.y);
                    acc_mac(aim, x.im, ed.y);
                }
                cw[(st.out_off + row.y) * CB + c] = mkF(acc_reduce(are), acc_reduce(aim));
            }
        } else {
            const unsigned slots = (unsigned)st.R * P;                       // <= tstep: one pass
            const unsigned slot = t0, t = slot / P, part = slot % P;
            if ((t0 & ~(SPW - 1)) < slots) {                                  // warp-uniform: the slots of a warp are SPW*a .. SPW*a + SPW - 1
                Acc are = {0, 0, 0, 0, 0, 0}, aim = {0, 0, 0, 0, 0, 0};
                unsigned target = 0;
                if (slot < slots) {
                    const uint2 row = ld_edge(&rp[t]);
                    const int e0 = (int)row.x, e1 = ld_rowptr(&rp[t + 1]);
                    target = row.y;
                    int e = e0 + (int)part;
                    if (HB_ENC_MAC4 && mac4_ok) mac4_run(are, aim, e, e1, (int)P);
#pragma unroll 2
                    for (; e < e1; e += (int)P) {
                        uint2 ed = ld_edge(&edges[e]);
                        F x = cw[ed.x * CB + c];
                        acc_mac(are, x.re, ed.y);
                        acc_mac(aim, x.im, ed.y);
                    }
                }
                u64 re = acc_reduce(are), im = acc_reduce(aim);
                re = add61(re, __shfl_xor_sync(0xffffffffu, re, CB)); im = add61(im, __shfl_xor_sync(0xffffffffu, im, CB));
                if (P >= 4) { re = add61(re, __shfl_xor_sync(0xffffffffu, re, 2 * CB)); im = add61(im, __shfl_xor_sync(0xffffffffu, im, 2 * CB)); }
                if (P >= 8) { re = add61(re, __shfl_xor_sync(0xffffffffu, re, 4 * CB)); im = add61(im, __shfl_xor_sync(0xffffffffu, im, 4 * CB)); }
                if (slot < slots && part == 0) cw[(st.out_off + target) * CB + c] = mkF(re, im);
            }
        }
        // Warps without a row slot in this stage (the small stages of the recursion keep only a few warps busy, and those are bound by
        // load latency, not by issue slots) hash quads of MESSAGE rows meanwhile — the stages only write parity rows.  At most
        // kHelpUnits units per warp and stage so that the stage barrier is not held up; what is left is done after the last stage.
        // (Also letting the warps that sit out the LAST pass of a big stage help was measured slower: 3.45 vs 3.39 ms per launch.)
        if (INNER && help_ok) {
            const unsigned nslots = (unsigned)st.R * P;
            if (nslots <= tstep && (threadIdx.x & ~31u) / CB >= nslots) {
                for (int it = 0; it < kHelpUnits; it++) {
                    unsigned u = 0;
                    if (lane == 0) u = atomicAdd(&hash_next, 1u);
                    u = __shfl_sync(0xffffffffu, u, 0);
                    if (u >= msg_units) break;
                    if (32 * u + lane < msg_pairs) hash_pair(32 * u + lane);
                }
            }
        }
        __syncthreads();
        mark(1 + s);
    }

    // rows [n, cwlen) are new; rows [cwlen, 2n) are the zero tail of the reference's 2n-sized buffer
    unsigned first_plain = n;
    if (box_rows) {
        // whole boxes of parity rows leave by TMA straight from shared memory (the last stage's barrier made them visible to this thread;
        // the proxy fence makes them visible to the async proxy); the ragged rest and the zero tail are plain stores
        const unsigned full = ((unsigned)(cwlen - n) / box_rows) * box_rows;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            for (unsigned r = n; r < n + full; r += box_rows) tma_store_2d(&tmap, (int)(2 * col0), (int)(blockIdx.y * 2 * n + r), &cw[r * CB]);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        first_plain = n + full;
    }
    for (unsigned r = first_plain + t0; r < 2u * n; r += tstep)
        T[(size_t)r * cols + col0 + c] = (r < (unsigned)cwlen) ? cw[r * CB + c] : mkF(0, 0);
    if (prof) __syncthreads();
    mark(14);

    if (INNER) {
        const unsigned done = min((unsigned)hash_next, msg_units);                   // message units already hashed by idle warps
        for (unsigned u = done + warp; u < msg_units; u += nwarps) if (32 * u + lane < msg_pairs) hash_pair(32 * u + lane);
        for (unsigned idx = msg_pairs + threadIdx.x; idx < (unsigned)(n / 2) * CB; idx += blockDim.x) hash_pair(idx);
        if (prof) __syncthreads();
        mark(15);
    }
    // shared memory must stay alive until the bulk stores have read it
    if (box_rows && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr; static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
        cudaGetLastError();
    }
    return fn;
}

// H1 of 64 zero bytes (blake3 KAT 4d006976...): the inner digest of every all-zero quad
static Digest zero_quad_digest() {
    static const uint8_t kat[32] = {0x4d, 0x00, 0x69, 0x76, 0x63, 0x6a, 0x86, 0x96, 0xd9, 0x09, 0xa6, 0x30, 0xa4, 0x08, 0x1a, 0xad,
                                    0x4d, 0x7c, 0x50, 0xf8, 0x1a, 0xfd, 0xee, 0x04, 0x02, 0x0b, 0xf0, 0x50, 0x86, 0xab, 0x6a, 0x55};
    Digest d; memcpy(d.w, kat, 32); return d;
}

template <int CB>
static int launch_encode(hb_ctx *ctx, F *T, long long n, size_t cols, size_t nchunks, size_t chunk_stride, uint8_t *inner, InnerLayout lay) {
    const ExpanderDev &ex = ctx->exp;
    size_t smem = (size_t)ex.cwlen * CB * sizeof(F);
    dim3 grid((unsigned)(cols / CB), (unsigned)nchunks);
    // small codes: several CTAs per SM, 256 threads each; the big code gets 1024 threads when one CTA fills the SM (8 columns, 220 KB) and
    // 512 when two CTAs share it (4 columns, 2 x 110 KB: one CTA's tile load, small stages and barriers overlap the other's big stages)
    const bool two_per_sm = smem > 100 * 1024 && 2 * (smem + 1024) <= 228 * 1024;
    unsigned threads = two_per_sm ? 512 : smem > 100 * 1024 ? 1024 : 256;
    static unsigned long long *prof = nullptr;
    static const int help_ok = getenv("HB_ENCODE_HELP") ? atoi(getenv("HB_ENCODE_HELP")) : 1;      // experiment switch
    static const int split_ok = getenv("HB_ENCODE_SPLIT") ? atoi(getenv("HB_ENCODE_SPLIT")) : 1;   // experiment switch
    if (!prof && getenv("HB_ENCODE_PROF")) { cudaMalloc(&prof, 16 * 8); cudaMemset(prof, 0, 16 * 8); }
    if (const char *e = getenv("HB_ENCODE_THREADS")) threads = (unsigned)atoi(e);          // experiment switch
    // TMA tile movement when the launch is one dense (2*cols u64) x (2n * nchunks) matrix and the message rows split into whole boxes
    CUtensorMap tmap; memset(&tmap, 0, sizeof(tmap));
    int box_rows = 0;
    static const int mac4_env = getenv("HB_ENCODE_MAC4") ? atoi(getenv("HB_ENCODE_MAC4")) : 1;   // experiment switch
    const int mac4_ok = mac4_env && ex.w31;
    static const int tma_ok = getenv("HB_ENCODE_TMA") ? atoi(getenv("HB_ENCODE_TMA")) : 1;   // experiment switch
    if (tma_ok && encode_tiled_fn() && (nchunks == 1 || chunk_stride == (size_t)(2 * n) * cols) && (n % 256 == 0 || n <= 256) && ((uintptr_t)T % 16) == 0 &&
        (size_t)(2 * n) * nchunks < ((size_t)1 << 31)) {
        const int br = n <= 256 ? (int)n : 256;
        cuuint64_t dims[2] = {(cuuint64_t)(2 * cols), (cuuint64_t)((size_t)(2 * n) * nchunks)};
        cuuint64_t strides[1] = {(cuuint64_t)(cols * sizeof(F))};
        cuuint32_t box[2] = {(cuuint32_t)(2 * CB), (cuuint32_t)br}, estr[2] = {1, 1};
        if (encode_tiled_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void *)T, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) box_rows = br;
    }
    if (inner) {
        HB_CHECK(ctx, cudaFuncSetAttribute(encode_cols_kernel<CB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (two_per_sm) HB_CHECK(ctx, cudaFuncSetAttribute(encode_cols_kernel<CB, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        HB_LAUNCH(ctx, (encode_cols_kernel<CB, true>), grid, threads, smem, T, chunk_stride, cols, (int)n, ex.cwlen, ex.d_stages, (int)ex.stages.size(),
                  ex.d_rowptr, ex.d_edges, inner, zero_quad_digest(), lay, prof, split_ok, help_ok, tmap, box_rows, mac4_ok);
    } else {
        HB_CHECK(ctx, cudaFuncSetAttribute(encode_cols_kernel<CB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (two_per_sm) HB_CHECK(ctx, cudaFuncSetAttribute(encode_cols_kernel<CB, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        HB_LAUNCH(ctx, (encode_cols_kernel<CB, false>), grid, threads, smem, T, chunk_stride, cols, (int)n, ex.cwlen, ex.d_stages, (int)ex.stages.size(),
                  ex.d_rowptr, ex.d_edges, inner, zero_quad_digest(), lay, prof, split_ok, help_ok, tmap, box_rows, mac4_ok);
    }
    if (prof) {                                                   // development aid: cumulative cycles per phase (thread 0 of every CTA)
        unsigned long long h[16];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h, prof, sizeof h, cudaMemcpyDeviceToHost);
        fprintf(stderr, "encode phases (load | stages.. | store | hash):");
        for (int i = 0; i < 16; i++) if (h[i]) fprintf(stderr, " %d:%llu", i, h[i]);
        fprintf(stderr, "\n");
    }
    return 0;
}

int encode_cols_dev(hb_ctx *ctx, F *T, long long n, size_t cols, size_t nchunks, size_t chunk_stride, uint8_t *inner, InnerLayout lay) {
    if (lay.part_leaves == 0) lay = InnerLayout::plain((size_t)(n / 2) * cols, nchunks);
    const ExpanderDev &ex = ctx->exp;
    if (ex.n != n) HB_FAIL(ctx, "encode: no expander installed for this message length (call hb_expander_set / expander_init_store first)");
    if (nchunks > 65535) HB_FAIL(ctx, "encode: too many chunks in one launch");
    const size_t kMaxSmem = 227 * 1024;
    size_t per_col = (size_t)ex.cwlen * sizeof(F);
    // widest column block that fits; prefer <= ~100 KB tiles when the code is small so several CTAs share an SM
    if (cols % 32 == 0 && per_col * 32 <= 100 * 1024) return launch_encode<32>(ctx, T, n, cols, nchunks, chunk_stride, inner, lay);
    if (cols % 16 == 0 && per_col * 16 <= 100 * 1024) return launch_encode<16>(ctx, T, n, cols, nchunks, chunk_stride, inner, lay);
    // Measured on B200 at n = 1024 (bench.py, ms per launch of 16 chunks).  Round 1: 8 columns per CTA (one 220 KB CTA per SM) 3.58; 4 columns
    // (two CTAs per SM) 3.62; 2 columns (four CTAs) 4.43; fetching each edge once per 8-lane group and passing it round with shuffles 5.01.
    // Round 2, after the four-edge accumulation cut the instruction count by 16 %: 8 columns x 1024 threads 3.32, 4 columns x 512 threads
    // x two CTAs per SM 3.17 — with fewer instructions the stage barriers, the tile load and the edge-load latency of one CTA are worth
    // overlapping with the other CTA's arithmetic.  So: two CTAs of 4 columns when both tiles fit, else one CTA of 8.
    if (const char *e = getenv("HB_ENCODE_CB")) {                                            // experiment switch
        if (atoi(e) == 8 && cols % 8 == 0 && per_col * 8 <= kMaxSmem) return launch_encode<8>(ctx, T, n, cols, nchunks, chunk_stride, inner, lay);
        if (atoi(e) == 4 && cols % 4 == 0) return launch_encode<4>(ctx, T, n, cols, nchunks, chunk_stride, inner, lay);
        if (atoi(e) == 2 && cols % 2 == 0) return launch_encode<2>(ctx, T, n, cols, nchunks, chunk_stride, inner, lay);
    }
    if (cols % 4 == 0 && per_col * 4 > 100 * 1024 && 2 * (per_col * 4 + 1024) <= 228 * 1024) return launch_encode<4>(ctx, T, n, cols, nchunks, chunk_stride, inner, lay);
    if (cols % 8 == 0 && per_col * 8 <= kMaxSmem) return launch_encode<8>(ctx, T, n, cols, nchunks, chunk_stride, inner, lay);
    if (cols % 4 == 0 && per_col * 4 <= kMaxSmem) return launch_encode<4>(ctx, T, n, cols, nchunks, chunk_stride, inner, lay);
    if (cols % 2 == 0 && per_col * 2 <= kMaxSmem) return launch_encode<2>(ctx, T, n, cols, nchunks, chunk_stride, inner, lay);
    if (per_col <= kMaxSmem) return launch_encode<1>(ctx, T, n, cols, nchunks, chunk_stride, inner, lay);
    HB_FAIL(ctx, "encode: codeword does not fit in shared memory (message length too large for the column kernel)");
}

}  // namespace hb

// ---------------------------------------------------------------------------------------------------------
extern "C" int hb_expander_set(hb_ctx *ctx, long long n, int levels, int deg_C, int deg_D,
                               const long long *L_C, const long long *R_C, const uint32_t *const *nbr_C, const uint64_t *const *w_C,
                               const long long *L_D, const long long *R_D, const uint32_t *const *nbr_D, const uint64_t *const *w_D) { HB_DEV(ctx);
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

    // CSR by target.  The rows of a stage are stored in PROCESSING order, not in target order: a warp walks the edge lists of its 32 / CB
    // rows in lockstep, so it runs as long as its longest row (in-degrees of these random graphs spread +-40 %: natural order costs 25 %
    // more lockstep steps than the edges need).  Rows are therefore sorted by in-degree, longest first, and laid out in snake order over
    // the passes of kRowSlots rows so that every row slot gets a similar total; each row record carries its target.  Within a row the
    // edges are ordered by the parity of their source row (even rows first for even positions, odd first for odd ones): with 4-column
    // tiles two adjacent rows share one shared-memory wavefront and collide when their sources have the same parity.  The sums are exact
    // in the field, so neither order changes a bit of the result.  (Padding every list to whole four-edge groups with zero-weight edges,
    // so that no row needs the one-edge tail loop, was measured too: 2.884 vs 2.887 ms per launch, not kept.)
    static const int sort_rows = getenv("HB_ENCODE_ROWSORT") ? atoi(getenv("HB_ENCODE_ROWSORT")) : 1;     // experiment switches
    static const int sort_edges = getenv("HB_ENCODE_EDGEPAR") ? atoi(getenv("HB_ENCODE_EDGEPAR")) : 1;
    constexpr long long kRowSlots = 128;                  // row slots of the big-code launches (1024 threads x 8 columns, 512 x 4)
    std::vector<uint2> rowptr; std::vector<uint2> edges;
    auto add_stage = [&](long long in_off, long long L, long long out_off, long long R, int deg, const uint32_t *nbr, const uint64_t *w) -> int {
        EncStage st; st.in_off = (int)in_off; st.out_off = (int)out_off; st.L = (int)L; st.R = (int)R; st.rowptr_base = (int)rowptr.size();
        std::vector<int> cnt(R + 1, 0);
        for (long long i = 0; i < L * deg; i++) { if (nbr[i] >= (uint32_t)R) return 1; if (w[i] >> 32) return 1; if (w[i] >> 31) ex.w31 = false; cnt[nbr[i] + 1]++; }
        // processing position of every target
        std::vector<int> by_deg(R), pos_of(R);
        for (long long t = 0; t < R; t++) { by_deg[t] = (int)t; ex.max_indeg = std::max(ex.max_indeg, cnt[t + 1]); }
        if (sort_rows) std::stable_sort(by_deg.begin(), by_deg.end(), [&](int a, int b) { return cnt[a + 1] > cnt[b + 1]; });
        for (long long r = 0; r < R; r++) {
            long long pass = r / kRowSlots, within = r % kRowSlots, in_pass = std::min(kRowSlots, R - pass * kRowSlots);
            if (sort_rows && (pass & 1)) within = in_pass - 1 - within;
            pos_of[by_deg[r]] = (int)(pass * kRowSlots + within);
        }
        std::vector<int> target_at(R), start(R + 1, 0);
        for (long long t = 0; t < R; t++) target_at[pos_of[t]] = (int)t;
        // every edge list holds an even number of records (one zero-weight record appended where needed) and starts at an even index, so
        // that the kernel fetches two records per 16-byte load: half the L1 wavefronts of the edge loads
        for (long long v = 0; v < R; v++) start[v + 1] = start[v] + ((cnt[target_at[v] + 1] + 1) & ~1);
        size_t ebase = edges.size();                                   // even: every stage appends an even number of records
        edges.resize(ebase + (size_t)start[R], make_uint2((unsigned)in_off, 0u));
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (long long i = 0; i < L; i++)
            for (int j = 0; j < deg; j++) {
                const int v = pos_of[nbr[i * deg + j]];
                edges[ebase + fill[v]++] = make_uint2((unsigned)(in_off + i), (unsigned)w[i * deg + j]);
            }
        if (sort_edges)
            for (long long v = 0; v < R; v++)
                std::stable_partition(edges.begin() + ebase + start[v], edges.begin() + ebase + fill[v],
                                      [&](const uint2 &e) { return ((e.x ^ (unsigned)v) & 1u) == 0; });
        for (long long v = 0; v <= R; v++) rowptr.push_back(make_uint2((unsigned)(ebase + start[v]), v < R ? (unsigned)target_at[v] : 0u));
        ex.stages.push_back(st);
        return 0;
    };
    for (int d = 0; d < levels; d++)
        if (add_stage(off[d], nd[d], off[d + 1], R_C[d], deg_C, nbr_C[d], w_C[d])) HB_FAIL(ctx, "hb_expander_set: bad C graph (target >= R or weight >= 2^32)");
    for (int d = levels - 1; d >= 0; d--)
        if (add_stage(off[d + 1], lenc[d + 1], off[d + 1] + lenc[d + 1], R_D[d], deg_D, nbr_D[d], w_D[d])) HB_FAIL(ctx, "hb_expander_set: bad D graph (target >= R or weight >= 2^32)");
    if (ex.max_indeg >= (1 << 29)) HB_FAIL(ctx, "hb_expander_set: in-degree above 2^29 (headroom of the 128-bit row accumulators)");
    ex.n_edges = edges.size();
    HB_CHECK(ctx, cudaMalloc(&ex.d_stages, ex.stages.size() * sizeof(EncStage)));
    HB_CHECK(ctx, cudaMalloc(&ex.d_rowptr, rowptr.size() * sizeof(uint2)));
    HB_CHECK(ctx, cudaMalloc(&ex.d_edges, edges.size() * sizeof(uint2)));
    HB_CHECK(ctx, cudaMemcpy(ex.d_stages, ex.stages.data(), ex.stages.size() * sizeof(EncStage), cudaMemcpyHostToDevice));
    HB_CHECK(ctx, cudaMemcpy(ex.d_rowptr, rowptr.data(), rowptr.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    HB_CHECK(ctx, cudaMemcpy(ex.d_edges, edges.data(), edges.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" long long hb_expander_codeword_len(hb_ctx *ctx) { return ctx->exp.cwlen; }

extern "C" int hb_encode_batch(hb_ctx *ctx, const hb_F *src, hb_F *dst, long long n, size_t ncols) { HB_DEV(ctx);
    using namespace hb;
    if (ncols == 0) return 0;
    Staged d(ctx);
    HB_TRY(d.outbuf(dst, 2 * (size_t)n * ncols * sizeof(F)));
    HB_CHECK(ctx, cudaMemcpyAsync(d.dev, src, (size_t)n * ncols * sizeof(F), cudaMemcpyDefault, ctx->stream));
    if (n <= 13 && ctx->exp.n != n) {
        // base case of the recursion: the codeword is the message (linear_code_encode.h:73-78)
        HB_CHECK(ctx, cudaMemsetAsync(d.as<F>() + (size_t)n * ncols, 0, (size_t)n * ncols * sizeof(F), ctx->stream));
    } else {
        HB_TRY(encode_cols_dev(ctx, d.as<F>(), n, ncols, 1, 0, nullptr, InnerLayout()));
    }
    HB_TRY(d.finish());
    HB_TRY(end_call(ctx));
    return 0;
}
