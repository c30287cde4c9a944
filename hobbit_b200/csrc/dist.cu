// Multi-GPU layer of the C ABI: one process per GPU, all GPUs of one NVSwitch box.
//
// Every rank owns a WINDOW of device memory that every other rank maps through CUDA IPC, so kernels exchange data with plain stores over
// NVLink and synchronise through flags in the window — no host round trip and no separate collective launch on the data path:
//   * the expander-encode kernel stores each inner leaf digest straight into the receive array of the rank that owns the leaf range
//     (InnerLayout::peer), so the digest exchange of a sharded commitment is part of the encode kernel;
//   * the last CTA of a sumcheck round kernel adds the round sums of all ranks through per-rank mail slots (reduce.cuh, grid_reduce);
//   * subtree levels are scattered to every rank's copy of the global tree by one kernel.
// Bootstrap is the caller's: hb_dist_local_info() returns a 256-byte blob, the caller all-gathers the blobs of all ranks by whatever
// means it has (torch.distributed in bench.py / the tests, a TCP store in the C++ host mirror) and hands them to hb_dist_connect().
//
// Reference: the reference is single-process; this file implements the partitioning §8e derives from Our_PC.cpp:146-171
// (chunks are independent up to the Merkle–Damgård chain of each leaf), Elastic_PC.cpp:174-285 and sumcheck.cpp:1974-2058.
#include "common.cuh"
#include "reduce.cuh"
#include <algorithm>
#include <unistd.h>

namespace hb {

struct PeerPtrs { uint8_t *p[kMaxRanks]; };
static PeerPtrs peer_ptrs(hb_ctx *ctx) { PeerPtrs q; for (int h = 0; h < kMaxRanks; h++) q.p[h] = ctx->dist.peer[h]; return q; }

struct DistBlob {                   // what hb_dist_local_info hands out (256 bytes)
    uint64_t magic; int32_t device, pid; uint64_t win_bytes; cudaIpcMemHandle_t handle; uint8_t pad[256 - 24 - sizeof(cudaIpcMemHandle_t)];
};
static_assert(sizeof(DistBlob) == 256, "blob layout");
static constexpr uint64_t kBlobMagic = 0x484f424249543032ULL;     // "HOBBIT02"

// monotonic epochs: rank r stores `epoch` into flag[r] of every window, then waits until every flag of its own window reached it
__global__ void dist_barrier_kernel(PeerPtrs pp, int rank, int world, u64 epoch) {
    const int h = threadIdx.x;
    if (h < world) {
        __threadfence_system();
        volatile u64 *dst = reinterpret_cast<u64 *>(pp.p[h] + kDistBarOff) + rank;
        *dst = epoch;
        const volatile u64 *mine = reinterpret_cast<const u64 *>(pp.p[rank] + kDistBarOff) + h;
        const long long t0 = clock64();
        while (*mine < epoch) {
            if (clock64() - t0 > 60000000000LL) { *reinterpret_cast<volatile u64 *>(pp.p[rank] + kDistErrOff) = 1; break; }
        }
        __threadfence_system();
    }
}

// small all-gather: value v = k * cnt + e is tabs[k][e]; out[k * cnt * world + g * cnt + e] = rank g's value
struct SmallTabs { const F *t[16]; };
__global__ void dist_xchg_kernel(PeerPtrs pp, int rank, int world, u64 seq, SmallTabs tabs, int nt, int cnt, F *__restrict__ out) {
    const int nv = nt * cnt, v = threadIdx.x;
    const size_t slot_u64 = 64 * 2;                                                     // 64 F per (parity, rank)
    const size_t my_slot = ((seq & 1) * kMaxRanks + rank) * slot_u64;
    if (v < nv) {
        const F x = tabs.t[v / cnt][v % cnt];
        for (int h = 0; h < world; h++) {
            volatile u64 *dst = reinterpret_cast<u64 *>(pp.p[h] + kDistXchgOff) + my_slot;
            dst[2 * v] = x.re; dst[2 * v + 1] = x.im;
        }
        __threadfence_system();
    }
    __syncthreads();
    bool ok = true;
    if (v < world) {
        __threadfence_system();
        volatile u64 *dst = reinterpret_cast<u64 *>(pp.p[v] + kDistXchgOff) + my_slot;
        dst[2 * 63] = seq;
        const volatile u64 *mine = reinterpret_cast<const u64 *>(pp.p[rank] + kDistXchgOff) + ((seq & 1) * kMaxRanks + v) * slot_u64;
        ok = wait_flag(mine + 2 * 63, seq);
        __threadfence_system();
    }
    ok = __syncthreads_and(ok);
    if (!ok && v == 0) *reinterpret_cast<volatile u64 *>(pp.p[rank] + kDistErrOff) = 1;
    if (v < nv) {
        const int k = v / cnt, e = v % cnt;
        for (int g = 0; g < world; g++) {
            const volatile u64 *src = reinterpret_cast<const u64 *>(pp.p[rank] + kDistXchgOff) + ((seq & 1) * kMaxRanks + g) * slot_u64;
            out[(size_t)k * cnt * world + (size_t)g * cnt + e] = mkF(src[2 * v], src[2 * v + 1]);
        }
    }
}

// my subtree (every level, 2*Bp-1 digests, leaves first) -> the same nodes of the GLOBAL tree (B = Bp*world leaves) in every rank's window
__global__ void __launch_bounds__(256) dist_scatter_tree_kernel(PeerPtrs pp, size_t tree_off, int rank, int world, const uint4 *__restrict__ sub, size_t Bp) {
    const size_t total = 2 * Bp - 1, B = Bp * world;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        // level l of the subtree starts at 2Bp - (2Bp >> l); the global level l starts at 2B - (2B >> l)
        int l = 0; size_t lvl_n = Bp, start = 0;
        while (i >= start + lvl_n) { start += lvl_n; lvl_n >>= 1; l++; }
        const size_t pos = i - start;
        const size_t gidx = (2 * B - ((2 * B) >> l)) + (size_t)rank * lvl_n + pos;
        const uint4 a = sub[2 * i], b = sub[2 * i + 1];
        for (int h = 0; h < world; h++) {
            uint4 *dst = reinterpret_cast<uint4 *>(pp.p[h] + tree_off) + 2 * gidx;
            dst[0] = a; dst[1] = b;
        }
    }
}

// field all-reduce of a vector through the data region: put (my vector into slot[rank] of every window) | barrier | sum
__global__ void __launch_bounds__(256) dist_vec_put_kernel(PeerPtrs pp, size_t off, int rank, int world, const F *__restrict__ v, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const F x = v[i];
        for (int h = 0; h < world; h++) reinterpret_cast<F *>(pp.p[h] + off)[(size_t)rank * n + i] = x;
    }
}
__global__ void __launch_bounds__(256) dist_vec_sum_kernel(const F *__restrict__ slots, int world, F *__restrict__ v, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        F s = slots[i];
        for (int g = 1; g < world; g++) s = fadd(s, slots[(size_t)g * n + i]);
        v[i] = s;
    }
}

static int check_peer_error(hb_ctx *ctx) {
    u64 e = 0;
    HB_CHECK(ctx, cudaMemcpyAsync(&e, ctx->dist.win + kDistErrOff, sizeof(e), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (e) HB_FAIL(ctx, "multi-GPU: a peer rank did not reach a barrier / exchange within 30 s");
    return 0;
}

Slice dist_slice(hb_ctx *ctx, size_t n) {
    const DistState &d = ctx->dist;
    Slice s{0, n, false};
    if (d.world > 1 && d.shard && n % d.world == 0 && n / d.world >= std::max<size_t>(d.min_slice, 1)) { s.len = n / d.world; s.off = s.len * d.rank; s.on = true; }
    return s;
}

int dist_barrier_dev(hb_ctx *ctx) {
    DistState &d = ctx->dist;
    if (d.world <= 1) return 0;
    HB_LAUNCH(ctx, dist_barrier_kernel, 1, 32, 0, peer_ptrs(ctx), d.rank, d.world, (u64)++d.epoch);
    return 0;
}

int dist_gather_small(hb_ctx *ctx, const F *const *tabs, int nt, int cnt, F *out_dev) {
    DistState &d = ctx->dist;
    if (nt * cnt > 32 || nt > 16) HB_FAIL(ctx, "dist_gather_small: at most 32 values per rank");
    SmallTabs st; for (int k = 0; k < 16; k++) st.t[k] = k < nt ? tabs[k] : nullptr;
    HB_LAUNCH(ctx, dist_xchg_kernel, 1, 64, 0, peer_ptrs(ctx), d.rank, d.world, (u64)++d.xseq, st, nt, cnt, out_dev);
    return 0;
}

int dist_allreduce_vec(hb_ctx *ctx, F *vec_dev, size_t n) {
    DistState &d = ctx->dist;
    if (d.world <= 1 || n == 0) return 0;
    if (kDistCtrlBytes + n * sizeof(F) * d.world > d.win_bytes) HB_FAIL(ctx, "dist_allreduce_vec: window too small");
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 4);
    HB_TRY(dist_barrier_dev(ctx));                                        // nobody still reads the data region of the previous operation
    HB_LAUNCH(ctx, dist_vec_put_kernel, grid, 256, 0, peer_ptrs(ctx), kDistCtrlBytes, d.rank, d.world, vec_dev, n);
    HB_TRY(dist_barrier_dev(ctx));
    HB_LAUNCH(ctx, dist_vec_sum_kernel, grid, 256, 0, reinterpret_cast<const F *>(d.win + kDistCtrlBytes), d.world, vec_dev, n);
    return 0;
}

void dist_release(hb_ctx *ctx) {
    DistState &d = ctx->dist;
    for (int h = 0; h < kMaxRanks; h++) if (d.peer[h] && d.peer[h] != d.win) cudaIpcCloseMemHandle(d.peer[h]);
    if (d.win) cudaFree(d.win);
    d = DistState();
}

}  // namespace hb

using namespace hb;

// =====================================================================================================================================
extern "C" int hb_dist_local_info(hb_ctx *ctx, size_t data_bytes, void *blob256) {
    cudaSetDevice(ctx->device);
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    dist_release(ctx);
    DistState &d = ctx->dist;
    d.win_bytes = kDistCtrlBytes + ((data_bytes + 255) & ~(size_t)255);
    HB_CHECK(ctx, cudaMalloc(&d.win, d.win_bytes));
    HB_CHECK(ctx, cudaMemset(d.win, 0, kDistCtrlBytes));
    HB_CHECK(ctx, cudaDeviceSynchronize());
    DistBlob b; memset(&b, 0, sizeof(b));
    b.magic = kBlobMagic; b.device = ctx->device; b.pid = (int32_t)getpid(); b.win_bytes = d.win_bytes;
    HB_CHECK(ctx, cudaIpcGetMemHandle(&b.handle, d.win));
    memcpy(blob256, &b, sizeof(b));
    return 0;
}

extern "C" int hb_dist_connect(hb_ctx *ctx, int rank, int world, const void *blobs) {
    cudaSetDevice(ctx->device);
    DistState &d = ctx->dist;
    if (!d.win) HB_FAIL(ctx, "hb_dist_connect: call hb_dist_local_info first");
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world || (world & (world - 1))) HB_FAIL(ctx, "hb_dist_connect: world must be 1, 2, 4 or 8");
    const DistBlob *bl = reinterpret_cast<const DistBlob *>(blobs);
    for (int h = 0; h < world; h++) {
        if (bl[h].magic != kBlobMagic) HB_FAIL(ctx, "hb_dist_connect: bad blob");
        if (bl[h].win_bytes != d.win_bytes) HB_FAIL(ctx, "hb_dist_connect: every rank must ask for the same window size");
        if (h == rank) { d.peer[h] = d.win; continue; }
        void *q = nullptr;
        HB_CHECK(ctx, cudaIpcOpenMemHandle(&q, bl[h].handle, cudaIpcMemLazyEnablePeerAccess));
        d.peer[h] = (uint8_t *)q;
    }
    d.rank = rank; d.world = world; d.epoch = d.rseq = d.xseq = 0;
    if (const char *e = getenv("HB_DIST_MIN_SLICE")) d.min_slice = std::max<size_t>(1, (size_t)atoll(e));
    HB_TRY(ensure_scratch(ctx));
    return 0;
}

extern "C" int hb_dist_rank(hb_ctx *ctx) { HB_DEV(ctx);  return ctx->dist.rank; }
extern "C" int hb_dist_world(hb_ctx *ctx) { HB_DEV(ctx);  return ctx->dist.world; }
extern "C" int hb_dist_shard(hb_ctx *ctx, int on) { HB_DEV(ctx);  ctx->dist.shard = on != 0 && ctx->dist.world > 1; return 0; }
extern "C" int hb_dist_disconnect(hb_ctx *ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); dist_release(ctx); return 0; }

extern "C" void hb_dist_stats(hb_ctx *ctx, uint64_t *out3) { out3[0] = ctx->dist.rseq; out3[1] = ctx->dist.xseq; out3[2] = ctx->dist.epoch; }

extern "C" int hb_dist_barrier(hb_ctx *ctx) { HB_DEV(ctx);
    HB_TRY(dist_barrier_dev(ctx));
    if (ctx->dist.world > 1) HB_TRY(check_peer_error(ctx));
    return 0;
}

extern "C" int hb_dist_allreduce(hb_ctx *ctx, hb_F *vec, size_t n) { HB_DEV(ctx);
    Staged v(ctx);
    HB_TRY(v.outbuf(vec, n * sizeof(F), true));
    HB_TRY(dist_allreduce_vec(ctx, v.as<F>(), n));
    HB_TRY(dist_barrier_dev(ctx));
    HB_TRY(v.finish());
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- sharded commitments ------------------------------------------------------------------------------------------------------------
// units: `units_total` chunks (commit_standard) or groups of 4 chunks (Elastic_PC) in the global order, rank g owns the contiguous range
// [g * units/G, (g+1) * units/G) for the chunk-independent part and the leaf range [g * leaves/G, (g+1) * leaves/G) for the chain + subtree.
// first half: checks, barrier, and the layout that sends every inner digest to the rank owning its leaf
int hb::sharded_begin(hb_ctx *ctx, size_t units_total, size_t leaves, InnerLayout *lay_out) {
    DistState &d = ctx->dist;
    const size_t G = d.world;
    if (units_total % G || leaves % G) HB_FAIL(ctx, "sharded commit: units and leaves must split evenly over the ranks");
    const size_t ul = units_total / G, Lp = leaves / G;
    const size_t recv_bytes = (units_total * Lp * 32 + 255) & ~(size_t)255, tree_bytes = (2 * leaves - 1) * 32;
    if (kDistCtrlBytes + recv_bytes + tree_bytes > d.win_bytes) {
        ctx->err = "sharded commit: window too small, need " + std::to_string(recv_bytes + tree_bytes) + " data bytes (hb_dist_local_info)";
        return 2;
    }
    HB_TRY(dist_barrier_dev(ctx));                                        // every rank is done with the window contents of the previous call
    InnerLayout lay; lay.part_leaves = Lp; lay.chunks_total = units_total; lay.chunk0 = (size_t)d.rank * ul;
    for (size_t h = 0; h < G; h++) lay.peer[h] = d.peer[h] + kDistCtrlBytes;
    *lay_out = lay;
    return 0;
}
// second half: chain of my leaf range over all units, my subtree, scatter to every rank, top levels; levels_out may be NULL
int hb::sharded_finish(hb_ctx *ctx, size_t units_total, size_t leaves, uint8_t *levels_out) {
    DistState &d = ctx->dist;
    const size_t G = d.world, Lp = leaves / G;
    const size_t recv_bytes = (units_total * Lp * 32 + 255) & ~(size_t)255, tree_bytes = (2 * leaves - 1) * 32;
    const size_t recv_off = kDistCtrlBytes, tree_off = kDistCtrlBytes + recv_bytes;
    HB_TRY(dist_barrier_dev(ctx));                                        // every rank's digests of MY leaf range have landed
    uint8_t *sub; HB_CHECK(ctx, cudaMallocAsync(&sub, (2 * Lp - 1) * 32, ctx->stream));
    HB_CHECK(ctx, cudaMemsetAsync(sub, 0, Lp * 32, ctx->stream));         // the chain starts from all-zero digests
    int rc = md_chain_dev(ctx, d.win + recv_off, units_total, Lp, sub);
    if (!rc) rc = merkle_tree_dev(ctx, sub, Lp);
    if (rc) { cudaFreeAsync(sub, ctx->stream); return rc; }
    HB_LAUNCH(ctx, dist_scatter_tree_kernel, (unsigned)std::min<size_t>((2 * Lp + 255) / 256, (size_t)ctx->sm_count * 8), 256, 0,
              peer_ptrs(ctx), tree_off, d.rank, d.world, reinterpret_cast<const uint4 *>(sub), Lp);
    cudaFreeAsync(sub, ctx->stream);
    HB_TRY(dist_barrier_dev(ctx));                                        // all subtrees are in my copy of the global tree
    uint8_t *tree = d.win + tree_off;
    if (G > 1) {                                                          // the top log2 G levels from the G subtree roots, on every rank
        int lsub = ilog2(Lp);
        size_t off = 2 * leaves - ((2 * leaves) >> lsub);
        HB_TRY(merkle_tree_dev(ctx, tree + off * 32, G));
    }
    if (levels_out) {
        if (!is_device_ptr(levels_out) && !is_pinned_host_ptr(levels_out) && tree_bytes >= kPageableDirect) { HB_TRY(copy_to_host(ctx, levels_out, tree, tree_bytes, ctx->stream)); }
        else HB_CHECK(ctx, cudaMemcpyAsync(levels_out, tree, tree_bytes, cudaMemcpyDefault, ctx->stream));
    }
    HB_TRY(check_peer_error(ctx));                                        // synchronises the stream
    return 0;
}
const uint8_t *hb::sharded_tree(hb_ctx *ctx, size_t units_total, size_t leaves) {
    const size_t Lp = leaves / ctx->dist.world;
    return ctx->dist.win + kDistCtrlBytes + ((units_total * Lp * 32 + 255) & ~(size_t)255);
}

static int sharded_commit(hb_ctx *ctx, const hb_F *local, size_t units_total, size_t leaves, size_t B, int trs, int lin, bool elastic, uint8_t *levels_out) {
    DistState &d = ctx->dist;
    InnerLayout lay;
    HB_TRY(sharded_begin(ctx, units_total, leaves, &lay));
    const size_t ul = units_total / d.world;
    if (elastic) HB_TRY(elastic_encode_groups_impl(ctx, local, ul, B, trs, lin, d.win + kDistCtrlBytes, lay));
    else HB_TRY(commit_encode_chunks_impl(ctx, local, ul, B, trs, lin, d.win + kDistCtrlBytes, lay, 0, ul));
    return sharded_finish(ctx, units_total, leaves, levels_out);
}

extern "C" int hb_dist_commit_standard(hb_ctx *ctx, const hb_F *poly_local, size_t K_total, size_t B, int trs, int linear_time, uint8_t *levels_out) {
    cudaSetDevice(ctx->device);
    if (B == 0 || (B & (B - 1))) HB_FAIL(ctx, "hb_dist_commit_standard: chunk size must be a power of two");
    if (ctx->dist.world <= 1) return hb_commit_standard(ctx, poly_local, K_total * B, (int)K_total, trs, linear_time, levels_out, nullptr);
    return sharded_commit(ctx, poly_local, K_total, B, B, trs, linear_time, false, levels_out);
}

extern "C" int hb_dist_elastic_commit(hb_ctx *ctx, const hb_F *chunks_local, size_t groups_total, size_t B, int trs, int linear_time, uint8_t *levels_out) {
    cudaSetDevice(ctx->device);
    if (B == 0 || (B & (B - 1))) HB_FAIL(ctx, "hb_dist_elastic_commit: BUFFER_SPACE must be a power of two");
    if (ctx->dist.world <= 1) {
        uint8_t *inner; HB_CHECK(ctx, cudaMallocAsync(&inner, groups_total * 4 * B * 32, ctx->stream));
        uint8_t *lv; HB_CHECK(ctx, cudaMallocAsync(&lv, (8 * B - 1) * 32, ctx->stream));
        HB_CHECK(ctx, cudaMemsetAsync(lv, 0, 4 * B * 32, ctx->stream));
        InnerLayout lay; lay.part_leaves = 4 * B; lay.chunks_total = groups_total; lay.chunk0 = 0;
        int rc = elastic_encode_groups_impl(ctx, chunks_local, groups_total, B, trs, linear_time, inner, lay);
        if (!rc) rc = md_chain_dev(ctx, inner, groups_total, 4 * B, lv);
        if (!rc) rc = merkle_tree_dev(ctx, lv, 4 * B);
        if (!rc && levels_out) HB_CHECK(ctx, cudaMemcpyAsync(levels_out, lv, (8 * B - 1) * 32, cudaMemcpyDefault, ctx->stream));
        cudaFreeAsync(inner, ctx->stream); cudaFreeAsync(lv, ctx->stream);
        HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        return rc;
    }
    return sharded_commit(ctx, chunks_local, groups_total, 4 * B, B, trs, linear_time, true, levels_out);
}
