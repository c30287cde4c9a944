// Internal glue shared by the .cu translation units: the context object, error handling, host<->device
// staging of caller buffers, and a launch counter.  Nothing here is part of the public C ABI.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <deque>
#include <cuda_runtime.h>
#include "../../include/hobbit_b200.h"
#include "field.cuh"

namespace hb {

static constexpr int kMaxRanks = 8;     // GPUs of one NVSwitch box

struct EncStage {                       // one sparse mat-vec of the expander encode: out[0..R) = G * in[0..L)
    int in_off, out_off, L, R;          // offsets/sizes in codeword coordinates
    int rowptr_base;                    // index into rows[] (R+1 entries)
};

// Where the inner leaf digest of (chunk c of this launch, leaf position p) goes.  Plain: [chunk][leaf].  Sharded exchange layout
// (hobbit_b200/dist.py): [leaf part][chunk of the whole call][leaf inside the part], so that the slice for destination rank h is
// contiguous and no permute pass is needed before the all_to_all.
struct InnerLayout {
    size_t part_leaves = 0;      // leaves per part (0 = unset)
    size_t chunks_total = 0;     // chunks in the whole call
    size_t chunk0 = 0;           // index of this launch's first chunk inside the call
    // Fused exchange (dist.cu): peer[h] != nullptr -> the digests of leaf part h are stored straight into rank h's receive array
    // [global chunk][leaf inside the part] through its NVLink-mapped window; chunk0 then counts GLOBAL chunks.
    uint8_t *peer[kMaxRanks] = {};
    __host__ __device__ size_t offset(size_t chunk_in_launch, size_t p) const {
        size_t part = p / part_leaves, off = p - part * part_leaves;
        return (part * chunks_total + chunk0 + chunk_in_launch) * part_leaves + off;
    }
    __device__ __forceinline__ uint8_t *addr(uint8_t *base, size_t chunk_in_launch, size_t p) const {
        size_t part = p / part_leaves, off = p - part * part_leaves;
        if (peer[0]) return peer[part] + ((chunk0 + chunk_in_launch) * part_leaves + off) * 32;
        return base + ((part * chunks_total + chunk0 + chunk_in_launch) * part_leaves + off) * 32;
    }
    static InnerLayout plain(size_t leaves, size_t chunks) { InnerLayout l; l.part_leaves = leaves; l.chunks_total = chunks; l.chunk0 = 0; return l; }
};

// Multi-GPU state (dist.cu): one process per GPU; every rank owns a WINDOW of device memory that all ranks of the box map through CUDA IPC
// (NVLink peer stores).  Control region (first kDistCtrlBytes): reduction mail slots, barrier flags, small-exchange slots; the rest is
// the data region the sharded commitments exchange digests and tree levels through.
static constexpr size_t kDistCtrlBytes = 64 << 10;
static constexpr size_t kDistMailOff = 0;            // [2 parities][8 ranks][32 F]: per-round sums of the sharded sumchecks
static constexpr size_t kDistBarOff = 16 << 10;      // [8] u64 epochs
static constexpr size_t kDistXchgOff = 24 << 10;     // [2 parities][8 ranks][64 F]: small all-gathers (table heads)
static constexpr size_t kDistErrOff = 60 << 10;      // u64: set when a peer wait timed out
struct DistState {
    int world = 1, rank = 0;
    uint8_t *win = nullptr; size_t win_bytes = 0;    // my window (cudaMalloc, IPC-exported)
    uint8_t *peer[kMaxRanks] = {};                   // every rank's window as mapped here (peer[rank] == win)
    unsigned long long epoch = 0;                    // barrier epochs (identical call sequence on every rank)
    unsigned long long rseq = 0;                     // sharded-reduction sequence
    unsigned long long xseq = 0;                     // small-exchange sequence
    bool shard = false;                              // provers slice their tables by rank and all-reduce the round sums (hb_dist_shard)
    bool reduce_on = false;                          // set by a prover around its sliced phase: grid reductions span all ranks
    size_t min_slice = 2;                            // a table is sliced only while its per-rank part keeps at least this many entries
};

struct ExpanderDev {
    long long n = 0;
    int cwlen = 0;
    std::vector<EncStage> stages;       // execution order: C0..C_last, D_last..D0
    EncStage *d_stages = nullptr;
    uint2 *d_rowptr = nullptr;          // CSR by target, rows in processing order: {first edge, target row within the stage}; R+1 entries per stage
    uint2 *d_edges = nullptr;           // {absolute source index in the codeword, 32-bit weight}
    size_t n_edges = 0;
    int max_indeg = 0;
    bool w31 = true;                    // every weight < 2^31 (the reference draws them with random()): enables the four-edges-at-a-time accumulation
};

// Double-buffered staging of host chunks for the chunk-at-a-time calls (hb_elastic_push, hb_elastic_open_push): the H2D of chunk i runs on
// the copy stream under the encode of chunk i-1, and the call returns only once the caller's buffer has been READ — a streaming
// producer may refill a (pinned) chunk buffer as soon as the push returns.
struct ChunkStager {
    F *buf[2] = {nullptr, nullptr}; cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr}; bool used[2] = {false, false};
    size_t count = 0;
};

struct ElasticState {
    bool active = false;
    size_t B = 0; int trs = 0; int lin = 0;
    size_t chunk_idx = 0;
    F *park[3] = {nullptr, nullptr, nullptr};   // 3 parked encoded chunks (4B each), Elastic_PC.cpp:179-183
    F *tensor = nullptr;                        // current encoded chunk (4B)
    ChunkStager stg;                            // staging for host chunks
    uint8_t *leaves = nullptr;                  // (2*4B-1)*32
    int *nz_flag = nullptr;
    // multi-GPU streaming commit (hb_dist_elastic_begin): this rank pushes its groups_total/world consecutive groups; the four encoded
    // chunks of a group are contiguous (park[0..2], tensor) and their inner digests go straight to the owner ranks
    size_t dist_groups_total = 0; InnerLayout dist_lay;
};

struct ElasticOpen {
    bool active = false;
    size_t B = 0, queries = 0, nchunks = 0, idx = 0; int trs = 0, lin = 0;
    void *buf = nullptr; F *agg = nullptr, *tensor = nullptr, *reply = nullptr; uint32_t *col = nullptr, *row = nullptr;
    size_t chunk_first = 0, chunk_total = 0;    // this pass covers chunks [chunk_first, chunk_first + nchunks) of chunk_total (multi-GPU: a rank's chunk range)
    ChunkStager stg;
};

// one pass of the circuit evaluator's trace, resident in HBM (trace.cu)
struct TraceState {
    void *tuples = nullptr; size_t capacity = 0, n = 0, n_ops = 0, n_del = 0;
    bool done = false, indexed = false;
    unsigned *pos = nullptr;            // is_op | is_del | op_pos | del_pos, n entries each
    size_t pos_capacity = 0;            // entries allocated at pos (kept across traces: cudaFree / cudaMalloc of it cost more than the evaluator)
};

int levels_copy_wait(hb_ctx *ctx);                                   // all queued background level copies are done (commit.cu)
// the chunk-independent half of the commitments (commit.cu), shared with the multi-GPU entry points (dist.cu)
int commit_encode_chunks_impl(hb_ctx *ctx, const hb_F *poly, size_t nchunks, size_t B, int trs, int linear_time, uint8_t *inner_dev,
                              InnerLayout lay0, size_t first_chunk, size_t total_chunks);
int elastic_encode_groups_impl(hb_ctx *ctx, const hb_F *chunks, size_t ngroups, size_t B, int trs, int linear_time, uint8_t *inner_dev, InnerLayout lay0);

// ---- multi-GPU helpers (dist.cu); all no-ops / identity on a single-GPU context -----------------------------------------------------
struct Slice { size_t off, len; bool on; };
// the part of an n-entry table this rank works on inside a prover: [rank * n/G, (rank+1) * n/G) when sharding is enabled and the part
// keeps at least min_slice entries, otherwise the whole table (on = false: every rank then computes the same thing redundantly)
Slice dist_slice(hb_ctx *ctx, size_t n);
void dist_release(hb_ctx *ctx);                                     // unmap the peers, free the window
int dist_barrier_dev(hb_ctx *ctx);                                   // cross-rank barrier, stream-ordered (a kernel on ctx->stream)
// every rank contributes `cnt` leading entries of each of its nt tables; out_dev: nt small tables of cnt * world entries (rank order)
int dist_gather_small(hb_ctx *ctx, const F *const *tabs, int nt, int cnt, F *out_dev);
int dist_allreduce_vec(hb_ctx *ctx, F *vec_dev, size_t n);           // field sum over the ranks, in place, through the data region
// the two halves of a sharded commitment (see dist.cu) and the global tree it leaves in the window
int sharded_begin(hb_ctx *ctx, size_t units_total, size_t leaves, InnerLayout *lay_out);
int sharded_finish(hb_ctx *ctx, size_t units_total, size_t leaves, uint8_t *levels_out);
const uint8_t *sharded_tree(hb_ctx *ctx, size_t units_total, size_t leaves);

// Background copy of Merkle levels to caller (pageable) memory: a worker thread with its own stream and pinned double buffer drains a queue
// of jobs while the context's stream goes on proving (hb_elastic_finish_levels_async / hb_levels_wait).
struct LevelsJob {
    uint8_t *dev = nullptr; bool owned = false;          // flat levels on the device (owned: freed when the job is done)
    cudaEvent_t ready = nullptr;                         // recorded on the context's stream once the tree is complete
    std::vector<uint8_t *> dst; std::vector<size_t> off, bytes;
};
struct LevelsCopier {
    std::thread th; std::mutex mu; std::condition_variable cv, cv_idle;
    std::deque<LevelsJob> q; bool busy = false, stop = false, started = false;
    cudaStream_t stream = nullptr; void *pin[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr};
    std::string err;
};

}  // namespace hb

struct hb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;      // compute
    cudaStream_t copy_stream = nullptr; // H2D prefetch of the next chunk
    std::string err;
    uint64_t launches = 0;
    void *pin[2] = {nullptr, nullptr};  // pinned staging for large PAGEABLE host buffers (hb::copy_from_host / copy_to_host)
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    bool sync_needed = false;           // set while a call has host-visible outputs (or pinned host inputs) in flight: see hb::end_call
    int sm_count = 148;                 // overwritten from cudaDeviceProp in hb_ctx_create; grids of grid-stride kernels are multiples of it
    // twiddle tables w[k] = omega_len^k, k < len/2, cached per log2(len)
    hb::F *tw[32] = {};
    hb::F *tw_pass[32] = {};            // per-pass twiddle tables of the tile kernel (ntt.cu: pass_tables), lane-contiguous
    bool tw_j_neg[32] = {};             // omega_len^(len/4) == -i (else +i)
    uint8_t tw_w8[32] = {};             // omega_len^(len/8) = 2^30 (+-1 +- i): bit 0 = real part negative, bit 1 = imaginary part negative
    hb::ExpanderDev exp;
    // resident tensor of the last commit_standard
    hb::F *tensor = nullptr; size_t tensor_elems = 0; size_t tensor_N = 0; int tensor_K = 0; int tensor_trs = 0;
    // device copy of the polynomial staged by the last commit_standard (open_standard's aggregate reads it again)
    hb::F *poly = nullptr; size_t poly_elems = 0; bool poly_valid = false;
    hb::ElasticState el;
    hb::ElasticOpen eo;
    hb::TraceState trace;
    // sumcheck reduction scratch: per-CTA partial coefficients + ticket counter (device), result mailbox (pinned host)
    hb::F *red = nullptr; unsigned *ticket = nullptr; hb::F *mailbox = nullptr; hb::F *mailbox_dev = nullptr; unsigned long long seq = 0;
    hb::LevelsCopier *lvl = nullptr;    // created on first use
    hb::DistState dist;
    // FNV-1a over every value a prover read back from the GPU (round sums, table heads): a digest of the whole Fiat–Shamir transcript
    uint64_t transcript = 0xcbf29ce484222325ULL;
    // optional per-kernel timing (hb_profile_*): CUDA events around every launch, on this context's stream
    bool prof = false;
    struct ProfRec { const char *name; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
};
namespace hb {
inline cudaEvent_t prof_event(hb_ctx *ctx) {
    cudaEvent_t e;
    if (!ctx->prof_pool.empty()) { e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}
}

#define HB_CHECK(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" + \
                         std::to_string(__LINE__) + ")";                                      \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)
// every C-ABI entry binds the calling thread to the context's device first (a context may be used from any host thread, and other
// libraries in the process may have changed the current device)
#define HB_DEV(ctx) do { cudaSetDevice((ctx)->device); } while (0)
#define HB_TRY(expr) do { int r__ = (expr); if (r__) return r__; } while (0)
#define HB_FAIL(ctx, msg) do { (ctx)->err = (msg); return 2; } while (0)
// every kernel launch goes through this so the launch counter is honest
#define HB_LAUNCH(ctx, kernel, grid, block, smem, ...)                                        \
    do {                                                                                      \
        cudaEvent_t pe0__ = nullptr, pe1__ = nullptr;                                         \
        if ((ctx)->prof) { pe0__ = hb::prof_event(ctx); pe1__ = hb::prof_event(ctx); cudaEventRecord(pe0__, (ctx)->stream); } \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                      \
        (ctx)->launches++;                                                                    \
        if ((ctx)->prof) { cudaEventRecord(pe1__, (ctx)->stream); (ctx)->prof_recs.push_back({#kernel, pe0__, pe1__}); } \
        HB_CHECK(ctx, cudaGetLastError());                                                    \
    } while (0)

namespace hb {

inline bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
inline bool is_pinned_host_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
// End of a C-ABI call whose results may all live in device memory: the stream is synchronised only if something host-visible is in flight
// (a host output buffer, or a pinned host input that a truly asynchronous copy is still reading; pageable inputs are copied to the driver's
// staging buffer before cudaMemcpyAsync returns).  Calls on HBM-resident tables therefore stay stream-ordered and return immediately.
inline int end_call(hb_ctx *ctx) {
    if (ctx->sync_needed) {
        ctx->sync_needed = false;
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e); return 1; }
    }
    return 0;
}

// Large PAGEABLE host buffers (the reference's std::vector storage) cross PCIe through the context's pinned double buffer: the host-side
// memcpy of one piece (worker threads) overlaps the DMA of the other.  cudaMemcpy on pageable memory stages through the driver at
// 3-6 GB/s (measured: 128 MiB of Merkle levels into a std::vector in 44 ms); this path is bound by the host memcpy instead.
constexpr size_t kPageableDirect = (size_t)1 << 20;      // below this the plain cudaMemcpyAsync is used
int copy_from_host(hb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, cudaStream_t stream);   // returns once src has been read
int copy_to_host(hb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes, cudaStream_t stream);     // returns once dst is filled

// RAII staging of a caller buffer: device pointers pass through, host pointers get a stream-ordered temporary.
struct Staged {
    hb_ctx *ctx; void *dev = nullptr; void *host = nullptr; size_t bytes = 0; bool owned = false; bool out = false;
    Staged(hb_ctx *c) : ctx(c) {}
    int in(const void *p, size_t n) {
        bytes = n;
        if (n == 0) { dev = nullptr; return 0; }
        if (is_device_ptr(p)) { dev = const_cast<void *>(p); return 0; }
        owned = true; host = const_cast<void *>(p);
        const bool pinned = is_pinned_host_ptr(p);
        if (pinned) ctx->sync_needed = true;
        HB_CHECK(ctx, cudaMallocAsync(&dev, n, ctx->stream));
        if (!pinned && n >= kPageableDirect) return copy_from_host(ctx, dev, p, n, ctx->stream);
        HB_CHECK(ctx, cudaMemcpyAsync(dev, p, n, cudaMemcpyHostToDevice, ctx->stream));
        return 0;
    }
    int outbuf(void *p, size_t n, bool copy_in = false) {
        bytes = n; out = true;
        if (n == 0) { dev = nullptr; return 0; }
        if (is_device_ptr(p)) { dev = p; return 0; }
        owned = true; host = p; ctx->sync_needed = true;
        HB_CHECK(ctx, cudaMallocAsync(&dev, n, ctx->stream));
        if (copy_in) HB_CHECK(ctx, cudaMemcpyAsync(dev, p, n, cudaMemcpyHostToDevice, ctx->stream));
        return 0;
    }
    // copy back (if host output) and synchronise
    int finish() {
        if (owned && out && bytes) {
            if (bytes >= kPageableDirect && !is_pinned_host_ptr(host)) return copy_to_host(ctx, host, dev, bytes, ctx->stream);
            HB_CHECK(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        }
        return 0;
    }
    ~Staged() { if (owned && dev) cudaFreeAsync(dev, ctx->stream); }
    template <class T> T *as() { return reinterpret_cast<T *>(dev); }
};

inline void transcript_absorb(hb_ctx *ctx, const F *v, int n) {
    uint64_t h = ctx->transcript;
    for (int i = 0; i < n; i++) { h = (h ^ v[i].re) * 0x100000001b3ULL; h = (h ^ v[i].im) * 0x100000001b3ULL; }
    ctx->transcript = h;
}
inline int ilog2(size_t x) { int l = 0; while (x >>= 1) l++; return l; }

// internal entry points implemented across the .cu files (all take DEVICE pointers)
int get_twiddles(hb_ctx *ctx, int logn, const F **out);
int ntt_rows_dev(hb_ctx *ctx, F *data, int logn, size_t batch, size_t stride);
// rows of `in_len` elements (row r at src + r*in_len) zero-extended to 2^logn and transformed into dst + r*dst_stride
// nchunks chunks of rows_per_chunk rows each; chunk c starts at src + c*src_chunk_stride / dst + c*dst_chunk_stride
int ntt_rows_padded_dev(hb_ctx *ctx, const F *src, size_t in_len, F *dst, size_t dst_stride, int logn, size_t rows_per_chunk,
                        size_t nchunks, size_t src_chunk_stride, size_t dst_chunk_stride);
// column NTT of a (2^logn x cols) row-major matrix whose rows >= nz_rows are implicitly zero on input
int ntt_cols_dev(hb_ctx *ctx, F *mat, int logn, size_t cols, size_t nz_rows);
// expander-encode every column of nchunks matrices T + c*chunk_stride (rows [0,n) hold the messages); writes rows [n, 2n).
// inner != nullptr: also the inner leaf digests of commit_standard, chunk c at inner + c*(n/2*cols)*32.
int encode_cols_dev(hb_ctx *ctx, F *T, long long n, size_t cols, size_t nchunks, size_t chunk_stride, uint8_t *inner, InnerLayout lay);
// inner digests H1(T[4j][k] | T[4j+1][k] | T[4j+2][k] | T[4j+3][k]) of nchunks matrices (rows x cols each)
int md_inner_standard_dev(hb_ctx *ctx, const F *T, size_t rows, size_t cols, size_t nchunks, size_t chunk_stride, uint8_t *inner, InnerLayout lay);
// leaf[p] <- H1(inner[c][p] | leaf[p]) for c = 0..nchunks-1 in order (the Merkle–Damgård chain over chunks)
int md_chain_dev(hb_ctx *ctx, const uint8_t *inner, size_t nchunks, size_t nleaves, uint8_t *leaves);
int md_inner_stream4_dev(hb_ctx *ctx, const F *T4, size_t cells, size_t ngroups, uint8_t *inner, InnerLayout lay);
int md_leaves_stream4_dev(hb_ctx *ctx, const F *c0, const F *c1, const F *c2, const F *T, size_t cells, uint8_t *leaves);
int blake3_64_dev(hb_ctx *ctx, const uint8_t *src, uint8_t *dst, size_t count);
int merkle_tree_dev(hb_ctx *ctx, uint8_t *levels, size_t nleaves);
// nchunks messages of n elements (contiguous) -> nchunks tensors of 4n elements (contiguous); inner: see encode_cols_dev
int tensorcode_dev(hb_ctx *ctx, const F *msg, size_t n, int trs, int lin, F *T, size_t nchunks, uint8_t *inner, InnerLayout lay = InnerLayout());

}  // namespace hb
