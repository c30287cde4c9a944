// Context, field primitives and the commitment pipelines (T1, C1, C2) of the C ABI.
// Reference: compute_tensorcode / _compute_tensorcode (src/PC_utils.cpp:9-123), commit_standard
// (src/Our_PC.cpp:146-171), Elastic_PC commit (src/Elastic_PC.cpp:174-285).
//
// Data layout in HBM: the encoded tensor of chunk i is a dense row-major (2*trs) x (2B/trs) matrix of 16-byte F
// at tensor + i*4B — the same order the reference's `_tensor[i][row][col]` flattens to.  Leaves/levels are one flat
// array of 32-byte digests, level after level, leaves first (== the reference's MT_hashes[lvl][i]).
#include "common.cuh"
#include <thread>
#include <sys/mman.h>
#include <new>
#include <algorithm>

namespace hb {

// ---- F1 ------------------------------------------------------------------------------------------------
__device__ F fpow_inv(F x) {                       // x^(p^2-2), fieldElement.cpp:206-209,322-334
    // p^2 - 2 = 2^122 - 2^62 - 1 : bits 0..61 set except bit 61?  (2^122 - 2^62 - 1) = [bits 62..121 set] - 1 ...
    // computed generically from the 128-bit exponent to stay obviously right
    unsigned __int128 e = (unsigned __int128)P61 * P61 - 2;
    F ret = mkF(1, 0), tmp = x;
    while (e) { if (e & 1) ret = fmul(ret, tmp); tmp = fmul(tmp, tmp); e >>= 1; }
    return ret;
}
__global__ void __launch_bounds__(256) field_binop_kernel(int op, const F *__restrict__ a, const F *__restrict__ b, F *__restrict__ c, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        F x = a[i], r;
        switch (op) {
            case 0: r = fadd(x, b[i]); break;
            case 1: r = fsub(x, b[i]); break;
            case 2: r = fmul(x, b[i]); break;
            case 3: r = fneg(x); break;
            default: r = fpow_inv(x); break;
        }
        c[i] = r;
    }
}

// ---- O1 helpers ----------------------------------------------------------------------------------------
// reply[q*K + i] = tensor[i][row[q]][col[q]]   (Our_PC.cpp:291-305)
__global__ void tensor_gather_kernel(const F *__restrict__ tensor, size_t chunk_elems, size_t cols, int K,
                                     const uint32_t *__restrict__ col, const uint32_t *__restrict__ row, size_t queries, F *__restrict__ reply) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= queries * K) return;
    size_t q = t / K; int i = (int)(t % K);
    reply[t] = tensor[(size_t)i * chunk_elems + (size_t)row[q] * cols + col[q]];
}
// agg[j] = sum_i beta[i] * poly[i*B + j]   (Our_PC.cpp:265-272); HBM-bound: 16 B read per coefficient
__global__ void __launch_bounds__(256) aggregate_kernel(const F *__restrict__ poly, size_t B, int K, const F *__restrict__ beta, F *__restrict__ agg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F *sb = reinterpret_cast<F *>(smem_raw);
    for (int i = threadIdx.x; i < K; i += blockDim.x) sb[i] = beta[i];
    __syncthreads();
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < B; j += (size_t)gridDim.x * blockDim.x) {
        F acc = mkF(0, 0);
        for (int i = 0; i < K; i++) acc = fadd(acc, fmul(sb[i], poly[(size_t)i * B + j]));
        agg[j] = acc;
    }
}

// agg[j] += beta * chunk[j]   (Elastic_PC.cpp:326-333, one chunk of the streaming aggregate)
__global__ void __launch_bounds__(256) axpy_kernel(F *__restrict__ agg, const F *__restrict__ chunk, F beta, size_t B) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < B; j += (size_t)gridDim.x * blockDim.x)
        agg[j] = fadd(agg[j], fmul(beta, chunk[j]));
}
// reply[q * nchunks + idx] = T[row[q]][col[q]]   (update_reply, Elastic_PC.cpp:59-110)
__global__ void reply_gather_kernel(const F *__restrict__ T, size_t cols, const uint32_t *__restrict__ col, const uint32_t *__restrict__ row,
                                    size_t queries, size_t nchunks, size_t idx, F *__restrict__ reply) {
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < queries) reply[q * nchunks + idx] = T[(size_t)row[q] * cols + col[q]];
}

__global__ void any_nonzero_kernel(const F *__restrict__ v, size_t n, int *flag) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (!fzero(v[i])) { *flag = 1; return; }
}

// ---- T1 ------------------------------------------------------------------------------------------------
int tensorcode_dev(hb_ctx *ctx, const F *msg, size_t n, int trs, int lin, F *T, size_t nchunks, uint8_t *inner, InnerLayout lay) {
    if (trs <= 0 || n % trs) HB_FAIL(ctx, "tensorcode: n must be a multiple of tensor_row_size");
    size_t cols = 2 * n / trs;
    if (cols & (cols - 1)) HB_FAIL(ctx, "tensorcode: 2n/trs must be a power of two");
    HB_TRY(ntt_rows_padded_dev(ctx, msg, n / trs, T, cols, ilog2(cols), trs, nchunks, n, 4 * n));
    if (lin) {
        HB_TRY(encode_cols_dev(ctx, T, trs, cols, nchunks, 4 * n, inner, lay));
    } else {
        size_t rows = 2 * (size_t)trs;
        if (rows & (rows - 1)) HB_FAIL(ctx, "tensorcode: RS columns need a power-of-two tensor_row_size");
        for (size_t c = 0; c < nchunks; c++) HB_TRY(ntt_cols_dev(ctx, T + c * 4 * n, ilog2(rows), cols, trs));
        if (inner) HB_TRY(md_inner_standard_dev(ctx, T, rows, cols, nchunks, 4 * n, inner, lay));
    }
    return 0;
}

// ---- pageable host memory <-> HBM (declared in common.cuh) -------------------------------------------------------------------
static constexpr size_t kPinPiece = (size_t)8 << 20;
static void par_memcpy(void *dst, const void *src, size_t n) {
    static const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
    if (n < ((size_t)1 << 20) || hw == 1) { memcpy(dst, src, n); return; }
    const size_t per = ((n + hw - 1) / hw + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    for (size_t off = per; off < n; off += per)
        th.emplace_back([=] { memcpy((char *)dst + off, (const char *)src + off, std::min(per, n - off)); });
    memcpy(dst, src, std::min(per, n));
    for (auto &t : th) t.join();
}
static int pin_ready(hb_ctx *ctx) {
    for (int b = 0; b < 2; b++) {
        if (!ctx->pin[b]) HB_CHECK(ctx, cudaHostAlloc(&ctx->pin[b], kPinPiece, cudaHostAllocDefault));
        if (!ctx->pin_ev[b]) { HB_CHECK(ctx, cudaEventCreateWithFlags(&ctx->pin_ev[b], cudaEventDisableTiming)); HB_CHECK(ctx, cudaEventRecord(ctx->pin_ev[b], ctx->stream)); }
    }
    return 0;
}
int copy_from_host(hb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, cudaStream_t stream) {
    HB_TRY(pin_ready(ctx));
    size_t i = 0;
    for (size_t off = 0; off < bytes; off += kPinPiece, i++) {
        const int b = (int)(i & 1); const size_t sz = std::min(kPinPiece, bytes - off);
        HB_CHECK(ctx, cudaEventSynchronize(ctx->pin_ev[b]));                    // the last DMA that used this half has finished
        par_memcpy(ctx->pin[b], (const char *)src_host + off, sz);
        HB_CHECK(ctx, cudaMemcpyAsync((char *)dst_dev + off, ctx->pin[b], sz, cudaMemcpyHostToDevice, stream));
        HB_CHECK(ctx, cudaEventRecord(ctx->pin_ev[b], stream));
    }
    return 0;
}
int copy_to_host(hb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes, cudaStream_t stream) {
    HB_TRY(pin_ready(ctx));
    // a freshly allocated destination (the usual case: a std::vector level sized just before the call) is first touched here; with
    // transparent huge pages the 2 MiB-aligned interior faults 512 x fewer times (advisory, ignored where THP is off)
    if (bytes >= ((size_t)4 << 20)) {
        const uintptr_t lo = ((uintptr_t)dst_host + ((size_t)2 << 20) - 1) & ~(((uintptr_t)2 << 20) - 1), hi = ((uintptr_t)dst_host + bytes) & ~(((uintptr_t)2 << 20) - 1);
        if (hi > lo) madvise((void *)lo, hi - lo, MADV_HUGEPAGE);
    }
    const size_t npieces = (bytes + kPinPiece - 1) / kPinPiece;
    for (size_t i = 0; i <= npieces; i++) {
        if (i < npieces) {
            const int b = (int)(i & 1); const size_t off = i * kPinPiece, sz = std::min(kPinPiece, bytes - off);
            HB_CHECK(ctx, cudaEventSynchronize(ctx->pin_ev[b]));
            HB_CHECK(ctx, cudaMemcpyAsync(ctx->pin[b], (const char *)src_dev + off, sz, cudaMemcpyDeviceToHost, stream));
            HB_CHECK(ctx, cudaEventRecord(ctx->pin_ev[b], stream));
        }
        if (i >= 1) {                                                            // piece i-1 has landed: hand it to the caller while piece i is in flight
            const int b = (int)((i - 1) & 1); const size_t off = (i - 1) * kPinPiece, sz = std::min(kPinPiece, bytes - off);
            HB_CHECK(ctx, cudaEventSynchronize(ctx->pin_ev[b]));
            par_memcpy((char *)dst_host + off, ctx->pin[b], sz);
        }
    }
    return 0;
}

// ---- background level copier ---------------------------------------------------------------------------------------------------------
static constexpr size_t kLvlPiece = (size_t)8 << 20;
static void levels_worker(hb_ctx *ctx) {
    LevelsCopier *L = ctx->lvl;
    cudaSetDevice(ctx->device);
    for (;;) {
        LevelsJob job;
        {
            std::unique_lock<std::mutex> lk(L->mu);
            L->cv.wait(lk, [&] { return L->stop || !L->q.empty(); });
            if (L->q.empty()) return;                                   // stop requested and nothing left
            job = std::move(L->q.front()); L->q.pop_front(); L->busy = true;
        }
        cudaError_t e = cudaStreamWaitEvent(L->stream, job.ready, 0);
        // every level, piece by piece: D2H into one half of the pinned double buffer while the previous piece is copied out (first touch of
        // the caller's fresh pages happens here, off the proving thread)
        struct Piece { uint8_t *dst; const uint8_t *src; size_t n; };
        std::vector<Piece> pieces;
        for (size_t l = 0; l < job.dst.size(); l++)
            if (job.bytes[l] >= ((size_t)4 << 20)) {
                const uintptr_t lo = ((uintptr_t)job.dst[l] + ((size_t)2 << 20) - 1) & ~(((uintptr_t)2 << 20) - 1), hi = ((uintptr_t)job.dst[l] + job.bytes[l]) & ~(((uintptr_t)2 << 20) - 1);
                if (hi > lo) madvise((void *)lo, hi - lo, MADV_HUGEPAGE);
            }
        for (size_t l = 0; l < job.dst.size(); l++)
            for (size_t o = 0; o < job.bytes[l]; o += kLvlPiece) pieces.push_back({job.dst[l] + o, job.dev + job.off[l] + o, std::min(kLvlPiece, job.bytes[l] - o)});
        for (size_t i = 0; i <= pieces.size() && e == cudaSuccess; i++) {
            if (i < pieces.size()) {
                const int b = (int)(i & 1);
                e = cudaMemcpyAsync(L->pin[b], pieces[i].src, pieces[i].n, cudaMemcpyDeviceToHost, L->stream);
                if (e == cudaSuccess) e = cudaEventRecord(L->ev[b], L->stream);
            }
            if (i >= 1 && e == cudaSuccess) {
                const int b = (int)((i - 1) & 1);
                e = cudaEventSynchronize(L->ev[b]);
                if (e == cudaSuccess) par_memcpy(pieces[i - 1].dst, L->pin[b], pieces[i - 1].n);      // fresh pages: faulted in by several threads
            }
        }
        if (job.owned) cudaFreeAsync(job.dev, L->stream);
        cudaStreamSynchronize(L->stream);
        cudaEventDestroy(job.ready);
        {
            std::lock_guard<std::mutex> lk(L->mu);
            if (e != cudaSuccess && L->err.empty()) L->err = cudaGetErrorString(e);
            L->busy = false;
        }
        L->cv_idle.notify_all();
    }
}
static int levels_copier_start(hb_ctx *ctx) {
    if (!ctx->lvl) ctx->lvl = new LevelsCopier();
    LevelsCopier *L = ctx->lvl;
    if (L->started) return 0;
    HB_CHECK(ctx, cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
        HB_CHECK(ctx, cudaHostAlloc(&L->pin[b], kLvlPiece, cudaHostAllocDefault));
        HB_CHECK(ctx, cudaEventCreateWithFlags(&L->ev[b], cudaEventDisableTiming));
    }
    L->th = std::thread(levels_worker, ctx);
    L->started = true;
    return 0;
}
int levels_copy_wait(hb_ctx *ctx) {
    LevelsCopier *L = ctx->lvl;
    if (!L || !L->started) return 0;
    std::unique_lock<std::mutex> lk(L->mu);
    L->cv_idle.wait(lk, [&] { return L->q.empty() && !L->busy; });
    if (!L->err.empty()) { ctx->err = "background level copy: " + L->err; L->err.clear(); return 1; }
    return 0;
}
static void levels_copier_stop(hb_ctx *ctx) {
    LevelsCopier *L = ctx->lvl;
    if (!L) return;
    if (L->started) {
        { std::lock_guard<std::mutex> lk(L->mu); L->stop = true; }
        L->cv.notify_all();
        L->th.join();
        for (int b = 0; b < 2; b++) { cudaFreeHost(L->pin[b]); cudaEventDestroy(L->ev[b]); }
        cudaStreamDestroy(L->stream);
    }
    delete L; ctx->lvl = nullptr;
}
// queue: levels l = 0.. (n >> l digests each, flat in `dev` leaves first) -> level_ptrs[l] (NULL: skipped); the tree must be complete in stream order
static int levels_copy_enqueue(hb_ctx *ctx, uint8_t *dev, bool owned, uint8_t *const *level_ptrs, int nlevels, size_t nleaves) {
    HB_TRY(levels_copier_start(ctx));
    LevelsJob job; job.dev = dev; job.owned = owned;
    size_t off = 0, n = nleaves;
    for (int l = 0; l < nlevels; l++, n /= 2) {
        if (level_ptrs[l]) { job.dst.push_back(level_ptrs[l]); job.off.push_back(off * 32); job.bytes.push_back(n * 32); }
        off += n;
    }
    HB_CHECK(ctx, cudaEventCreateWithFlags(&job.ready, cudaEventDisableTiming));
    HB_CHECK(ctx, cudaEventRecord(job.ready, ctx->stream));
    { std::lock_guard<std::mutex> lk(ctx->lvl->mu); ctx->lvl->q.push_back(std::move(job)); }
    ctx->lvl->cv.notify_all();
    return 0;
}

}  // namespace hb

using namespace hb;

// =========================================================================================================
extern "C" int hb_ctx_create(hb_ctx **out, int device) {
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fprintf(stderr, "hobbit_b200: no CUDA device — this backend has no CPU fallback\n");
        return 100;
    }
    if (device < 0 || device >= ndev) return 101;
    if (cudaSetDevice(device) != cudaSuccess) return 102;
    hb_ctx *c = new (std::nothrow) hb_ctx();
    if (!c) return 103;
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return 104; }
    // keep freed stream-ordered allocations cached instead of returning them to the driver on every sync
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = c;
    return 0;
}

static int stager_init(hb_ctx *ctx, ChunkStager &st, size_t B) {
    for (int q = 0; q < 2; q++) {
        HB_CHECK(ctx, cudaMallocAsync(&st.buf[q], B * sizeof(F), ctx->stream));
        HB_CHECK(ctx, cudaEventCreateWithFlags(&st.copied[q], cudaEventDisableTiming));
        HB_CHECK(ctx, cudaEventCreateWithFlags(&st.consumed[q], cudaEventDisableTiming));
        st.used[q] = false;
    }
    st.count = 0;
    // the buffers are stream-ordered allocations of the compute stream: the copy stream may touch them only after this point
    HB_CHECK(ctx, cudaEventRecord(st.consumed[0], ctx->stream));
    HB_CHECK(ctx, cudaStreamWaitEvent(ctx->copy_stream, st.consumed[0], 0));
    return 0;
}
static void stager_free(hb_ctx *ctx, ChunkStager &st) {
    for (int q = 0; q < 2; q++) {
        if (st.buf[q]) { cudaStreamSynchronize(ctx->copy_stream); cudaFreeAsync(st.buf[q], ctx->stream); }
        if (st.copied[q]) cudaEventDestroy(st.copied[q]);
        if (st.consumed[q]) cudaEventDestroy(st.consumed[q]);
    }
    st = ChunkStager();
}
// host chunk -> device staging buffer; *src is valid for kernels launched on ctx->stream after this call; call stager_consumed after them
static int stager_push(hb_ctx *ctx, ChunkStager &st, const void *chunk, size_t B, const F **src, int *slot) {
    const int q = (int)(st.count++ & 1);
    if (st.used[q]) HB_CHECK(ctx, cudaStreamWaitEvent(ctx->copy_stream, st.consumed[q], 0));      // the encode of chunk i-2 read this buffer
    HB_CHECK(ctx, cudaMemcpyAsync(st.buf[q], chunk, B * sizeof(F), cudaMemcpyHostToDevice, ctx->copy_stream));
    HB_CHECK(ctx, cudaEventRecord(st.copied[q], ctx->copy_stream));
    // pageable memory has been staged by the driver when cudaMemcpyAsync returns; a pinned buffer is read by the DMA engine later
    if (is_pinned_host_ptr(chunk)) HB_CHECK(ctx, cudaEventSynchronize(st.copied[q]));
    HB_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, st.copied[q], 0));
    st.used[q] = true;
    *src = st.buf[q]; *slot = q;
    return 0;
}
static int stager_consumed(hb_ctx *ctx, ChunkStager &st, int slot) { HB_CHECK(ctx, cudaEventRecord(st.consumed[slot], ctx->stream)); return 0; }

static void elastic_free(hb_ctx *ctx) {
    ElasticState &el = ctx->el;
    if (el.park[0]) cudaFreeAsync(el.park[0], ctx->stream);
    stager_free(ctx, el.stg);
    if (el.leaves) cudaFreeAsync(el.leaves, ctx->stream);
    el = ElasticState();
}

extern "C" void hb_ctx_destroy(hb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &t : ctx->tw) if (t) cudaFree(t);
    for (auto &t : ctx->tw_pass) if (t) cudaFree(t);
    if (ctx->exp.d_stages) { cudaFree(ctx->exp.d_stages); cudaFree(ctx->exp.d_rowptr); cudaFree(ctx->exp.d_edges); }
    if (ctx->tensor) cudaFree(ctx->tensor);
    if (ctx->poly) cudaFree(ctx->poly);
    if (ctx->trace.tuples) cudaFree(ctx->trace.tuples);
    if (ctx->trace.pos) cudaFree(ctx->trace.pos);
    if (ctx->red) cudaFree(ctx->red);
    if (ctx->ticket) cudaFree(ctx->ticket);
    if (ctx->mailbox) cudaFreeHost(ctx->mailbox);
    levels_copier_stop(ctx);
    elastic_free(ctx);
    dist_release(ctx);
    for (auto &r : ctx->prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto &e : ctx->prof_pool) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    for (int b = 0; b < 2; b++) { if (ctx->pin[b]) cudaFreeHost(ctx->pin[b]); if (ctx->pin_ev[b]) cudaEventDestroy(ctx->pin_ev[b]); }
    cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

extern "C" const char *hb_last_error(hb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int hb_sync(hb_ctx *ctx) { HB_DEV(ctx);  HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream)); return 0; }
extern "C" uint64_t hb_launch_count(hb_ctx *ctx) { return ctx->launches; }
extern "C" uint64_t hb_transcript_digest(hb_ctx *ctx, int reset) {
    const uint64_t h = ctx->transcript;
    if (reset) ctx->transcript = 0xcbf29ce484222325ULL;
    return h;
}
extern "C" void *hb_stream(hb_ctx *ctx) { return (void *)ctx->stream; }
// ---- per-kernel timing -------------------------------------------------------------------------------------
extern "C" int hb_profile_enable(hb_ctx *ctx, int on) { HB_DEV(ctx);
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof_recs) { ctx->prof_pool.push_back(r.e0); ctx->prof_pool.push_back(r.e1); }
    ctx->prof_recs.clear();
    ctx->prof = on != 0;
    return 0;
}
// JSON: {"kernel": {"launches": n, "total_ms": t}, ...}; returns the number of bytes needed (incl. NUL)
extern "C" size_t hb_profile_report(hb_ctx *ctx, char *buf, size_t cap) {
    cudaStreamSynchronize(ctx->stream);
    struct Agg { std::string name; long n; double ms; };
    std::vector<Agg> agg;
    for (auto &r : ctx->prof_recs) {
        float ms = 0; cudaEventElapsedTime(&ms, r.e0, r.e1);
        std::string nm = r.name;
        size_t lt = nm.find('<'); if (nm[0] == '(') nm = nm.substr(1, nm.size() - 2);
        lt = nm.find('<'); if (lt != std::string::npos) nm = nm.substr(0, lt);
        size_t ns = nm.rfind("::"); if (ns != std::string::npos) nm = nm.substr(ns + 2);
        bool found = false;
        for (auto &a : agg) if (a.name == nm) { a.n++; a.ms += ms; found = true; break; }
        if (!found) agg.push_back({nm, 1, ms});
    }
    std::string js = "{";
    for (size_t i = 0; i < agg.size(); i++) {
        char tmp[256];
        snprintf(tmp, sizeof tmp, "%s\"%s\": {\"launches\": %ld, \"total_ms\": %.6f}", i ? ", " : "", agg[i].name.c_str(), agg[i].n, agg[i].ms);
        js += tmp;
    }
    js += "}";
    if (buf && cap) { size_t n = std::min(cap - 1, js.size()); memcpy(buf, js.data(), n); buf[n] = 0; }
    return js.size() + 1;
}

extern "C" int hb_malloc_device(hb_ctx *ctx, void **p, size_t bytes) { HB_DEV(ctx);  HB_CHECK(ctx, cudaMalloc(p, bytes)); return 0; }
extern "C" int hb_free_device(hb_ctx *ctx, void *p) { HB_DEV(ctx);  HB_CHECK(ctx, cudaFree(p)); return 0; }
// stream-ordered scratch from the context's pool (release threshold = never): no device synchronisation, microseconds per call
extern "C" int hb_malloc_stream(hb_ctx *ctx, void **p, size_t bytes) { HB_DEV(ctx);  HB_CHECK(ctx, cudaMallocAsync(p, bytes ? bytes : 16, ctx->stream)); return 0; }
extern "C" int hb_free_stream(hb_ctx *ctx, void *p) { HB_DEV(ctx);  if (p) HB_CHECK(ctx, cudaFreeAsync(p, ctx->stream)); return 0; }
extern "C" int hb_malloc_pinned(hb_ctx *ctx, void **p, size_t bytes) { HB_DEV(ctx);  HB_CHECK(ctx, cudaMallocHost(p, bytes)); return 0; }
extern "C" int hb_free_pinned(hb_ctx *ctx, void *p) { HB_DEV(ctx);  HB_CHECK(ctx, cudaFreeHost(p)); return 0; }
extern "C" int hb_memcpy(hb_ctx *ctx, void *dst, const void *src, size_t bytes) { HB_DEV(ctx);
    const bool dd = is_device_ptr(dst), sd = is_device_ptr(src);
    if (bytes >= kPageableDirect && dd && !sd && !is_pinned_host_ptr(src)) { HB_TRY(copy_from_host(ctx, dst, src, bytes, ctx->stream)); }
    else if (bytes >= kPageableDirect && sd && !dd && !is_pinned_host_ptr(dst)) { HB_TRY(copy_to_host(ctx, dst, src, bytes, ctx->stream)); }
    else HB_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    if (!(dd && sd)) HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));                                   // device-to-device stays stream-ordered
    return 0;
}

// ---- F1/F2 ---------------------------------------------------------------------------------------------
extern "C" int hb_field_binop(hb_ctx *ctx, int op, const hb_F *a, const hb_F *b, hb_F *c, size_t n) { HB_DEV(ctx);
    if (op < 0 || op > 4) HB_FAIL(ctx, "hb_field_binop: unknown op");
    if (n == 0) return 0;
    Staged sa(ctx), sb(ctx), sc(ctx);
    HB_TRY(sa.in(a, n * sizeof(F)));
    HB_TRY(sb.in(op < 3 ? b : a, n * sizeof(F)));
    HB_TRY(sc.outbuf(c, n * sizeof(F)));
    unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8);
    HB_LAUNCH(ctx, field_binop_kernel, grid, 256, 0, op, sa.as<F>(), sb.as<F>(), sc.as<F>(), n);
    HB_TRY(sc.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" void hb_root_of_unity(int logn, hb_F *out) {          // utils.cpp:452-463
    F rou = mkF(2147483648ULL, 1033321771269002680ULL);
    for (int i = 0; i < 62 - logn; i++) rou = h_fmul(rou, rou);
    out->real = rou.re; out->img = rou.im;
}

// MiMC sits on the critical path between sumcheck rounds (3-5 hashes per round, each 161 dependent cubings), so the host version is
// written for latency: (a + bi)^3 = a (a^2 - 3 b^2) + i b (3 a^2 - b^2) is FOUR 64x64 products per cubing (two squarings, two
// products) instead of the eight of two generic F multiplications, and limbs stay lazily reduced (< 2^62) until the very end.
// (One GPU thread needs 31 us per hash, tools/ubench_fmul.cu; this takes 2.3 us.)
static inline u64 mimc_red(unsigned __int128 x) {              // x < 2^124 + 2^72 -> congruent value < 2^61 + 8
    u64 s = (u64)(x >> 61) + ((u64)x & P61);                    // < 2^63 + 2^11 + 2^61: fits
    return (s & P61) + (s >> 61);
}
extern "C" void hb_mimc_hash(const hb_F *input, const hb_F *k, hb_F *out) {   // mimc.cpp:95-107, constants c_i = F(i) (:11-19)
    typedef unsigned __int128 u128;
    const u64 kr = k->real, ki = k->img;
    u64 a = input->real + kr, b = input->img + ki;              // t_0 = in + k            (limbs < 2^62 throughout)
    u64 hr = 0, hi = 0;
    for (int i = 0; i < 161; i++) {
        if (i) { a = hr + kr + (u64)(i - 1); b = hi + ki; }     // t_i = h + k + c_{i-1}
        const u64 s = mimc_red((u128)a * a), u = mimc_red((u128)b * b);          // a^2, b^2 < 2^61 + 4
        const u64 m = fold61(s + 4 * P61 - 3 * u);               // a^2 - 3 b^2   (3u < 4p)
        const u64 n = fold61(3 * s + 2 * P61 - u);               // 3 a^2 - b^2
        hr = mimc_red((u128)a * m); hi = mimc_red((u128)b * n);
    }
    out->real = canon61(fold61(hr + kr)); out->img = canon61(fold61(hi + ki));
}

// ---- T1 ------------------------------------------------------------------------------------------------
extern "C" int hb_tensorcode(hb_ctx *ctx, const hb_F *msg, size_t n, int trs, int linear_time, hb_F *tensor) { HB_DEV(ctx);
    Staged m(ctx), t(ctx);
    HB_TRY(m.in(msg, n * sizeof(F)));
    HB_TRY(t.outbuf(tensor, 4 * n * sizeof(F)));
    HB_TRY(tensorcode_dev(ctx, m.as<F>(), n, trs, linear_time, t.as<F>(), 1, nullptr));
    HB_TRY(t.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

// ---- C1 ------------------------------------------------------------------------------------------------
extern "C" int hb_commit_standard(hb_ctx *ctx, const hb_F *poly, size_t N, int K, int trs, int linear_time,
                                  uint8_t *levels_out, hb_F *tensor_out) { HB_DEV(ctx);
    if (K <= 0 || N % K) HB_FAIL(ctx, "hb_commit_standard: N must be a multiple of K");
    const size_t B = N / K;
    if (B & (B - 1)) HB_FAIL(ctx, "hb_commit_standard: N/K must be a power of two");
    if (trs < 2 || (2 * trs) % 4) HB_FAIL(ctx, "hb_commit_standard: tensor_row_size must be even");
    // resident tensor (K chunks x 4B) — open_standard gathers from it
    if (ctx->tensor_elems != 4 * N) {
        if (ctx->tensor) cudaFree(ctx->tensor);
        ctx->tensor = nullptr; ctx->tensor_elems = 0;
        HB_CHECK(ctx, cudaMalloc(&ctx->tensor, 4 * N * sizeof(F)));
        ctx->tensor_elems = 4 * N;
    }
    ctx->tensor_N = N; ctx->tensor_K = K; ctx->tensor_trs = trs;
    Staged lv(ctx);
    HB_TRY(lv.outbuf(levels_out, (2 * B - 1) * 32));
    HB_CHECK(ctx, cudaMemsetAsync(lv.dev, 0, B * 32, ctx->stream));        // chain starts from all-zero digests (Our_PC.cpp:153)

    const bool on_dev = is_device_ptr(poly);
    if (!on_dev) {
        // stage the polynomial chunk by chunk on the copy stream so the H2D of chunk i+1 overlaps the encode of chunk i;
        // the device copy is kept for hb_aggregate (open_standard reads the polynomial again)
        if (ctx->poly_elems != N) {
            if (ctx->poly) cudaFree(ctx->poly);
            ctx->poly = nullptr; ctx->poly_elems = 0;
            HB_CHECK(ctx, cudaMalloc(&ctx->poly, N * sizeof(F)));
            ctx->poly_elems = N;
        }
        ctx->poly_valid = false;                           // set again only once the whole polynomial has arrived (end of this call)
    }
    // Chunks are processed in groups of G through ONE launch per kernel (grid.y = chunk): no per-chunk grid tail.
    // The inner leaf digests of a group are staged in HBM (32 B per coefficient) and chained in chunk order afterwards.
    // Host input: small groups so the H2D of the next group overlaps this group's encode; resident input: up to 1 GiB of
    // inner digests per group.
    // host input: groups of >= 2^21 coefficients (32 MiB).  Measured at N = 2^26, K = 32 (e2e ms): 1 chunk per group 22.44, 2: 22.71, 4: 23.56,
    // 8: 24.98 — the H2D stream is the floor, a smaller group only shortens the tail after the last copy
    const int host_group = (int)std::max<size_t>(1, (((size_t)1 << 21) + B - 1) / B);
    int G = on_dev ? (int)std::max<size_t>(1, std::min<size_t>((size_t)K, ((size_t)1 << 30) / (B * 32))) : std::min(K, host_group);
    uint8_t *inner;
    HB_CHECK(ctx, cudaMallocAsync(&inner, (size_t)G * B * 32, ctx->stream));
    const int ngroups = (K + G - 1) / G;
    std::vector<cudaEvent_t> ev(on_dev ? 0 : ngroups);
    if (!on_dev) {
        cudaEvent_t start;
        HB_CHECK(ctx, cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
        HB_CHECK(ctx, cudaEventRecord(start, ctx->stream));
        HB_CHECK(ctx, cudaStreamWaitEvent(ctx->copy_stream, start, 0));    // do not overwrite ctx->poly under earlier work
        cudaEventDestroy(start);
        for (int g = 0; g < ngroups; g++) HB_CHECK(ctx, cudaEventCreateWithFlags(&ev[g], cudaEventDisableTiming));
    }
    // pinned polynomial: every group's DMA is queued up front; pageable (a std::vector): the host stages group g+1 through the pinned
    // double buffer while the GPU encodes group g (copy_from_host returns when the host memory has been read)
    const bool pageable = !on_dev && !is_pinned_host_ptr(poly);
    auto upload = [&](int g) -> int {
        size_t c0 = (size_t)g * G, nc = std::min<size_t>(G, K - c0);
        if (pageable) { HB_TRY(copy_from_host(ctx, ctx->poly + c0 * B, (const F *)poly + c0 * B, nc * B * sizeof(F), ctx->copy_stream)); }
        else HB_CHECK(ctx, cudaMemcpyAsync(ctx->poly + c0 * B, (const F *)poly + c0 * B, nc * B * sizeof(F), cudaMemcpyHostToDevice, ctx->copy_stream));
        HB_CHECK(ctx, cudaEventRecord(ev[g], ctx->copy_stream));
        return 0;
    };
    if (!on_dev && !pageable) for (int g = 0; g < ngroups; g++) HB_TRY(upload(g));
    const F *src = on_dev ? (const F *)poly : ctx->poly;
    for (int g = 0; g < ngroups; g++) {
        size_t c0 = (size_t)g * G, nc = std::min<size_t>(G, K - c0);
        if (pageable) HB_TRY(upload(g));
        if (!on_dev) HB_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ev[g], 0));
        HB_TRY(tensorcode_dev(ctx, src + c0 * B, B, trs, linear_time, ctx->tensor + c0 * 4 * B, nc, inner));
        HB_TRY(md_chain_dev(ctx, inner, nc, B, lv.as<uint8_t>()));
    }
    cudaFreeAsync(inner, ctx->stream);
    HB_TRY(merkle_tree_dev(ctx, lv.as<uint8_t>(), B));
    HB_TRY(lv.finish());
    if (tensor_out) HB_CHECK(ctx, cudaMemcpyAsync(tensor_out, ctx->tensor, 4 * N * sizeof(F), cudaMemcpyDefault, ctx->stream));
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &e : ev) cudaEventDestroy(e);
    if (!on_dev) ctx->poly_valid = true;                  // the whole polynomial is resident: hb_aggregate(poly = NULL) may use it
    return 0;
}

// ---- C1, sharded form: the chunk-independent part and the chunk-ordered chain as two calls -----------------------
// shared by hb_commit_encode_chunks and the fused multi-GPU commit (dist.cu): `inner_dev` device memory (or, with lay0.peer set, the base of
// this rank's own receive array); lay0.chunk0 = index of the first chunk of this call in the digest layout
int hb::commit_encode_chunks_impl(hb_ctx *ctx, const hb_F *poly, size_t nchunks, size_t B, int trs, int linear_time, uint8_t *inner_dev,
                                  InnerLayout lay0, size_t first_chunk, size_t total_chunks) {
    // the resident tensor covers ALL of this rank's chunks (total_chunks); this call fills [first_chunk, first_chunk + nchunks)
    const size_t Nt = total_chunks * B;
    if (ctx->tensor_elems != 4 * Nt) {
        if (ctx->tensor) cudaFree(ctx->tensor);
        ctx->tensor = nullptr; ctx->tensor_elems = 0;
        HB_CHECK(ctx, cudaMalloc(&ctx->tensor, 4 * Nt * sizeof(F)));
        ctx->tensor_elems = 4 * Nt;
    }
    ctx->tensor_N = Nt; ctx->tensor_K = (int)total_chunks; ctx->tensor_trs = trs;
    const size_t N = nchunks * B;
    const bool on_dev = is_device_ptr(poly);
    // groups bound the size of one launch (grid.y).  Host input: groups of 2 chunks, every group's H2D queued on the copy stream up
    // front so that the copy of group g+1 runs under the encode of group g (as in hb_commit).
    const size_t G = on_dev ? std::max<size_t>(1, std::min<size_t>(nchunks, ((size_t)1 << 30) / (B * 32))) : std::min<size_t>(nchunks, 2);
    const size_t ngroups = (nchunks + G - 1) / G;
    F *stage = nullptr;
    std::vector<cudaEvent_t> ev(on_dev ? 0 : ngroups);
    if (!on_dev) {
        const bool pinned = is_pinned_host_ptr(poly);
        HB_CHECK(ctx, cudaMallocAsync(&stage, N * sizeof(F), ctx->stream));
        cudaEvent_t start;
        HB_CHECK(ctx, cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
        HB_CHECK(ctx, cudaEventRecord(start, ctx->stream));                 // the staging buffer exists from here on
        HB_CHECK(ctx, cudaStreamWaitEvent(ctx->copy_stream, start, 0));
        cudaEventDestroy(start);
        for (size_t g = 0; g < ngroups; g++) {
            size_t c0 = g * G, nc = std::min(G, nchunks - c0);
            HB_CHECK(ctx, cudaEventCreateWithFlags(&ev[g], cudaEventDisableTiming));
            if (!pinned && nc * B * sizeof(F) >= kPageableDirect) { HB_TRY(copy_from_host(ctx, stage + c0 * B, (const F *)poly + c0 * B, nc * B * sizeof(F), ctx->copy_stream)); }
            else HB_CHECK(ctx, cudaMemcpyAsync(stage + c0 * B, (const F *)poly + c0 * B, nc * B * sizeof(F), cudaMemcpyHostToDevice, ctx->copy_stream));
            HB_CHECK(ctx, cudaEventRecord(ev[g], ctx->copy_stream));
        }
    }
    const F *src = on_dev ? (const F *)poly : stage;
    int rc = 0;
    for (size_t g = 0; g < ngroups && !rc; g++) {
        size_t c0 = g * G, nc = std::min(G, nchunks - c0);
        if (!on_dev) HB_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ev[g], 0));
        InnerLayout lay = lay0; lay.chunk0 = lay0.chunk0 + c0;
        rc = tensorcode_dev(ctx, src + c0 * B, B, trs, linear_time, ctx->tensor + (first_chunk + c0) * 4 * B, nc, inner_dev, lay);
    }
    if (!on_dev) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaFreeAsync(stage, ctx->stream);
        for (auto &e : ev) cudaEventDestroy(e);
    }
    return rc;
}

extern "C" int hb_commit_encode_chunks(hb_ctx *ctx, const hb_F *poly, size_t nchunks, size_t B, int trs, int linear_time, uint8_t *inner_out,
                                       size_t leaf_parts, size_t first_chunk, size_t total_chunks) { HB_DEV(ctx);
    if (nchunks == 0) return 0;
    if (B == 0 || (B & (B - 1))) HB_FAIL(ctx, "hb_commit_encode_chunks: chunk size must be a power of two");
    if (leaf_parts == 0) leaf_parts = 1;
    if (total_chunks == 0) { total_chunks = nchunks; first_chunk = 0; }
    if (B % leaf_parts || first_chunk + nchunks > total_chunks) HB_FAIL(ctx, "hb_commit_encode_chunks: bad leaf_parts / chunk range");
    Staged in(ctx);
    HB_TRY(in.outbuf(inner_out, nchunks * B * 32));
    InnerLayout lay; lay.part_leaves = B / leaf_parts; lay.chunks_total = nchunks; lay.chunk0 = 0;
    HB_TRY(commit_encode_chunks_impl(ctx, poly, nchunks, B, trs, linear_time, in.as<uint8_t>(), lay, first_chunk, total_chunks));
    HB_TRY(in.finish());
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Elastic_PC commit split for sharding: `ngroups` groups of 4 consecutive chunks -> inner digests of the 4B positions of each group
int hb::elastic_encode_groups_impl(hb_ctx *ctx, const hb_F *chunks, size_t ngroups, size_t B, int trs, int linear_time, uint8_t *inner_dev, InnerLayout lay0) {
    const size_t cells = 4 * B;
    // a launch covers up to 1 GiB of encoded tensors (64 B per coefficient)
    size_t G = std::max<size_t>(1, std::min<size_t>(ngroups, ((size_t)1 << 30) / (16 * B * sizeof(F))));
    if (const char *cap = getenv("HB_ELASTIC_GROUPS_PER_LAUNCH")) G = std::max<size_t>(1, std::min<size_t>(G, (size_t)atoll(cap)));   // tests: force several launches
    F *T4; HB_CHECK(ctx, cudaMallocAsync(&T4, G * 16 * B * sizeof(F), ctx->stream));
    const bool on_dev = is_device_ptr(chunks);
    // Host stream (pinned memory for full PCIe rate): the chunks of launch g+1 are copied on the copy stream into the other half of a
    // double buffer while launch g is encoded — the witness never has to be resident, 16 B per coefficient cross PCIe exactly once.
    F *stage[2] = {nullptr, nullptr}; cudaEvent_t copied[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
    if (!on_dev) {
        for (int q = 0; q < 2; q++) {
            HB_CHECK(ctx, cudaMallocAsync(&stage[q], G * 4 * B * sizeof(F), ctx->stream));
            HB_CHECK(ctx, cudaEventCreateWithFlags(&copied[q], cudaEventDisableTiming));
            HB_CHECK(ctx, cudaEventCreateWithFlags(&freed[q], cudaEventDisableTiming));
            HB_CHECK(ctx, cudaEventRecord(freed[q], ctx->stream));                       // the buffers exist (stream-ordered allocation) from here on
        }
        ctx->sync_needed = true;
    }
    auto prefetch = [&](size_t g0) -> int {
        const int q = (int)((g0 / G) & 1); const size_t ng = std::min(G, ngroups - g0);
        HB_CHECK(ctx, cudaStreamWaitEvent(ctx->copy_stream, freed[q], 0));
        HB_CHECK(ctx, cudaMemcpyAsync(stage[q], (const F *)chunks + g0 * 4 * B, ng * 4 * B * sizeof(F), cudaMemcpyHostToDevice, ctx->copy_stream));
        HB_CHECK(ctx, cudaEventRecord(copied[q], ctx->copy_stream));
        return 0;
    };
    if (!on_dev) HB_TRY(prefetch(0));
    for (size_t g0 = 0; g0 < ngroups; g0 += G) {
        size_t ng = std::min(G, ngroups - g0);
        const int q = (int)((g0 / G) & 1);
        const F *src = on_dev ? (const F *)chunks + g0 * 4 * B : stage[q];
        if (!on_dev) {
            HB_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, copied[q], 0));
            if (g0 + G < ngroups) HB_TRY(prefetch(g0 + G));                              // overlaps the encode below
        }
        int rc = tensorcode_dev(ctx, src, B, trs, linear_time, T4, 4 * ng, nullptr);
        if (!on_dev) HB_CHECK(ctx, cudaEventRecord(freed[q], ctx->stream));              // stage[q] may be overwritten once this encode has read it
        InnerLayout lay = lay0; lay.chunk0 = lay0.chunk0 + g0;
        if (!rc) rc = md_inner_stream4_dev(ctx, T4, cells, ng, inner_dev, lay);
        if (rc) { cudaFreeAsync(T4, ctx->stream); return rc; }
    }
    if (!on_dev) {
        HB_CHECK(ctx, cudaStreamSynchronize(ctx->copy_stream));
        for (int q = 0; q < 2; q++) { cudaFreeAsync(stage[q], ctx->stream); cudaEventDestroy(copied[q]); cudaEventDestroy(freed[q]); }
    }
    cudaFreeAsync(T4, ctx->stream);
    return 0;
}

extern "C" int hb_elastic_encode_groups(hb_ctx *ctx, const hb_F *chunks, size_t ngroups, size_t B, int trs, int linear_time, uint8_t *inner_out,
                                        size_t leaf_parts, size_t first_group, size_t total_groups) { HB_DEV(ctx);
    if (ngroups == 0) return 0;
    if (B == 0 || (B & (B - 1))) HB_FAIL(ctx, "hb_elastic_encode_groups: BUFFER_SPACE must be a power of two");
    if (leaf_parts == 0) leaf_parts = 1;
    if (total_groups == 0) { total_groups = ngroups; first_group = 0; }
    const size_t cells = 4 * B;
    if (cells % leaf_parts || first_group + ngroups > total_groups) HB_FAIL(ctx, "hb_elastic_encode_groups: bad leaf_parts / group range");
    Staged in(ctx);
    HB_TRY(in.outbuf(inner_out, ngroups * cells * 32));
    InnerLayout lay; lay.part_leaves = cells / leaf_parts; lay.chunks_total = ngroups; lay.chunk0 = 0;
    HB_TRY(elastic_encode_groups_impl(ctx, chunks, ngroups, B, trs, linear_time, in.as<uint8_t>(), lay));
    HB_TRY(in.finish());
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int hb_md_chain(hb_ctx *ctx, const uint8_t *inner, size_t nchunks, size_t nleaves, uint8_t *leaves) { HB_DEV(ctx);
    Staged in(ctx), lv(ctx);
    HB_TRY(in.in(inner, nchunks * nleaves * 32));
    HB_TRY(lv.outbuf(leaves, nleaves * 32, true));
    HB_TRY(md_chain_dev(ctx, in.as<uint8_t>(), nchunks, nleaves, lv.as<uint8_t>()));
    HB_TRY(lv.finish());
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" const hb_F *hb_tensor_device(hb_ctx *ctx) { return (const hb_F *)ctx->tensor; }

extern "C" int hb_tensor_gather(hb_ctx *ctx, const uint32_t *col, const uint32_t *row, size_t queries, hb_F *reply) { HB_DEV(ctx);
    if (!ctx->tensor) HB_FAIL(ctx, "hb_tensor_gather: no committed tensor in this context");
    if (queries == 0) return 0;
    const size_t B = ctx->tensor_N / ctx->tensor_K, cols = 2 * B / ctx->tensor_trs;
    Staged c(ctx), r(ctx), o(ctx);
    HB_TRY(c.in(col, queries * 4)); HB_TRY(r.in(row, queries * 4));
    HB_TRY(o.outbuf(reply, queries * ctx->tensor_K * sizeof(F)));
    size_t tot = queries * ctx->tensor_K;
    HB_LAUNCH(ctx, tensor_gather_kernel, (unsigned)((tot + 255) / 256), 256, 0, ctx->tensor, 4 * B, cols, ctx->tensor_K,
              c.as<uint32_t>(), r.as<uint32_t>(), queries, o.as<F>());
    HB_TRY(o.finish());
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int hb_aggregate(hb_ctx *ctx, const hb_F *poly, size_t N, int K, const hb_F *beta, hb_F *agg) { HB_DEV(ctx);
    if (K <= 0 || N % K) HB_FAIL(ctx, "hb_aggregate: N must be a multiple of K");
    const size_t B = N / K;
    Staged p(ctx), b(ctx), o(ctx);
    const F *src;
    if (poly == nullptr) {                                 // explicit request only: identity is never inferred from a host address
        if (!ctx->poly || !ctx->poly_valid || ctx->poly_elems != N) HB_FAIL(ctx, "hb_aggregate: no resident polynomial of this size (pass poly)");
        src = ctx->poly;                                   // device copy kept by the last successful hb_commit_standard from host memory
    } else {
        HB_TRY(p.in(poly, N * sizeof(F))); src = p.as<F>();
    }
    HB_TRY(b.in(beta, (size_t)K * sizeof(F)));
    HB_TRY(o.outbuf(agg, B * sizeof(F)));
    unsigned grid = (unsigned)std::min<size_t>((B + 255) / 256, (size_t)ctx->sm_count * 8);
    HB_LAUNCH(ctx, aggregate_kernel, grid, 256, (size_t)K * sizeof(F), src, B, K, b.as<F>(), o.as<F>());
    HB_TRY(o.finish());
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- C2 ------------------------------------------------------------------------------------------------
extern "C" int hb_elastic_begin(hb_ctx *ctx, size_t B, int trs, int linear_time) { HB_DEV(ctx);
    if (B == 0 || (B & (B - 1))) HB_FAIL(ctx, "hb_elastic_begin: BUFFER_SPACE must be a power of two");
    elastic_free(ctx);
    ElasticState &el = ctx->el;
    el.B = B; el.trs = trs; el.lin = linear_time; el.chunk_idx = 0;
    // stream-ordered pool allocations: a commit per call must not pay cudaMalloc/cudaFree (milliseconds each)
    // one allocation: park[0..2] and the current tensor are contiguous (a group of 4 encoded chunks, as md_inner_stream4 wants it)
    HB_CHECK(ctx, cudaMallocAsync(&el.park[0], 16 * B * sizeof(F), ctx->stream));
    el.park[1] = el.park[0] + 4 * B; el.park[2] = el.park[1] + 4 * B; el.tensor = el.park[2] + 4 * B;
    HB_TRY(stager_init(ctx, el.stg, B));
    HB_CHECK(ctx, cudaMallocAsync(&el.leaves, (8 * B - 1) * 32, ctx->stream));
    HB_CHECK(ctx, cudaMemsetAsync(el.leaves, 0, 4 * B * 32, ctx->stream));   // Elastic_PC.cpp:195-199
    el.active = true;
    return 0;
}

extern "C" int hb_elastic_push(hb_ctx *ctx, const hb_F *chunk) { HB_DEV(ctx);
    ElasticState &el = ctx->el;
    if (!el.active) HB_FAIL(ctx, "hb_elastic_push: call hb_elastic_begin first");
    const size_t B = el.B;
    const F *src = (const F *)chunk;
    int stg_slot = -1;
    if (!is_device_ptr(chunk)) HB_TRY(stager_push(ctx, el.stg, chunk, B, &src, &stg_slot));
    const unsigned slot = (unsigned)(el.chunk_idx % 4);
    F *T = slot == 3 ? el.tensor : el.park[slot];          // encode straight into the parking slot: no copy
    // The reference skips the encode of an all-zero chunk and zero-fills the tensor (Elastic_PC.cpp:206-222).  Both codes are
    // linear, so encoding the zero chunk yields exactly that all-zero tensor: no test, no host round trip, same bits.
    HB_TRY(tensorcode_dev(ctx, src, B, el.trs, el.lin, T, 1, nullptr));
    if (stg_slot >= 0) HB_TRY(stager_consumed(ctx, el.stg, stg_slot));
    if (slot == 3) {
        if (el.dist_groups_total) {                        // sharded: inner digests of this group -> the owner ranks' windows (NVLink stores)
            InnerLayout lay = el.dist_lay; lay.chunk0 += el.chunk_idx / 4;
            HB_TRY(md_inner_stream4_dev(ctx, el.park[0], 4 * B, 1, ctx->dist.win + kDistCtrlBytes, lay));
        } else HB_TRY(md_leaves_stream4_dev(ctx, el.park[0], el.park[1], el.park[2], el.tensor, 4 * B, el.leaves));
    }
    el.chunk_idx++;
    return 0;
}

// Streaming form of the sharded Elastic_PC commit: begin, then hb_elastic_push for this rank's 4 * groups_total / world chunks (its
// consecutive groups, in order), then hb_elastic_finish / hb_elastic_finish_levels as usual — every rank receives the whole tree.
extern "C" int hb_dist_elastic_begin(hb_ctx *ctx, size_t B, int trs, int linear_time, size_t groups_total) { HB_DEV(ctx);
    HB_TRY(hb_elastic_begin(ctx, B, trs, linear_time));
    if (ctx->dist.world <= 1) return 0;
    ElasticState &el = ctx->el;
    HB_TRY(sharded_begin(ctx, groups_total, 4 * B, &el.dist_lay));
    el.dist_groups_total = groups_total;
    return 0;
}

extern "C" int hb_elastic_finish(hb_ctx *ctx, uint8_t *levels_out) { HB_DEV(ctx);
    ElasticState &el = ctx->el;
    if (!el.active) HB_FAIL(ctx, "hb_elastic_finish: no commit in progress");
    if (el.dist_groups_total) {
        if (el.chunk_idx * (size_t)ctx->dist.world != 4 * el.dist_groups_total) HB_FAIL(ctx, "hb_elastic_finish: this rank must push exactly its 4 * groups_total / world chunks");
        int rc = sharded_finish(ctx, el.dist_groups_total, 4 * el.B, levels_out);
        elastic_free(ctx);
        return rc;
    }
    HB_TRY(merkle_tree_dev(ctx, el.leaves, 4 * el.B));
    HB_CHECK(ctx, cudaMemcpyAsync(levels_out, el.leaves, (8 * el.B - 1) * 32, cudaMemcpyDefault, ctx->stream));
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    elastic_free(ctx);
    return 0;
}

// The same, written level by level straight into the caller's per-level arrays (== MT_hashes[l].data(): no intermediate flat copy of the
// 8B digests on the host).  level_ptrs[l] receives 4B >> l digests, l = 0 .. log2(4B).
extern "C" int hb_elastic_finish_levels(hb_ctx *ctx, uint8_t *const *level_ptrs, int nlevels) { HB_DEV(ctx);
    ElasticState &el = ctx->el;
    if (!el.active) HB_FAIL(ctx, "hb_elastic_finish_levels: no commit in progress");
    if (nlevels != ilog2(4 * el.B) + 1) HB_FAIL(ctx, "hb_elastic_finish_levels: expected log2(4B)+1 levels");
    const uint8_t *tree = el.leaves;
    if (el.dist_groups_total) {
        if (el.chunk_idx * (size_t)ctx->dist.world != 4 * el.dist_groups_total) HB_FAIL(ctx, "hb_elastic_finish_levels: this rank must push exactly its 4 * groups_total / world chunks");
        HB_TRY(sharded_finish(ctx, el.dist_groups_total, 4 * el.B, nullptr));
        tree = sharded_tree(ctx, el.dist_groups_total, 4 * el.B);
    } else HB_TRY(merkle_tree_dev(ctx, el.leaves, 4 * el.B));
    size_t off = 0, n = 4 * el.B;
    for (int l = 0; l < nlevels; l++, n /= 2) {
        if (!level_ptrs[l]) { off += n; continue; }
        if (n * 32 >= kPageableDirect && !is_device_ptr(level_ptrs[l]) && !is_pinned_host_ptr(level_ptrs[l])) { HB_TRY(copy_to_host(ctx, level_ptrs[l], tree + off * 32, n * 32, ctx->stream)); }
        else HB_CHECK(ctx, cudaMemcpyAsync(level_ptrs[l], tree + off * 32, n * 32, cudaMemcpyDefault, ctx->stream));
        off += n;
    }
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    elastic_free(ctx);
    return 0;
}

// The same, but the copies into level_ptrs happen in the BACKGROUND (a worker thread with its own stream and pinned double buffer) while the
// context goes on proving: the call returns as soon as the tree kernels are queued.  The caller must not read or free the destination
// arrays before hb_levels_wait() returned (the host mirror's open() waits before it frees the tree; hobbit::commit_levels_async).
extern "C" int hb_elastic_finish_levels_async(hb_ctx *ctx, uint8_t *const *level_ptrs, int nlevels) { HB_DEV(ctx);
    ElasticState &el = ctx->el;
    if (!el.active) HB_FAIL(ctx, "hb_elastic_finish_levels_async: no commit in progress");
    if (nlevels != ilog2(4 * el.B) + 1) HB_FAIL(ctx, "hb_elastic_finish_levels_async: expected log2(4B)+1 levels");
    if (el.dist_groups_total) {
        // the global tree lives in the multi-GPU window, which the next sharded call overwrites: copy out of a private snapshot
        if (el.chunk_idx * (size_t)ctx->dist.world != 4 * el.dist_groups_total) HB_FAIL(ctx, "hb_elastic_finish_levels_async: this rank must push exactly its 4 * groups_total / world chunks");
        HB_TRY(sharded_finish(ctx, el.dist_groups_total, 4 * el.B, nullptr));
        uint8_t *snap; HB_CHECK(ctx, cudaMallocAsync(&snap, (8 * el.B - 1) * 32, ctx->stream));
        HB_CHECK(ctx, cudaMemcpyAsync(snap, sharded_tree(ctx, el.dist_groups_total, 4 * el.B), (8 * el.B - 1) * 32, cudaMemcpyDeviceToDevice, ctx->stream));
        HB_TRY(levels_copy_enqueue(ctx, snap, true, level_ptrs, nlevels, 4 * el.B));
    } else {
        HB_TRY(merkle_tree_dev(ctx, el.leaves, 4 * el.B));
        HB_TRY(levels_copy_enqueue(ctx, el.leaves, true, level_ptrs, nlevels, 4 * el.B));
        el.leaves = nullptr;                                   // ownership moved to the job
    }
    elastic_free(ctx);
    return 0;
}
// any flat tree in device memory (leaves first, nleaves >> l digests per level) -> level_ptrs in the background; take_ownership != 0: `dev_levels`
// (a stream-ordered allocation, hb_malloc_stream) is freed when the copy is done
extern "C" int hb_levels_copy_async(hb_ctx *ctx, uint8_t *dev_levels, int take_ownership, uint8_t *const *level_ptrs, int nlevels, size_t nleaves) { HB_DEV(ctx);
    if (!is_device_ptr(dev_levels)) HB_FAIL(ctx, "hb_levels_copy_async: the tree must be in device memory");
    return levels_copy_enqueue(ctx, dev_levels, take_ownership != 0, level_ptrs, nlevels, nleaves);
}
extern "C" int hb_levels_wait(hb_ctx *ctx) { HB_DEV(ctx); return levels_copy_wait(ctx); }

// W1: synthetic default stream of read_stream_PC (witness_stream.cpp:2405-2411).  The recurrence is inherently
// sequential (x <- x^2 + i), it is the INPUT GENERATOR of test_Elastic_PC, so it is evaluated once on the host.
extern "C" int hb_stream_pc_test(hb_ctx *ctx, hb_F *out, size_t n) { HB_DEV(ctx);
    std::vector<F> v(n);
    F x = mkF(322322, 0);
    for (size_t i = 0; i < n; i++) { v[i] = x; x = fadd(h_fmul(x, x), mkF((u64)i, 0)); }
    HB_CHECK(ctx, cudaMemcpyAsync(out, v.data(), n * sizeof(F), cudaMemcpyDefault, ctx->stream));
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- O2 (front half): Elastic_PC open = aggregate + compute_aggregation_reply (Elastic_PC.cpp:316-333, 487-533), chunk at a time ----
extern "C" int hb_elastic_open_begin(hb_ctx *ctx, size_t B, int trs, int linear_time, const uint32_t *col, const uint32_t *row, size_t queries, size_t nchunks) { HB_DEV(ctx);
    if (B == 0 || (B & (B - 1))) HB_FAIL(ctx, "hb_elastic_open_begin: BUFFER_SPACE must be a power of two");
    ElasticOpen &eo = ctx->eo;
    if (eo.active) { cudaFreeAsync(eo.buf, ctx->stream); stager_free(ctx, eo.stg); eo = ElasticOpen(); }
    eo.B = B; eo.trs = trs; eo.lin = linear_time; eo.queries = queries; eo.nchunks = nchunks; eo.idx = 0;
    size_t bytes = (B + 4 * B + queries * nchunks) * sizeof(F) + 2 * queries * sizeof(uint32_t);
    HB_CHECK(ctx, cudaMallocAsync(&eo.buf, bytes, ctx->stream));
    HB_TRY(stager_init(ctx, eo.stg, B));
    eo.agg = (F *)eo.buf; eo.tensor = eo.agg + B; eo.reply = eo.tensor + 4 * B;
    eo.col = (uint32_t *)(eo.reply + queries * nchunks); eo.row = eo.col + queries;
    HB_CHECK(ctx, cudaMemsetAsync(eo.agg, 0, B * sizeof(F), ctx->stream));
    HB_CHECK(ctx, cudaMemsetAsync(eo.reply, 0, queries * nchunks * sizeof(F), ctx->stream));
    HB_CHECK(ctx, cudaMemcpyAsync(eo.col, col, queries * sizeof(uint32_t), cudaMemcpyDefault, ctx->stream));
    HB_CHECK(ctx, cudaMemcpyAsync(eo.row, row, queries * sizeof(uint32_t), cudaMemcpyDefault, ctx->stream));
    eo.active = true;
    return 0;
}
// Multi-GPU: the next pushes are chunks [first, first + nchunks) of `total`; the reply array then has queries * total entries with
// reply[q * total + first + i] filled by this pass and zeros elsewhere, so a field all-reduce over the ranks assembles it.
extern "C" int hb_elastic_open_range(hb_ctx *ctx, size_t first, size_t total) { HB_DEV(ctx);
    ElasticOpen &eo = ctx->eo;
    if (!eo.active || eo.idx || first + eo.nchunks > total) HB_FAIL(ctx, "hb_elastic_open_range: call right after hb_elastic_open_begin with a valid range");
    if (eo.queries) {
        void *old = eo.buf;
        const size_t B = eo.B, bytes = (B + 4 * B + eo.queries * total) * sizeof(F) + 2 * eo.queries * sizeof(uint32_t);
        uint32_t *oc = eo.col;                                               // the queries were uploaded into the old buffer: move them
        void *nb; HB_CHECK(ctx, cudaMallocAsync(&nb, bytes, ctx->stream));
        F *agg = (F *)nb, *tensor = agg + B, *reply = tensor + 4 * B; uint32_t *col = (uint32_t *)(reply + eo.queries * total);
        HB_CHECK(ctx, cudaMemcpyAsync(col, oc, 2 * eo.queries * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
        HB_CHECK(ctx, cudaMemsetAsync(agg, 0, B * sizeof(F), ctx->stream));
        HB_CHECK(ctx, cudaMemsetAsync(reply, 0, eo.queries * total * sizeof(F), ctx->stream));
        cudaFreeAsync(old, ctx->stream);
        eo.buf = nb; eo.agg = agg; eo.tensor = tensor; eo.reply = reply; eo.col = col; eo.row = col + eo.queries;
    }
    eo.chunk_first = first; eo.chunk_total = total;
    return 0;
}
extern "C" int hb_elastic_open_push(hb_ctx *ctx, const hb_F *chunk, const hb_F *beta_i) { HB_DEV(ctx);
    ElasticOpen &eo = ctx->eo;
    if (!eo.active || eo.idx >= eo.nchunks) HB_FAIL(ctx, "hb_elastic_open_push: no open in progress / too many chunks");
    const size_t B = eo.B;
    const F *src = (const F *)chunk;
    int stg_slot = -1;
    if (!is_device_ptr(chunk)) HB_TRY(stager_push(ctx, eo.stg, chunk, B, &src, &stg_slot));
    F beta = mkF(beta_i->real, beta_i->img);
    unsigned grid = (unsigned)std::min<size_t>((B + 255) / 256, (size_t)ctx->sm_count * 8);
    HB_LAUNCH(ctx, axpy_kernel, grid, 256, 0, eo.agg, src, beta, B);
    HB_TRY(tensorcode_dev(ctx, src, B, eo.trs, eo.lin, eo.tensor, 1, nullptr));
    if (stg_slot >= 0) HB_TRY(stager_consumed(ctx, eo.stg, stg_slot));
    if (eo.queries) HB_LAUNCH(ctx, reply_gather_kernel, (unsigned)((eo.queries + 255) / 256), 256, 0, eo.tensor, 2 * B / eo.trs, eo.col, eo.row,
                              eo.queries, eo.chunk_total ? eo.chunk_total : eo.nchunks, eo.chunk_first + eo.idx, eo.reply);
    eo.idx++;
    return 0;
}
extern "C" int hb_elastic_open_finish(hb_ctx *ctx, hb_F *agg_out, hb_F *reply_out) { HB_DEV(ctx);
    ElasticOpen &eo = ctx->eo;
    if (!eo.active) HB_FAIL(ctx, "hb_elastic_open_finish: no open in progress");
    HB_CHECK(ctx, cudaMemcpyAsync(agg_out, eo.agg, eo.B * sizeof(F), cudaMemcpyDefault, ctx->stream));
    if (eo.queries) HB_CHECK(ctx, cudaMemcpyAsync(reply_out, eo.reply, eo.queries * (eo.chunk_total ? eo.chunk_total : eo.nchunks) * sizeof(F), cudaMemcpyDefault, ctx->stream));
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(eo.buf, ctx->stream);
    stager_free(ctx, eo.stg);
    eo = ElasticOpen();
    return 0;
}
