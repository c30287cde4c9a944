// TEST INFRASTRUCTURE ONLY — CPU emulation of (a subset of) the C ABI in include/hobbit_b200.h on top of the C oracle.
//
// Purpose: (1) lets the host-side logic of the C++ mirror (hobbit_b200/host/*.cpp: libc RNG order, Fiat–Shamir scalars, proof-size
// accounting, container marshalling) be tested against the unmodified reference on a machine WITHOUT a GPU (tests/cpp/open_test.cpp
// built against this library instead of libhobbit_b200.so); (2) serves as the plain restatement of the 8f.1 building blocks
// (hb_matvec_*, hb_phi_g_init, hb_change_form, hb_whir_*, hb_shockwave_leaves ...) that tests/test_gpu_open.py checks the CUDA
// kernels against.  Never linked into the product: libhobbit_host.so links libhobbit_b200.so, which fails without a CUDA device.
//
// Each function follows the reference lines cited next to its declaration in include/hobbit_b200.h.
#include "../include/hobbit_b200.h"
#include "hobbit_oracle.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <array>

typedef orc_F F;
struct hb_ctx {
    std::string err;
    std::vector<F> tensor; size_t tN = 0; int tK = 0, ttrs = 0;
    std::vector<F> poly; const void *poly_host = nullptr;
    long long exp_n = 0, exp_cw = 0;
    std::vector<unsigned char> trace; size_t tr_n = 0; bool tr_done = false;
    size_t eB = 0; int etrs = 0, elin = 0; std::vector<F> estream, ebeta; std::vector<uint32_t> ecol, erow;
};
static const uint64_t P = 2305843009213693951ULL;
static inline F fadd(F a, F b) { F c; orc_field_binop(0, &a, &b, &c, 1); return c; }
static inline F fsub(F a, F b) { F c; orc_field_binop(1, &a, &b, &c, 1); return c; }
static inline F fmul(F a, F b) { F c; orc_field_binop(2, &a, &b, &c, 1); return c; }
static inline F mk(uint64_t re, uint64_t im = 0) { F r; r.re = re; r.im = im; return r; }
static inline const F *cF(const hb_F *p) { return reinterpret_cast<const F *>(p); }
static inline F *mF(hb_F *p) { return reinterpret_cast<F *>(p); }
static int ilog2(size_t x) { int l = 0; while (x >>= 1) l++; return l; }
#define FAIL(ctx, msg) do { (ctx)->err = (msg); return 2; } while (0)

extern "C" {

int hb_ctx_create(hb_ctx **out, int) { *out = new hb_ctx; return 0; }
void hb_ctx_destroy(hb_ctx *ctx) { delete ctx; }
const char *hb_last_error(hb_ctx *ctx) { return ctx->err.c_str(); }
int hb_sync(hb_ctx *) { return 0; }
uint64_t hb_launch_count(hb_ctx *) { return 0; }
int hb_malloc_device(hb_ctx *, void **p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? 0 : 1; }
int hb_free_device(hb_ctx *, void *p) { free(p); return 0; }
int hb_malloc_stream(hb_ctx *, void **p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? 0 : 1; }
int hb_free_stream(hb_ctx *, void *p) { free(p); return 0; }
int hb_memcpy(hb_ctx *, void *dst, const void *src, size_t bytes) { memmove(dst, src, bytes); return 0; }

void hb_root_of_unity(int logn, hb_F *out) { orc_root_of_unity(logn, mF(out)); }
void hb_mimc_hash(const hb_F *input, const hb_F *k, hb_F *out) { orc_mimc_hash(cF(input), cF(k), mF(out)); }
int hb_field_binop(hb_ctx *, int op, const hb_F *a, const hb_F *b, hb_F *c, size_t n) { orc_field_binop(op, cF(a), cF(b), mF(c), n); return 0; }

int hb_ntt_batch(hb_ctx *, hb_F *data, int logn, size_t batch, size_t stride) {
    for (size_t r = 0; r < batch; r++) orc_fft(mF(data) + r * stride, logn);
    return 0;
}
int hb_expander_set(hb_ctx *ctx, long long n, int levels, int deg_C, int deg_D, const long long *L_C, const long long *R_C,
                    const uint32_t *const *nbr_C, const uint64_t *const *w_C, const long long *L_D, const long long *R_D,
                    const uint32_t *const *nbr_D, const uint64_t *const *w_D) {
    for (int d = 0; d < levels; d++) {
        orc_expander_install(0, d, L_C[d], R_C[d], deg_C, nbr_C[d], w_C[d]);
        orc_expander_install(1, d, L_D[d], R_D[d], deg_D, nbr_D[d], w_D[d]);
    }
    ctx->exp_n = n;
    std::vector<F> src(n), dst(2 * n);
    memset(src.data(), 0, n * sizeof(F)); memset(dst.data(), 0, 2 * n * sizeof(F));
    ctx->exp_cw = orc_encode_monolithic(src.data(), dst.data(), n);
    return 0;
}
long long hb_expander_codeword_len(hb_ctx *ctx) { return ctx->exp_cw; }

int hb_blake3_64(hb_ctx *, const uint8_t *src, uint8_t *dst, size_t count) { for (size_t i = 0; i < count; i++) orc_blake3_hash(src + 64 * i, dst + 32 * i); return 0; }
int hb_merkle_tree(hb_ctx *, uint8_t *levels, size_t nleaves) {
    std::vector<uint8_t> lv(nleaves * 32); memcpy(lv.data(), levels, nleaves * 32);
    orc_create_tree_blake(lv.data(), (int)nleaves, levels);
    return 0;
}
int hb_mt_commit(hb_ctx *, const hb_F *leafs, size_t N, uint8_t *levels) { orc_mt_commit_blake(cF(leafs), (int)N, levels); return 0; }
int hb_tensorcode(hb_ctx *, const hb_F *msg, size_t n, int trs, int lin, hb_F *tensor) { orc_compute_tensorcode(cF(msg), n, trs, lin, mF(tensor)); return 0; }
int hb_commit_standard(hb_ctx *ctx, const hb_F *poly, size_t N, int K, int trs, int lin, uint8_t *levels_out, hb_F *tensor_out) {
    ctx->tensor.resize(4 * N); ctx->tN = N; ctx->tK = K; ctx->ttrs = trs;
    orc_commit_standard(cF(poly), N, K, trs, lin, levels_out, ctx->tensor.data());
    if (tensor_out) memcpy(tensor_out, ctx->tensor.data(), 4 * N * sizeof(F));
    ctx->poly.assign(cF(poly), cF(poly) + N); ctx->poly_host = poly;
    return 0;
}
const hb_F *hb_tensor_device(hb_ctx *ctx) { return ctx->tensor.empty() ? nullptr : reinterpret_cast<const hb_F *>(ctx->tensor.data()); }
int hb_tensor_gather(hb_ctx *ctx, const uint32_t *col, const uint32_t *row, size_t queries, hb_F *reply) {
    if (ctx->tensor.empty()) FAIL(ctx, "hb_tensor_gather: no committed tensor in this context");
    size_t B = ctx->tN / ctx->tK, cols = 2 * B / ctx->ttrs;
    for (size_t q = 0; q < queries; q++)
        for (int i = 0; i < ctx->tK; i++) mF(reply)[q * ctx->tK + i] = ctx->tensor[(size_t)i * 4 * B + (size_t)row[q] * cols + col[q]];
    return 0;
}
int hb_aggregate(hb_ctx *ctx, const hb_F *poly, size_t N, int K, const hb_F *beta, hb_F *agg) {
    const F *src = poly ? cF(poly) : ctx->poly.data();
    size_t B = N / K;
    for (size_t j = 0; j < B; j++) {
        F a = mk(0);
        for (int i = 0; i < K; i++) a = fadd(a, fmul(cF(beta)[i], src[(size_t)i * B + j]));
        mF(agg)[j] = a;
    }
    return 0;
}
int hb_stream_pc_test(hb_ctx *, hb_F *out, size_t n) { orc_read_stream_pc_test(mF(out), n); return 0; }

int hb_precompute_beta(hb_ctx *, const hb_F *r, int nr, hb_F *out) { orc_precompute_beta(cF(r), nr, mF(out)); return 0; }
int hb_evaluate_vector(hb_ctx *, const hb_F *v, size_t n, const hb_F *r, hb_F *out) { orc_evaluate_vector(cF(v), n, cF(r), ilog2(n), mF(out)); return 0; }
int hb_sumcheck2(hb_ctx *, const hb_F *v1, const hb_F *v2, size_t n, const hb_F *prev_r, hb_F *proof, double *ps) {
    *ps += orc_sumcheck2(cF(v1), cF(v2), n, cF(prev_r), mF(proof)); return 0;
}

int hb_sumcheck3(hb_ctx *, const hb_F *v1, const hb_F *v2, const hb_F *v3, size_t n, const hb_F *prev_r, hb_F *proof, double *ps) {
    *ps += orc_sumcheck3(cF(v1), cF(v2), cF(v3), n, cF(prev_r), mF(proof)); return 0;
}
int hb_batch_sumcheck3(hb_ctx *, const hb_F *t1, const hb_F *t2, const hb_F *t3, const size_t *sizes, int batches, const hb_F *a, hb_F *proof, double *ps) {
    *ps += orc_batch_sumcheck3(cF(t1), cF(t2), cF(t3), sizes, batches, cF(a), mF(proof)); return 0;
}
/* The host mirror draws the libc values in the reference's order and hands them over; the oracle's provers replay them (injection). */
int hb_mul_tree(hb_ctx *, const hb_F *input, int vectors, size_t n, const hb_F *prev_r, const hb_F *x_rand, hb_F *out, size_t *written, int *nfr, double *ps) {
    orc_inject_randomness(cF(x_rand), vectors > 1 ? (size_t)ilog2((size_t)vectors) : 0);
    double p = 0; *written = orc_mul_tree(cF(input), vectors, n, cF(prev_r), mF(out), nfr, &p); *ps += p;
    orc_inject_randomness(nullptr, 0);
    return 0;
}
int hb_mul_tree_stream(hb_ctx *, const hb_F *xy, size_t total, int vectors, size_t B, int distance, int naive, const hb_F *prev_r, const hb_F *x_rand,
                       const hb_F *rnd, hb_F *out, int *layers_out, double *ps) {
    int layers = 0;
    if (total > 2 * B) { layers = ilog2(total / (2 * B)); if (layers % distance != 0 && layers > distance) layers = distance + layers - (layers % distance); }
    std::vector<F> q(cF(x_rand), cF(x_rand) + ilog2((size_t)vectors));
    size_t nrnd = 4 * (size_t)layers;                                        /* per streamed layer: a, b0, b1, pad */
    if (layers > distance && !naive) { int batches = layers / distance; nrnd = (size_t)(layers - distance) + (size_t)distance * (3 * batches + 1); }
    q.insert(q.end(), cF(rnd), cF(rnd) + nrnd);
    orc_inject_randomness(q.data(), q.size());
    *ps += orc_mul_tree_stream(cF(xy), total, vectors, B, distance, naive, cF(prev_r), mF(out));
    orc_inject_randomness(nullptr, 0);
    if (layers_out) *layers_out = layers;
    return 0;
}
int hb_gate_consistency_stream(hb_ctx *, const hb_F *L, const hb_F *R, const hb_F *O, const hb_F *S, size_t cs, size_t B, const hb_F *r, const hb_F *rnd10,
                               hb_F *out, double *ps) {
    orc_inject_randomness(cF(rnd10), 10);
    *ps += orc_gate_consistency_stream(cF(L), cF(R), cF(O), cF(S), cs, B, cF(r), mF(out));
    orc_inject_randomness(nullptr, 0);
    return 0;
}
/* Elastic_PC commit / open front: chunks are buffered and handed to the oracle's restatement at finish */
int hb_elastic_begin(hb_ctx *ctx, size_t B, int trs, int lin) { ctx->eB = B; ctx->etrs = trs; ctx->elin = lin; ctx->estream.clear(); return 0; }
int hb_elastic_push(hb_ctx *ctx, const hb_F *chunk) { ctx->estream.insert(ctx->estream.end(), cF(chunk), cF(chunk) + ctx->eB); return 0; }
int hb_elastic_finish(hb_ctx *ctx, uint8_t *levels_out) {
    orc_elastic_commit_stream(ctx->estream.data(), ctx->estream.size(), ctx->eB, ctx->etrs, ctx->elin, levels_out);
    ctx->estream.clear(); return 0;
}
int hb_elastic_finish_levels(hb_ctx *ctx, uint8_t *const *level_ptrs, int nlevels) {
    std::vector<uint8_t> flat((8 * ctx->eB - 1) * 32);
    orc_elastic_commit_stream(ctx->estream.data(), ctx->estream.size(), ctx->eB, ctx->etrs, ctx->elin, flat.data());
    size_t off = 0, n = 4 * ctx->eB;
    for (int l = 0; l < nlevels; l++, n /= 2) { if (level_ptrs[l]) memcpy(level_ptrs[l], flat.data() + off * 32, n * 32); off += n; }
    ctx->estream.clear(); return 0;
}
int hb_elastic_open_begin(hb_ctx *ctx, size_t B, int trs, int lin, const uint32_t *col, const uint32_t *row, size_t queries, size_t nchunks) {
    ctx->eB = B; ctx->etrs = trs; ctx->elin = lin; ctx->estream.clear(); ctx->ebeta.clear();
    ctx->ecol.assign(col, col + queries); ctx->erow.assign(row, row + queries); (void)nchunks; return 0;
}
int hb_elastic_open_push(hb_ctx *ctx, const hb_F *chunk, const hb_F *beta_i) {
    ctx->estream.insert(ctx->estream.end(), cF(chunk), cF(chunk) + ctx->eB); ctx->ebeta.push_back(*cF(beta_i)); return 0;
}
int hb_elastic_open_finish(hb_ctx *ctx, hb_F *agg_out, hb_F *reply_out) {
    orc_elastic_open_front(ctx->estream.data(), ctx->ebeta.size(), ctx->eB, ctx->etrs, ctx->elin, ctx->ebeta.data(), ctx->ecol.data(),
                           ctx->erow.data(), ctx->ecol.size(), mF(agg_out), mF(reply_out));
    ctx->estream.clear(); return 0;
}

/* ---- 8f.1 building blocks: plain restatements ------------------------------------------------------------------------------------ */
int hb_vec_zero(hb_ctx *, hb_F *v, size_t n) { memset(v, 0, n * sizeof(F)); return 0; }
int hb_rs_encode_rows(hb_ctx *, const hb_F *src, size_t in_len, size_t rows, hb_F *dst, int logn) {
    size_t len = (size_t)1 << logn;
    for (size_t r = 0; r < rows; r++) {
        F *d = mF(dst) + r * len;
        memset(d, 0, len * sizeof(F)); memcpy(d, cF(src) + r * in_len, in_len * sizeof(F));
        orc_fft(d, logn);
    }
    return 0;
}
int hb_matvec_cols(hb_ctx *, const hb_F *M, size_t rows, size_t cols, size_t stride, const hb_F *w, hb_F *out) {
    for (size_t j = 0; j < cols; j++) {
        F a = mk(0);
        for (size_t i = 0; i < rows; i++) a = fadd(a, fmul(cF(w)[i], cF(M)[i * stride + j]));
        mF(out)[j] = a;
    }
    return 0;
}
int hb_matvec_rows(hb_ctx *, const hb_F *M, size_t rows, size_t cols, size_t stride, const hb_F *s, hb_F *out) {
    for (size_t i = 0; i < rows; i++) {
        F a = mk(0);
        for (size_t j = 0; j < cols; j++) a = fadd(a, fmul(cF(s)[j], cF(M)[i * stride + j]));
        mF(out)[i] = a;
    }
    return 0;
}
int hb_axpy(hb_ctx *, hb_F *y, const hb_F *x, const hb_F *a, size_t n) {
    for (size_t i = 0; i < n; i++) mF(y)[i] = fadd(mF(y)[i], fmul(*cF(a), cF(x)[i]));
    return 0;
}
int hb_scatter(hb_ctx *, hb_F *out, size_t n, const uint64_t *idx, const hb_F *val, size_t m) {
    memset(out, 0, n * sizeof(F));
    for (size_t k = 0; k < m; k++) out[idx[k]] = val[k];
    return 0;
}
int hb_gather_cols(hb_ctx *, const hb_F *M, size_t rows, size_t, size_t stride, const uint64_t *col, size_t m, hb_F *out) {
    for (size_t q = 0; q < m; q++) for (size_t j = 0; j < rows; j++) out[q * rows + j] = M[j * stride + col[q]];
    return 0;
}
int hb_select_cols(hb_ctx *, const hb_F *M, size_t rows, size_t, size_t stride, const uint64_t *col, size_t m, hb_F *out) {
    for (size_t j = 0; j < rows; j++) for (size_t q = 0; q < m; q++) out[j * m + q] = M[j * stride + col[q]];
    return 0;
}
int hb_transpose(hb_ctx *, const hb_F *in, size_t rows, size_t cols, hb_F *out) {
    for (size_t i = 0; i < rows; i++) for (size_t j = 0; j < cols; j++) out[j * rows + i] = in[i * cols + j];
    return 0;
}
int hb_any_nonzero(hb_ctx *, const hb_F *v, size_t n, int *out) {
    *out = 0;
    for (size_t i = 0; i < n; i++) if (v[i].real | v[i].img) { *out = 1; break; }
    return 0;
}
int hb_encode_batch(hb_ctx *, const hb_F *src, hb_F *dst, long long n, size_t ncols) {
    std::vector<F> m(n), cw(2 * n);
    for (size_t c = 0; c < ncols; c++) {
        for (long long j = 0; j < n; j++) m[j] = cF(src)[(size_t)j * ncols + c];
        memset(cw.data(), 0, 2 * n * sizeof(F));
        orc_encode_monolithic(m.data(), cw.data(), n);
        for (long long j = 0; j < 2 * n; j++) mF(dst)[(size_t)j * ncols + c] = cw[j];
    }
    return 0;
}
/* utils.cpp:677-755 (forward transform), loop for loop */
int hb_phi_g_init(hb_ctx *, const hb_F *r, int n, hb_F *out) {
    size_t N = (size_t)1 << n;
    std::vector<F> pm(N);
    F rou; orc_root_of_unity(n, &rou);
    pm[0] = mk(1);
    for (size_t i = 1; i < N; i++) pm[i] = fmul(pm[i - 1], rou);
    F *g = mF(out); const F *rx = cF(r);
    memset(g, 0, N * sizeof(F));
    g[0] = mk(1);
    for (int i = 1; i < n; i++)
        for (size_t b = 0; b < ((size_t)1 << (i - 1)); b++) {
            size_t l = b, rr = b ^ ((size_t)1 << (i - 1)); int m = n - i;
            F t1 = fsub(mk(1), rx[m]), t2 = fmul(rx[m], pm[b << m]);
            g[rr] = fmul(g[l], fsub(t1, t2));
            g[l] = fmul(g[l], fadd(t1, t2));
        }
    for (size_t b = 0; b < ((size_t)1 << (n - 1)); b++) {
        F t1 = fsub(mk(1), rx[0]), t2 = fmul(rx[0], pm[b]);
        g[b] = fmul(g[b], fadd(t1, t2));
    }
    return 0;
}
/* Virgo.cpp:140-152: the full per-column MT_commit_Blake, root taken */
int hb_shockwave_leaves(hb_ctx *, const hb_F *enc, int k, size_t cols, uint8_t *leaves) {
    std::vector<F> buf(k); std::vector<uint8_t> lv((size_t)(2 * (k / 4) - 1) * 32);
    for (size_t c = 0; c < cols; c++) {
        for (int j = 0; j < k; j++) buf[j] = cF(enc)[(size_t)j * cols + c];
        orc_mt_commit_blake(buf.data(), k, lv.data());
        memcpy(leaves + c * 32, lv.data() + (size_t)(2 * (k / 4) - 2) * 32, 32);
    }
    return 0;
}
/* Virgo.cpp:104-118, recursively as written */
static void change_form_rec(F *poly, int logn, int l, size_t pos) {
    size_t S = (size_t)1 << (logn - l);
    std::vector<F> buff(S);
    for (size_t i = 0; i < S / 2; i++) { buff[i] = poly[pos + 2 * i]; buff[i + S / 2] = fsub(poly[pos + 2 * i + 1], poly[pos + 2 * i]); }
    memcpy(poly + pos, buff.data(), S * sizeof(F));
    if (l + 1 == logn) return;
    change_form_rec(poly, logn, l + 1, pos);
    change_form_rec(poly, logn, l + 1, pos + S / 2);
}
int hb_change_form(hb_ctx *, hb_F *poly, int logn) { change_form_rec(mF(poly), logn, 0, 0); return 0; }
int hb_regroup(hb_ctx *, const hb_F *in, size_t n, int k, hb_F *out) {
    size_t K = (size_t)1 << k, c = 0;
    for (size_t j = 0; j < n / K; j++) for (size_t t = 0; t < K; t++) out[c++] = in[j + t * (n / K)];
    return 0;
}
int hb_whir_poly(hb_ctx *, const hb_F *poly, const hb_F *beta, size_t L, hb_F *coeffs3) {
    F a = mk(0), b = mk(0), c = mk(0); const F *p = cF(poly), *q = cF(beta);
    for (size_t j = 0; j < L; j++) {
        F d1 = fsub(p[j + L], p[j]), d2 = fsub(q[j + L], q[j]);
        a = fadd(a, fmul(d1, d2)); b = fadd(b, fadd(fmul(d1, q[j]), fmul(d2, p[j]))); c = fadd(c, fmul(p[j], q[j]));
    }
    mF(coeffs3)[0] = a; mF(coeffs3)[1] = b; mF(coeffs3)[2] = c;
    return 0;
}
int hb_whir_fold(hb_ctx *, hb_F *poly, hb_F *beta, size_t L, const hb_F *a) {
    F *p = mF(poly), *q = mF(beta);
    for (size_t j = 0; j < L; j++) { p[j] = fadd(p[j], fmul(*cF(a), fsub(p[j + L], p[j]))); q[j] = fadd(q[j], fmul(*cF(a), fsub(q[j + L], q[j]))); }
    return 0;
}
int hb_whir_zeta(hb_ctx *, const hb_F *poly, hb_F *beta, int v, const hb_F *zetas, int repeats, const hb_F *pows, hb_F *y) {
    size_t n = (size_t)1 << v; std::vector<F> eq(n);
    for (int i = 0; i < repeats; i++) {
        orc_precompute_beta(cF(zetas) + (size_t)i * v, v, eq.data());
        F acc = mk(0);
        for (size_t j = 0; j < n; j++) acc = fadd(acc, fmul(eq[j], cF(poly)[j]));
        mF(y)[i] = acc;
        for (size_t j = 0; j < n; j++) mF(beta)[j] = fadd(mF(beta)[j], fmul(cF(pows)[i], eq[j]));
    }
    return 0;
}

/* ---- W1/W2: the named circuit streams as the reference's readers emit them front to back (witness_stream.cpp:768-874, 1055-1338,
 * 1620-1807, 2276-2311), restated sequentially over one pass of the trace */
struct emu_tuple { F value_o, value_l, value_r; int idx_o, idx_l, idx_r; int access_o, access_l, access_r; uint8_t type; };
static_assert(sizeof(emu_tuple) == 80, "tr_tuple layout");
static F f_int(long long x) { return mk(x >= 0 ? (uint64_t)x : P - (uint64_t)(-x)); }
int hb_trace_begin(hb_ctx *ctx, size_t) { ctx->trace.clear(); ctx->tr_n = 0; ctx->tr_done = false; return 0; }
int hb_trace_push(hb_ctx *ctx, const void *records, size_t n, int *done) {
    const emu_tuple *t = (const emu_tuple *)records;
    for (size_t i = 0; i < n && !ctx->tr_done; i++) {
        if (t[i].type == 255) { ctx->tr_done = true; break; }
        ctx->trace.insert(ctx->trace.end(), (const unsigned char *)&t[i], (const unsigned char *)&t[i] + 80); ctx->tr_n++;
    }
    if (done) *done = ctx->tr_done ? 1 : 0;
    return 0;
}
/* Seval.cpp:20-24, 97-170, 1238-1286, 1462-1489 restated gate by gate */
namespace {
struct emu_gt { F value; int idx; int access; };
struct MlpEval {
    hb_ctx *ctx; int labels = 1;
    void emit(const emu_tuple &t) { ctx->trace.insert(ctx->trace.end(), (const unsigned char *)&t, (const unsigned char *)&t + 80); ctx->tr_n++; }
    void init(emu_gt &g, F v) { g.value = v; g.idx = labels++; g.access = 0; }
    void del(emu_gt &g) { emu_tuple t; memset(&t, 0, sizeof t); t.type = 0; t.idx_o = g.idx; t.value_o = g.value; t.access_o = g.access; emit(t); }
    emu_gt op(emu_gt &g1, emu_gt &g2, int type) {
        emu_gt g; g.value = type == 1 ? fadd(g1.value, g2.value) : fmul(g1.value, g2.value); g.idx = labels++;
        emu_tuple t; memset(&t, 0, sizeof t);
        t.type = (uint8_t)type; t.idx_o = g.idx; t.value_o = g.value; t.idx_l = g1.idx; t.value_l = g1.value; t.idx_r = g2.idx; t.value_r = g2.value;
        t.access_l = g1.access; g1.access++; t.access_r = g2.access; t.access_o = 0; g2.access++;
        g.access = 1; emit(t);
        return g;
    }
};
}
int hb_trace_generate_mlp(hb_ctx *ctx, const int *layer_size, int nsizes, size_t *n_records) {
    ctx->trace.clear(); ctx->tr_n = 0;
    MlpEval E; E.ctx = ctx;
    std::vector<emu_gt> input(layer_size[0]);
    std::vector<std::vector<std::vector<emu_gt>>> W(nsizes - 1);
    for (int i = 0; i + 1 < nsizes; i++) {
        W[i].assign(layer_size[i + 1], std::vector<emu_gt>(layer_size[i]));
        for (int j = 0; j < layer_size[i + 1]; j++) for (int k = 0; k < layer_size[i]; k++) E.init(W[i][j][k], mk((uint64_t)((j + i + k) % 256)));
    }
    for (int k = 0; k < layer_size[0]; k++) E.init(input[k], mk((uint64_t)((k + 1) % 256)));
    emu_gt zero; E.init(zero, mk(0));
    std::vector<emu_gt> inp = input;
    for (size_t i = 0; i < W.size(); i++) {
        std::vector<emu_gt> hidden(W[i].size());
        for (size_t j = 0; j < W[i].size(); j++)
            for (size_t k = 0; k < W[i][0].size(); k++) {
                if (k == 0) hidden[j] = E.op(W[i][j][k], inp[k], 2);
                else { emu_gt temp = E.op(W[i][j][k], inp[k], 2); emu_gt ts = E.op(hidden[j], temp, 1); E.del(temp); E.del(hidden[j]); hidden[j] = ts; }
            }
        for (auto &g : inp) E.del(g);
        inp = hidden;
    }
    for (auto &Wi : W) for (auto &Wj : Wi) for (auto &g : Wj) E.del(g);
    for (auto &g : inp) E.del(g);
    E.del(zero);
    ctx->tr_done = true;
    if (n_records) *n_records = ctx->tr_n;
    return 0;
}
/* Seval.cpp:324-354 xor_gate, :957-989 lookup_box, :991-1084 encrypt / AES, :1353-1396 the fun == 5 driver, restated gate by gate */
namespace {
struct AesEval : MlpEval {
    emu_gt table_gate(emu_gt &l, emu_gt &r, int table, uint64_t value) {       /* a lookup record: type = table + 3 */
        emu_gt g; init(g, mk(value));
        emu_tuple t; memset(&t, 0, sizeof t);
        t.type = (uint8_t)(table + 3); t.idx_o = g.idx; t.value_o = g.value; t.idx_l = l.idx; t.value_l = l.value; t.idx_r = r.idx; t.value_r = r.value;
        t.access_l = l.access; t.access_r = r.access; t.access_o = 0; l.access++; r.access++;
        g.access = 1; emit(t);
        return g;
    }
    emu_gt x(emu_gt &a, emu_gt &b) { return table_gate(a, b, 1, a.value.re ^ b.value.re); }
    emu_gt box(emu_gt &a, emu_gt &zero, int table) {
        const uint64_t shift = table == 2 ? 21 : (uint64_t)table;                     /* S-box (x+21)%256, mix boxes (x+idx)%256 */
        return table_gate(a, zero, table, (a.value.re + shift) % 256);
    }
};
}
int hb_trace_generate_aes(hb_ctx *ctx, int input_size, size_t *n_records) {
    ctx->trace.clear(); ctx->tr_n = 0;
    AesEval E; E.ctx = ctx;
    std::vector<std::vector<emu_gt>> in(input_size, std::vector<emu_gt>(16)), key(10, std::vector<emu_gt>(16));
    for (int i = 0; i < input_size; i++) for (int j = 0; j < 16; j++) E.init(in[i][j], mk((uint64_t)((i * 122 + j) % 256)));
    for (int i = 0; i < 10; i++) for (int j = 0; j < 16; j++) E.init(key[i][j], mk((uint64_t)((i + j + 1) % 256)));
    emu_gt zero; E.init(zero, mk(0));
    for (int b = 0; b < input_size; b++) {
        std::vector<emu_gt> out(16);
        for (int j = 0; j < 16; j++) out[j] = E.x(in[b][j], key[0][j]);
        for (int round = 1; round < 9; round++) {
            std::vector<emu_gt> sb(16), mixed(16);
            std::vector<std::array<emu_gt, 3>> t2(16);
            for (int j = 0; j < 16; j++) sb[j] = E.box(out[j], zero, 2);
            for (int j = 0; j < 16; j++) { t2[j][1] = E.box(sb[j], zero, 3); t2[j][2] = E.box(sb[j], zero, 4); t2[j][0] = sb[j]; }   /* a COPY of the S-box output */
            static const int pick[4][4] = {{1, 2, 0, 0}, {0, 1, 2, 0}, {0, 0, 1, 2}, {2, 0, 0, 1}};        /* [output byte][group byte] */
            for (int k = 0; k < 4; k++)
                for (int m = 0; m < 4; m++) {
                    emu_gt x1 = E.x(t2[4 * k][pick[m][0]], t2[4 * k + 1][pick[m][1]]);
                    emu_gt x2 = E.x(t2[4 * k + 2][pick[m][2]], t2[4 * k + 3][pick[m][3]]);
                    mixed[4 * k + m] = E.x(x1, x2);
                    E.del(x1); E.del(x2);
                }
            for (int j = 0; j < 16; j++) for (int v = 0; v < 3; v++) E.del(t2[j][v]);
            std::vector<emu_gt> next(16);
            for (int j = 0; j < 16; j++) next[j] = E.x(mixed[j], key[round][j]);
            for (int j = 0; j < 16; j++) E.del(mixed[j]);
            for (int j = 0; j < 16; j++) { E.del(out[j]); out[j] = next[j]; }
        }
        for (int j = 0; j < 16; j++) E.del(out[j]);
    }
    for (auto &row : in) for (auto &g : row) E.del(g);
    for (auto &row : key) for (auto &g : row) E.del(g);
    E.del(zero);
    ctx->tr_done = true;
    if (n_records) *n_records = ctx->tr_n;
    return 0;
}
/* Seval.cpp:1170-1236 `inference` under the fun == 8 driver (:1424-1461), restated gate by gate: two sparse layers, neuron i of layer l reads
 * the entries cols_l[rowptr_l[i] .. rowptr_l[i+1]) of the previous layer; a neuron without inputs is a COPY of the gate `zero` */
int hb_trace_generate_pruned_mlp(hb_ctx *ctx, int n_inputs, int n_hidden, int n_out, const int *rowptr0, const int *cols0, const int *rowptr1,
                                 const int *cols1, size_t *n_records) {
    ctx->trace.clear(); ctx->tr_n = 0;
    MlpEval E; E.ctx = ctx;
    std::vector<emu_gt> input(n_inputs), hidden(n_hidden), outl(n_out);
    std::vector<std::vector<emu_gt>> W0(n_hidden), W1(n_out);
    for (int i = 0; i < n_inputs; i++) E.init(input[i], mk((uint64_t)((i + 1) % 256)));
    for (int i = 0; i < n_hidden; i++) { W0[i].resize(rowptr0[i + 1] - rowptr0[i]); for (size_t j = 0; j < W0[i].size(); j++) E.init(W0[i][j], mk((uint64_t)((j + i) % 256))); }
    for (int i = 0; i < n_out; i++) { W1[i].resize(rowptr1[i + 1] - rowptr1[i]); for (size_t j = 0; j < W1[i].size(); j++) E.init(W1[i][j], mk((uint64_t)((j + i) % 256))); }
    emu_gt zero; E.init(zero, mk(0));
    auto layer = [&](std::vector<std::vector<emu_gt>> &W, const int *rowptr, const int *cols, std::vector<emu_gt> &in, std::vector<emu_gt> &out) {
        for (size_t i = 0; i < W.size(); i++) {
            for (size_t j = 0; j < W[i].size(); j++) {
                emu_gt &x = in[cols[rowptr[i] + j]];
                if (j == 0) out[i] = E.op(W[i][j], x, 2);
                else { emu_gt t = E.op(W[i][j], x, 2); emu_gt sum = E.op(out[i], t, 1); E.del(t); E.del(out[i]); out[i] = sum; }
            }
            if (W[i].empty()) out[i] = zero;
        }
    };
    layer(W0, rowptr0, cols0, input, hidden);
    layer(W1, rowptr1, cols1, hidden, outl);
    for (auto &Wi : W0) for (auto &g : Wi) E.del(g);
    for (auto &Wi : W1) for (auto &g : Wi) E.del(g);
    for (auto &g : input) E.del(g);
    for (auto &g : hidden) E.del(g);
    for (auto &g : outl) E.del(g);
    E.del(zero);
    ctx->tr_done = true;
    if (n_records) *n_records = ctx->tr_n;
    return 0;
}
/* Seval.cpp:190-221 lookup_gate, :223-293 ltu_gate, :667-687 get_bytes, :1085-1166 range_query under the fun == 6 driver (:1398-1416),
 * restated gate by gate; domain = 16 (two bytes per word) */
int hb_trace_generate_sql(hb_ctx *ctx, int input_size, size_t *n_records) {
    ctx->trace.clear(); ctx->tr_n = 0;
    AesEval E; E.ctx = ctx;
    const int NB = 2;
    std::vector<emu_gt> DB(input_size), range_table(256), pows(NB);
    for (int i = 0; i < input_size; i++) E.init(DB[i], mk((uint64_t)((i + 12) % 16)));
    emu_gt R, L, zero, minus_one;
    E.init(R, mk(321)); E.init(L, mk(21));
    for (int i = 0; i < 256; i++) E.init(range_table[i], mk((uint64_t)i));
    for (int i = 0; i < NB; i++) E.init(pows[i], mk((uint64_t)1 << (8 * i)));
    E.init(zero, mk(0)); E.init(minus_one, mk(P - 1));
    auto get_bytes = [&](emu_gt &word) {
        std::vector<emu_gt> bytes(NB);
        unsigned v = (unsigned)word.value.re;
        for (int i = 0; i < NB; i++) { emu_gt &e = range_table[v % 256]; bytes[i] = E.table_gate(e, zero, 0, e.value.re); v >>= 8; }
        emu_gt check = E.op(pows[0], bytes[0], 2);
        for (int i = 1; i < NB; i++) { emu_gt tp = E.op(pows[i], bytes[i], 2); emu_gt tc = E.op(check, tp, 1); E.del(check); E.del(tp); check = tc; }
        E.del(check);
        return bytes;
    };
    auto ltu = [&](std::vector<emu_gt> &b1, std::vector<emu_gt> &b2) {
        std::vector<emu_gt> lt(NB), eq(NB - 1);
        for (int i = 0; i < NB; i++) lt[i] = E.table_gate(b1[i], b2[i], 1, b2[i].value.re < b1[i].value.re ? 1 : 0);
        for (int i = 0; i < NB - 1; i++) eq[i] = E.table_gate(b1[i], b2[i], 2, b2[i].value.re == b1[i].value.re ? 1 : 0);
        for (int i = NB - 3; i >= 0; i--) { emu_gt tp = E.op(eq[i], eq[i + 1], 2); E.del(eq[i]); eq[i] = tp; }
        emu_gt out = E.op(eq[0], lt[0], 2);
        for (int i = 1; i < NB - 2; i++) { emu_gt tp = E.op(eq[i], lt[i], 2); emu_gt to = E.op(out, tp, 1); E.del(out); E.del(tp); out = to; }
        emu_gt to = E.op(out, lt[NB - 1], 1); E.del(out); out = to;
        for (auto &g : lt) E.del(g);
        for (auto &g : eq) E.del(g);
        return out;
    };
    std::vector<emu_gt> Lb = get_bytes(L), Rb = get_bytes(L), mx(NB);            /* sic: both from L (Seval.cpp:1099) */
    for (int i = 0; i < input_size; i++) {
        std::vector<emu_gt> bytes = get_bytes(DB[i]);
        emu_gt bit1 = ltu(bytes, Rb), bit2 = ltu(Lb, bytes);
        emu_gt bit = E.op(bit1, bit2, 2);
        E.del(bit1); E.del(bit2);
        std::vector<emu_gt> tm(NB);
        for (int j = 0; j < NB; j++) tm[j] = E.op(bytes[j], bit, 2);
        E.del(bit);
        for (int j = 0; j < NB; j++) E.del(bytes[j]);
        if (i == 0) mx = tm;
        else {
            emu_gt mb = ltu(tm, mx);
            emu_gt mbn = E.op(mb, minus_one, 1);
            for (int j = 0; j < NB; j++) {
                emu_gt p1 = E.op(mx[j], mb, 2), p2 = E.op(tm[j], mbn, 2);
                E.del(mx[j]); E.del(tm[j]);
                mx[j] = E.op(p1, p2, 1);
                E.del(p1); E.del(p2);
            }
            E.del(mb); E.del(mbn);
        }
    }
    for (auto &g : mx) E.del(g);
    E.del(zero);
    for (auto &g : pows) E.del(g);
    for (auto &g : range_table) E.del(g);
    for (int i = 0; i < NB; i++) { E.del(Lb[i]); E.del(Rb[i]); }
    for (auto &g : DB) E.del(g);
    E.del(R); E.del(L); E.del(minus_one);
    ctx->tr_done = true;
    if (n_records) *n_records = ctx->tr_n;
    return 0;
}
int hb_trace_finish(hb_ctx *ctx, size_t *n_records, size_t *n_ops, size_t *n_deletes) {
    const emu_tuple *t = (const emu_tuple *)ctx->trace.data(); size_t o = 0, d = 0;
    for (size_t i = 0; i < ctx->tr_n; i++) { if (t[i].type == 0) d++; else o++; }
    if (n_records) *n_records = ctx->tr_n; if (n_ops) *n_ops = o; if (n_deletes) *n_deletes = d;
    return 0;
}
int hb_trace_witness(hb_ctx *ctx, size_t cs, hb_F *out) {
    const emu_tuple *t = (const emu_tuple *)ctx->trace.data(); F *w = mF(out);
    memset(w, 0, 4 * cs * sizeof(F));
    size_t c = 0;
    for (size_t i = 0; i < ctx->tr_n; i++) if (t[i].type > 0) { w[c++] = t[i].value_l; w[c++] = t[i].value_r; w[c++] = t[i].value_o; }   /* stage 0 */
    c = 3 * cs;
    for (size_t i = 0; i < ctx->tr_n; i++) if (t[i].type == 0) w[c++] = t[i].value_o;                                                  /* stage 1 */
    return 0;
}
int hb_trace_transcript(hb_ctx *ctx, size_t cs, int has_lookups, hb_F *L, hb_F *R, hb_F *O, hb_F *S) {
    const emu_tuple *t = (const emu_tuple *)ctx->trace.data();
    for (hb_F *p : {L, R, O, S}) memset(p, 0, cs * sizeof(F));
    size_t c = 0;
    for (size_t i = 0; i < ctx->tr_n; i++) if (t[i].type > 0) {
        mF(L)[c] = t[i].value_l; mF(R)[c] = t[i].value_r; mF(O)[c] = t[i].value_o;
        int s = t[i].type == 1 ? (has_lookups ? 0 : 1) : t[i].type == 2 ? (has_lookups ? 1 : 0) : 2;
        mF(S)[c++] = mk((uint64_t)s);
    }
    return 0;
}
int hb_trace_circuit(hb_ctx *ctx, size_t cs, hb_F *out) {
    const emu_tuple *t = (const emu_tuple *)ctx->trace.data(); F *v = mF(out);
    memset(v, 0, 16 * cs * sizeof(F));
    size_t c = 0;
    for (size_t i = 0; i < ctx->tr_n; i++) if (t[i].type > 0) v[c++] = mk(t[i].type == 1 ? 1 : 0);              /* read_circuit_trace */
    c = cs;
    for (size_t i = 0; i < ctx->tr_n; i++) if (t[i].type > 0) {                                                      /* read_circuit_memory_transcript */
        v[c++] = f_int(t[i].idx_l); v[c++] = f_int(t[i].access_l); v[c++] = f_int(t[i].idx_r); v[c++] = f_int(t[i].access_r);
        v[c++] = f_int(t[i].idx_o); v[c++] = f_int(t[i].access_o);
    }
    c = 7 * cs;
    for (size_t i = 0; i < ctx->tr_n; i++) if (t[i].type == 0) { v[c++] = f_int(t[i].idx_o); v[c++] = f_int(t[i].access_o); }  /* read_circuit_memory_final */
    return 0;
}
/* access_table semantics (witness_stream.cpp:920-1053): the count of EARLIER lookups of the same table entry in this pass */
static std::vector<uint64_t> lookup_access_counts(const emu_tuple *t, size_t n) {
    std::vector<uint64_t> acc(n, 0);
    std::vector<std::vector<uint64_t>> table(256);
    for (size_t i = 0; i < n; i++) if (t[i].type >= 3) {
        size_t key = t[i].type == 3 ? t[i].value_l.re : t[i].value_l.re + 256 * t[i].value_r.re;
        auto &tb = table[t[i].type - 3];
        if (tb.size() <= key) tb.resize(key + 1, 0);
        acc[i] = tb[key]++;
    }
    return acc;
}
int hb_trace_lookup_basic(hb_ctx *ctx, size_t cs, const hb_F *lookup_rand4, hb_F *xy) {
    const emu_tuple *t = (const emu_tuple *)ctx->trace.data(); const F *lr = cF(lookup_rand4);
    std::vector<uint64_t> acc = lookup_access_counts(t, ctx->tr_n);
    F *X = mF(xy), *Y = X + cs; const F one = mk(1);
    for (size_t i = 0; i < 2 * cs; i++) X[i] = one;
    size_t c = 0;
    for (size_t i = 0; i < ctx->tr_n; i++) if (t[i].type > 0) {
        if (t[i].type >= 3) {
            F v = fadd(one, t[i].value_l);
            v = fadd(v, fmul(lr[0], t[i].value_r)); v = fadd(v, fmul(lr[1], t[i].value_o));
            v = fadd(v, fmul(lr[2], mk(acc[i]))); v = fadd(v, fmul(lr[3], mk(t[i].type)));
            X[c] = v; Y[c] = (v.re == 1 && v.im == 0) ? v : fadd(v, lr[2]);
        }
        c++;
    }
    return 0;
}
int hb_trace_lookup_witness(hb_ctx *ctx, size_t cs, const hb_F *lookup_rand2, hb_F *out) {
    const emu_tuple *t = (const emu_tuple *)ctx->trace.data(); const F *lr = cF(lookup_rand2);
    std::vector<uint64_t> acc = lookup_access_counts(t, ctx->tr_n);
    memset(out, 0, 2 * cs * sizeof(F));
    size_t q = 0;
    for (size_t i = 0; i < ctx->tr_n; i++) if (t[i].type >= 3) {
        mF(out)[2 * q] = fadd(fadd(t[i].value_o, fmul(lr[0], t[i].value_l)), fmul(lr[1], t[i].value_r));
        mF(out)[2 * q + 1] = mk(acc[i]);
        q++;
    }
    return 0;
}
int hb_gate_consistency_lookups_stream(hb_ctx *, const hb_F *L, const hb_F *R, const hb_F *O, const hb_F *S, size_t cs, size_t B, const hb_F *r,
                                       const hb_F *lookup_rand2, const hb_F *rnd13, hb_F *out, double *ps) {
    orc_inject_randomness(cF(rnd13), 13);
    *ps += orc_gate_consistency_lookups_stream(cF(L), cF(R), cF(O), cF(S), cs, B, cF(r), cF(lookup_rand2), mF(out));
    orc_inject_randomness(nullptr, 0);
    return 0;
}
int hb_trace_wiring(hb_ctx *ctx, size_t cs, const hb_F *a_w, const hb_F *b_w, hb_F *xy) {
    const emu_tuple *t = (const emu_tuple *)ctx->trace.data();
    std::vector<F> addr(4 * cs, mk(0)), val(4 * cs, mk(0)), freq(4 * cs, mk(0));
    size_t c = 0;
    for (size_t i = 0; i < ctx->tr_n; i++) if (t[i].type > 0) {
        addr[c] = f_int(t[i].idx_l); val[c] = t[i].value_l; freq[c++] = f_int(t[i].access_l);
        addr[c] = f_int(t[i].idx_r); val[c] = t[i].value_r; freq[c++] = f_int(t[i].access_r);
        addr[c] = f_int(t[i].idx_o); val[c] = t[i].value_o; freq[c++] = f_int(t[i].access_o);
    }
    c = 3 * cs;
    for (size_t i = 0; i < ctx->tr_n; i++) if (t[i].type == 0) { addr[c] = f_int(t[i].idx_o); val[c] = t[i].value_o; freq[c++] = f_int(t[i].access_o); }
    F *X = mF(xy), *Y = X + 4 * cs; const F a = *cF(a_w), b = *cF(b_w), one = mk(1);
    for (size_t i = 0; i < 3 * cs; i++) {                                  /* stage 0 (:2291-2301) */
        X[i] = fadd(fadd(fadd(addr[i], one), fmul(a, val[i])), fmul(b, freq[i]));
        Y[i] = (X[i].re == 1 && X[i].im == 0) ? X[i] : fadd(X[i], b);
    }
    for (size_t i = 3 * cs; i < 4 * cs; i++) {                             /* stage 1 (:2302-2309) */
        X[i] = fadd(fadd(addr[i], one), fmul(a, val[i]));
        Y[i] = fadd(X[i], fmul(b, freq[i]));
    }
    return 0;
}


/* ---- single-process stand-ins for the multi-GPU / pinned-memory entry points the host mirror links against: the emulation is one rank ---- */
int hb_malloc_pinned(hb_ctx *, void **p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? 0 : 1; }
int hb_free_pinned(hb_ctx *, void *p) { free(p); return 0; }
uint64_t hb_transcript_digest(hb_ctx *, int) { return 0; }
int hb_dist_local_info(hb_ctx *, size_t, void *blob) { memset(blob, 0, 256); return 0; }
int hb_dist_connect(hb_ctx *, int, int world, const void *) { return world == 1 ? 0 : 1; }
int hb_dist_disconnect(hb_ctx *) { return 0; }
int hb_dist_rank(hb_ctx *) { return 0; }
int hb_dist_world(hb_ctx *) { return 1; }
int hb_dist_barrier(hb_ctx *) { return 0; }
int hb_dist_shard(hb_ctx *, int) { return 0; }
int hb_dist_allreduce(hb_ctx *, hb_F *, size_t) { return 0; }
int hb_dist_elastic_commit(hb_ctx *, const hb_F *, size_t, size_t, int, int, uint8_t *) { return 1; }   /* only reached with world > 1 */
int hb_dist_elastic_begin(hb_ctx *ctx, size_t B, int trs, int lin, size_t) { return hb_elastic_begin(ctx, B, trs, lin); }
int hb_elastic_open_range(hb_ctx *, size_t, size_t) { return 1; }
int hb_elastic_finish_levels_async(hb_ctx *ctx, uint8_t *const *level_ptrs, int nlevels) { return hb_elastic_finish_levels(ctx, level_ptrs, nlevels); }
int hb_levels_wait(hb_ctx *) { return 0; }
int hb_levels_copy_async(hb_ctx *, uint8_t *, int, uint8_t *const *, int, size_t) { return 1; }                                /* only reached with world > 1 */                                       /* only reached with world > 1 */

}  // extern "C"
