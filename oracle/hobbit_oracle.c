/* TEST INFRASTRUCTURE ONLY — see hobbit_oracle.h.  Plain-C restatement of the
 * reference hot path; pinned against oracle/_ref (the unmodified reference). */
#include "hobbit_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef orc_F F;
typedef unsigned __int128 u128;
#define P61 2305843009213693951ULL

/* ------------------------------------------------------------------ F1 ---
 * fieldElement.cpp:34-104 (+ - * unary-), :336-360 (myMod, mymult).
 * Outputs are canonical (limbs in [0,p)) for canonical inputs, so any exact
 * formula reproduces the reference bits. */
static inline uint64_t red(u128 x) {           /* x < 2^125 -> [0,p) */
    uint64_t lo = (uint64_t)x & P61, hi = (uint64_t)(x >> 61);
    uint64_t s = lo + (hi & P61) + (hi >> 61);
    s = (s & P61) + (s >> 61);
    return s >= P61 ? s - P61 : s;
}
static inline uint64_t addm(uint64_t a, uint64_t b) { uint64_t s = a + b; return s >= P61 ? s - P61 : s; }
static inline uint64_t subm(uint64_t a, uint64_t b) { return a >= b ? a - b : a + P61 - b; }
static inline uint64_t mulm(uint64_t a, uint64_t b) { return red((u128)a * b); }

static inline F f_add(F a, F b) { F r = { addm(a.re, b.re), addm(a.im, b.im) }; return r; }
static inline F f_sub(F a, F b) { F r = { subm(a.re, b.re), subm(a.im, b.im) }; return r; }
static inline F f_neg(F a) { F z = {0, 0}; return f_sub(z, a); }
static inline F f_mul(F a, F b) {              /* (a.re + i a.im)(b.re + i b.im), i^2 = -1 */
    uint64_t ac = mulm(a.re, b.re), bd = mulm(a.im, b.im);
    uint64_t ad = mulm(a.re, b.im), bc = mulm(a.im, b.re);
    F r = { subm(ac, bd), addm(ad, bc) };
    return r;
}
static inline int f_eq(F a, F b) { return a.re == b.re && a.im == b.im; }
static inline F f_int(long long x) { F r = { x >= 0 ? (uint64_t)x : P61 + x, 0 }; return r; } /* fieldElement.cpp:25-28 */
static const F F0 = {0, 0}, F1 = {1, 0};

/* fieldElement.cpp:206-209, 322-334: inv = x^(p^2-2) by square-and-multiply */
static F f_inv(F x) {
    u128 e = (u128)P61 * P61 - 2;
    F ret = F1, tmp = x;
    while (e) { if (e & 1) ret = f_mul(ret, tmp); tmp = f_mul(tmp, tmp); e >>= 1; }
    return ret;
}

void orc_field_binop(int op, const F *a, const F *b, F *c, size_t n) {
    for (size_t i = 0; i < n; i++) {
        switch (op) {
            case 0: c[i] = f_add(a[i], b[i]); break;
            case 1: c[i] = f_sub(a[i], b[i]); break;
            case 2: c[i] = f_mul(a[i], b[i]); break;
            case 3: c[i] = f_neg(a[i]); break;
            case 4: c[i] = f_inv(a[i]); break;
        }
    }
}

/* utils.cpp:452-463 */
static F root_of_unity(int n) {
    F rou = { 2147483648ULL, 1033321771269002680ULL };
    for (int i = 0; i < 62 - n; i++) rou = f_mul(rou, rou);
    return rou;
}
void orc_root_of_unity(int n, F *out) { *out = root_of_unity(n); }

/* ------------------------------------------------------------------ M1 ---
 * mimc.cpp:11-19 (constants c_i = F(i), key unused here), :95-107. */
static F mimc(F input, F k) {
    F t, h = F0;
    for (int i = 0; i < 161; i++) {
        if (i == 0) t = f_add(input, k);
        else t = f_add(f_add(h, k), f_int(i - 1));
        h = f_mul(f_mul(t, t), t);
    }
    return f_add(h, k);
}
void orc_mimc_hash(const F *in, const F *k, F *out) { *out = mimc(*in, *k); }

/* ------------------------------------------------------------------ N1 ---
 * utils.cpp:605-673 (_fft, flag=false): bit-reverse, twiddles w[k]=rou^k by
 * repeated multiplication, radix-2 DIT. */
void orc_fft(F *arr, int logn) {
    uint32_t len = 1u << logn;
    F rou = root_of_unity(logn);
    uint32_t *rev = (uint32_t *)malloc(len * sizeof(uint32_t));
    rev[0] = 0;
    for (uint32_t i = 1; i < len; i++) rev[i] = rev[i >> 1] >> 1 | (i & 1) << (logn - 1);
    for (uint32_t i = 0; i < len; i++) if (rev[i] < i) { F t = arr[i]; arr[i] = arr[rev[i]]; arr[rev[i]] = t; }
    F *w = (F *)malloc(len * sizeof(F));
    w[0] = F1; if (len > 1) w[1] = rou;
    for (uint32_t i = 2; i < len; i++) w[i] = f_mul(w[i - 1], w[1]);
    for (uint32_t i = 2; i <= len; i <<= 1)
        for (uint32_t j = 0; j < len; j += i)
            for (uint32_t k = 0; k < (i >> 1); k++) {
                F u = arr[j + k], v = f_mul(arr[j + k + (i >> 1)], w[len / i * k]);
                arr[j + k] = f_add(u, v);
                arr[j + k + (i >> 1)] = f_sub(u, v);
            }
    free(rev); free(w);
}

/* utils.cpp:873-883: c = random() every 100 elements; x_i = c + F(rand()) */
/* Injected randomness (used only by the CPU emulation of the C ABI, oracle/hb_emul.cpp): the host mirror draws the libc values itself,
 * in the reference's order, and hands them to the C ABI; the provers below then take their draws from this queue instead of libc. */
static const F *inj_q = NULL; static size_t inj_n = 0, inj_pos = 0;
void orc_inject_randomness(const F *q, size_t n) { inj_q = q; inj_n = n; inj_pos = 0; }
static void prover_randomness(int n, F *out) {
    if (inj_q) { for (int i = 0; i < n; i++) { if (inj_pos >= inj_n) { printf("orc: injected randomness exhausted\n"); exit(-1); } out[i] = inj_q[inj_pos++]; } return; }
    orc_generate_randomness(n, out);
}
static F prover_random(void) {
    if (inj_q) { F o; prover_randomness(1, &o); return o; }
    F o = { (uint64_t)random(), 0 }; return o;
}
void orc_generate_randomness(int n, F *out);
void orc_generate_randomness(int n, F *out) {
    F c = F0;
    for (int i = 0; i < n; i++) {
        if (i % 100 == 0) c = f_int(random());
        out[i] = f_add(c, f_int(rand()));
    }
}

/* --------------------------------------------------------------- E1/E2 ---
 * parameter.h:4-8, expanders.h:20-47,78-92, linear_code_encode.h:62-119. */
static const double kAlpha = 0.211, kR = 1.72;
enum { kCn = 9, kDn = 12, kMaxDep = 64 };
static const int kDistThreshold = (int)(1.0 / 0.07) - 1;      /* 13 */
typedef struct { long long L, R; int deg; uint32_t *nbr; uint64_t *w; } orc_graph;
static orc_graph gC[kMaxDep], gD[kMaxDep];

static void gen_graph(orc_graph *g, long long L, long long R, int d) {
    free(g->nbr); free(g->w);
    g->L = L; g->R = R; g->deg = d;
    g->nbr = (uint32_t *)malloc((size_t)L * d * sizeof(uint32_t));
    g->w = (uint64_t *)malloc((size_t)L * d * sizeof(uint64_t));
    for (long long i = 0; i < L; i++)
        for (int j = 0; j < d; j++) {
            g->nbr[i * d + j] = (uint32_t)(rand() % R);   /* target first ... */
            g->w[i * d + j] = (uint64_t)random();          /* ... then weight = F(random()) */
        }
}
static long long expander_init(long long n, int dep) {
    if (n <= kDistThreshold) return n;
    gen_graph(&gC[dep], n, (long long)(kAlpha * n), kCn);
    long long L = expander_init((long long)(kAlpha * n), dep + 1);
    gen_graph(&gD[dep], L, (long long)(n * (kR - 1) - L), kDn);
    return n + L + (long long)(n * (kR - 1) - L);
}
long long orc_expander_init_store(long long n) { return expander_init(n, 0); }
int orc_expander_levels(long long n) { int d = 0; while (n > kDistThreshold) { n = (long long)(kAlpha * n); d++; } return d; }
long long orc_expander_dims(int which, int dep, long long *R, int *deg) {
    orc_graph *g = which ? &gD[dep] : &gC[dep]; *R = g->R; *deg = g->deg; return g->L;
}
void orc_expander_dump(int which, int dep, uint32_t *nbr, uint64_t *w) {
    orc_graph *g = which ? &gD[dep] : &gC[dep];
    memcpy(nbr, g->nbr, (size_t)g->L * g->deg * sizeof(uint32_t));
    memcpy(w, g->w, (size_t)g->L * g->deg * sizeof(uint64_t));
}
/* install host-generated graphs (used by the CPU emulation of the C ABI, oracle/hb_emul.cpp: the host mirror draws them itself) */
void orc_expander_install(int which, int dep, long long L, long long R, int deg, const uint32_t *nbr, const uint64_t *w) {
    orc_graph *g = which ? &gD[dep] : &gC[dep];
    free(g->nbr); free(g->w);
    g->L = L; g->R = R; g->deg = deg;
    g->nbr = (uint32_t *)malloc((size_t)L * deg * sizeof(uint32_t));
    g->w = (uint64_t *)malloc((size_t)L * deg * sizeof(uint64_t));
    memcpy(g->nbr, nbr, (size_t)L * deg * sizeof(uint32_t));
    memcpy(g->w, w, (size_t)L * deg * sizeof(uint64_t));
}
static long long encode_rec(const F *src, F *dst, long long n, int dep) {
    if (n <= kDistThreshold) { for (long long i = 0; i < n; i++) dst[i] = src[i]; return n; }
    for (long long i = 0; i < n; i++) dst[i] = src[i];
    long long R = (long long)(kAlpha * n);
    F *y = (F *)calloc((size_t)R, sizeof(F));
    const orc_graph *c = &gC[dep];
    for (long long i = 0; i < n; i++)
        for (int d = 0; d < c->deg; d++) {
            uint32_t t = c->nbr[i * c->deg + d];
            F wv = { c->w[i * c->deg + d], 0 };
            y[t] = f_add(y[t], f_mul(wv, src[i]));
        }
    long long L = encode_rec(y, dst + n, R, dep + 1);
    free(y);
    const orc_graph *g = &gD[dep];
    R = g->R;
    for (long long i = 0; i < R; i++) dst[n + L + i] = F0;
    for (long long i = 0; i < L; i++)
        for (int d = 0; d < g->deg; d++) {
            uint32_t t = g->nbr[i * g->deg + d];
            F wv = { g->w[i * g->deg + d], 0 };
            dst[n + L + t] = f_add(dst[n + L + t], f_mul(dst[n + i], wv));
        }
    return n + L + R;
}
int orc_encode_monolithic(const F *src, F *dst, long long n) { return (int)encode_rec(src, dst, n, 0); }

/* E3: encode() (linear_code_encode.h:122-191) — the variant that re-draws its graph on every call from FIXED libc seeds: the C stage of
 * depth dep after srand(666 + dep) (per edge: target = rand() % R, then weight = rand()), the D stage after srand(2 * (666 + dep)) (per
 * edge: weight = rand(), THEN target = rand() % R).  Sizes as in expander_init (R_C = alpha n; D: L x (n (r - 1) - L)).  The libc state
 * a caller sees afterwards is the one the depth-0 D stage leaves behind. */
static long long encode_reseed_rec(const F *src, F *dst, long long n, int dep) {
    if (n <= kDistThreshold) { for (long long i = 0; i < n; i++) dst[i] = src[i]; return n; }
    for (long long i = 0; i < n; i++) dst[i] = src[i];
    long long R = (long long)(kAlpha * n);
    F *y = (F *)calloc((size_t)R, sizeof(F));
    srand(666 + dep);
    for (long long i = 0; i < n; i++)
        for (int d = 0; d < kCn; d++) {
            int t = rand() % (int)R;
            F wv = { (uint64_t)rand(), 0 };
            y[t] = f_add(y[t], f_mul(wv, src[i]));
        }
    long long L = encode_reseed_rec(y, dst + n, R, dep + 1);
    free(y);
    long long RD = (long long)(n * (kR - 1) - L);
    for (long long i = 0; i < RD; i++) dst[n + L + i] = f_int(0);
    srand(2 * (666 + dep));
    for (long long i = 0; i < L; i++)
        for (int d = 0; d < kDn; d++) {
            F wv = { (uint64_t)rand(), 0 };
            long long t = rand() % RD;
            dst[n + L + t] = f_add(dst[n + L + t], f_mul(dst[n + i], wv));
        }
    return n + L + RD;
}
int orc_encode_reseed(const F *src, F *dst, long long n) { return (int)encode_reseed_rec(src, dst, n, 0); }

/* ------------------------------------------------------------------ H1 ---
 * Blake3_hash.cpp:5-10 = BLAKE3 of exactly 64 bytes -> one compression with
 * cv = IV, counter 0, block_len 64, flags CHUNK_START|CHUNK_END|ROOT
 * (Blake/blake3_impl.h:18-21; compression per Blake/blake3_portable.c). */
static const uint32_t B3_IV[8] = { 0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u };
static const uint8_t B3_PERM[16] = { 2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8 };
static inline uint32_t rotr32(uint32_t x, int c) { return (x >> c) | (x << (32 - c)); }
#define B3G(a, b, c, d, x, y) do { \
    v[a] = v[a] + v[b] + (x); v[d] = rotr32(v[d] ^ v[a], 16); v[c] = v[c] + v[d]; v[b] = rotr32(v[b] ^ v[c], 12); \
    v[a] = v[a] + v[b] + (y); v[d] = rotr32(v[d] ^ v[a], 8);  v[c] = v[c] + v[d]; v[b] = rotr32(v[b] ^ v[c], 7); } while (0)
void orc_blake3_hash(const uint8_t *src, uint8_t *dst) {
    uint32_t m[16], v[16], t[16];
    memcpy(m, src, 64);                                   /* little-endian host */
    for (int i = 0; i < 8; i++) v[i] = B3_IV[i];
    v[8] = B3_IV[0]; v[9] = B3_IV[1]; v[10] = B3_IV[2]; v[11] = B3_IV[3];
    v[12] = 0; v[13] = 0; v[14] = 64; v[15] = 1 | 2 | 8;
    for (int r = 0; r < 7; r++) {
        B3G(0, 4, 8, 12, m[0], m[1]);  B3G(1, 5, 9, 13, m[2], m[3]);
        B3G(2, 6, 10, 14, m[4], m[5]); B3G(3, 7, 11, 15, m[6], m[7]);
        B3G(0, 5, 10, 15, m[8], m[9]); B3G(1, 6, 11, 12, m[10], m[11]);
        B3G(2, 7, 8, 13, m[12], m[13]); B3G(3, 4, 9, 14, m[14], m[15]);
        for (int i = 0; i < 16; i++) t[i] = m[B3_PERM[i]];
        memcpy(m, t, sizeof m);
    }
    for (int i = 0; i < 8; i++) v[i] ^= v[i + 8];
    memcpy(dst, v, 32);
}

/* H2: merkle_tree.cpp:62-87 */
void orc_md_leaf(const F *xyzw, const uint8_t *prev, uint8_t *out) {
    uint8_t data[64];
    memcpy(data, xyzw, 64);
    orc_blake3_hash(data, data);          /* first 32 bytes <- H1(x|y|z|w) */
    memcpy(data + 32, prev, 32);
    orc_blake3_hash(data, out);
}
/* H3: merkle_tree.cpp:255-287 — parent = H1(left ‖ LEFT) (the right child is never read) */
static int create_tree(int n, uint8_t *flat) {          /* flat holds the leaves at offset 0 */
    int lvls = 1; size_t prev = 0, cur = (size_t)n * 32;
    for (int sz = n / 2; sz >= 1; sz /= 2) {
        for (int i = 0; i < sz; i++) {
            uint8_t data[64];
            memcpy(data, flat + prev + (size_t)2 * i * 32, 32);
            memcpy(data + 32, flat + prev + (size_t)2 * i * 32, 32);
            orc_blake3_hash(data, flat + cur + (size_t)i * 32);
        }
        prev = cur; cur += (size_t)sz * 32; lvls++;
    }
    return lvls;
}
int orc_create_tree_blake(const uint8_t *leaves, int n, uint8_t *out) {
    memcpy(out, leaves, (size_t)n * 32);
    return create_tree(n, out);
}
/* H4: merkle_tree.cpp:193-221 */
int orc_mt_commit_blake(const F *leafs, int N, uint8_t *out) {
    for (int i = 0; i < N / 4; i++) orc_blake3_hash((const uint8_t *)(leafs + 4 * i), out + (size_t)i * 32);
    return create_tree(N / 4, out);
}

/* ------------------------------------------------------------------ T1 ---
 * PC_utils.cpp:66-123 (and its pointer twin :9-64): 2trs x (2n/trs) matrix,
 * message in the top-left trs x n/trs block, rows NTT'd at length 2n/trs,
 * then every column either NTT'd at length 2trs (RS) or expander-encoded. */
static int ilog2(size_t x) { int l = 0; while (x >>= 1) l++; return l; }
void orc_compute_tensorcode(const F *msg, size_t n, int trs, int lin, F *T) {
    size_t cols = 2 * n / trs, rows = 2 * (size_t)trs;
    memset(T, 0, rows * cols * sizeof(F));
    for (int i = 0; i < trs; i++) memcpy(T + i * cols, msg + (size_t)i * (n / trs), (n / trs) * sizeof(F));
    for (int i = 0; i < trs; i++) orc_fft(T + i * cols, ilog2(cols));
    F *buf = (F *)malloc(rows * sizeof(F)), *buf2 = (F *)malloc(rows * sizeof(F));
    for (size_t c = 0; c < cols; c++) {
        if (!lin) {
            for (size_t j = 0; j < rows; j++) buf[j] = j < (size_t)trs ? T[j * cols + c] : F0;
            orc_fft(buf, ilog2(rows));
            for (size_t j = 0; j < rows; j++) T[j * cols + c] = buf[j];
        } else {
            for (int j = 0; j < trs; j++) buf[j] = T[j * cols + c];
            memset(buf2, 0, rows * sizeof(F));
            orc_encode_monolithic(buf, buf2, trs);
            for (size_t j = 0; j < rows; j++) T[j * cols + c] = buf2[j];
        }
    }
    free(buf); free(buf2);
}

/* ------------------------------------------------------------------ C1 ---
 * Our_PC.cpp:146-171 */
void orc_commit_standard(const F *poly, size_t N, int K, int trs, int lin, uint8_t *levels_out, F *tensor_out) {
    size_t B = N / K, cols = 2 * B / trs;
    F *T = tensor_out ? NULL : (F *)malloc(4 * B * sizeof(F));
    memset(levels_out, 0, B * 32);
    for (int i = 0; i < K; i++) {
        F *Ti = tensor_out ? tensor_out + (size_t)i * 4 * B : T;
        orc_compute_tensorcode(poly + (size_t)i * B, B, trs, lin, Ti);
        for (int j = 0; j < trs / 2; j++)
            for (size_t k = 0; k < cols; k++) {
                F q[4] = { Ti[(4 * j) * cols + k], Ti[(4 * j + 1) * cols + k], Ti[(4 * j + 2) * cols + k], Ti[(4 * j + 3) * cols + k] };
                uint8_t *leaf = levels_out + (j * cols + k) * 32;
                orc_md_leaf(q, leaf, leaf);
            }
    }
    free(T);
    create_tree((int)B, levels_out);
}

/* witness_stream.cpp:2405-2411: synthetic stream used by test_Elastic_PC */
void orc_read_stream_pc_test(F *v, size_t n) {
    F x = f_int(322322);
    for (size_t i = 0; i < n; i++) { v[i] = x; x = f_add(f_mul(x, x), f_int((long long)i)); }
}
/* ------------------------------------------------------------------ C2 ---
 * Elastic_PC.cpp:174-285 on stream "test".  Chunks i%4 in {0,1,2} are parked;
 * chunk i%4==3 triggers h[pos] = H2(c0[pos],c1[pos],c2[pos],T[pos], h[pos]);
 * trailing chunks (K%4 != 0) are never hashed. */
void orc_elastic_commit_stream(const F *stream, size_t N, size_t B, int trs, int lin, uint8_t *levels_out);
void orc_elastic_commit(size_t N, size_t B, int trs, int lin, uint8_t *levels_out) { orc_elastic_commit_stream(NULL, N, B, trs, lin, levels_out); }
/* stream == NULL: the synthetic "test" stream (every chunk restarts the recurrence); else N consecutive elements */
void orc_elastic_commit_stream(const F *stream, size_t N, size_t B, int trs, int lin, uint8_t *levels_out) {
    F *buff = (F *)malloc(B * sizeof(F));
    F *park[3], *T = (F *)malloc(4 * B * sizeof(F));
    for (int i = 0; i < 3; i++) park[i] = (F *)malloc(4 * B * sizeof(F));
    memset(levels_out, 0, 4 * B * 32);
    for (size_t i = 0; i < N / B; i++) {
        if (stream) memcpy(buff, stream + i * B, B * sizeof(F)); else orc_read_stream_pc_test(buff, B);
        int nz = 0;
        for (size_t j = 0; j < B; j++) if (!f_eq(buff[j], F0)) { nz = 1; break; }
        if (nz) orc_compute_tensorcode(buff, B, trs, lin, T); else memset(T, 0, 4 * B * sizeof(F));
        if (i % 4 != 3) memcpy(park[i % 4], T, 4 * B * sizeof(F));
        else for (size_t p = 0; p < 4 * B; p++) {
            /* Elastic_PC.cpp:234-236 passes (c0[counter], c1[counter], c2[counter++], T[j][k]) as
             * by-value arguments.  GCC evaluates them right to left, so the increment lands BEFORE
             * c1 and c0 are read: the hashed tuple is (c0[p+1], c1[p+1], c2[p], T[p]).  At the last
             * position the reference reads one element past both arrays (heap contents: zero when the
             * blocks are mmap'd, allocator metadata otherwise); we define it as zero.  That leaf is a
             * right child and, with the left‖left parent rule, influences nothing above it. */
            F q[4] = { p + 1 < 4 * B ? park[0][p + 1] : F0, p + 1 < 4 * B ? park[1][p + 1] : F0, park[2][p], T[p] };
            orc_md_leaf(q, levels_out + p * 32, levels_out + p * 32);
        }
    }
    for (int i = 0; i < 3; i++) free(park[i]);
    free(T); free(buff);
    create_tree((int)(4 * B), levels_out);
}

/* Elastic commit split for sharding: inner digests of `ngroups` groups of 4 chunks, plain [group][position] order */
void orc_elastic_encode_groups(const F *chunks, size_t ngroups, size_t B, int trs, int lin, uint8_t *inner_out) {
    size_t cells = 4 * B;
    F *T = (F *)malloc(4 * cells * sizeof(F));
    for (size_t g = 0; g < ngroups; g++) {
        for (int c = 0; c < 4; c++) orc_compute_tensorcode(chunks + (g * 4 + c) * B, B, trs, lin, T + c * cells);
        for (size_t p = 0; p < cells; p++) {
            F q[4] = { p + 1 < cells ? T[p + 1] : F0, p + 1 < cells ? T[cells + p + 1] : F0, T[2 * cells + p], T[3 * cells + p] };
            orc_blake3_hash((const uint8_t *)q, inner_out + (g * cells + p) * 32);
        }
    }
    free(T);
}

/* ------------------------------------------------------------------ S9 ---
 * utils.cpp:251-296: eq table; step i uses r[size-1-i] (last variable = LSB) */
void orc_precompute_beta(const F *r, int nr, F *B) {
    F *tmp = (F *)malloc(((size_t)1 << nr) * sizeof(F));
    B[0] = F1;
    for (int i = 0; i < nr; i++) {
        memcpy(tmp, B, ((size_t)1 << i) * sizeof(F));
        for (size_t j = 0; j < ((size_t)1 << i); j++) {
            F t = f_mul(r[nr - 1 - i], tmp[j]);
            B[2 * j] = f_sub(tmp[j], t);
            B[2 * j + 1] = t;
        }
    }
    free(tmp);
}
/* utils.cpp:789-802 */
static F evaluate_vector(const F *vin, size_t n, const F *r) {
    int nr = ilog2(n);
    F *v = (F *)malloc(n * sizeof(F)); memcpy(v, vin, n * sizeof(F));
    for (int i = 0; i < nr; i++) {
        size_t L = (size_t)1 << (nr - 1 - i);
        for (size_t j = 0; j < L; j++) v[j] = f_add(f_mul(f_sub(F1, r[i]), v[2 * j]), f_mul(r[i], v[2 * j + 1]));
    }
    F e = v[0]; free(v); return e;
}
void orc_evaluate_vector(const F *v, size_t n, const F *r, int nr, F *out) { (void)nr; *out = evaluate_vector(v, n, r); }

/* ------------------------------------------------------------------ S1 ---
 * sumcheck.cpp:2391-2460.  Round poly (a,b,c) of sum_j (d1 t + v1[2j])(d2 t + v2[2j]);
 * challenge AFTER the polynomial: rand = mimc(mimc(mimc(rand,a),b),c); fold with it. */
double orc_sumcheck2(const F *_v1, const F *_v2, size_t n, const F *prev_r, F *out) {
    int rounds = ilog2(n); double ps = 0; size_t k = 0;
    F *v1 = (F *)malloc(n * sizeof(F)), *v2 = (F *)malloc(n * sizeof(F));
    memcpy(v1, _v1, n * sizeof(F)); memcpy(v2, _v2, n * sizeof(F));
    F rand = *prev_r; F *rs = out + 3 * rounds;
    for (int i = 0; i < rounds; i++) {
        size_t L = (size_t)1 << (rounds - 1 - i);
        F a = F0, b = F0, c = F0;
        for (size_t j = 0; j < L; j++) {
            F d1 = f_sub(v1[2 * j + 1], v1[2 * j]), d2 = f_sub(v2[2 * j + 1], v2[2 * j]);
            a = f_add(a, f_mul(d1, d2));
            b = f_add(b, f_add(f_mul(d1, v2[2 * j]), f_mul(d2, v1[2 * j])));
            c = f_add(c, f_mul(v1[2 * j], v2[2 * j]));
        }
        rand = mimc(rand, a); rand = mimc(rand, b); rand = mimc(rand, c);
        out[k++] = a; out[k++] = b; out[k++] = c; rs[i] = rand;
        ps += 3 * 16 / 1024.0;
        for (size_t j = 0; j < L; j++) {
            v1[j] = f_add(v1[2 * j], f_mul(rand, f_sub(v1[2 * j + 1], v1[2 * j])));
            v2[j] = f_add(v2[2 * j], f_mul(rand, f_sub(v2[2 * j + 1], v2[2 * j])));
        }
    }
    rand = mimc(rand, v1[0]); rand = mimc(rand, v2[0]);
    ps += 2 * 16 / 1024.0;
    k += rounds; out[k++] = v1[0]; out[k++] = v2[0]; out[k++] = rand;
    free(v1); free(v2);
    return ps;
}

/* cubic coefficients of (d1 t + x1)(d2 t + x2)(d3 t + x3) accumulated into acc[4] = (a,b,c,d) */
static inline void cubic_acc(F *acc, F x1, F y1, F x2, F y2, F x3, F y3) {
    F d1 = f_sub(y1, x1), d2 = f_sub(y2, x2), d3 = f_sub(y3, x3);
    F qa = f_mul(d1, d2), qb = f_add(f_mul(d1, x2), f_mul(d2, x1)), qc = f_mul(x1, x2);
    acc[0] = f_add(acc[0], f_mul(qa, d3));
    acc[1] = f_add(acc[1], f_add(f_mul(qa, x3), f_mul(qb, d3)));
    acc[2] = f_add(acc[2], f_add(f_mul(qb, x3), f_mul(qc, d3)));
    acc[3] = f_add(acc[3], f_mul(qc, x3));
}
/* ------------------------------------------------------------------ S2 ---
 * sumcheck.cpp:1974-2058.  The fold of round i happens in the same loop as the
 * polynomial accumulation and uses the INCOMING rand; randomness[0][i] is the
 * rand before round i; 4 mimc per round afterwards. Tables are destroyed. */
static double sumcheck3_inplace(F *v1, F *v2, F *v3, size_t n, F prev_r, F *out) {
    int rounds = ilog2(n); double ps = 0; size_t k = 0;
    F rand = prev_r; F *rs = out + 4 * rounds;
    for (int i = 0; i < rounds; i++) {
        size_t L = (size_t)1 << (rounds - 1 - i);
        F acc[4] = { F0, F0, F0, F0 };
        for (size_t j = 0; j < L; j++) {
            cubic_acc(acc, v1[2 * j], v1[2 * j + 1], v2[2 * j], v2[2 * j + 1], v3[2 * j], v3[2 * j + 1]);
            v1[j] = f_add(v1[2 * j], f_mul(rand, f_sub(v1[2 * j + 1], v1[2 * j])));
            v2[j] = f_add(v2[2 * j], f_mul(rand, f_sub(v2[2 * j + 1], v2[2 * j])));
            v3[j] = f_add(v3[2 * j], f_mul(rand, f_sub(v3[2 * j + 1], v3[2 * j])));
        }
        rs[i] = rand;
        rand = mimc(rand, acc[0]); rand = mimc(rand, acc[1]); rand = mimc(rand, acc[2]); rand = mimc(rand, acc[3]);
        ps += 5 * 16 / 1024.0;
        for (int c = 0; c < 4; c++) out[k++] = acc[c];
    }
    rand = mimc(rand, v1[0]); rand = mimc(rand, v2[0]);
    ps += 3 * 16 / 1024.0;
    k += rounds; out[k++] = v1[0]; out[k++] = v2[0]; out[k++] = v3[0]; out[k++] = rand;
    return ps;
}
double orc_sumcheck3(const F *_v1, const F *_v2, const F *_v3, size_t n, const F *prev_r, F *out) {
    F *v = (F *)malloc(3 * n * sizeof(F));
    memcpy(v, _v1, n * sizeof(F)); memcpy(v + n, _v2, n * sizeof(F)); memcpy(v + 2 * n, _v3, n * sizeof(F));
    double ps = sumcheck3_inplace(v, v + n, v + 2 * n, n, *prev_r, out);
    free(v); return ps;
}

/* ------------------------------------------------------------------ S3 ---
 * sumcheck.cpp:275-372.  rand starts at F(312); exhausted batches continue as
 * (1-rand)*v[0] with linear factors (-v0 t + v0); vr uses F(-1) as "unset". */
double orc_batch_sumcheck3(const F *t1, const F *t2, const F *t3, const size_t *sizes, int batches, const F *a, F *out) {
    size_t tot = 0, Lmax = 0;
    for (int b = 0; b < batches; b++) { tot += sizes[b]; if (sizes[b] > Lmax) Lmax = sizes[b]; }
    F *A = (F *)malloc(3 * tot * sizeof(F)), *B = A + tot, *C = A + 2 * tot;
    memcpy(A, t1, tot * sizeof(F)); memcpy(B, t2, tot * sizeof(F)); memcpy(C, t3, tot * sizeof(F));
    size_t *off = (size_t *)malloc(batches * sizeof(size_t));
    for (int b = 0, o = 0; b < batches; b++) { off[b] = o; o += sizes[b]; }
    int rounds = ilog2(Lmax); double ps = 0; size_t k = 0;
    F rand = f_int(312), unset = f_int(-1);
    F *rs = out + 4 * rounds, *vr = out + 5 * rounds;
    for (int j = 0; j < 3 * batches; j++) vr[j] = unset;
    for (int i = 0; i < rounds; i++) {
        F poly[4] = { F0, F0, F0, F0 };
        for (int j = 0; j < batches; j++) {
            F *x1 = A + off[j], *x2 = B + off[j], *x3 = C + off[j];
            int lg = ilog2(sizes[j]) - 1 - i;
            F p[4] = { F0, F0, F0, F0 };
            if (lg >= 0) {
                size_t L = (size_t)1 << lg;
                for (size_t q = 0; q < L; q++) cubic_acc(p, x1[2 * q], x1[2 * q + 1], x2[2 * q], x2[2 * q + 1], x3[2 * q], x3[2 * q + 1]);
            } else {
                if (f_eq(vr[3 * j], unset)) { vr[3 * j] = x1[0]; vr[3 * j + 1] = x2[0]; vr[3 * j + 2] = x3[0]; }
                /* linear_poly(-v0, v0): value v0 at t=0 and 0 at t=1, i.e. pair (v0, 0) */
                cubic_acc(p, x1[0], F0, x2[0], F0, x3[0], F0);
            }
            for (int c = 0; c < 4; c++) poly[c] = f_add(poly[c], f_mul(a[j], p[c]));
        }
        rand = mimc(rand, poly[0]); rand = mimc(rand, poly[1]); rand = mimc(rand, poly[2]); rand = mimc(rand, poly[3]);
        rs[i] = rand;
        for (int c = 0; c < 4; c++) out[k++] = poly[c];
        ps += 4 * 16 / 1024.0;
        for (int j = 0; j < batches; j++) {
            F *x1 = A + off[j], *x2 = B + off[j], *x3 = C + off[j];
            int lg = ilog2(sizes[j]) - 1 - i;
            if (lg >= 0) {
                size_t L = (size_t)1 << lg;
                for (size_t q = 0; q < L; q++) {
                    x1[q] = f_add(x1[2 * q], f_mul(rand, f_sub(x1[2 * q + 1], x1[2 * q])));
                    x2[q] = f_add(x2[2 * q], f_mul(rand, f_sub(x2[2 * q + 1], x2[2 * q])));
                    x3[q] = f_add(x3[2 * q], f_mul(rand, f_sub(x3[2 * q + 1], x3[2 * q])));
                }
            } else {
                F om = f_sub(F1, rand);
                x1[0] = f_mul(om, x1[0]); x2[0] = f_mul(om, x2[0]); x3[0] = f_mul(om, x3[0]);
            }
        }
    }
    for (int j = 0; j < batches; j++)
        if (f_eq(vr[3 * j], unset)) { vr[3 * j] = A[off[j]]; vr[3 * j + 1] = B[off[j]]; vr[3 * j + 2] = C[off[j]]; }
    ps += (3 * batches - batches) * 16 / 1024.0;
    free(A); free(off);
    return ps;
}

/* ------------------------------------------------------------------ S5 ---
 * sumcheck.cpp:35-257 (prev_x empty).  Power-of-two inputs only here (the
 * reference pads with F(1) / zero vectors otherwise). */
size_t orc_mul_tree(const F *input, int vectors, size_t n, const F *prev_r, F *out, int *nfr_out, double *ps_out) {
    int depth = ilog2(n); size_t total = (size_t)vectors * n; double ps = 0;
    F **tr = (F **)malloc(depth * sizeof(F *)), **in1 = (F **)malloc(depth * sizeof(F *)), **in2 = (F **)malloc(depth * sizeof(F *));
    for (int i = 0; i < depth; i++) {
        size_t sz = total >> (i + 1);
        tr[i] = (F *)malloc(sz * sizeof(F)); in1[i] = (F *)malloc(sz * sizeof(F)); in2[i] = (F *)malloc(sz * sizeof(F));
        const F *src = i ? tr[i - 1] : input;
        for (size_t j = 0; j < sz; j++) { in1[i][j] = src[2 * j]; in2[i][j] = src[2 * j + 1]; tr[i][j] = f_mul(src[2 * j], src[2 * j + 1]); }
    }
    size_t k = 0; F previous_r = *prev_r, sum;
    int maxr = ilog2(total);
    F *r = (F *)malloc((maxr + 1) * sizeof(F)); int nr = 0;
    F *beta = (F *)malloc(total * sizeof(F));
    F *pbuf = (F *)malloc((5 * (size_t)maxr + 8) * sizeof(F));
    for (int v = 0; v < vectors; v++) out[k++] = tr[depth - 1][v];
    size_t hdr = k;                 /* out_eval, final_r, final_eval are filled at the end */
    size_t pk = 0; F *proofs = (F *)malloc(((size_t)depth * (5 * (size_t)maxr + 8)) * sizeof(F));
    F out_eval;
    if (vectors == 1) {
        previous_r = mimc(previous_r, tr[depth - 1][0]);
        sum = tr[depth - 1][0]; out_eval = sum;
        for (int i = depth - 1; i >= 0; i--) {
            if (nr == 0) {
                F num = mimc(previous_r, in1[i][0]);
                previous_r = mimc(num, in2[i][0]);
                sum = f_add(f_mul(f_sub(F1, previous_r), in1[i][0]), f_mul(previous_r, in2[i][0]));
                r[nr++] = previous_r;
            } else {
                orc_precompute_beta(r, nr, beta);
                size_t sz = total >> (i + 1); int rounds = ilog2(sz);
                ps += sumcheck3_inplace(in1[i], in2[i], beta, sz, previous_r, pbuf);
                memcpy(proofs + pk, pbuf, 4 * rounds * sizeof(F)); pk += 4 * rounds;
                memcpy(proofs + pk, pbuf + 5 * rounds, 4 * sizeof(F)); pk += 4;
                previous_r = pbuf[5 * rounds + 3];
                sum = f_add(f_mul(pbuf[5 * rounds], f_sub(F1, previous_r)), f_mul(pbuf[5 * rounds + 1], previous_r));
                r[0] = previous_r; memcpy(r + 1, pbuf + 4 * rounds, rounds * sizeof(F)); nr = rounds + 1;
            }
        }
    } else {
        nr = ilog2(vectors);
        prover_randomness(nr, r);
        sum = evaluate_vector(tr[depth - 1], vectors, r); out_eval = sum;
        previous_r = mimc(r[nr - 1], sum);
        for (int i = depth - 1; i >= 0; i--) {
            orc_precompute_beta(r, nr, beta);
            size_t sz = total >> (i + 1); int rounds = ilog2(sz);
            ps += sumcheck3_inplace(in1[i], in2[i], beta, sz, previous_r, pbuf);
            memcpy(proofs + pk, pbuf, 4 * rounds * sizeof(F)); pk += 4 * rounds;
            memcpy(proofs + pk, pbuf + 5 * rounds, 4 * sizeof(F)); pk += 4;
            previous_r = pbuf[5 * rounds + 3];
            sum = f_add(f_mul(pbuf[5 * rounds], f_sub(F1, previous_r)), f_mul(pbuf[5 * rounds + 1], previous_r));
            r[0] = previous_r; memcpy(r + 1, pbuf + 4 * rounds, rounds * sizeof(F)); nr = rounds + 1;
        }
        if (!f_eq(evaluate_vector(input, total, r), sum)) { printf("Error in mul tree final\n"); exit(-1); }
    }
    k = hdr; out[k++] = out_eval;
    for (int i = 0; i < nr; i++) out[k++] = r[i];
    out[k++] = sum;
    memcpy(out + k, proofs, pk * sizeof(F)); k += pk;
    *nfr_out = nr; *ps_out = ps;
    for (int i = 0; i < depth; i++) { free(tr[i]); free(in1[i]); free(in2[i]); }
    free(tr); free(in1); free(in2); free(r); free(beta); free(pbuf); free(proofs);
    return k;
}

/* ------------------------------------------------------------------ S4 ---
 * generate_3product_sumcheck_beta_stream_batch_optimized (sumcheck.cpp:1150-1392) with batches = 1, distance = 1 — the
 * configuration prove_multiplication_tree_stream_shallow uses whenever layers <= distance (all MLP/AES configs).
 * The witness stream is given in its logical two-half form xy = [X | Y] (witness_stream.cpp:2276-2311 emits X-block | Y-block per
 * read; read_mul_tree_data (:2461-2510) multiplies 2^layer-element segments inside each half).  With A = [seg(X) | seg(Y)]
 * (S = total >> layer entries), E[j] = A[2j], O[j] = A[2j+1] split in blocks of B, the reference processes the blocks in the
 * order X0 | Y0, X1, Y1, ... but indexes eq(r_high) by the NATURAL block number, and de-interleaves R afterwards
 * (permute_partial_evals, :1138-1148).  libc draws: a = generate_randomness(1), b = generate_randomness(2), pad = random(). */
static F *layer_array(const F *xy, size_t total, int layer) {
    size_t S = total >> layer, seg = (size_t)1 << layer;
    F *A = (F *)malloc(S * sizeof(F));
    for (size_t i = 0; i < S; i++) { F p = xy[i * seg]; for (size_t j = 1; j < seg; j++) p = f_mul(p, xy[i * seg + j]); A[i] = p; }
    return A;       /* segments never straddle the X|Y boundary: total/2 is a multiple of seg */
}
int orc_stream_sumcheck_layer(const F *xy, size_t total, size_t B, int layer_id, const F *r, int nr, const F *old_claim,
                              F *new_claim, F *new_r, double *ps_out) {
    (void)nr;
    size_t S = total >> layer_id, nb = S / (2 * B);
    int lgB = ilog2(B), lgnb = ilog2(nb);
    F *A = layer_array(xy, total, layer_id);
    F *eq_low = (F *)malloc(B * sizeof(F)), *eq_high = (F *)malloc(nb * sizeof(F));
    orc_precompute_beta(r, lgB, eq_low); orc_precompute_beta(r + lgB, lgnb, eq_high);
    F *f1 = (F *)malloc(3 * B * sizeof(F)), *f2 = f1 + B, *f3 = f2 + B;
    double ps = 0;
    F Kp = F0;
    for (size_t j = 0; j < B; j++) { f1[j] = A[2 * j]; f2[j] = A[2 * j + 1]; f3[j] = eq_low[j]; Kp = f_add(Kp, f_mul(f_mul(f1[j], f2[j]), f3[j])); }
    F a; prover_randomness(1, &a);
    F Kf = f_mul(a, Kp);
    Kp = f_mul(Kp, eq_high[0]);
    ps += 2 * 16 / 1024.0;
    F *R = (F *)malloc(nb * sizeof(F)); size_t nR = 0; R[nR++] = F1;
    for (size_t step = 1; step < nb; step++) {
        /* processing order X0 | Y0, X1, Y1, ...: step 1 -> Y0, then (Xc, Yc) */
        size_t c = (step + 1) / 2, g = (step % 2) ? nb / 2 + (step - 1) / 2 : step / 2;
        (void)c;
        const F *blk = A + 2 * g * B;
        F K1 = F0, K2 = F0, K3 = F0;
        for (size_t k = 0; k < B; k++) {
            F b1 = blk[2 * k], b2 = blk[2 * k + 1], b3 = eq_low[k];
            F t1 = f_add(f_mul(b1, f2[k]), f_mul(b2, f1[k])), t2 = f_mul(b1, b2);
            K1 = f_add(K1, f_add(f_mul(f3[k], t1), f_mul(f_mul(b3, f1[k]), f2[k])));
            K2 = f_add(K2, f_add(f_mul(b3, t1), f_mul(f3[k], t2)));
            K3 = f_add(K3, f_mul(t2, b3));
        }
        K1 = f_mul(a, K1); K2 = f_mul(a, K2);
        F rand = R[nR - 1];
        rand = mimc(K1, rand); rand = mimc(K2, rand); rand = mimc(K3, rand);
        F x1 = rand, x2 = f_mul(rand, x1), x3 = f_mul(rand, x2);
        Kp = f_add(Kp, f_mul(eq_high[g], K3));
        Kf = f_add(Kf, f_mul(f_mul(x3, a), K3));
        Kf = f_add(Kf, f_add(f_mul(x2, K2), f_mul(x1, K1)));
        R[nR++] = rand;
        ps += 2 * 16 / 1024.0;
        for (size_t k = 0; k < B; k++) {
            f1[k] = f_add(f1[k], f_mul(rand, blk[2 * k])); f2[k] = f_add(f2[k], f_mul(rand, blk[2 * k + 1])); f3[k] = f_add(f3[k], f_mul(rand, eq_low[k]));
        }
    }
    if (!f_eq(Kp, *old_claim)) printf("Error in sumcheck 0 %d\n", 0);
    F *p1 = (F *)malloc((5 * (size_t)lgB + 8) * sizeof(F));
    size_t szB = B;
    ps += orc_batch_sumcheck3(f1, f2, f3, &szB, 1, &a, p1);
    {   /* P1.c_poly[0].eval(0) + eval(1) == Kf */
        F s = f_add(f_add(f_add(p1[0], p1[1]), f_add(p1[2], p1[3])), p1[3]);
        if (!f_eq(s, Kf)) { printf("Error in sumcheck 1\n"); exit(-1); }
    }
    const F *P1r = p1 + 4 * lgB, *P1vr = p1 + 5 * lgB;
    F *beta = (F *)malloc(B * sizeof(F));
    orc_precompute_beta(P1r, lgB, beta);
    F *PE0 = (F *)malloc(2 * nb * sizeof(F)), *PE1 = PE0 + nb;
    for (size_t g = 0; g < nb; g++) {
        F s0 = F0, s1 = F0; const F *blk = A + 2 * g * B;
        for (size_t j = 0; j < B; j++) { s0 = f_add(s0, f_mul(beta[j], blk[2 * j])); s1 = f_add(s1, f_mul(beta[j], blk[2 * j + 1])); }
        PE0[g] = s0; PE1[g] = s1;
    }
    F *Rp = (F *)malloc(nb * sizeof(F)); size_t cnt = 0;
    for (size_t i = 0; i < nb / 2; i++) Rp[cnt++] = R[2 * i];
    for (size_t i = 0; i < nb / 2; i++) Rp[cnt++] = R[2 * i + 1];
    F b[2]; prover_randomness(2, b);
    F *aggr = (F *)malloc(nb * sizeof(F));
    for (size_t j = 0; j < nb; j++) aggr[j] = f_add(f_mul(b[0], PE0[j]), f_mul(b[1], PE1[j]));
    F *p2 = (F *)malloc((4 * (size_t)lgnb + 8) * sizeof(F));
    F zero = F0;
    ps += orc_sumcheck2(Rp, aggr, nb, &zero, p2);
    {
        F sum = f_add(f_mul(b[0], P1vr[0]), f_mul(b[1], P1vr[1]));
        F q = f_add(f_add(p2[0], p2[1]), f_add(p2[2], p2[2]));
        if (!f_eq(sum, q)) { printf("Error in sumcheck 2\n"); exit(-1); }
    }
    F pad = prover_random();
    size_t k = 0; new_r[k++] = pad;
    for (int j = 0; j < lgB; j++) new_r[k++] = P1r[j];
    const F *P2r = p2 + 3 * lgnb;
    for (int j = 0; j < lgnb; j++) new_r[k++] = P2r[j];
    *new_claim = f_add(f_mul(f_sub(F1, pad), evaluate_vector(PE0, nb, P2r)), f_mul(pad, evaluate_vector(PE1, nb, P2r)));
    *ps_out = ps;
    free(A); free(eq_low); free(eq_high); free(f1); free(R); free(p1); free(beta); free(PE0); free(Rp); free(aggr); free(p2);
    return (int)k;
}

/* ------------------------------------------------------------------ S4, batched ---
 * generate_3product_sumcheck_beta_stream_batch_optimized (sumcheck.cpp:1150-1392) with `batches` layers proven together: batch j is the
 * product-tree layer layer_id + j*distance with block size B >> (j*distance) (read_mul_tree_data, witness_stream.cpp:2461-2510: every
 * batch has the same number of half-chunks nb).  r[j]: the point of batch j (log2(B_j) + log2(nb) entries used).
 * libc draws: a = generate_randomness(batches), b = generate_randomness(2*batches), pad = random().
 * new_r[j] gets 1 + log2(B_j) + log2(nb) entries (row stride rs).  Returns the entries written for batch 0. */
static int stream_sumcheck_batch(const F *xy, size_t total, size_t B, int layer_id, int distance, int batches, const F *r, int rs,
                                 const F *old_claims, F *new_claims, F *new_r, double *ps_out) {
    enum { MAXB = 8 };
    if (batches > MAXB) { printf("stream_sumcheck_batch: too many batches\n"); exit(-1); }
    size_t S0 = total >> layer_id, nb = S0 / (2 * B);
    int lgnb = ilog2(nb);
    F *A[MAXB], *eq_low[MAXB], *eq_high[MAXB], *f1[MAXB], *f2[MAXB], *f3[MAXB];
    size_t Bj[MAXB]; int lgBj[MAXB];
    F Kp[MAXB], a[MAXB];
    double ps = 0;
    for (int j = 0; j < batches; j++) {
        Bj[j] = B >> (j * distance); lgBj[j] = ilog2(Bj[j]);
        A[j] = layer_array(xy, total, layer_id + j * distance);
        eq_low[j] = (F *)malloc(Bj[j] * sizeof(F)); eq_high[j] = (F *)malloc(nb * sizeof(F));
        orc_precompute_beta(r + (size_t)j * rs, lgBj[j], eq_low[j]); orc_precompute_beta(r + (size_t)j * rs + lgBj[j], lgnb, eq_high[j]);
        f1[j] = (F *)malloc(3 * Bj[j] * sizeof(F)); f2[j] = f1[j] + Bj[j]; f3[j] = f2[j] + Bj[j];
        Kp[j] = F0;
        for (size_t k = 0; k < Bj[j]; k++) {
            f1[j][k] = A[j][2 * k]; f2[j][k] = A[j][2 * k + 1]; f3[j][k] = eq_low[j][k];
            Kp[j] = f_add(Kp[j], f_mul(f_mul(f1[j][k], f2[j][k]), f3[j][k]));
        }
    }
    prover_randomness(batches, a);
    F Kf = F0;
    for (int j = 0; j < batches; j++) { Kf = f_add(Kf, f_mul(a[j], Kp[j])); Kp[j] = f_mul(Kp[j], eq_high[j][0]); }
    ps += (1 + batches) * 16 / 1024.0;
    F *R = (F *)malloc(nb * sizeof(F)); size_t nR = 0; R[nR++] = F1;
    for (size_t step = 1; step < nb; step++) {
        size_t g = (step % 2) ? nb / 2 + (step - 1) / 2 : step / 2;       /* X0 | Y0, X1, Y1, ...: natural half-chunk index */
        F K1 = F0, K2 = F0, K3[MAXB];
        for (int j = 0; j < batches; j++) {
            const F *blk = A[j] + 2 * g * Bj[j];
            F k1 = F0, k2 = F0; K3[j] = F0;
            for (size_t k = 0; k < Bj[j]; k++) {
                F b1 = blk[2 * k], b2 = blk[2 * k + 1], b3 = eq_low[j][k];
                F t1 = f_add(f_mul(b1, f2[j][k]), f_mul(b2, f1[j][k])), t2 = f_mul(b1, b2);
                k1 = f_add(k1, f_add(f_mul(f3[j][k], t1), f_mul(f_mul(b3, f1[j][k]), f2[j][k])));
                k2 = f_add(k2, f_add(f_mul(b3, t1), f_mul(f3[j][k], t2)));
                K3[j] = f_add(K3[j], f_mul(t2, b3));
            }
            K1 = f_add(K1, f_mul(a[j], k1)); K2 = f_add(K2, f_mul(a[j], k2));
        }
        F rand = R[nR - 1];
        rand = mimc(K1, rand); rand = mimc(K2, rand);
        for (int j = 0; j < batches; j++) rand = mimc(K3[j], rand);
        F x1 = rand, x2 = f_mul(rand, x1), x3 = f_mul(rand, x2);
        for (int j = 0; j < batches; j++) { Kp[j] = f_add(Kp[j], f_mul(eq_high[j][g], K3[j])); Kf = f_add(Kf, f_mul(f_mul(x3, a[j]), K3[j])); }
        Kf = f_add(Kf, f_add(f_mul(x2, K2), f_mul(x1, K1)));
        R[nR++] = rand;
        ps += (1 + batches) * 16 / 1024.0;
        for (int j = 0; j < batches; j++) {
            const F *blk = A[j] + 2 * g * Bj[j];
            for (size_t k = 0; k < Bj[j]; k++) {
                f1[j][k] = f_add(f1[j][k], f_mul(rand, blk[2 * k])); f2[j][k] = f_add(f2[j][k], f_mul(rand, blk[2 * k + 1]));
                f3[j][k] = f_add(f3[j][k], f_mul(rand, eq_low[j][k]));
            }
        }
    }
    for (int j = 0; j < batches; j++) if (!f_eq(Kp[j], old_claims[j])) printf("Error in sumcheck 0 %d\n", j);
    /* batch_3product_sumcheck over the folds */
    size_t tot = 0, sizes[MAXB]; for (int j = 0; j < batches; j++) { sizes[j] = Bj[j]; tot += Bj[j]; }
    F *t1 = (F *)malloc(3 * tot * sizeof(F)), *t2 = t1 + tot, *t3 = t2 + tot; size_t off = 0;
    for (int j = 0; j < batches; j++) { memcpy(t1 + off, f1[j], Bj[j] * sizeof(F)); memcpy(t2 + off, f2[j], Bj[j] * sizeof(F)); memcpy(t3 + off, f3[j], Bj[j] * sizeof(F)); off += Bj[j]; }
    int lgB = lgBj[0];
    F *p1 = (F *)malloc((5 * (size_t)lgB + 3 * batches + 8) * sizeof(F));
    ps += orc_batch_sumcheck3(t1, t2, t3, sizes, batches, a, p1);
    free(t1);
    {
        F s = f_add(f_add(f_add(p1[0], p1[1]), f_add(p1[2], p1[3])), p1[3]);
        if (!f_eq(s, Kf)) { printf("Error in sumcheck 1\n"); exit(-1); }
    }
    const F *P1r = p1 + 4 * lgB, *P1vr = p1 + 5 * lgB;
    if (nb < 2) { printf("stream_sumcheck_batch: single-chunk layers are handled by the in-memory prover\n"); exit(-1); }
    F *PE = (F *)malloc(2 * (size_t)batches * nb * sizeof(F));        /* PE[(2j+h)*nb + g] */
    for (int j = 0; j < batches; j++) {
        F *beta = (F *)malloc(Bj[j] * sizeof(F));
        orc_precompute_beta(P1r, lgBj[j], beta);
        for (size_t g = 0; g < nb; g++) {
            F s0 = F0, s1 = F0; const F *blk = A[j] + 2 * g * Bj[j];
            for (size_t k = 0; k < Bj[j]; k++) { s0 = f_add(s0, f_mul(beta[k], blk[2 * k])); s1 = f_add(s1, f_mul(beta[k], blk[2 * k + 1])); }
            PE[(2 * j) * nb + g] = s0; PE[(2 * j + 1) * nb + g] = s1;
        }
        free(beta);
    }
    F *Rp = (F *)malloc(nb * sizeof(F)); size_t cnt = 0;
    for (size_t i = 0; i < nb / 2; i++) Rp[cnt++] = R[2 * i];
    for (size_t i = 0; i < nb / 2; i++) Rp[cnt++] = R[2 * i + 1];
    F b[2 * MAXB]; prover_randomness(2 * batches, b);
    F *aggr = (F *)malloc(nb * sizeof(F));
    for (size_t g = 0; g < nb; g++) { aggr[g] = F0; for (int i = 0; i < 2 * batches; i++) aggr[g] = f_add(aggr[g], f_mul(b[i], PE[(size_t)i * nb + g])); }
    F *p2 = (F *)malloc((4 * (size_t)lgnb + 8) * sizeof(F));
    F zero = F0;
    ps += orc_sumcheck2(Rp, aggr, nb, &zero, p2);
    {
        F sum = F0;
        for (int i = 0; i < batches; i++) sum = f_add(sum, f_add(f_mul(b[2 * i], P1vr[3 * i]), f_mul(b[2 * i + 1], P1vr[3 * i + 1])));
        F q = f_add(f_add(p2[0], p2[1]), f_add(p2[2], p2[2]));
        if (!f_eq(sum, q)) { printf("Error in sumcheck 2\n"); exit(-1); }
    }
    F pad = prover_random();
    const F *P2r = p2 + 3 * lgnb;
    int n0 = 0;
    for (int j = 0; j < batches; j++) {
        F *nr = new_r + (size_t)j * rs; int k = 0;
        nr[k++] = pad;
        for (int q = 0; q < lgBj[j]; q++) nr[k++] = P1r[q];
        for (int q = 0; q < lgnb; q++) nr[k++] = P2r[q];
        if (j == 0) n0 = k;
        new_claims[j] = f_add(f_mul(f_sub(F1, pad), evaluate_vector(PE + (size_t)(2 * j) * nb, nb, P2r)), f_mul(pad, evaluate_vector(PE + (size_t)(2 * j + 1) * nb, nb, P2r)));
    }
    *ps_out = ps;
    for (int j = 0; j < batches; j++) { free(A[j]); free(eq_low[j]); free(eq_high[j]); free(f1[j]); }
    free(R); free(p1); free(PE); free(Rp); free(aggr); free(p2);
    return n0;
}
int orc_stream_sumcheck_batch(const F *xy, size_t total, size_t B, int layer_id, int distance, int batches, const F *r, int rs, const int *rlen,
                              const F *old_claims, F *new_claims, F *new_r, double *ps_out) {
    (void)rlen;
    return stream_sumcheck_batch(xy, total, B, layer_id, distance, batches, r, rs, old_claims, new_claims, new_r, ps_out);
}
/* generate_claims_opt (sumcheck.cpp:1014-1055): claims[j] = sum over the layer layer_id + j*distance of eq(r)(x) * A[2x] * A[2x+1], with the
 * half-chunk order of the two-half stream (chunk i <-> eq_high[i] for X halves, eq_high[i + nb/2] for Y halves == natural index) */
static void generate_claims_opt(const F *xy, size_t total, size_t B, int layer_id, int distance, int batches, const F *r, F *claims) {
    size_t S0 = total >> layer_id, nb = S0 / (2 * B); int lgnb = ilog2(nb);
    for (int j = 0; j < batches; j++) {
        size_t Bj = B >> (j * distance); int lgBj = ilog2(Bj);
        F *A = layer_array(xy, total, layer_id + j * distance);
        F *lo = (F *)malloc(Bj * sizeof(F)), *hi = (F *)malloc(nb * sizeof(F));
        orc_precompute_beta(r, lgBj, lo); orc_precompute_beta(r + lgBj, lgnb, hi);
        F c = F0;
        for (size_t g = 0; g < nb; g++) {
            F s = F0; const F *blk = A + 2 * g * Bj;
            for (size_t k = 0; k < Bj; k++) s = f_add(s, f_mul(f_mul(lo[k], blk[2 * k]), blk[2 * k + 1]));
            c = f_add(c, f_mul(hi[g], s));
        }
        claims[j] = c;
        free(A); free(lo); free(hi);
    }
}

/* ------------------------------------------------------------------ S6 ---
 * prove_multiplication_tree_stream_shallow (sumcheck.cpp:1746-1915), branches: whole stream fits (size*vectors <= 2B) and
 * layers <= distance (or naive).  Returns ps; out = the `vectors` products. */
double orc_mul_tree_stream(const F *xy, size_t total, int vectors, size_t B, int distance, int naive, const F *prev_r, F *out) {
    int maxr = ilog2(total);
    F *buf = (F *)malloc((64 + (size_t)vectors + 8 * (size_t)(maxr + 2) * (maxr + 2)) * sizeof(F));
    int nfr; double ps = 0;
    if (total <= 2 * B) {
        orc_mul_tree(xy, vectors, total / vectors, prev_r, buf, &nfr, &ps);
        memcpy(out, buf, vectors * sizeof(F)); free(buf); return ps;
    }
    int layers = ilog2(total / (2 * B));
    if (layers % distance != 0 && layers > distance) layers = distance + layers - (layers % distance);
    F *top = layer_array(xy, total, layers);
    size_t St = total >> layers;
    orc_mul_tree(top, vectors, St / vectors, prev_r, buf, &nfr, &ps);
    free(top);
    memcpy(out, buf, vectors * sizeof(F));
    F claim = buf[vectors + 1 + nfr];                 /* final_eval */
    F *r = (F *)malloc((maxr + 2) * sizeof(F)), *nr = (F *)malloc((maxr + 2) * sizeof(F));
    memcpy(r, buf + vectors + 1, nfr * sizeof(F));    /* vectors > 1: individual ++ global == final_r */
    int n = nfr;
    if (layers <= distance || naive) {
        for (int i = layers - 1; i >= 0; i--) {
            F nc; double p = 0;
            n = orc_stream_sumcheck_layer(xy, total, B, i, r, n, &claim, &nc, nr, &p);
            ps += p; claim = nc; memcpy(r, nr, n * sizeof(F));
        }
    } else {
        /* layers > distance (:1871-1908): `batches` layers, `distance` apart, are proven together.  NOT included here: commit_layers /
         * open_layers (:983-1011), i.e. the Elastic_PC commit + open of the intermediate layers — their only trace in the reference's
         * outputs is ps and libc draws; the host mirror runs them around this call (hobbit_host.cpp). */
        int batches = layers / distance, rs = maxr + 2;
        int extra = ilog2(total >> distance) - nfr;                  /* r_temp = individual | global | generate_randomness(extra) */
        prover_randomness(extra, r + nfr);
        F *rb = (F *)calloc((size_t)batches * rs, sizeof(F)), *nrb = (F *)calloc((size_t)batches * rs, sizeof(F));
        for (int j = 0; j < batches; j++) memcpy(rb + (size_t)j * rs, r, (nfr + extra) * sizeof(F));
        F claims[8], nclaims[8];
        generate_claims_opt(xy, total, B, distance - 1, distance, batches, r, claims);
        for (int i = distance - 1; i >= 0; i--) {
            double p = 0;
            stream_sumcheck_batch(xy, total, B, i, distance, batches, rb, rs, claims, nclaims, nrb, &p);
            ps += p; memcpy(claims, nclaims, sizeof claims); memcpy(rb, nrb, (size_t)batches * rs * sizeof(F));
        }
        free(rb); free(nrb);
    }
    free(r); free(nr); free(buf);
    return ps;
}

/* C1 split for sharding (same arithmetic as orc_commit_standard): inner digests of `nchunks` chunks, and the chain. */
void orc_commit_encode_chunks(const F *poly, size_t nchunks, size_t B, int trs, int lin, uint8_t *inner_out) {
    size_t cols = 2 * B / trs;
    F *T = (F *)malloc(4 * B * sizeof(F));
    for (size_t c = 0; c < nchunks; c++) {
        orc_compute_tensorcode(poly + c * B, B, trs, lin, T);
        for (int j = 0; j < trs / 2; j++)
            for (size_t k = 0; k < cols; k++) {
                F q[4] = { T[(4 * j) * cols + k], T[(4 * j + 1) * cols + k], T[(4 * j + 2) * cols + k], T[(4 * j + 3) * cols + k] };
                orc_blake3_hash((const uint8_t *)q, inner_out + (c * B + j * cols + k) * 32);
            }
    }
    free(T);
}
void orc_md_chain(const uint8_t *inner, size_t nchunks, size_t nleaves, uint8_t *leaves) {
    for (size_t c = 0; c < nchunks; c++)
        for (size_t p = 0; p < nleaves; p++) {
            uint8_t data[64];
            memcpy(data, inner + (c * nleaves + p) * 32, 32);
            memcpy(data + 32, leaves + p * 32, 32);
            orc_blake3_hash(data, leaves + p * 32);
        }
}

/* -------------------------------------------------------- gate consistency ---
 * prove_gate_consistency_standard (sumcheck.cpp:434-501): degree-4 sumcheck of
 *     sum_x beta(x) * ( mul(x) L(x) R(x) + add(x) (L(x) + R(x)) - O(x) ),  mul = 1 - add, beta = eq(r),
 * rand starts at F(213); per round rand = mimc(e, mimc(d, mimc(c, mimc(b, mimc(a, rand))))) with the (value, rand) argument order;
 * the tables are folded with the new rand.  The reference returns nothing (tables are folded in place); out gets, per round,
 * (a,b,c,d,e) and rand [6*rounds], then the final add, L, R, O, mul, beta values [6]. */
void orc_gate_consistency_standard(const F *L_in, const F *R_in, const F *O_in, const F *add_in, size_t n, const F *r, F *out) {
    int rounds = ilog2(n);
    F *t = (F *)malloc(6 * n * sizeof(F));
    F *A = t, *Bt = t + n, *L = t + 2 * n, *R = t + 3 * n, *O = t + 4 * n, *M = t + 5 * n;
    memcpy(A, add_in, n * sizeof(F)); memcpy(L, L_in, n * sizeof(F)); memcpy(R, R_in, n * sizeof(F)); memcpy(O, O_in, n * sizeof(F));
    for (size_t i = 0; i < n; i++) M[i] = f_sub(F1, A[i]);
    orc_precompute_beta(r, rounds, Bt);
    F rand = f_int(213);
    size_t k = 0;
    for (int i = 0; i < rounds; i++) {
        size_t half = (size_t)1 << (rounds - 1 - i);
        F p[5] = { F0, F0, F0, F0, F0 };             /* a (X^4) .. e (X^0) */
        for (size_t j = 0; j < half; j++) {
            /* linear factors: value at 0 and slope */
            F a0 = A[2 * j], a1 = f_sub(A[2 * j + 1], a0), m0 = M[2 * j], m1 = f_sub(M[2 * j + 1], m0);
            F b0 = Bt[2 * j], b1 = f_sub(Bt[2 * j + 1], b0), l0 = L[2 * j], l1 = f_sub(L[2 * j + 1], l0);
            F r0 = R[2 * j], r1 = f_sub(R[2 * j + 1], r0), o0 = O[2 * j], o1 = f_sub(O[2 * j + 1], o0);
            /* q(X) = m(X) l(X) r(X) + a(X) (l(X) + r(X)) - o(X)  (cubic: q3..q0) */
            F ml2 = f_mul(m1, l1), ml1 = f_add(f_mul(m1, l0), f_mul(m0, l1)), ml0 = f_mul(m0, l0);
            F q3 = f_mul(ml2, r1);
            F q2 = f_add(f_mul(ml2, r0), f_mul(ml1, r1));
            F q1 = f_add(f_mul(ml1, r0), f_mul(ml0, r1));
            F q0 = f_mul(ml0, r0);
            F s0 = f_add(l0, r0), s1 = f_add(l1, r1);
            q2 = f_add(q2, f_mul(a1, s1));
            q1 = f_add(q1, f_add(f_mul(a1, s0), f_mul(a0, s1)));
            q0 = f_add(q0, f_mul(a0, s0));
            q1 = f_sub(q1, o1); q0 = f_sub(q0, o0);
            /* times beta(X) */
            p[0] = f_add(p[0], f_mul(b1, q3));
            p[1] = f_add(p[1], f_add(f_mul(b1, q2), f_mul(b0, q3)));
            p[2] = f_add(p[2], f_add(f_mul(b1, q1), f_mul(b0, q2)));
            p[3] = f_add(p[3], f_add(f_mul(b1, q0), f_mul(b0, q1)));
            p[4] = f_add(p[4], f_mul(b0, q0));
        }
        for (int c = 0; c < 5; c++) { rand = mimc(p[c], rand); out[k++] = p[c]; }
        out[k++] = rand;
        for (size_t j = 0; j < half; j++) {
            A[j] = f_add(A[2 * j], f_mul(rand, f_sub(A[2 * j + 1], A[2 * j])));
            L[j] = f_add(L[2 * j], f_mul(rand, f_sub(L[2 * j + 1], L[2 * j])));
            R[j] = f_add(R[2 * j], f_mul(rand, f_sub(R[2 * j + 1], R[2 * j])));
            O[j] = f_add(O[2 * j], f_mul(rand, f_sub(O[2 * j + 1], O[2 * j])));
            M[j] = f_add(M[2 * j], f_mul(rand, f_sub(M[2 * j + 1], M[2 * j])));
            Bt[j] = f_add(Bt[2 * j], f_mul(rand, f_sub(Bt[2 * j + 1], Bt[2 * j])));
        }
    }
    out[k++] = A[0]; out[k++] = L[0]; out[k++] = R[0]; out[k++] = O[0]; out[k++] = M[0]; out[k++] = Bt[0];
    free(t);
}

/* ------------------------------------------------------------------ S7 ---
 * prove_gate_consistency (sumcheck.cpp:796-981, has_lookups = false) with the transcript stream given as resident arrays
 * L, R, O, S (S[i] = F(1) for an add gate, F(0) for a mul gate; read_trace, witness_stream.cpp:1701-1807), cs entries, buffer B.
 * libc draws: a = generate_randomness(4) after the streaming pass, b = generate_randomness(6) after the partial evaluations.
 * The reference returns nothing; out gets  R[nch] | (a,b,c,d,e,rand) x log2 B | final L,R,O,add,mul,beta | Peval[6][nch] |
 * P2 flat proof (4*log2 nch + 3).  Returns ps. */
double orc_gate_consistency_stream(const F *L, const F *R, const F *O, const F *S, size_t cs, size_t B, const F *r, F *out) {
    size_t nch = cs / B; int lgB = ilog2(B), lgn = ilog2(nch);
    double ps = 0; size_t k = 0;
    F *beta = (F *)malloc(B * sizeof(F));
    orc_precompute_beta(r, lgB, beta);
    F *fL = (F *)malloc(6 * B * sizeof(F)), *fR = fL + B, *fO = fR + B, *fA = fO + B, *fM = fA + B, *fB = fM + B;
    F KfO = F0, KfL = F0, KfR = F0, KfM = F0;
    for (size_t i = 0; i < B; i++) {
        fL[i] = L[i]; fR[i] = R[i]; fO[i] = O[i]; fA[i] = S[i]; fM[i] = f_sub(F1, S[i]); fB[i] = beta[i];
        KfO = f_add(KfO, f_mul(beta[i], fO[i]));
        KfL = f_add(KfL, f_mul(f_mul(beta[i], fL[i]), fA[i]));
        KfR = f_add(KfR, f_mul(f_mul(beta[i], fR[i]), fA[i]));
        KfM = f_add(KfM, f_mul(f_mul(f_mul(beta[i], fR[i]), fL[i]), fM[i]));
    }
    ps += 4 * 16 / 1024.0;
    F *Rv = (F *)malloc(nch * sizeof(F)); size_t nR = 0; Rv[nR++] = F1;
    F rand = F0;
    for (size_t c = 1; c < nch; c++) {
        const F *bL = L + c * B, *bR = R + c * B, *bO = O + c * B, *bS = S + c * B;
        F K1O = F0, K2O = F0, K1L = F0, K2L = F0, K3L = F0, K1R = F0, K2R = F0, K3R = F0, K1M = F0, K2M = F0, K3M = F0, K4M = F0;
        for (size_t i = 0; i < B; i++) {
            F gate = bS[i], ngate = f_sub(F1, bS[i]), t1, t2, t3, t4, t5, t6;
            K1O = f_add(K1O, f_add(f_mul(bO[i], fB[i]), f_mul(beta[i], fO[i])));
            K2O = f_add(K2O, f_mul(bO[i], beta[i]));
            t1 = f_add(f_mul(bL[i], fA[i]), f_mul(gate, fL[i])); t2 = f_mul(bL[i], gate);
            K1L = f_add(K1L, f_add(f_mul(fB[i], t1), f_mul(f_mul(beta[i], fL[i]), fA[i])));
            K2L = f_add(K2L, f_add(f_mul(beta[i], t1), f_mul(fB[i], t2)));
            K3L = f_add(K3L, f_mul(t2, beta[i]));
            t1 = f_add(f_mul(bR[i], fA[i]), f_mul(gate, fR[i])); t2 = f_mul(bR[i], gate);
            K1R = f_add(K1R, f_add(f_mul(fB[i], t1), f_mul(f_mul(beta[i], fR[i]), fA[i])));
            K2R = f_add(K2R, f_add(f_mul(beta[i], t1), f_mul(fB[i], t2)));
            K3R = f_add(K3R, f_mul(t2, beta[i]));
            t1 = f_add(f_mul(fL[i], bR[i]), f_mul(fR[i], bL[i]));
            t2 = f_add(f_mul(fB[i], ngate), f_mul(fM[i], beta[i]));
            t3 = f_mul(bL[i], bR[i]); t4 = f_mul(ngate, beta[i]); t5 = f_mul(fL[i], fR[i]); t6 = f_mul(fB[i], fM[i]);
            K1M = f_add(K1M, f_add(f_mul(t1, t6), f_mul(t2, t5)));
            K2M = f_add(K2M, f_add(f_add(f_mul(t1, t2), f_mul(t3, t6)), f_mul(t4, t5)));
            K3M = f_add(K3M, f_add(f_mul(t1, t4), f_mul(t2, t3)));
            K4M = f_add(K4M, f_mul(t3, t4));
        }
        if (!f_eq(f_sub(f_add(f_add(K4M, K3L), K3R), K2O), F0)) { printf("Error in gate consistency 1 : %d\n", (int)c); exit(-1); }
        rand = mimc(K1O, rand); rand = mimc(K2O, rand); rand = mimc(K1L, rand); rand = mimc(K2L, rand); rand = mimc(K3L, rand);
        rand = mimc(K1R, rand); rand = mimc(K2R, rand); rand = mimc(K3R, rand);
        Rv[nR++] = rand;
        F x1 = rand, x2 = f_mul(rand, x1), x3 = f_mul(rand, x2), x4 = f_mul(rand, x3);
        KfO = f_add(KfO, f_add(f_mul(x1, K1O), f_mul(x2, K2O)));
        KfL = f_add(KfL, f_add(f_add(f_mul(x1, K1L), f_mul(x2, K2L)), f_mul(x3, K3L)));
        KfR = f_add(KfR, f_add(f_add(f_mul(x1, K1R), f_mul(x2, K2R)), f_mul(x3, K3R)));
        KfM = f_add(KfM, f_add(f_add(f_add(f_mul(x1, K1M), f_mul(x2, K2M)), f_mul(x3, K3M)), f_mul(x4, K4M)));
        ps += 12 * 16 / 1024.0;
        for (size_t i = 0; i < B; i++) {
            fA[i] = f_add(fA[i], f_mul(rand, bS[i])); fL[i] = f_add(fL[i], f_mul(rand, bL[i])); fR[i] = f_add(fR[i], f_mul(rand, bR[i]));
            fO[i] = f_add(fO[i], f_mul(rand, bO[i])); fM[i] = f_add(fM[i], f_mul(rand, f_sub(F1, bS[i]))); fB[i] = f_add(fB[i], f_mul(rand, beta[i]));
        }
    }
    for (size_t i = 0; i < nch; i++) out[k++] = Rv[i];
    F a[4]; prover_randomness(4, a);
    F sum = f_add(f_add(f_mul(a[0], KfL), f_mul(a[1], KfR)), f_add(f_mul(a[2], KfM), f_mul(KfO, a[3])));
    F *srand_ = (F *)malloc(lgB * sizeof(F));
    for (int i = lgB - 1, q = 0; i >= 0; i--, q++) {
        F p[5] = { F0, F0, F0, F0, F0 };
        for (size_t j = 0; j < ((size_t)1 << i); j++) {
            F a0 = fA[2 * j], a1 = f_sub(fA[2 * j + 1], a0), m0 = fM[2 * j], m1 = f_sub(fM[2 * j + 1], m0);
            F b0 = fB[2 * j], b1 = f_sub(fB[2 * j + 1], b0), l0 = fL[2 * j], l1 = f_sub(fL[2 * j + 1], l0);
            F r0 = fR[2 * j], r1 = f_sub(fR[2 * j + 1], r0), o0 = fO[2 * j], o1 = f_sub(fO[2 * j + 1], o0);
            F ml2 = f_mul(m1, l1), ml1 = f_add(f_mul(m1, l0), f_mul(m0, l1)), ml0 = f_mul(m0, l0);
            F q3 = f_mul(a[2], f_mul(ml2, r1));
            F q2 = f_mul(a[2], f_add(f_mul(ml2, r0), f_mul(ml1, r1)));
            F q1 = f_mul(a[2], f_add(f_mul(ml1, r0), f_mul(ml0, r1)));
            F q0 = f_mul(a[2], f_mul(ml0, r0));
            F s0 = f_add(f_mul(a[0], l0), f_mul(a[1], r0)), s1 = f_add(f_mul(a[0], l1), f_mul(a[1], r1));
            q2 = f_add(q2, f_mul(a1, s1));
            q1 = f_add(q1, f_add(f_mul(a1, s0), f_mul(a0, s1)));
            q0 = f_add(q0, f_mul(a0, s0));
            q1 = f_add(q1, f_mul(a[3], o1)); q0 = f_add(q0, f_mul(a[3], o0));
            p[0] = f_add(p[0], f_mul(b1, q3));
            p[1] = f_add(p[1], f_add(f_mul(b1, q2), f_mul(b0, q3)));
            p[2] = f_add(p[2], f_add(f_mul(b1, q1), f_mul(b0, q2)));
            p[3] = f_add(p[3], f_add(f_mul(b1, q0), f_mul(b0, q1)));
            p[4] = f_add(p[4], f_mul(b0, q0));
        }
        for (int c = 0; c < 5; c++) { rand = mimc(p[c], rand); out[k++] = p[c]; }
        out[k++] = rand;
        {   /* sum == poly(0) + poly(1) */
            F s = f_add(f_add(f_add(p[0], p[1]), f_add(p[2], p[3])), f_add(p[4], p[4]));
            if (!f_eq(s, sum)) { printf("Error in gate consistency 2: %d\n", i); exit(-1); }
        }
        sum = f_add(f_mul(f_add(f_mul(f_add(f_mul(f_add(f_mul(p[0], rand), p[1]), rand), p[2]), rand), p[3]), rand), p[4]);
        srand_[q] = rand;
        ps += 5 * 16 / 1024.0;
        for (size_t j = 0; j < ((size_t)1 << i); j++) {
            fA[j] = f_add(fA[2 * j], f_mul(rand, f_sub(fA[2 * j + 1], fA[2 * j]))); fL[j] = f_add(fL[2 * j], f_mul(rand, f_sub(fL[2 * j + 1], fL[2 * j])));
            fR[j] = f_add(fR[2 * j], f_mul(rand, f_sub(fR[2 * j + 1], fR[2 * j]))); fO[j] = f_add(fO[2 * j], f_mul(rand, f_sub(fO[2 * j + 1], fO[2 * j])));
            fM[j] = f_add(fM[2 * j], f_mul(rand, f_sub(fM[2 * j + 1], fM[2 * j]))); fB[j] = f_add(fB[2 * j], f_mul(rand, f_sub(fB[2 * j + 1], fB[2 * j])));
        }
    }
    out[k++] = fL[0]; out[k++] = fR[0]; out[k++] = fO[0]; out[k++] = fA[0]; out[k++] = fM[0]; out[k++] = fB[0];
    F *beta1 = (F *)malloc(B * sizeof(F));
    orc_precompute_beta(srand_, lgB, beta1);
    F *Pe = (F *)calloc(6 * nch, sizeof(F));
    for (size_t c = 0; c < nch; c++)
        for (size_t j = 0; j < B; j++) {
            F s = S[c * B + j];
            Pe[0 * nch + c] = f_add(Pe[0 * nch + c], f_mul(beta1[j], L[c * B + j]));
            Pe[1 * nch + c] = f_add(Pe[1 * nch + c], f_mul(beta1[j], R[c * B + j]));
            Pe[2 * nch + c] = f_add(Pe[2 * nch + c], f_mul(beta1[j], O[c * B + j]));
            Pe[3 * nch + c] = f_add(Pe[3 * nch + c], f_mul(beta1[j], s));
            Pe[4 * nch + c] = f_add(Pe[4 * nch + c], f_mul(beta1[j], f_sub(F1, s)));
            Pe[5 * nch + c] = f_add(Pe[5 * nch + c], f_mul(beta1[j], beta[j]));
        }
    for (size_t i = 0; i < 6 * nch; i++) out[k++] = Pe[i];
    F b[6]; prover_randomness(6, b);
    F *pe = (F *)malloc(nch * sizeof(F));
    for (size_t j = 0; j < nch; j++) { pe[j] = F0; for (int i = 0; i < 6; i++) pe[j] = f_add(pe[j], f_mul(b[i], Pe[i * nch + j])); }
    F *p2 = (F *)malloc((4 * (size_t)lgn + 8) * sizeof(F));
    ps += orc_sumcheck2(Rv, pe, nch, &rand, p2);
    ps += 5 * 16 / 1024.0;
    {
        F s2 = f_add(f_add(f_add(f_mul(fL[0], b[0]), f_mul(fR[0], b[1])), f_add(f_mul(fO[0], b[2]), f_mul(b[3], fA[0]))), f_add(f_mul(b[4], fM[0]), f_mul(b[5], fB[0])));
        F qv = lgn ? f_add(f_add(p2[0], p2[1]), f_add(p2[2], p2[2])) : s2;
        if (!f_eq(qv, s2)) { printf("Error in gate consistency 3\n"); exit(-1); }
    }
    for (int i = 0; i < 4 * lgn + 3; i++) out[k++] = p2[i];
    free(beta); free(fL); free(Rv); free(srand_); free(beta1); free(Pe); free(pe); free(p2);
    return ps;
}

/* ------------------------------------------------------------------ S8 ---
 * prove_gate_consistency_lookups (sumcheck.cpp:503-794, has_lookups = true) on a resident transcript: L, R, O and the selector
 * S (F(0) add, F(1) mul, F(2) lookup; read_trace with has_lookups).  lr = lookup_rand[0..1].  Per element the reference derives
 *   gate_L = 1 / 0 / lr0,  gate_R = 1 / 0 / lr1,  gate_mul = 0 / 1 / 0,  gate_lkp = 0 / 0 / 1   for S = 0 / 1 / 2
 *   lkp_O = lr0 L + lr1 R - O on lookup rows, 0 elsewhere
 * (the in-place rewrites of buff_S to 3, -1, 4 at :568-586 select exactly these through compute3p/4p_error_terms, :382-432).
 * libc draws: a = generate_randomness(5) after the streaming pass, b = generate_randomness(8) after the partial evaluations.
 * out: R[nch] | (a,b,c,d,e,rand) x log2 B | final L,R,O,add_L,add_R,mul,lkp,lkp_O,beta | Peval[8][nch] | P2 flat (4*log2 nch + 3). */
static inline void s8_gates(F s, const F *lr, F *gL, F *gR, F *gM, F *gK) {
    if (s.re == 0) { *gL = F1; *gR = F1; *gM = F0; *gK = F0; }
    else if (s.re == 1) { *gL = F0; *gR = F0; *gM = F1; *gK = F0; }
    else { *gL = lr[0]; *gR = lr[1]; *gM = F0; *gK = F1; }
}
static inline void err3(F b1, F gate, F f1, F f2, F fb, F be, F *K) {        /* compute3p_error_terms */
    F t1 = f_add(f_mul(b1, f2), f_mul(gate, f1)), t2 = f_mul(b1, gate);
    K[0] = f_add(K[0], f_add(f_mul(fb, t1), f_mul(f_mul(be, f1), f2)));
    K[1] = f_add(K[1], f_add(f_mul(be, t1), f_mul(fb, t2)));
    K[2] = f_add(K[2], f_mul(t2, be));
}
double orc_gate_consistency_lookups_stream(const F *L, const F *R, const F *O, const F *S, size_t cs, size_t B, const F *r, const F *lr, F *out) {
    size_t nch = cs / B; int lgB = ilog2(B), lgn = ilog2(nch);
    double ps = 0; size_t k = 0;
    F *beta = (F *)malloc(B * sizeof(F));
    orc_precompute_beta(r, lgB, beta);
    F *fL = (F *)malloc(9 * B * sizeof(F)), *fR = fL + B, *fO = fR + B, *fAL = fO + B, *fAR = fAL + B, *fM = fAR + B, *fK = fM + B, *fKO = fK + B, *fB = fKO + B;
    F Kf[5] = { F0, F0, F0, F0, F0 };               /* O, L, R, M, lkp */
    for (size_t i = 0; i < B; i++) {
        F gL, gR, gM, gK; s8_gates(S[i], lr, &gL, &gR, &gM, &gK);
        fL[i] = L[i]; fR[i] = R[i]; fO[i] = O[i]; fAL[i] = gL; fAR[i] = gR; fM[i] = gM; fK[i] = gK; fB[i] = beta[i];
        fKO[i] = gK.re ? f_sub(f_add(f_mul(lr[0], L[i]), f_mul(lr[1], R[i])), O[i]) : F0;
        Kf[0] = f_add(Kf[0], f_mul(beta[i], fO[i]));
        Kf[1] = f_add(Kf[1], f_mul(f_mul(beta[i], fL[i]), fAL[i]));
        Kf[2] = f_add(Kf[2], f_mul(f_mul(beta[i], fR[i]), fAR[i]));
        Kf[4] = f_add(Kf[4], f_mul(f_mul(beta[i], fKO[i]), fK[i]));
        Kf[3] = f_add(Kf[3], f_mul(f_mul(f_mul(beta[i], fR[i]), fL[i]), fM[i]));
    }
    if (!f_eq(f_sub(f_sub(f_add(f_add(Kf[3], Kf[1]), Kf[2]), Kf[4]), Kf[0]), F0)) { printf("Error\n"); exit(-1); }
    ps += 5 * 16 / 1024.0;
    F *Rv = (F *)malloc(nch * sizeof(F)); size_t nR = 0; Rv[nR++] = F1;
    F rand = F0;
    for (size_t c = 1; c < nch; c++) {
        const F *bL = L + c * B, *bR = R + c * B, *bO = O + c * B, *bS = S + c * B;
        F KO[2] = { F0, F0 }, KL[3] = { F0, F0, F0 }, KR[3] = { F0, F0, F0 }, KK[3] = { F0, F0, F0 }, KM[4] = { F0, F0, F0, F0 };
        for (size_t i = 0; i < B; i++) {
            F gL, gR, gM, gK; s8_gates(bS[i], lr, &gL, &gR, &gM, &gK);
            F bK = gK.re ? f_sub(f_add(f_mul(lr[0], bL[i]), f_mul(lr[1], bR[i])), bO[i]) : F0;
            KO[0] = f_add(KO[0], f_add(f_mul(bO[i], fB[i]), f_mul(beta[i], fO[i])));
            KO[1] = f_add(KO[1], f_mul(bO[i], beta[i]));
            err3(bL[i], gL, fL[i], fAL[i], fB[i], beta[i], KL);
            err3(bR[i], gR, fR[i], fAR[i], fB[i], beta[i], KR);
            err3(bK, gK, fKO[i], fK[i], fB[i], beta[i], KK);
            F t1 = f_add(f_mul(fL[i], bR[i]), f_mul(fR[i], bL[i])), t2 = f_add(f_mul(fB[i], gM), f_mul(fM[i], beta[i]));
            F t3 = f_mul(bL[i], bR[i]), t4 = f_mul(gM, beta[i]), t5 = f_mul(fL[i], fR[i]), t6 = f_mul(fB[i], fM[i]);
            KM[0] = f_add(KM[0], f_add(f_mul(t1, t6), f_mul(t2, t5)));
            KM[1] = f_add(KM[1], f_add(f_add(f_mul(t1, t2), f_mul(t3, t6)), f_mul(t4, t5)));
            KM[2] = f_add(KM[2], f_add(f_mul(t1, t4), f_mul(t2, t3)));
            KM[3] = f_add(KM[3], f_mul(t3, t4));
        }
        if (!f_eq(f_sub(f_sub(f_add(f_add(KM[3], KL[2]), KR[2]), KK[2]), KO[1]), F0)) { printf("Error in gate consistency 1 : %d\n", (int)c); exit(-1); }
        rand = mimc(KO[0], rand); rand = mimc(KO[1], rand); rand = mimc(KL[0], rand); rand = mimc(KL[1], rand); rand = mimc(KL[2], rand);
        rand = mimc(KR[0], rand); rand = mimc(KR[1], rand); rand = mimc(KR[2], rand);
        Rv[nR++] = rand;
        F x1 = rand, x2 = f_mul(rand, x1), x3 = f_mul(rand, x2), x4 = f_mul(rand, x3);
        Kf[0] = f_add(Kf[0], f_add(f_mul(x1, KO[0]), f_mul(x2, KO[1])));
        Kf[4] = f_add(Kf[4], f_add(f_add(f_mul(x1, KK[0]), f_mul(x2, KK[1])), f_mul(x3, KK[2])));
        Kf[1] = f_add(Kf[1], f_add(f_add(f_mul(x1, KL[0]), f_mul(x2, KL[1])), f_mul(x3, KL[2])));
        Kf[2] = f_add(Kf[2], f_add(f_add(f_mul(x1, KR[0]), f_mul(x2, KR[1])), f_mul(x3, KR[2])));
        Kf[3] = f_add(Kf[3], f_add(f_add(f_add(f_mul(x1, KM[0]), f_mul(x2, KM[1])), f_mul(x3, KM[2])), f_mul(x4, KM[3])));
        ps += 15 * 16 / 1024.0;
        F chk = F0;
        for (size_t i = 0; i < B; i++) {
            F gL, gR, gM, gK; s8_gates(bS[i], lr, &gL, &gR, &gM, &gK);
            F bK = gK.re ? f_sub(f_add(f_mul(lr[0], bL[i]), f_mul(lr[1], bR[i])), bO[i]) : F0;
            fAL[i] = f_add(fAL[i], f_mul(rand, gL)); fAR[i] = f_add(fAR[i], f_mul(rand, gR)); fM[i] = f_add(fM[i], f_mul(rand, gM)); fK[i] = f_add(fK[i], f_mul(rand, gK));
            fL[i] = f_add(fL[i], f_mul(rand, bL[i])); fR[i] = f_add(fR[i], f_mul(rand, bR[i])); fO[i] = f_add(fO[i], f_mul(rand, bO[i]));
            fKO[i] = f_add(fKO[i], f_mul(rand, bK)); fB[i] = f_add(fB[i], f_mul(rand, beta[i]));
            chk = f_add(chk, f_mul(f_mul(f_mul(fB[i], fM[i]), fR[i]), fL[i]));
        }
        if (!f_eq(chk, Kf[3])) { printf("ERRRROR %d\n", (int)c); exit(-1); }
    }
    for (size_t i = 0; i < nch; i++) out[k++] = Rv[i];
    F a[5]; prover_randomness(5, a);
    F sum = f_add(f_add(f_add(f_mul(a[0], Kf[1]), f_mul(a[1], Kf[2])), f_add(f_mul(a[2], Kf[3]), f_mul(Kf[0], a[3]))), f_mul(Kf[4], a[4]));
    F *srand_ = (F *)malloc((lgB + 1) * sizeof(F));
    F *tabs[9] = { fAL, fAR, fM, fK, fL, fR, fO, fKO, fB };
    for (int i = lgB - 1, q = 0; i >= 0; i--, q++) {
        F p[5] = { F0, F0, F0, F0, F0 };
        for (size_t j = 0; j < ((size_t)1 << i); j++) {
            F x[9], d[9];
            for (int t = 0; t < 9; t++) { x[t] = tabs[t][2 * j]; d[t] = f_sub(tabs[t][2 * j + 1], x[t]); }
            /* q(X) = a2 mul L R + a0 addL L + a1 addR R + a4 lkp lkpO + a3 O  (degree 3), then times beta(X) */
            F m0 = f_mul(a[2], x[2]), m1 = f_mul(a[2], d[2]);
            F ml2 = f_mul(m1, d[4]), ml1 = f_add(f_mul(m1, x[4]), f_mul(m0, d[4])), ml0 = f_mul(m0, x[4]);
            F q3 = f_mul(ml2, d[5]);
            F q2 = f_add(f_mul(ml2, x[5]), f_mul(ml1, d[5]));
            F q1 = f_add(f_mul(ml1, x[5]), f_mul(ml0, d[5]));
            F q0 = f_mul(ml0, x[5]);
            const int sel[3] = { 0, 1, 3 }, val[3] = { 4, 5, 7 }, wi[3] = { 0, 1, 4 };
            for (int u = 0; u < 3; u++) {
                F s0 = f_mul(a[wi[u]], x[sel[u]]), s1 = f_mul(a[wi[u]], d[sel[u]]);
                q2 = f_add(q2, f_mul(s1, d[val[u]]));
                q1 = f_add(q1, f_add(f_mul(s1, x[val[u]]), f_mul(s0, d[val[u]])));
                q0 = f_add(q0, f_mul(s0, x[val[u]]));
            }
            q1 = f_add(q1, f_mul(a[3], d[6])); q0 = f_add(q0, f_mul(a[3], x[6]));
            p[0] = f_add(p[0], f_mul(d[8], q3));
            p[1] = f_add(p[1], f_add(f_mul(d[8], q2), f_mul(x[8], q3)));
            p[2] = f_add(p[2], f_add(f_mul(d[8], q1), f_mul(x[8], q2)));
            p[3] = f_add(p[3], f_add(f_mul(d[8], q0), f_mul(x[8], q1)));
            p[4] = f_add(p[4], f_mul(x[8], q0));
        }
        for (int c = 0; c < 5; c++) { rand = mimc(p[c], rand); out[k++] = p[c]; }
        out[k++] = rand;
        {
            F s = f_add(f_add(f_add(p[0], p[1]), f_add(p[2], p[3])), f_add(p[4], p[4]));
            if (!f_eq(s, sum)) { printf("Error in gate consistency 2: %d\n", i); exit(-1); }
        }
        sum = f_add(f_mul(f_add(f_mul(f_add(f_mul(f_add(f_mul(p[0], rand), p[1]), rand), p[2]), rand), p[3]), rand), p[4]);
        srand_[q] = rand;
        ps += 5 * 16 / 1024.0;
        for (size_t j = 0; j < ((size_t)1 << i); j++)
            for (int t = 0; t < 9; t++) tabs[t][j] = f_add(tabs[t][2 * j], f_mul(rand, f_sub(tabs[t][2 * j + 1], tabs[t][2 * j])));
    }
    out[k++] = fL[0]; out[k++] = fR[0]; out[k++] = fO[0]; out[k++] = fAL[0]; out[k++] = fAR[0]; out[k++] = fM[0]; out[k++] = fK[0]; out[k++] = fKO[0]; out[k++] = fB[0];
    F *beta1 = (F *)malloc(B * sizeof(F));
    orc_precompute_beta(srand_, lgB, beta1);
    F *Pe = (F *)calloc(8 * nch, sizeof(F));
    for (size_t c = 0; c < nch; c++)
        for (size_t j = 0; j < B; j++) {
            F l = L[c * B + j], rr = R[c * B + j], o = O[c * B + j], gL, gR, gM, gK;
            s8_gates(S[c * B + j], lr, &gL, &gR, &gM, &gK);
            F bK = gK.re ? f_sub(f_add(f_mul(lr[0], l), f_mul(lr[1], rr)), o) : F0;
            const F v[8] = { l, rr, o, gL, gR, gM, gK, bK };
            for (int t = 0; t < 8; t++) Pe[t * nch + c] = f_add(Pe[t * nch + c], f_mul(beta1[j], v[t]));
        }
    for (size_t i = 0; i < 8 * nch; i++) out[k++] = Pe[i];
    F b[8]; prover_randomness(8, b);
    F *pe = (F *)malloc(nch * sizeof(F));
    for (size_t j = 0; j < nch; j++) { pe[j] = F0; for (int i = 0; i < 8; i++) pe[j] = f_add(pe[j], f_mul(b[i], Pe[i * nch + j])); }
    F *p2 = (F *)malloc((4 * (size_t)lgn + 8) * sizeof(F));
    ps += orc_sumcheck2(Rv, pe, nch, &rand, p2);
    ps += 5 * 16 / 1024.0;
    {
        const F fin[8] = { fL[0], fR[0], fO[0], fAL[0], fAR[0], fM[0], fK[0], fKO[0] };
        F s2 = F0; for (int t = 0; t < 8; t++) s2 = f_add(s2, f_mul(fin[t], b[t]));
        F qv = lgn ? f_add(f_add(p2[0], p2[1]), f_add(p2[2], p2[2])) : s2;
        if (!f_eq(qv, s2)) { printf("Error in gate consistency 3\n"); exit(-1); }
    }
    for (int i = 0; i < 4 * lgn + 3; i++) out[k++] = p2[i];
    free(beta); free(fL); free(Rv); free(srand_); free(beta1); free(Pe); free(pe); free(p2);
    return ps;
}

/* O2 front (Elastic_PC.cpp:316-333 aggregate's axpy, :487-533 compute_aggregation_reply + update_reply :59-110):
 * stream = nchunks chunks of B elements; agg[j] = sum_i beta[i] stream[i][j]; reply[q*nchunks + i] = tensorcode(chunk i)[row[q]][col[q]]. */
void orc_elastic_open_front(const F *stream, size_t nchunks, size_t B, int trs, int lin, const F *beta, const uint32_t *col, const uint32_t *row,
                            size_t queries, F *agg, F *reply) {
    size_t cols = 2 * B / trs;
    F *T = (F *)malloc(4 * B * sizeof(F));
    for (size_t j = 0; j < B; j++) agg[j] = F0;
    for (size_t i = 0; i < nchunks; i++) {
        const F *c = stream + i * B;
        for (size_t j = 0; j < B; j++) agg[j] = f_add(agg[j], f_mul(beta[i], c[j]));
        orc_compute_tensorcode(c, B, trs, lin, T);
        for (size_t q = 0; q < queries; q++) reply[q * nchunks + i] = T[(size_t)row[q] * cols + col[q]];
    }
    free(T);
}

/* one S2 round on a slice (for the sharded-sumcheck orchestration tests): coefficients of the slice + fold with `rand` */
void orc_sc3_round(const F *v1, const F *v2, const F *v3, F *o1, F *o2, F *o3, size_t L, const F *rand, F *coeffs4) {
    F acc[4] = { F0, F0, F0, F0 };
    for (size_t j = 0; j < L; j++) {
        cubic_acc(acc, v1[2 * j], v1[2 * j + 1], v2[2 * j], v2[2 * j + 1], v3[2 * j], v3[2 * j + 1]);
        o1[j] = f_add(v1[2 * j], f_mul(*rand, f_sub(v1[2 * j + 1], v1[2 * j])));
        o2[j] = f_add(v2[2 * j], f_mul(*rand, f_sub(v2[2 * j + 1], v2[2 * j])));
        o3[j] = f_add(v3[2 * j], f_mul(*rand, f_sub(v3[2 * j + 1], v3[2 * j])));
    }
    memcpy(coeffs4, acc, sizeof acc);
}
