// TEST INFRASTRUCTURE ONLY — flat C wrappers over the UNMODIFIED reference
// (compiled in place from /root/reference by oracle/Makefile into
// oracle/_ref/libhobbit_ref.so).  Used by tests/ to pin the C restatement
// (oracle/hobbit_oracle.c) and the CUDA path, by tools that generate
// tests/golden/, and by bench.py's cpu_baseline / --impl reference leg.
// Never linked into or called by the product library.
//
// Every wrapper only marshals flat buffers <-> the reference's STL containers;
// all arithmetic is the reference's own.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <chrono>
#include <mutex>
#include <thread>
#include "config_pc.hpp"
#include "utils.hpp"
#include "mimc.h"
#include "merkle_tree.h"
#include "expanders.h"
#include "linear_code_encode.h"
#include "Our_PC.hpp"
#include "witness_stream.h"
#include "Elastic_PC.hpp"
#include "sumcheck.h"
#include "Virgo.h"
#include "PC_utils.h"

extern bool linear_time;
extern int tensor_row_size;
extern size_t BUFFER_SPACE;
void _fft(F *arr, int logn, bool flag);
proof batch_3product_sumcheck(vector<vector<F>> &arr1, vector<vector<F>> &arr2, vector<vector<F>> &arr3, vector<F> a, double &vt, double &ps);
void generate_3product_sumcheck_beta_stream_batch_optimized(stream_descriptor fd, vector<vector<F>> r, int batches, int distance, int layer_id,
		vector<F> old_claims, vector<F> &new_claims, vector<vector<F>> &new_r, double &vt, double &ps);
extern int BUFFER_SPACE_tr;
extern int aggregation_queries;
// circuit front end (main.cpp globals / Seval.cpp producer thread)
extern int fun;
extern size_t circuit_size;
extern F a_w, b_w;
extern std::mutex mtx, mtx2;
extern std::vector<int> layer_size;
void Seval_Oracle();
void init_stream(int b, int n, int d);

void compute_aggregation_reply(stream_descriptor fd, vector<vector<size_t>> &I, vector<vector<F>> &reply);
void aggregate(stream_descriptor fd, vector<F> beta1, vector<F> random_points, vector<vector<_hash>> &MT_hashes, vector<F> &aggregated_vector, vector<vector<F>> &aggregated_tensor);

static_assert(sizeof(F) == 16, "F must be the 16-byte {real,img} POD");

static void reset_encode_scratch() {
    // linear_code_encode.h:64-72 sizes the static scratch from the FIRST call's n;
    // force re-allocation so a larger n in the same process is safe (leaks the old one).
    __encode_initialized = false;
}

static void levels_to_flat(const vector<vector<_hash>> &MT, uint8_t *out) {
    size_t off = 0;
    for (size_t l = 0; l < MT.size(); l++) {
        memcpy(out + off, MT[l].data(), MT[l].size() * 32);
        off += MT[l].size() * 32;
    }
}

extern "C" {

void ref_init() { init_hash(); }

// ---- F1/F2 field ----------------------------------------------------------
void ref_field_binop(int op, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t n) {
    const F *A = (const F *)a, *B = (const F *)b; F *C = (F *)c;
    for (size_t i = 0; i < n; i++) {
        switch (op) {
            case 0: C[i] = A[i] + B[i]; break;
            case 1: C[i] = A[i] - B[i]; break;
            case 2: C[i] = A[i] * B[i]; break;
            case 3: C[i] = -A[i]; break;
            case 4: C[i] = A[i].inv(); break;
        }
    }
}
void ref_root_of_unity(int n, uint64_t *out) { F r = getRootOfUnity(n); memcpy(out, &r, 16); }
void ref_mimc_hash(const uint64_t *in, const uint64_t *k, uint64_t *out) {
    F r = mimc_hash(*(const F *)in, *(const F *)k); memcpy(out, &r, 16);
}

// ---- N1 NTT ---------------------------------------------------------------
void ref_fft(uint64_t *arr, int logn) { _fft((F *)arr, logn, false); }

// ---- RNG-driven generators ------------------------------------------------
void ref_generate_randomness(int n, uint64_t *out) {
    vector<F> v = generate_randomness(n);
    memcpy(out, v.data(), (size_t)n * 16);
}

// ---- E1/E2 expander -------------------------------------------------------
long long ref_expander_init_store(long long n) { reset_encode_scratch(); return expander_init_store(n); }
// number of levels that expander_init_store(n) filled
int ref_expander_levels(long long n) {
    int d = 0; while (n > distance_threshold) { n = (long long)(alpha * n); d++; } return d;
}
// which: 0 = _C[dep], 1 = D[dep].  Returns L, writes R and degree.
long long ref_expander_dims(int which, int dep, long long *R, int *deg) {
    graph &g = which ? D[dep] : _C[dep]; *R = g.R; *deg = g.degree; return g.L;
}
// neighbors: L*deg u32 ; weights: L*deg u64 (real part; img is always 0)
void ref_expander_dump(int which, int dep, uint32_t *nbr, uint64_t *w) {
    graph &g = which ? D[dep] : _C[dep];
    for (long long i = 0; i < g.L; i++)
        for (int j = 0; j < g.degree; j++) {
            nbr[i * g.degree + j] = (uint32_t)g.neighbor[i][j];
            w[i * g.degree + j] = g.weight[i][j].real;
            if (g.weight[i][j].img != 0) { printf("ref_shim: complex weight?\n"); exit(-1); }
        }
}
int ref_encode_monolithic(const uint64_t *src, uint64_t *dst, long long n) {
    return encode_monolithic((const F *)src, (F *)dst, n);
}
// E3: encode() (linear_code_encode.h:122-191); needs the graph DIMENSIONS of expander_init / expander_init_store(n) in _C[] / D[]
int ref_encode_reseed(const uint64_t *src, uint64_t *dst, long long n) {
    return encode((const F *)src, (F *)dst, n);
}

// ---- H1..H4 ---------------------------------------------------------------
void ref_blake3_hash(const uint8_t *src, uint8_t *dst) { blake3_hash((uint8_t *)src, dst); }
void ref_md_leaf(const uint64_t *xyzw, const uint8_t *prev, uint8_t *out) {
    const F *f = (const F *)xyzw; _hash p; memcpy(p.arr, prev, 32);
    _hash r = merkle_tree::hash_double_field_element_merkle_damgard_blake(f[0], f[1], f[2], f[3], p);
    memcpy(out, r.arr, 32);
}
// out must hold (2*N/4 - 1)*32 bytes: all levels, leaves first
int ref_mt_commit_blake(const uint64_t *leafs, int N, uint8_t *out) {
    vector<vector<_hash>> H;
    merkle_tree::merkle_tree_prover::MT_commit_Blake((F *)leafs, H, N);
    levels_to_flat(H, out);
    return (int)H.size();
}
int ref_create_tree_blake(const uint8_t *leaves, int n, uint8_t *out) {
    vector<vector<_hash>> H((int)log2(n) + 1);
    H[0].resize(n); memcpy(H[0].data(), leaves, (size_t)n * 32);
    merkle_tree::merkle_tree_prover::create_tree_blake(n, H, 32, true);
    levels_to_flat(H, out);
    return (int)H.size();
}

// ---- T1 tensor code -------------------------------------------------------
// msg: n F ; tensor_out: 4n F row-major (2*trs rows x 2n/trs cols)
void ref_compute_tensorcode(const uint64_t *msg, size_t n, int trs, int lin, uint64_t *tensor_out) {
    tensor_row_size = trs; linear_time = lin;
    vector<F> m((const F *)msg, (const F *)msg + n);
    vector<vector<F>> T;
    compute_tensorcode(m, T);
    size_t cols = T[0].size();
    for (size_t i = 0; i < T.size(); i++) memcpy(tensor_out + 2 * i * cols, T[i].data(), cols * 16);
}

// ---- C1 commit_standard ---------------------------------------------------
// The caller must have prepared the RNG/expander state (ref_expander_init_store) when lin=1.
// levels_out: (2*(N/K)-1)*32 bytes.  tensor_out (optional): K * 4*(N/K) F, chunk-major then row-major.
// Returns seconds spent inside commit_standard.
static vector<vector<vector<F>>> g_tensor;
static vector<vector<_hash>> g_MT;
static vector<F> g_poly;
double ref_commit_standard(const uint64_t *poly, size_t N, int K, int trs, int lin,
                           uint8_t *levels_out, uint64_t *tensor_out) {
    tensor_row_size = trs; linear_time = lin;
    g_poly.assign((const F *)poly, (const F *)poly + N);
    g_tensor.clear(); g_MT.clear();
    _hash comm;
    auto t0 = std::chrono::steady_clock::now();
    commit_standard(g_poly, comm, g_MT, g_tensor, K);
    auto t1 = std::chrono::steady_clock::now();
    if (levels_out) levels_to_flat(g_MT, levels_out);
    if (tensor_out) {
        size_t off = 0;
        for (auto &T : g_tensor) for (auto &row : T) { memcpy(tensor_out + 2 * off, row.data(), row.size() * 16); off += row.size(); }
    }
    return std::chrono::duration_cast<std::chrono::duration<double>>(t1 - t0).count();
}
// open_standard on the state left by ref_commit_standard.  x: log2(N) F.  Returns seconds; writes ps.
double ref_open_standard(const uint64_t *x, int nx, int K, double *ps_out) {
    vector<F> xv((const F *)x, (const F *)x + nx);
    double vt = 0, ps = 0;
    auto t0 = std::chrono::steady_clock::now();
    open_standard(g_poly, xv, g_MT, g_tensor, K, vt, ps);
    auto t1 = std::chrono::steady_clock::now();
    *ps_out = ps;
    return std::chrono::duration_cast<std::chrono::duration<double>>(t1 - t0).count();
}
void ref_release() { g_tensor.clear(); g_tensor.shrink_to_fit(); g_MT.clear(); g_poly.clear(); g_poly.shrink_to_fit(); }

// ---- C2 Elastic commit (synthetic stream "test") -------------------------
// levels_out: (2*4B-1)*32 bytes.  Returns seconds.
double ref_elastic_commit(size_t N, size_t B, int trs, int lin, uint8_t *levels_out) {
    BUFFER_SPACE = B; tensor_row_size = trs; linear_time = lin;
    stream_descriptor fd; fd.name = "test"; fd.size = N; fd.pos = 0;
    vector<vector<_hash>> MT; _hash comm;
    auto t0 = std::chrono::steady_clock::now();
    commit(fd, comm, MT);
    auto t1 = std::chrono::steady_clock::now();
    if (levels_out) levels_to_flat(MT, levels_out);
    return std::chrono::duration_cast<std::chrono::duration<double>>(t1 - t0).count();
}
void ref_read_stream_pc_test(uint64_t *out, size_t n) {
    stream_descriptor fd; fd.name = "test"; fd.size = n; fd.pos = 0;
    read_stream_PC(fd, (F *)out, n);
}

// ---- S9 helpers -----------------------------------------------------------
void ref_precompute_beta(const uint64_t *r, int nr, uint64_t *out) {
    vector<F> rv((const F *)r, (const F *)r + nr), B;
    precompute_beta(rv, B);
    memcpy(out, B.data(), B.size() * 16);
}
void ref_evaluate_vector(const uint64_t *v, size_t n, const uint64_t *r, int nr, uint64_t *out) {
    vector<F> vv((const F *)v, (const F *)v + n), rv((const F *)r, (const F *)r + nr);
    F e = evaluate_vector(vv, rv); memcpy(out, &e, 16);
}

// ---- S1/S2/S3 sumchecks ---------------------------------------------------
// Flat proof layout written to `out` (F units):
//   S1: per round (a,b,c) [3*rounds] | randomness[0] [rounds] | vr [2] | final_rand [1]
//   S2: per round (a,b,c,d) [4*rounds] | randomness[0] [rounds] | vr [3] | final_rand [1]
//   S3: per round (a,b,c,d) [4*rounds] | randomness[0] [rounds] | vr [3*batches]
// Return value: ps increment (KB) so the deterministic proof-size counter is checked too.
double ref_sumcheck2(const uint64_t *v1, const uint64_t *v2, size_t n, const uint64_t *prev_r, uint64_t *out) {
    vector<F> a((const F *)v1, (const F *)v1 + n), b((const F *)v2, (const F *)v2 + n);
    double vt = 0, ps = 0;
    proof P = generate_2product_sumcheck_proof(a, b, *(const F *)prev_r, vt, ps);
    F *o = (F *)out; size_t k = 0;
    for (auto &q : P.q_poly) { o[k++] = q.a; o[k++] = q.b; o[k++] = q.c; }
    for (auto &r : P.randomness[0]) o[k++] = r;
    for (auto &v : P.vr) o[k++] = v;
    o[k++] = P.final_rand;
    return ps;
}
double ref_sumcheck3(const uint64_t *v1, const uint64_t *v2, const uint64_t *v3, size_t n, const uint64_t *prev_r, uint64_t *out) {
    vector<F> a((const F *)v1, (const F *)v1 + n), b((const F *)v2, (const F *)v2 + n), c((const F *)v3, (const F *)v3 + n);
    double vt = 0, ps = 0;
    proof P = _generate_3product_sumcheck_proof(a, b, c, *(const F *)prev_r, vt, ps);
    F *o = (F *)out; size_t k = 0;
    for (auto &q : P.c_poly) { o[k++] = q.a; o[k++] = q.b; o[k++] = q.c; o[k++] = q.d; }
    for (auto &r : P.randomness[0]) o[k++] = r;
    for (auto &v : P.vr) o[k++] = v;
    o[k++] = P.final_rand;
    return ps;
}
// sizes[b] = length of batch b (power of two, non-increasing); tables concatenated batch after batch.
double ref_batch_sumcheck3(const uint64_t *t1, const uint64_t *t2, const uint64_t *t3, const size_t *sizes, int batches,
                           const uint64_t *a, uint64_t *out) {
    vector<vector<F>> A(batches), B(batches), C(batches);
    size_t off = 0;
    for (int b = 0; b < batches; b++) {
        A[b].assign((const F *)t1 + off, (const F *)t1 + off + sizes[b]);
        B[b].assign((const F *)t2 + off, (const F *)t2 + off + sizes[b]);
        C[b].assign((const F *)t3 + off, (const F *)t3 + off + sizes[b]);
        off += sizes[b];
    }
    vector<F> av((const F *)a, (const F *)a + batches);
    double vt = 0, ps = 0;
    proof P = batch_3product_sumcheck(A, B, C, av, vt, ps);
    F *o = (F *)out; size_t k = 0;
    for (auto &q : P.c_poly) { o[k++] = q.a; o[k++] = q.b; o[k++] = q.c; o[k++] = q.d; }
    for (auto &r : P.randomness[0]) o[k++] = r;
    for (auto &v : P.vr) o[k++] = v;
    return ps;
}

// ---- S5 product-tree GKR --------------------------------------------------
// input: `vectors` tables of `n` F each, concatenated.  Output (F units):
//   output [vectors] | out_eval [1] | final_r [nfr] | final_eval [1] ; per layer proof: c_poly(4/round) | vr[3] | final_rand
// *nfr_out = final_r.size(); returns number of F written.
size_t ref_mul_tree(const uint64_t *input, int vectors, size_t n, const uint64_t *prev_r, uint64_t *out, int *nfr_out, double *ps_out) {
    vector<vector<F>> in(vectors);
    for (int i = 0; i < vectors; i++) in[i].assign((const F *)input + i * n, (const F *)input + (i + 1) * n);
    vector<F> px; double vt = 0, ps = 0;
    mul_tree_proof P = prove_multiplication_tree_new(in, *(const F *)prev_r, px, vt, ps);
    F *o = (F *)out; size_t k = 0;
    for (auto &v : P.output) o[k++] = v;
    o[k++] = P.out_eval;
    for (auto &v : P.final_r) o[k++] = v;
    o[k++] = P.final_eval;
    for (auto &pr : P.proofs) {
        for (auto &q : pr.c_poly) { o[k++] = q.a; o[k++] = q.b; o[k++] = q.c; o[k++] = q.d; }
        for (auto &v : pr.vr) o[k++] = v;
        o[k++] = pr.final_rand;
    }
    *nfr_out = (int)P.final_r.size(); *ps_out = ps;
    return k;
}

// ---- S4 / S6 on the synthetic default stream (read_stream default branch: v[i] = F(i%1024+1), witness_stream.cpp:2348-2352) ----
// One layer of the streaming folding sumcheck (sumcheck.cpp:1150-1392), batches = 1, distance = 1.
// r: log2((total >> layer_id)/2) points.  Writes new_claim (1 F) and new_r (returns its length).
int ref_stream_sumcheck_layer(size_t total, size_t B, int layer_id, const uint64_t *r, int nr, const uint64_t *old_claim,
                              uint64_t *new_claim, uint64_t *new_r, double *ps_out) {
    BUFFER_SPACE = B; BUFFER_SPACE_tr = B / 8;
    stream_descriptor fd; fd.name = "test"; fd.size = total; reset_stream(fd);
    vector<vector<F>> rv(1), nr_out; rv[0].assign((const F *)r, (const F *)r + nr);
    vector<F> oc(1, *(const F *)old_claim), nc(1);
    double vt = 0, ps = 0;
    generate_3product_sumcheck_beta_stream_batch_optimized(fd, rv, 1, 1, layer_id, oc, nc, nr_out, vt, ps);
    memcpy(new_claim, &nc[0], 16);
    memcpy(new_r, nr_out[0].data(), nr_out[0].size() * 16);
    *ps_out = ps;
    return (int)nr_out[0].size();
}
// The batched form (batches > 1, as used when layers > distance): r = batches rows of `rs` points (row j: log2(B >> j*distance) + log2(nb)
// used); writes new_claims[batches] and new_r (batches rows of rs); returns the length of row 0.
int ref_stream_sumcheck_batch(size_t total, size_t B, int layer_id, int distance, int batches, const uint64_t *r, int rs, const int *rlen,
                              const uint64_t *old_claims, uint64_t *new_claims, uint64_t *new_r, double *ps_out) {
    BUFFER_SPACE = B; BUFFER_SPACE_tr = B / 8;
    stream_descriptor fd; fd.name = "test"; fd.size = total; reset_stream(fd);
    vector<vector<F>> rv(batches), nr_out;
    for (int j = 0; j < batches; j++) rv[j].assign((const F *)r + (size_t)j * rs, (const F *)r + (size_t)j * rs + rlen[j]);
    vector<F> oc((const F *)old_claims, (const F *)old_claims + batches), nc(batches);
    double vt = 0, ps = 0;
    generate_3product_sumcheck_beta_stream_batch_optimized(fd, rv, batches, distance, layer_id, oc, nc, nr_out, vt, ps);
    memcpy(new_claims, nc.data(), batches * 16);
    for (int j = 0; j < batches; j++) memcpy(new_r + 2 * (size_t)j * rs, nr_out[j].data(), nr_out[j].size() * 16);
    *ps_out = ps;
    return (int)nr_out[0].size();
}
// prove_multiplication_tree_stream_shallow (sumcheck.cpp:1746-1915) on stream "test"; out: the `vectors` products.
double ref_mul_tree_stream(size_t total, int vectors, size_t B, int distance, int naive, const uint64_t *prev_r, uint64_t *out) {
    BUFFER_SPACE = B; BUFFER_SPACE_tr = B / 8;
    stream_descriptor fd; fd.name = "test"; fd.size = total; reset_stream(fd);
    vector<F> px; double vt = 0, ps = 0;
    vector<F> o = prove_multiplication_tree_stream_shallow(fd, vectors, total / vectors, *(const F *)prev_r, distance, px, naive, vt, ps);
    memcpy(out, o.data(), o.size() * 16);
    return ps;
}

// prove_gate_consistency_standard (sumcheck.cpp:434-501) folds its arguments in place and returns nothing: out4 = the final
// add_gate[0], arr_L[0], arr_R[0], arr_O[0] (they depend on every round polynomial through the challenge chain).
void ref_gate_consistency_standard(const uint64_t *L, const uint64_t *R, const uint64_t *O, const uint64_t *add_gate, size_t n, const uint64_t *r, uint64_t *out4) {
    vector<F> l((const F *)L, (const F *)L + n), rr((const F *)R, (const F *)R + n), o((const F *)O, (const F *)O + n), a((const F *)add_gate, (const F *)add_gate + n);
    int nr = 0; while (((size_t)1 << nr) < n) nr++;
    vector<F> rv((const F *)r, (const F *)r + nr);
    double vt = 0, ps = 0;
    prove_gate_consistency_standard(l, rr, o, a, rv, vt, ps);
    F res[4] = { a[0], l[0], rr[0], o[0] };
    memcpy(out4, res, 64);
}

// O2 front on the synthetic default stream (read_stream: v[i] = F(i%1024+1)), RS columns (linear_time = false):
// compute_aggregation_reply (Elastic_PC.cpp:487-533) and the aggregated vector of aggregate (:316-345; it also runs shockwave_commit).
void ref_elastic_open_front_rs(size_t N, size_t B, int trs, const uint64_t *beta, const uint32_t *col, const uint32_t *row, size_t queries,
                               uint64_t *agg, uint64_t *reply) {
    BUFFER_SPACE = B; BUFFER_SPACE_tr = B / 8; tensor_row_size = trs; linear_time = false; aggregation_queries = (int)queries;
    size_t nch = N / B;
    stream_descriptor fd; fd.name = "test"; fd.size = N; reset_stream(fd);
    vector<vector<size_t>> I(queries);
    for (size_t q = 0; q < queries; q++) { I[q].push_back(col[q]); I[q].push_back(row[q]); }
    vector<vector<F>> rep;
    compute_aggregation_reply(fd, I, rep);
    for (size_t q = 0; q < queries; q++) {
        if (rep[q].size() != nch) { printf("ref_shim: reply row %zu has %zu entries\n", q, rep[q].size()); exit(-1); }
        memcpy(reply + 2 * q * nch, rep[q].data(), nch * 16);
    }
    vector<F> b((const F *)beta, (const F *)beta + nch), rv(nch, F(1)), av; vector<vector<_hash>> MT; vector<vector<F>> at;
    reset_stream(fd);
    aggregate(fd, b, rv, MT, av, at);
    memcpy(agg, av.data(), B * 16);
}

// ---- circuit streams: the reference's producer thread (Seval_Oracle) + its named streams ------------------------------------
// One circuit per process (the reference keeps everything in globals and the producer thread never exits): callers run this in a
// subprocess.  Mirrors main() (main.cpp:1171-1233): init_hash, lock both mutexes, start the producer, init_stream, draw a_w/b_w.
static bool g_circuit_started = false;
size_t ref_circuit_start(int fun_, int b, int n, int d, const int *extra, int nextra) {
    if (g_circuit_started) { printf("ref_shim: one circuit per process\n"); exit(-1); }
    g_circuit_started = true;
    init_hash();
    mtx.lock(); mtx2.lock();
    fun = fun_;
    if (fun == 9) for (int i = 0; i < nextra; i++) layer_size.push_back(extra[i]);
    std::thread t(Seval_Oracle);
    t.detach();
    init_stream(b, n, d);
    a_w = random(); b_w = random();
    return circuit_size;
}
void ref_circuit_ab(uint64_t *out2) { memcpy(out2, &a_w, 16); memcpy(out2 + 2, &b_w, 16); }
size_t ref_buffer_space() { return BUFFER_SPACE; }
// read `total` elements of stream `name` in reads of `block` elements (read_stream, witness_stream.cpp:2106-2353)
void ref_dump_stream(const char *name, size_t total, size_t block, uint64_t *out) {
    stream_descriptor fd; fd.name = name; fd.size = total; reset_stream(fd);
    vector<F> v(block);
    for (size_t off = 0; off < total; off += block) { read_stream(fd, v, (int)block); memcpy(out + 2 * off, v.data(), block * 16); }
}
// the gate transcript (read_trace, witness_stream.cpp:1701-1807): cs entries in reads of B
void ref_dump_trace(size_t cs, size_t B, uint64_t *L, uint64_t *R, uint64_t *O, uint64_t *S) {
    stream_descriptor fd; fd.name = "transcript_stream"; fd.size = cs; reset_stream(fd);
    vector<F> l(B), r(B), o(B); vector<int> s(B);
    for (size_t off = 0; off < cs; off += B) {
        read_trace(fd, l, r, o, s);
        memcpy(L + 2 * off, l.data(), B * 16); memcpy(R + 2 * off, r.data(), B * 16); memcpy(O + 2 * off, o.data(), B * 16);
        for (size_t i = 0; i < B; i++) { S[2 * (off + i)] = (uint64_t)s[i]; S[2 * (off + i) + 1] = 0; }
    }
}
// the reference's own provers on the live streams (prove_circuit, main.cpp:862-886)
double ref_circuit_commit_witness(uint8_t *levels_out) {
    stream_descriptor fd; fd.name = "witness"; fd.size = 4 * circuit_size; reset_stream(fd);
    init_commitment(false);
    vector<vector<_hash>> MT; _hash comm;
    auto t0 = std::chrono::steady_clock::now();
    commit(fd, comm, MT);
    auto t1 = std::chrono::steady_clock::now();
    levels_to_flat(MT, levels_out);
    return std::chrono::duration_cast<std::chrono::duration<double>>(t1 - t0).count();
}
double ref_circuit_mul_tree(uint64_t *out8, double *ps_out) {
    stream_descriptor fd; fd.name = "wiring_consistency_check_opt"; fd.size = 8 * circuit_size; reset_stream(fd);
    vector<F> px; double vt = 0, ps = 0;
    auto t0 = std::chrono::steady_clock::now();
    vector<F> o = prove_multiplication_tree_stream_shallow(fd, 8, fd.size / 8, F(32), 5, px, 0, vt, ps);
    auto t1 = std::chrono::steady_clock::now();
    memcpy(out8, o.data(), o.size() * 16); *ps_out = ps;
    return std::chrono::duration_cast<std::chrono::duration<double>>(t1 - t0).count();
}
double ref_circuit_gate_consistency(const uint64_t *r, int nr, double *ps_out) {
    stream_descriptor fd; fd.name = "transcript_stream"; fd.size = circuit_size; reset_stream(fd);
    vector<F> rv((const F *)r, (const F *)r + nr);
    double vt = 0, ps = 0;
    auto t0 = std::chrono::steady_clock::now();
    prove_gate_consistency(fd, rv, vt, ps);
    auto t1 = std::chrono::steady_clock::now();
    *ps_out = ps;
    return std::chrono::duration_cast<std::chrono::duration<double>>(t1 - t0).count();
}

// ---- 8f.1 building blocks, for golden vectors (tests/golden/make_golden.py) ------------------------------------------------------------
void ref_phi_g_init(const uint64_t *r, int n, uint64_t *out) {                 // utils.cpp:694-755, forward transform, scale 1
    vector<F> rv((const F *)r, (const F *)r + n), g((size_t)1 << n, F(0));
    phiGInit(g, rv.begin(), F(1), n, false);
    memcpy(out, g.data(), g.size() * 16);
}
void ref_change_form(uint64_t *poly, int logn) {                                // Virgo.cpp:104-118
    vector<F> p((const F *)poly, (const F *)poly + ((size_t)1 << logn));
    change_form(p, logn, 0, 0);
    memcpy(poly, p.data(), p.size() * 16);
}
// shockwave_commit (Virgo.cpp:120-157): encoded matrix (k x 2N/k) and every Merkle level (leaves first)
void ref_shockwave_commit(const uint64_t *poly, size_t N, int k, uint64_t *encoded, uint8_t *levels) {
    vector<F> p((const F *)poly, (const F *)poly + N);
    shockwave_data *d = shockwave_commit(p, k);
    for (int i = 0; i < k; i++) memcpy(encoded + 2 * (size_t)i * (2 * N / k), d->encoded_matrix[i], (2 * N / k) * 16);
    levels_to_flat(d->MT, levels);
    delete d;
}
void ref_whir_commit(const uint64_t *poly, size_t N, uint8_t *levels) {          // Virgo.cpp:160-178
    vector<F> p((const F *)poly, (const F *)poly + N);
    Whir_data D; whir_commit(p, D);
    levels_to_flat(D.MT, levels);
}
// prove_fft (sumcheck.cpp:2975-2987): flat 2-product proof of 4*rounds+3 elements (rounds = log2(2n)); randomness as returned (last popped -> zeroed)
double ref_prove_fft(const uint64_t *m, size_t n, const uint64_t *r, const uint64_t *prev_sum, uint64_t *out) {
    vector<F> mv((const F *)m, (const F *)m + n); int rounds = 0; while (((size_t)1 << rounds) < 2 * n) rounds++;
    vector<F> rv((const F *)r, (const F *)r + rounds);
    double vt = 0, ps = 0;
    proof P = prove_fft(mv, rv, *(const F *)prev_sum, vt, ps);
    F *o = (F *)out; size_t k = 0;
    for (int i = 0; i < rounds; i++) { o[k++] = P.q_poly[i].a; o[k++] = P.q_poly[i].b; o[k++] = P.q_poly[i].c; }
    for (int i = 0; i < rounds; i++) o[k++] = i < (int)P.randomness[0].size() ? P.randomness[0][i] : F(0);
    o[k++] = P.vr[0]; o[k++] = P.vr[1]; o[k++] = P.final_rand;
    return ps;
}

} // extern "C"
