// hobbit_host — C++ host-side mirror of the reference's entry points for the GPU hot path.
//
// Every function below has the NAME, ARGUMENT MEANING and ERROR BEHAVIOUR (printf + exit(-1)) of the reference function
// it replaces and forwards to the C ABI in include/hobbit_b200.h.  They live in namespace `hobbit` so that a test
// binary can link the unmodified reference (global namespace) next to them and compare (tests/cpp/dropin_test.cpp).
// A maintainer who wants the drop-in inside the reference tree removes the namespace.
//
//   reference                                                   here
//   ---------------------------------------------------------   ------------------------------------------------
//   virgo::fieldElement (fieldElement.hpp:15-97)                 hobbit::F  (same 16-byte layout)
//   struct _hash (Blake3_hash.h:3-5)                             hobbit::_hash
//   generate_randomness (utils.cpp:873-883)                      hobbit::generate_randomness   (libc rand()/random())
//   expander_init_store (expanders.h:78-92) + _C/D               hobbit::expander_init_store   (same RNG order) + upload
//   commit_standard / open_standard (Our_PC.cpp:146-171,604-692) hobbit::commit_standard / open_standard_queries
//   init_commitment / commit (Elastic_PC.cpp:728-734,174-285)    hobbit::init_commitment / commit
//   read_stream_PC default stream (witness_stream.cpp:2405-2411) hobbit::read_stream_PC
//   merkle_tree_prover::{MT_commit_Blake,create_tree_blake,open_tree_blake}  hobbit::merkle_tree::merkle_tree_prover::…
//   mimc_hash, precompute_beta, evaluate_vector                  same names
//   generate_2product_sumcheck_proof, _generate_3product_sumcheck_proof, batch_3product_sumcheck,
//   prove_multiplication_tree_new (sumcheck.cpp)                 same names, same `proof` / `mul_tree_proof` fields
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/hobbit_b200.h"

namespace hobbit {

struct F {                                     // virgo::fieldElement: {real, img}, canonical limbs
    unsigned long long real = 0, img = 0;
    F() {}
    F(long long x) { real = x >= 0 ? (unsigned long long)x : 2305843009213693951ULL + x; img = 0; }
    F(long long x, long long y) { real = x >= 0 ? x : 2305843009213693951ULL + x; img = y >= 0 ? y : 2305843009213693951ULL + y; }
    F operator+(const F &o) const;
    F operator-(const F &o) const;
    F operator*(const F &o) const;
    F operator-() const;
    bool operator==(const F &o) const { return real == o.real && img == o.img; }
    bool operator!=(const F &o) const { return !(*this == o); }
};
static_assert(sizeof(F) == sizeof(hb_F), "layout must match the C ABI");

struct _hash {
    uint8_t arr[32];
    _hash() {}          // deliberately NOT zeroing: std::vector<_hash>(n) of a 128 MiB level must not be touched (page-faulted, zeroed) by one
                        // thread before the download fills it; every level the mirror returns is written in full
};

struct quadratic_poly { F a, b, c; F eval(const F &x) const { return (a * x + b) * x + c; } };
struct cubic_poly { F a, b, c, d; F eval(const F &x) const { return ((a * x + b) * x + c) * x + d; } };

struct proof {                                 // the fields of reference `struct proof` the hot path fills (sumcheck.h:21-43)
    std::vector<std::vector<F>> randomness;
    std::vector<quadratic_poly> q_poly;
    std::vector<cubic_poly> c_poly;
    std::vector<F> vr;
    F final_rand;
};
struct mul_tree_proof {                        // sumcheck.h:8-20
    F initial_randomness;
    size_t size = 0;
    F out_eval;
    std::vector<proof> proofs;
    std::vector<F> output, final_r, global_randomness, individual_randomness;
    F final_eval;
};

struct stream_descriptor {                     // witness_stream.h:6-16
    int idx = 0, offset = 0, stage = 0;
    bool finished = false;
    size_t pos = 0, pos_j = 0;
    size_t data_size = 0, row_size = 0, col_size = 0, size = 0, layer = 0, tree_pos = 0;
    std::string name;
};

// the reference's globals (main.cpp:31,38; Our_PC.cpp:21)
extern int tensor_row_size;
extern size_t BUFFER_SPACE;
extern bool linear_time;

// one-time setup: creates the GPU context (exits like the reference on failure: there is no CPU fallback)
void init_backend(int device = 0);
hb_ctx *backend();
// Multi-GPU (one process per GPU of one box, include/hobbit_b200.h "multi-GPU").  Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR /
// MASTER_PORT (what torchrun sets; HB_RDV_PORT overrides the rendezvous port, default MASTER_PORT + 29), creates the backend on device
// LOCAL_RANK (HB_SHARE_GPU=1: all ranks on device 0 — single-GPU boxes, tests), exchanges the window blobs over a TCP store that rank 0
// serves, and switches the provers to sharded mode.  Every rank must then run the same call sequence; all ranks obtain the same proof.
// data_bytes: size of the window's data region (commit() shards a commitment only when its digests fit: 32 B per coefficient / world
// + 256 B per leaf).  WORLD_SIZE unset or 1: plain init_backend.
void dist_init_from_env(size_t data_bytes = (size_t)1 << 30);
int dist_rank();
int dist_world();

std::vector<F> generate_randomness(int size);
void trace_rng(const char *where);           // HOBBIT_TRACE_RNG=1: fingerprint of the libc random() state on stderr
long long expander_init_store(long long n, int dep = 0);
// E3: encode() (linear_code_encode.h:122-191): the expander code whose graph is re-drawn from fixed seeds on every call; returns n + L + R
int encode(const F *src, F *dst, long long n, int dep = 0);
F mimc_hash(F input, F k);
void precompute_beta(std::vector<F> r, std::vector<F> &B);
F evaluate_vector(std::vector<F> v, std::vector<F> r);
void _fft(F *arr, int logn, bool flag);

namespace merkle_tree { namespace merkle_tree_prover {
void MT_commit_Blake(F *leafs, std::vector<std::vector<_hash>> &hashes, int N);
void create_tree_blake(int ele_num, std::vector<std::vector<_hash>> &hashes, const int element_size = 32, bool alloc_required = false);
std::vector<_hash> open_tree_blake(std::vector<std::vector<_hash>> &MT_hashes, std::vector<size_t> c, int collumns);
} }

// Our_PC.  `_tensor` is filled only when hobbit::materialize_tensor is true (the encoded tensor always stays resident
// in HBM for the open phase; copying 64 B per coefficient back to the host is what the reference's layout implies
// but not what its open needs).
extern bool materialize_tensor;
// false: Elastic_PC commit() sizes every MT_hashes level as the reference does but only FILLS the levels of <= 1024 digests (root included);
// the big levels stay in HBM.  The prover (open()) only uses the level sizes; a caller that hands the tree to a verifier keeps the default.
extern bool commit_levels_on_host;
// true: open_standard aggregates from the device copy of the polynomial that the LAST commit_standard staged instead of uploading
// `poly` again — the caller asserts that `poly` still is that polynomial, unmodified.  Default false: the library never infers the
// identity of a host buffer from its address.
extern bool open_reuses_committed_poly;
// true: streams that are produced chunk by chunk (everything that is not a resident circuit stream) are pushed from PINNED HOST memory, one
// PCIe transfer per chunk, double-buffered against the encode — the witness never has to fit HBM (BASELINE config 5).  Default false: the
// synthetic chunk is uploaded once and pushed from HBM.
extern bool stream_in_pinned_host;
// true: Elastic_PC commit() returns as soon as the tree kernels are queued; the Merkle levels are copied into MT_hashes in the BACKGROUND
// (worker thread, own stream) while the caller goes on proving.  MT_hashes must not be read before wait_levels() returned; open() waits on
// its own before it frees the tree, so the reference's commit -> prove -> open sequences (prove_circuit, test_Elastic_PC) need nothing else.
extern bool commit_levels_async;
void wait_levels();
extern size_t g_dist_data_bytes;               // data bytes of this rank's multi-GPU window (0: single GPU)
void commit_standard(std::vector<F> &poly, _hash &comm, std::vector<std::vector<_hash>> &MT_hashes,
                     std::vector<std::vector<std::vector<F>>> &_tensor, int K);
// The data-parallel front half of open_standard (Our_PC.cpp:604-660): beta = eq(x1), aggregate, the rand()-drawn
// queries I, the reply gather and the Merkle paths; open_standard (below) continues with the shockwave / WHIR / linear-code recursion.
struct open_front { std::vector<F> beta, aggr_vector; std::vector<std::vector<size_t>> I; std::vector<std::vector<F>> reply;
                    std::vector<std::vector<_hash>> commitment_paths; F r_v0; double ps = 0; };
open_front open_standard_front(std::vector<F> &poly, std::vector<F> x, std::vector<std::vector<_hash>> &Commitment_MT, int K);

// ---- the opening recursion (hobbit_open.cpp) -------------------------------------------------------------------------
// Same names and call order as the reference; the containers that hold TABLES are handles to HBM-resident data instead of host
// vectors (a maintainer dropping this into the reference keeps the call sites: they only pass these objects around).
struct host_graph { long long L = 0, R = 0; int deg = 0; const uint32_t *nbr = nullptr; const uint64_t *w = nullptr; };
const host_graph &expander_graph(int which /*0 = _C, 1 = D*/, int dep);      // the graphs expander_init_store drew (expanders.h:18)
// graphs drawn elsewhere (the reference's own expander_init_store, when this library is linked under the reference's callers):
// nbr[i*deg + j] / w[i*deg + j] as in hb_expander_set; C: degree 9, D: degree 12
struct ext_graph { long long L = 0, R = 0; std::vector<uint32_t> nbr; std::vector<uint64_t> w; };
void expander_adopt(long long n, int levels, const std::vector<ext_graph> &C, const std::vector<ext_graph> &D);
struct shockwave_data {                        // Virgo.h:27-51; matrix / encoded_matrix / MT live on the device
    int k = 0; size_t N = 0;
    F *matrix = nullptr, *encoded_matrix = nullptr;     // k x N/k and k x 2N/k, row-major
    uint8_t *MT = nullptr;                               // 2N/k leaves, every level, leaves first
    ~shockwave_data();
    std::vector<std::vector<_hash>> MT_host() const;     // for tests / serving paths
    std::vector<F> encoded_host() const;
};
struct Whir_data {                             // Virgo.h:52-60
    int k = 0; size_t N = 0;
    F *poly = nullptr, *poly_com = nullptr; uint8_t *MT = nullptr;
    std::vector<F *> FRI_poly; std::vector<uint8_t *> FRI_MT; std::vector<size_t> FRI_size;
    Whir_data() {}
    Whir_data(const Whir_data &) = delete; Whir_data &operator=(const Whir_data &) = delete;
    ~Whir_data();
    std::vector<std::vector<_hash>> MT_host() const;
    std::vector<std::vector<_hash>> FRI_MT_host(int i) const;
    std::vector<F> poly_host() const;
};
extern shockwave_data *C_f, *C_c;              // PC_utils.cpp:6-7
shockwave_data *shockwave_commit(std::vector<F> &poly, int k);
void shockwave_prove(shockwave_data *data, std::vector<F> x, double &vt, double &ps);        // deletes data, like the reference
void whir_commit(std::vector<F> &poly, Whir_data &data);
void _whir_prove(Whir_data &data, std::vector<F> x, double &vt, double &ps);
proof prove_fft(std::vector<F> &m, std::vector<F> r, F previous_sum, double &vt, double &ps);
proof prove_fft_matrix(std::vector<std::vector<F>> M, std::vector<F> r, F previous_sum, double &vt, double &ps);
proof prove_linear_code(std::vector<F> &codeword, int n, double &vt, double &ps);
void recursive_prover_Spielman(std::vector<F> &input, std::vector<std::vector<F>> &C, std::vector<size_t> I, double &vt, double &ps);
void recursive_prover_RS(std::vector<F> &aggregated_vector, std::vector<std::vector<size_t>> I, double &vt, double &ps);
void open_standard(std::vector<F> &poly, std::vector<F> x, std::vector<std::vector<_hash>> &Commitment_MT,
                   std::vector<std::vector<std::vector<F>>> &_tensor, int K, double &vt, double &ps);

// Elastic_PC
void init_commitment(bool mod);
void read_stream_PC(stream_descriptor &fd, F *v, int size);
void commit(stream_descriptor fd, _hash &comm, std::vector<std::vector<_hash>> &MT_hashes);
extern std::vector<std::vector<size_t>> I;                      // Elastic_PC.cpp:314
void open(stream_descriptor fd, std::vector<F> x, std::vector<std::vector<_hash>> &Commitment_MT, double &vt, double &ps);
// chunk i (BUFFER_SPACE elements) of a stream as read_stream emits it: a pointer into HBM for resident circuit streams, else `buff`
const F *stream_chunk(stream_descriptor &fd, size_t i, size_t B, std::vector<F> &buff);

// ---- circuit streams (hobbit_circuit.cpp): ONE pass of the evaluator's trace goes to HBM, every named stream is derived
// there and stays resident ------------------------------------------------------------------------------------------------------------
struct tr_tuple {                              // Seval.h:4-9
    F value_o, value_l, value_r;
    int idx_o, idx_l, idx_r;
    int access_o, access_l, access_r;
    uint8_t type;
};
extern size_t circuit_size;                    // main.cpp:36
extern F a_w, b_w;                             // main.cpp:63
extern bool has_lookups;                       // main.cpp:67
extern std::vector<F> lookup_rand;             // main.cpp:70
// the consumer side of the producer hand-off (what read_tr does, main.cpp:283-300): call trace_append with each refilled tr[] buffer until
// it returns true (type 255 seen); trace_end() sets and returns circuit_size (get_circuit_size, main.cpp:303-321)
void trace_begin(size_t capacity_hint = 0);
bool trace_append(const tr_tuple *buf, size_t n);
size_t trace_end();
void trace_generate_mlp(const std::vector<int> &layer_size);   // 8f.4: MLP_inference (Seval.cpp:1238-1286) evaluated on the GPU; then trace_end()
void trace_generate_aes(int input_size);                        // 8f.4: AES (Seval.cpp:991-1084, fun == 5 driver) evaluated on the GPU; then trace_end()
void trace_generate_sql(int input_size);                        // 8f.4: range_query (Seval.cpp:1085-1166, fun == 6 driver) evaluated on the GPU
void trace_generate_pruned_mlp(const std::vector<std::vector<std::vector<unsigned short>>> &indexes, int n_inputs = 128 * 128);   // 8f.4: `inference` (Seval.cpp:1170-1236, fun == 8 driver)
const F *resident_stream(const stream_descriptor &fd);          // the whole logical stream in HBM, nullptr for non-circuit streams
bool read_circuit_stream(stream_descriptor &fd, std::vector<F> &v, int size);
void read_trace(stream_descriptor &fd, std::vector<F> &buff_L, std::vector<F> &buff_R, std::vector<F> &buff_O, std::vector<int> &buff_S);
void prove_gate_consistency(stream_descriptor tr, std::vector<F> r, double &vt, double &ps);
void prove_gate_consistency_lookups(stream_descriptor tr, std::vector<F> r, double &vt, double &ps);

// sumcheck.h
proof generate_2product_sumcheck_proof(std::vector<F> &v1, std::vector<F> &v2, F previous_r, double &vt, double &ps);
proof _generate_3product_sumcheck_proof(std::vector<F> &v1, std::vector<F> &v2, std::vector<F> &v3, F previous_r, double &vt, double &ps);
proof batch_3product_sumcheck(std::vector<std::vector<F>> &arr1, std::vector<std::vector<F>> &arr2, std::vector<std::vector<F>> &arr3,
                              std::vector<F> a, double &vt, double &ps);
mul_tree_proof prove_multiplication_tree_new(std::vector<std::vector<F>> &input, F previous_r, std::vector<F> prev_x, double &vt, double &ps);
// witness_stream.h: name-dispatched producer.  Only the synthetic default stream (v[i] = F(i%1024+1)) exists without the
// circuit trace; the circuit streams ("witness", "wiring_consistency_check_opt", "circuit", the lookup streams) are served from the
// HBM-resident trace (hobbit_circuit.cpp).
void read_stream(stream_descriptor &fd, std::vector<F> &v, int size);
void reset_stream(stream_descriptor &fd);
// sumcheck.cpp:1746-1915.  The stream is materialised ONCE into HBM (instead of being re-generated per pass) and every layer is
// proven there; supported: size*vectors <= 2*BUFFER_SPACE, or layers <= distance, or naive.
std::vector<F> prove_multiplication_tree_stream_shallow(stream_descriptor fd, int vectors, int size, F previous_r, int distance,
                                                        std::vector<F> prev_x, bool naive, double &vt, double &ps);

}  // namespace hobbit
