// hobbit_circuit — the witness_stream interface for circuit streams, host side.
//
// Reference: the producer thread Seval_Oracle refills `tr[BUFFER_SPACE_tr]` and hands it over under mtx/mtx2 (src/main.cpp:283-321,
// src/Seval.cpp:1273-1283); the readers read_stream / read_trace (src/witness_stream.cpp:1701-1807, 2106-2353) pull from it and
// RE-EXECUTE the circuit once per pass (commit: 2 passes, product tree: 2 per layer, gate consistency: 3, open: 4).
// Here the consumer of that hand-off (trace_append) uploads ONE pass into HBM; every named stream is derived there by the kernels in
// csrc/trace.cu and stays resident, and commit / open / prove_multiplication_tree_stream_shallow / prove_gate_consistency take the
// resident arrays (stream_chunk / resident_stream) instead of calling the readers.  read_stream / read_trace remain available with
// the reference's block semantics for other callers (they copy from HBM).
#include "hobbit_host.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace hobbit {

[[noreturn]] static void die(const char *what) { printf("hobbit_b200: %s: %s\n", what, hb_last_error(backend())); exit(-1); }
#define CK(call) do { if (call) die(#call); } while (0)
static inline const hb_F *abi(const F *p) { return reinterpret_cast<const hb_F *>(p); }
static inline hb_F *abi(F *p) { return reinterpret_cast<hb_F *>(p); }

size_t circuit_size = 0;                       // main.cpp:36
F a_w, b_w;                                    // main.cpp:63
bool has_lookups = false;                      // main.cpp:67
std::vector<F> lookup_rand;                    // main.cpp:70 (4 values, drawn by prove_circuit after the witness commitment)
static_assert(sizeof(tr_tuple) == 80, "tr_tuple must match the reference layout (Seval.h:4-9)");

namespace {
struct Resident { F *p = nullptr; size_t n = 0; void ensure(size_t want) { if (n != want) { if (p) hb_free_device(backend(), p); void *q; CK(hb_malloc_device(backend(), &q, want * sizeof(F))); p = (F *)q; n = want; } } };
Resident g_witness, g_trL, g_trR, g_trO, g_trS, g_wiring, g_lkp_basic, g_lkp_wit, g_circuit;
bool have_circuit = false;
bool have_witness = false, have_transcript = false, have_wiring = false, tr_lookups = false, have_lkp_basic = false, have_lkp_wit = false;
F wiring_a, wiring_b;
std::vector<F> lkp_rand_basic, lkp_rand_wit;
bool trace_loaded = false;
}

void trace_begin(size_t capacity_hint) {
    CK(hb_trace_begin(backend(), capacity_hint));
    have_witness = have_transcript = have_wiring = trace_loaded = have_lkp_basic = have_lkp_wit = have_circuit = false;
}
bool trace_append(const tr_tuple *buf, size_t n) {
    int done = 0;
    CK(hb_trace_push(backend(), buf, n, &done));
    return done != 0;
}
// 8f.4: for the MLP circuit (fun == 9) the trace can be produced on the GPU instead of by the producer thread; follow with trace_end()
void trace_generate_mlp(const std::vector<int> &layer_size) {
    size_t n = 0;
    CK(hb_trace_generate_mlp(backend(), layer_size.data(), (int)layer_size.size(), &n));
    have_witness = have_transcript = have_wiring = trace_loaded = have_lkp_basic = have_lkp_wit = have_circuit = false;
}
// 8f.4: the AES circuit (fun == 5, input_size = 2^n blocks) likewise; follow with trace_end()
void trace_generate_aes(int input_size) {
    size_t n = 0;
    CK(hb_trace_generate_aes(backend(), input_size, &n));
    have_witness = have_transcript = have_wiring = trace_loaded = have_lkp_basic = have_lkp_wit = have_circuit = false;
}
// 8f.4: the pruned MLP (fun == 8) likewise; `indexes` is the reference driver's sparsity pattern (Seval.cpp:1427-1437, drawn from libc rand())
void trace_generate_pruned_mlp(const std::vector<std::vector<std::vector<unsigned short>>> &indexes, int n_inputs) {
    std::vector<int> rp[2], cols[2];
    for (int l = 0; l < 2; l++) {
        rp[l].push_back(0);
        for (const auto &row : indexes[l]) { for (unsigned short c : row) cols[l].push_back((int)c); rp[l].push_back((int)cols[l].size()); }
        if (cols[l].empty()) cols[l].push_back(0);
    }
    size_t n = 0;
    CK(hb_trace_generate_pruned_mlp(backend(), n_inputs, (int)indexes[0].size(), (int)indexes[1].size(), rp[0].data(), cols[0].data(), rp[1].data(),
                                    cols[1].data(), &n));
    have_witness = have_transcript = have_wiring = trace_loaded = have_lkp_basic = have_lkp_wit = have_circuit = false;
}
// 8f.4: the SQL range query (fun == 6, input_size = 2^n rows) likewise
void trace_generate_sql(int input_size) {
    size_t n = 0;
    CK(hb_trace_generate_sql(backend(), input_size, &n));
    have_witness = have_transcript = have_wiring = trace_loaded = have_lkp_basic = have_lkp_wit = have_circuit = false;
}
// get_circuit_size (main.cpp:303-321): the number of delete records, rounded up to a power of two
size_t trace_end() {
    size_t n = 0, ops = 0, dels = 0;
    CK(hb_trace_finish(backend(), &n, &ops, &dels));
    size_t cs = dels;
    if (cs == 0 || ((size_t)1 << (int)std::log2((double)cs)) != cs) cs = (size_t)1 << ((int)std::log2((double)(cs ? cs : 1)) + 1);
    circuit_size = cs;
    trace_loaded = true;
    return cs;
}

static void need_trace(const char *name) {
    if (!trace_loaded) { printf("hobbit_b200: stream '%s' needs the evaluator's trace (trace_begin / trace_append / trace_end)\n", name); exit(-1); }
}
static const F *witness_dev() {
    need_trace("witness");
    if (!have_witness) { g_witness.ensure(4 * circuit_size); CK(hb_trace_witness(backend(), circuit_size, abi(g_witness.p))); have_witness = true; }
    return g_witness.p;
}
static void transcript_dev() {
    need_trace("transcript_stream");
    if (!have_transcript || tr_lookups != has_lookups) {
        for (Resident *r : {&g_trL, &g_trR, &g_trO, &g_trS}) r->ensure(circuit_size);
        CK(hb_trace_transcript(backend(), circuit_size, has_lookups ? 1 : 0, abi(g_trL.p), abi(g_trR.p), abi(g_trO.p), abi(g_trS.p)));
        have_transcript = true; tr_lookups = has_lookups;
    }
}
static const F *wiring_dev() {
    need_trace("wiring_consistency_check_opt");
    if (!have_wiring || wiring_a != a_w || wiring_b != b_w) {
        g_wiring.ensure(8 * circuit_size);
        CK(hb_trace_wiring(backend(), circuit_size, abi(&a_w), abi(&b_w), abi(g_wiring.p)));
        have_wiring = true; wiring_a = a_w; wiring_b = b_w;
    }
    return g_wiring.p;
}

static const F *circuit_dev() {
    need_trace("circuit");
    if (has_lookups) { printf("hobbit_b200: stream 'circuit' with lookups is not built\n"); exit(-1); }
    if (!have_circuit) { g_circuit.ensure(16 * circuit_size); CK(hb_trace_circuit(backend(), circuit_size, abi(g_circuit.p))); have_circuit = true; }
    return g_circuit.p;
}
static void need_lookup_rand() {
    if (lookup_rand.size() < 4) { printf("hobbit_b200: lookup streams need lookup_rand (4 values, main.cpp:911)\n"); exit(-1); }
}
static const F *lookup_basic_dev() {
    need_trace("lookup_basic"); need_lookup_rand();
    if (!have_lkp_basic || lkp_rand_basic != lookup_rand) {
        g_lkp_basic.ensure(2 * circuit_size);
        CK(hb_trace_lookup_basic(backend(), circuit_size, abi(lookup_rand.data()), abi(g_lkp_basic.p)));
        have_lkp_basic = true; lkp_rand_basic = lookup_rand;
    }
    return g_lkp_basic.p;
}
static const F *lookup_witness_dev() {
    need_trace("lookup_witness_basic"); need_lookup_rand();
    if (!have_lkp_wit || lkp_rand_wit != lookup_rand) {
        g_lkp_wit.ensure(2 * circuit_size);
        CK(hb_trace_lookup_witness(backend(), circuit_size, abi(lookup_rand.data()), abi(g_lkp_wit.p)));
        have_lkp_wit = true; lkp_rand_wit = lookup_rand;
    }
    return g_lkp_wit.p;
}

// the whole logical stream in HBM, or nullptr when `fd` is not a circuit stream
const F *resident_stream(const stream_descriptor &fd) {
    if (fd.name == "witness") return witness_dev();
    if (fd.name == "wiring_consistency_check_opt") return wiring_dev();
    if (fd.name == "circuit") return circuit_dev();
    if (fd.name == "lookup_basic") return lookup_basic_dev();
    if (fd.name == "lookup_witness_basic") return lookup_witness_dev();
    return nullptr;
}

// read_stream for the circuit names (witness_stream.cpp:2163-2178, 2276-2311): block `fd.pos` of `size` elements, copied out of HBM.
// "wiring_consistency_check_opt" blocks are X-half | Y-half of size/2 each.
bool read_circuit_stream(stream_descriptor &fd, std::vector<F> &v, int size) {
    if (fd.name == "witness" || fd.name == "circuit") {
        const bool wit = fd.name == "witness";
        const F *w = wit ? witness_dev() : circuit_dev();
        const size_t len = (wit ? 4 : 16) * circuit_size;
        size_t off = fd.pos * (size_t)size;
        if (off + size > len) { printf("hobbit_b200: read past the end of stream '%s'\n", fd.name.c_str()); exit(-1); }
        CK(hb_memcpy(backend(), v.data(), w + off, (size_t)size * sizeof(F)));
        fd.pos = (fd.pos + 1) % (len / size);
        return true;
    }
    if (fd.name == "wiring_consistency_check_opt" || fd.name == "lookup_basic") {           // two-half streams: X block | Y block per read
        const bool wiring = fd.name == "wiring_consistency_check_opt";
        const F *xy = wiring ? wiring_dev() : lookup_basic_dev();
        const size_t half_len = wiring ? 4 * circuit_size : circuit_size;
        size_t h = (size_t)size / 2, off = fd.pos * h;
        if (off + h > half_len) { printf("hobbit_b200: read past the end of stream '%s'\n", fd.name.c_str()); exit(-1); }
        CK(hb_memcpy(backend(), v.data(), xy + off, h * sizeof(F)));
        CK(hb_memcpy(backend(), v.data() + h, xy + half_len + off, h * sizeof(F)));
        fd.pos = (fd.pos + 1) % (half_len / h);
        return true;
    }
    if (fd.name == "lookup_witness_basic") {
        const F *w = lookup_witness_dev();
        size_t off = fd.pos * (size_t)size;
        if (off + size > 2 * circuit_size) { printf("hobbit_b200: read past the end of stream 'lookup_witness_basic'\n"); exit(-1); }
        CK(hb_memcpy(backend(), v.data(), w + off, (size_t)size * sizeof(F)));
        fd.pos = (fd.pos + 1) % (2 * circuit_size / size);
        return true;
    }
    return false;
}
// read_trace (witness_stream.cpp:1701-1807): the next block of the gate transcript
void read_trace(stream_descriptor &fd, std::vector<F> &buff_L, std::vector<F> &buff_R, std::vector<F> &buff_O, std::vector<int> &buff_S) {
    transcript_dev();
    const size_t n = buff_L.size(), off = fd.pos * n;
    if (off + n > circuit_size) { printf("hobbit_b200: read past the end of the gate transcript\n"); exit(-1); }
    std::vector<F> s(n);
    CK(hb_memcpy(backend(), buff_L.data(), g_trL.p + off, n * sizeof(F)));
    CK(hb_memcpy(backend(), buff_R.data(), g_trR.p + off, n * sizeof(F)));
    CK(hb_memcpy(backend(), buff_O.data(), g_trO.p + off, n * sizeof(F)));
    CK(hb_memcpy(backend(), s.data(), g_trS.p + off, n * sizeof(F)));
    for (size_t i = 0; i < n; i++) buff_S[i] = (int)s[i].real;
    fd.pos = (fd.pos + 1) % (circuit_size / n);
}

// prove_gate_consistency (sumcheck.cpp:796-981) on the resident transcript.  libc draws in the reference's order: generate_randomness(4)
// for the batching of the degree-4 sumcheck (:877-880), generate_randomness(6) for the Peval combination (:958).
void prove_gate_consistency(stream_descriptor tr, std::vector<F> r, double &vt, double &ps) {
    trace_rng("prove_gate_consistency");
    (void)vt;
    if (has_lookups) { printf("hobbit_b200: with has_lookups call prove_gate_consistency_lookups\n"); exit(-1); }
    transcript_dev();
    const size_t cs = tr.size, B = BUFFER_SPACE, nch = cs / B;
    const int lgB = (int)std::log2((double)B), lgn = (int)std::log2((double)nch);
    std::vector<F> rnd = generate_randomness(4), b6 = generate_randomness(6);
    rnd.insert(rnd.end(), b6.begin(), b6.end());
    std::vector<F> out(nch + 6 * (size_t)lgB + 6 + 6 * nch + 4 * (size_t)lgn + 3);
    CK(hb_gate_consistency_stream(backend(), abi(g_trL.p), abi(g_trR.p), abi(g_trO.p), abi(g_trS.p), cs, B, abi(r.data()), abi(rnd.data()),
                                  abi(out.data()), &ps));
}

// prove_gate_consistency_lookups (sumcheck.cpp:503-794).  libc draws: generate_randomness(5) (:637), generate_randomness(8) (:770).
void prove_gate_consistency_lookups(stream_descriptor tr, std::vector<F> r, double &vt, double &ps) {
    trace_rng("prove_gate_consistency_lookups");
    (void)vt;
    if (!has_lookups) { printf("hobbit_b200: prove_gate_consistency_lookups needs has_lookups\n"); exit(-1); }
    need_lookup_rand();
    transcript_dev();
    const size_t cs = tr.size, B = BUFFER_SPACE, nch = cs / B;
    const int lgB = (int)std::log2((double)B), lgn = (int)std::log2((double)nch);
    std::vector<F> rnd = generate_randomness(5), b8 = generate_randomness(8);
    rnd.insert(rnd.end(), b8.begin(), b8.end());
    std::vector<F> out(nch + 6 * (size_t)lgB + 9 + 8 * nch + 4 * (size_t)lgn + 3);
    CK(hb_gate_consistency_lookups_stream(backend(), abi(g_trL.p), abi(g_trR.p), abi(g_trO.p), abi(g_trS.p), cs, B, abi(r.data()), abi(lookup_rand.data()),
                                          abi(rnd.data()), abi(out.data()), &ps));
}

}  // namespace hobbit
