// Link-level drop-in: the reference's entry points for the hot path, DEFINED IN THE GLOBAL NAMESPACE WITH THE REFERENCE'S OWN TYPES, each
// forwarding to the host mirror (hobbit::, hobbit_host.hpp) and through it to the CUDA library.  A maintainer of the reference adds this one
// file to the build, removes (or weakens) the reference's definitions of the same functions and links libhobbit_host.so + libhobbit_b200.so:
// main.cpp, test_PC and test_Elastic_PC then run on the GPU unchanged.  oracle/Makefile builds exactly that (target pigeon_gpu: the
// reference's objects with these symbols weakened by objcopy) and tests/test_link_dropin.py runs the reference's own `main` on it.
//
// Compiled against the REFERENCE's headers (-I<reference>/src); it is the only file of the product that includes them.
//   replaced function                         reference definition
//   commit_standard / open_standard           Our_PC.cpp:146-171, 604-692
//   commit / open                             Elastic_PC.cpp:174-285, 625-726
//   prove_multiplication_tree_stream_shallow  sumcheck.cpp:1746-1915
//   prove_gate_consistency / _lookups         sumcheck.cpp:796-981, 503-794
// The container types are layout-compatible by construction: virgo::fieldElement == hobbit::F (two u64), _hash == 32 bytes, std::vector
// is std::vector; stream_descriptor is copied member by member.
#include "hobbit_host.hpp"
namespace hobbit { typedef F Fe; }
#include "config_pc.hpp"
#include "utils.hpp"
#include "Our_PC.hpp"
#include "expanders.h"
#include "witness_stream.h"
#include "Elastic_PC.hpp"
#include "sumcheck.h"
#include "Seval.h"
#include <cmath>
#include <cstring>
#include <mutex>

extern bool linear_time;
extern int tensor_row_size;
extern size_t BUFFER_SPACE;
extern size_t circuit_size;
extern F a_w, b_w;
extern bool has_lookups;
extern std::vector<F> lookup_rand;
extern std::mutex mtx, mtx2;
extern tr_tuple *tr;
extern int BUFFER_SPACE_tr;
extern bool __encode_initialized;
extern std::vector<std::vector<int>> access_table, lookup_table;      // witness_stream.cpp:20, Seval.cpp:17

static_assert(sizeof(F) == sizeof(hobbit::Fe) && sizeof(_hash) == sizeof(hobbit::_hash), "field / digest layout");
static_assert(sizeof(tr_tuple) == sizeof(hobbit::tr_tuple), "trace tuple layout");

namespace {
template <class T, class U> T &as(U &x) { return reinterpret_cast<T &>(x); }
// the mirror's stream_descriptor is the reference's without the trailing input_* members (witness_stream.h:15-16, unused on this path)
hobbit::stream_descriptor fd_of(const stream_descriptor &fd) {
    hobbit::stream_descriptor h;
    // the reference's callers set name and size and reset the read position (reset_stream); data_size / row_size / col_size / layer /
    // tree_pos have NO initialiser in the reference (witness_stream.h:12) and hold garbage there — they stay zero here
    h.idx = fd.idx; h.offset = fd.offset; h.stage = fd.stage; h.finished = fd.finished; h.pos = fd.pos; h.pos_j = fd.pos_j;
    h.size = fd.size; h.name = fd.name;
    return h;
}
std::vector<hobbit::Fe> &fv(std::vector<F> &v) { return reinterpret_cast<std::vector<hobbit::Fe> &>(v); }
std::vector<std::vector<hobbit::_hash>> &hv(std::vector<std::vector<_hash>> &v) { return reinterpret_cast<std::vector<std::vector<hobbit::_hash>> &>(v); }

void sync_globals() {
    // the reference's callers only hand MT_hashes from commit() to open() (prove_circuit main.cpp:862-951, test_Elastic_PC): the levels may
    // arrive in the background
    hobbit::commit_levels_async = true;
    hobbit::linear_time = linear_time; hobbit::tensor_row_size = tensor_row_size; hobbit::BUFFER_SPACE = BUFFER_SPACE;
    hobbit::has_lookups = has_lookups; hobbit::circuit_size = circuit_size;
    hobbit::a_w = hobbit::Fe(a_w.real, a_w.img); hobbit::b_w = hobbit::Fe(b_w.real, b_w.img);
    hobbit::lookup_rand = fv(lookup_rand);
}
// The graphs the reference's expander_init_store has drawn (expanders.h:78-92 fills _C[] / D[]) are handed to the GPU as they are.
long long adopted_n = -1;
void adopt_expander(long long n) {
    if (adopted_n == n) return;
    int levels = 0; long long nn = n;
    while (nn > distance_threshold) { nn = _C[levels].R; levels++; }
    std::vector<hobbit::ext_graph> gc(levels), gd(levels);
    auto conv = [](const graph &g, hobbit::ext_graph &o) {
        o.L = g.L; o.R = g.R; o.nbr.resize((size_t)g.L * g.degree); o.w.resize((size_t)g.L * g.degree);
        for (long long i = 0; i < g.L; i++) for (int j = 0; j < g.degree; j++) { o.nbr[i * g.degree + j] = (uint32_t)g.neighbor[i][j]; o.w[i * g.degree + j] = g.weight[i][j].real; }
    };
    for (int d = 0; d < levels; d++) { conv(_C[d], gc[d]); conv(D[d], gd[d]); }
    hobbit::expander_adopt(n, levels, gc, gd);
    adopted_n = n;
}
// The circuit streams: ONE pass of the reference's evaluator thread (Seval_Oracle, started by main) is taken over the same mutex hand-off
// read_tr uses (main.cpp:283-300) and kept in HBM; every stream of every later call is derived from it on the GPU.
bool trace_taken = false;
void take_trace() {
    if (trace_taken) return;
    hobbit::trace_begin(2 * circuit_size);
    bool counted = false, ended = false;
    while (true) {
        mtx.unlock(); mtx2.lock();
        // The reference's lookup-stream readers leave the per-entry access counts of the pass in the global `access_table`
        // (witness_stream.cpp:931-936, 2199-2208), and prove_circuit's closing product check reads them (main.cpp:925-948): the same
        // counts are taken from this one pass.
        if (has_lookups && !ended) {
            if (!counted) {
                access_table.assign(lookup_table.size(), std::vector<int>());
                for (size_t t = 0; t < lookup_table.size(); t++) access_table[t].assign(lookup_table[t].size(), 0);
                counted = true;
            }
            for (int i = 0; i < BUFFER_SPACE_tr; i++) {
                if (tr[i].type == 255) { ended = true; break; }
                if (tr[i].type >= 3) access_table[tr[i].type - 3][tr[i].type == 3 ? tr[i].value_l.real : tr[i].value_l.real + 256 * tr[i].value_r.real]++;
            }
        }
        if (hobbit::trace_append(reinterpret_cast<const hobbit::tr_tuple *>(tr), (size_t)BUFFER_SPACE_tr)) break;
    }
    hobbit::trace_end();
    trace_taken = true;
}
bool circuit_stream(const std::string &n) {
    return n == "witness" || n == "wiring_consistency_check_opt" || n == "transcript_stream" || n == "lookup_basic" || n == "lookup_witness_basic" || n == "circuit";
}
}  // namespace

// ---- Our_PC.hpp:11-12 ---------------------------------------------------------------------------------------------------------------
void commit_standard(vector<F> &poly, _hash &comm, vector<vector<_hash>> &MT_hashes, vector<vector<vector<F>>> &_tensor, int K) {
    sync_globals();
    if (linear_time) adopt_expander(tensor_row_size);
    hobbit::commit_standard(fv(poly), as<hobbit::_hash>(comm), hv(MT_hashes), reinterpret_cast<std::vector<std::vector<std::vector<hobbit::Fe>>> &>(_tensor), K);
}
void open_standard(vector<F> &poly, vector<F> x, vector<vector<_hash>> &Commitment_MT, vector<vector<vector<F>>> &_tensor, int K, double &vt, double &ps) {
    sync_globals();
    if (linear_time) adopt_expander(tensor_row_size);
    hobbit::open_standard(fv(poly), fv(x), hv(Commitment_MT), reinterpret_cast<std::vector<std::vector<std::vector<hobbit::Fe>>> &>(_tensor), K, vt, ps);
}
// ---- Elastic_PC.hpp:12-14 -----------------------------------------------------------------------------------------------------------
void commit(stream_descriptor fd, _hash &comm, vector<vector<_hash>> &MT_hashes) {
    sync_globals();
    if (linear_time) adopt_expander(tensor_row_size);
    if (circuit_stream(fd.name)) take_trace();
    hobbit::commit(fd_of(fd), as<hobbit::_hash>(comm), hv(MT_hashes));
}
void open(stream_descriptor fd, vector<F> x, vector<vector<_hash>> &Commitment_MT, double &vt, double &ps) {
    sync_globals();
    if (linear_time) adopt_expander(tensor_row_size);
    if (circuit_stream(fd.name)) take_trace();
    hobbit::open(fd_of(fd), fv(x), hv(Commitment_MT), vt, ps);
}
// ---- sumcheck.h:83-95 ---------------------------------------------------------------------------------------------------------------
vector<F> prove_multiplication_tree_stream_shallow(stream_descriptor fd, int vectors, int size, F previous_r, int distance, vector<F> prev_x, bool naive,
                                                   double &vt, double &ps) {
    sync_globals();
    if (circuit_stream(fd.name)) take_trace();
    std::vector<hobbit::Fe> r = hobbit::prove_multiplication_tree_stream_shallow(fd_of(fd), vectors, size, hobbit::Fe(previous_r.real, previous_r.img), distance, fv(prev_x),
                                                                                  naive, vt, ps);
    vector<F> out(r.size());
    memcpy(out.data(), r.data(), r.size() * sizeof(F));
    return out;
}
void prove_gate_consistency(stream_descriptor tr_fd, vector<F> r, double &vt, double &ps) {
    sync_globals(); take_trace();
    hobbit::prove_gate_consistency(fd_of(tr_fd), fv(r), vt, ps);
}
void prove_gate_consistency_lookups(stream_descriptor tr_fd, vector<F> r, double &vt, double &ps) {
    sync_globals(); take_trace();
    hobbit::prove_gate_consistency_lookups(fd_of(tr_fd), fv(r), vt, ps);
}
