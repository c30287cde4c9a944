// Circuit drop-in test (SURVEY §8f.2 + BASELINE config 3): the UNMODIFIED reference proves a circuit the way prove_circuit does
// (src/main.cpp:862-886: commit(witness) -> prove_multiplication_tree_stream_shallow(wiring) -> prove_gate_consistency(transcript) ->
// open(witness)) with its live producer thread re-executing the circuit for every pass, and the host mirror does the same from ONE pass
// of the trace uploaded to HBM.  Every named stream, every Merkle level, the product-tree outputs, ps and the libc RNG state must agree.
// One circuit per process (the reference's producer thread never exits).  Built twice like open_test (GPU / CPU emulation).
// usage: circ_test <log2 BUFFER_SPACE> <layer sizes...>      e.g. circ_test 12 64 32 16   (MLP, `pigeon 9 b b 1 n l0 l1 ...`)
//        circ_test <log2 BUFFER_SPACE> aes <n> <d>           e.g. circ_test 19 aes 8 1    (`pigeon 5 19 8 1`, lookups; main.cpp:887-917)
//        circ_test <log2 BUFFER_SPACE> sql <n> <d>           e.g. circ_test 19 sql 17 1   (`pigeon 6 19 17 1`)
//        circ_test <log2 BUFFER_SPACE> pruned <n> <d> <rate> e.g. circ_test 20 pruned 20 1 10 (`pigeon 8 20 20 1 10`: pruned MLP through
//                                                            prove_arbitrary_circuit, main.cpp:812-858: also opens the "circuit" stream)
#include "../../hobbit_b200/host/hobbit_host.hpp"
namespace hobbit { typedef F Fe; }
#include "config_pc.hpp"
#include "utils.hpp"
#include "mimc.h"
#include "Our_PC.hpp"
#include "witness_stream.h"
#include "Elastic_PC.hpp"
#include "sumcheck.h"
#include "Seval.h"
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>

extern int fun;
extern size_t circuit_size;
extern F a_w, b_w;
extern std::mutex mtx, mtx2;
extern std::vector<int> layer_size;
extern tr_tuple *tr;
extern int BUFFER_SPACE_tr;
extern size_t BUFFER_SPACE;
extern bool has_lookups;
extern vector<F> lookup_rand;
extern double prune_rate;
void Seval_Oracle();
void prove_gate_consistency_lookups(stream_descriptor tr, vector<F> r, double &vt, double &ps);
void init_stream(int b, int n, int d);
extern bool linear_time;
extern int tensor_row_size;

static int failures = 0;
#define CHECK(cond, what) do { if (!(cond)) { printf("FAIL: %s (%s:%d)\n", what, __FILE__, __LINE__); failures++; } else printf("ok:   %s\n", what); fflush(stdout); } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static bool same_levels(const vector<vector<_hash>> &a, const vector<vector<hobbit::_hash>> &b) {
    if (a.size() != b.size()) return false;
    for (size_t l = 0; l < a.size(); l++) if (a[l].size() != b[l].size() || memcmp(a[l].data(), b[l].data(), a[l].size() * 32)) return false;
    return true;
}
static vector<hobbit::Fe> conv(const vector<F> &v) { vector<hobbit::Fe> o(v.size()); memcpy(o.data(), v.data(), v.size() * 16); return o; }

int main(int argc, char **argv) {
    int b = argc > 1 ? atoi(argv[1]) : 12;
    init_hash();
    hobbit::init_backend(0);
    mtx.lock(); mtx2.lock();                                       // main.cpp:1172-1174
    bool lookups = false; int n_arg = b, d_arg = 1;
    if (argc > 2 && (!strcmp(argv[2], "aes") || !strcmp(argv[2], "sql"))) {
        fun = !strcmp(argv[2], "aes") ? 5 : 6; lookups = true;
        n_arg = argc > 3 ? atoi(argv[3]) : 8; d_arg = argc > 4 ? atoi(argv[4]) : 1;
    } else if (argc > 2 && !strcmp(argv[2], "pruned")) {
        fun = 8; n_arg = argc > 3 ? atoi(argv[3]) : 20; d_arg = argc > 4 ? atoi(argv[4]) : 1;
        prune_rate = (argc > 5 ? atoi(argv[5]) : 1) / 100.0;
    } else {
        fun = 9;
        for (int i = 2; i < argc; i++) layer_size.push_back(atoi(argv[i]));
        if (layer_size.empty()) layer_size = {64, 32, 16};
    }
    // fun == 8: the producer thread draws the sparsity pattern from libc rand() (Seval.cpp:1427-1437); the same expressions, compiled with the
    // same flags, are replayed here from the same RNG state so that the GPU evaluator can be handed the pattern the reference will use
    vector<vector<vector<unsigned short>>> pruned_idx(2);
    if (fun == 8) {
        srand(1);
        pruned_idx[0].resize(1024); pruned_idx[1].resize(128);
        for (int i = 0; i < (int)(prune_rate * 1024 * 128 * 128); i++) pruned_idx[0][((unsigned int)rand()) % 1024].push_back(((unsigned int)rand()) % (128 * 128));
        for (int i = 0; i < (int)(prune_rate * 1024 * 128); i++) pruned_idx[1][((unsigned int)rand()) % 128].push_back(((unsigned int)rand()) % (256));
        srand(1);
    }
    std::thread t(Seval_Oracle); t.detach();
    init_stream(b, n_arg, d_arg);
    const size_t cs = circuit_size, B = BUFFER_SPACE;
    printf("circuit_size 2^%d, BUFFER_SPACE 2^%d\n", (int)log2(cs), (int)log2(B));

    // ---- one pass of the trace -> HBM (the consumer side of the hand-off, as read_tr does it, main.cpp:283-300) -----------------------
    double t0 = now();
    hobbit::trace_begin(2 * cs);
    while (true) {
        mtx.unlock(); mtx2.lock();
        if (hobbit::trace_append(reinterpret_cast<const hobbit::tr_tuple *>(tr), (size_t)BUFFER_SPACE_tr)) break;
    }
    size_t hcs = hobbit::trace_end();
    double t_trace = now() - t0;
    CHECK(hcs == cs, "circuit_size from the trace == get_circuit_size()");
    hobbit::BUFFER_SPACE = B; hobbit::has_lookups = lookups; has_lookups = lookups;

    srand(7); a_w = random(); b_w = random();
    hobbit::a_w = hobbit::Fe(a_w.real, a_w.img); hobbit::b_w = hobbit::Fe(b_w.real, b_w.img);
    if (lookups) { lookup_rand = generate_randomness(4); hobbit::lookup_rand = conv(lookup_rand); }
    double vt = 0, tr_ref[4] = {0, 0, 0, 0}, tr_gpu[4] = {0, 0, 0, 0};

    // ---- the named streams, element by element ---------------------------------------------------------------------------------------
    {
        stream_descriptor fd; fd.name = "witness"; fd.size = 4 * cs; reset_stream(fd);
        hobbit::stream_descriptor hfd; hfd.name = "witness"; hfd.size = 4 * cs;
        vector<F> v(B); vector<hobbit::Fe> hv(B); bool ok = true;
        for (size_t off = 0; off < 4 * cs; off += B) { read_stream(fd, v, (int)B); hobbit::read_stream(hfd, hv, (int)B); ok = ok && !memcmp(v.data(), hv.data(), B * 16); }
        CHECK(ok, "stream \"witness\" (4 cs), read in blocks of BUFFER_SPACE");
        stream_descriptor fw; fw.name = "wiring_consistency_check_opt"; fw.size = 8 * cs; reset_stream(fw);
        hobbit::stream_descriptor hfw; hfw.name = "wiring_consistency_check_opt"; hfw.size = 8 * cs;
        vector<F> w(2 * B); vector<hobbit::Fe> hw(2 * B); ok = true;
        for (size_t off = 0; off < 8 * cs; off += 2 * B) { read_stream(fw, w, (int)(2 * B)); hobbit::read_stream(hfw, hw, (int)(2 * B)); ok = ok && !memcmp(w.data(), hw.data(), 2 * B * 16); }
        CHECK(ok, "stream \"wiring_consistency_check_opt\" (8 cs), read in blocks of 2 BUFFER_SPACE (X half | Y half)");
        stream_descriptor ft; ft.name = "transcript_stream"; ft.size = cs; reset_stream(ft);
        hobbit::stream_descriptor hft; hft.name = "transcript_stream"; hft.size = cs;
        vector<F> l(B), r(B), o(B); vector<int> s(B); vector<hobbit::Fe> hl(B), hr(B), ho(B); vector<int> hs(B); ok = true;
        for (size_t off = 0; off < cs; off += B) {
            read_trace(ft, l, r, o, s); hobbit::read_trace(hft, hl, hr, ho, hs);
            ok = ok && !memcmp(l.data(), hl.data(), B * 16) && !memcmp(r.data(), hr.data(), B * 16) && !memcmp(o.data(), ho.data(), B * 16) && s == hs;
        }
        CHECK(ok, "gate transcript via read_trace (L, R, O, selector)");
        if (fun == 8) {
            // The reference cannot be the checker for "circuit": its reader read_memory_circuit (witness_stream.cpp:1944-2019) is declared int
            // and has no return statement — undefined behaviour that aborts at -O3 (the same path makes test_arb.sh crash).
            // The mirror's stream is checked structurally instead: selectors == the transcript selectors, zero tail, pair section non-empty.
            hobbit::stream_descriptor hfc; hfc.name = "circuit"; hfc.size = 16 * cs;
            vector<hobbit::Fe> all; all.reserve(16 * cs);
            for (size_t off = 0; off < 16 * cs; off += B) { hobbit::read_stream(hfc, hv, (int)B); all.insert(all.end(), hv.begin(), hv.end()); }
            hobbit::stream_descriptor hft2; hft2.name = "transcript_stream"; hft2.size = cs;
            ok = true; size_t nz_pairs = 0;
            for (size_t off = 0; off < cs; off += B) {
                hobbit::read_trace(hft2, hl, hr, ho, hs);
                for (size_t i = 0; ok && i < B; i++) ok = all[off + i].real == (unsigned long long)hs[i] && all[off + i].img == 0;
            }
            for (size_t i = cs; i < 9 * cs; i++) nz_pairs += all[i].real != 0;
            for (size_t i = 9 * cs; ok && i < 16 * cs; i++) ok = all[i].real == 0 && all[i].img == 0;
            CHECK(ok && nz_pairs > 0, "stream \"circuit\" (16 cs): selector section == transcript selectors, memory-transcript pairs, zero tail");
        }
        if (lookups) {
            stream_descriptor fl; fl.name = "lookup_basic"; fl.size = 2 * cs; reset_stream(fl);
            hobbit::stream_descriptor hfl; hfl.name = "lookup_basic"; hfl.size = 2 * cs;
            ok = true;
            for (size_t off = 0; off < 2 * cs; off += 2 * B) { read_stream(fl, w, (int)(2 * B)); hobbit::read_stream(hfl, hw, (int)(2 * B)); ok = ok && !memcmp(w.data(), hw.data(), 2 * B * 16); }
            CHECK(ok, "stream \"lookup_basic\" (2 cs), blocks of 2 BUFFER_SPACE (X half | Y half), access counters");
            stream_descriptor fq; fq.name = "lookup_witness_basic"; fq.size = 2 * cs; reset_stream(fq);
            hobbit::stream_descriptor hfq; hfq.name = "lookup_witness_basic"; hfq.size = 2 * cs;
            ok = true;
            for (size_t off = 0; off < 2 * cs; off += B) { read_stream(fq, v, (int)B); hobbit::read_stream(hfq, hv, (int)B); ok = ok && !memcmp(v.data(), hv.data(), B * 16); }
            CHECK(ok, "stream \"lookup_witness_basic\" (2 cs), blocks of BUFFER_SPACE");
        }
    }
    // ---- 8f.4: the same trace produced by the GPU evaluator instead of the producer thread (MLP, pruned MLP, AES): every stream again ----------------
    if (fun == 9 || fun == 5 || fun == 8 || fun == 6) {
        t0 = now();
        if (fun == 9) hobbit::trace_generate_mlp(layer_size); else if (fun == 5) hobbit::trace_generate_aes(1 << n_arg); else if (fun == 6) hobbit::trace_generate_sql(1 << n_arg); else hobbit::trace_generate_pruned_mlp(pruned_idx);
        size_t gcs = hobbit::trace_end();
        double t_eval = now() - t0;
        CHECK(gcs == cs, "GPU circuit evaluator: circuit_size");
        stream_descriptor fd; fd.name = "witness"; fd.size = 4 * cs; reset_stream(fd);
        hobbit::stream_descriptor hfd; hfd.name = "witness"; hfd.size = 4 * cs;
        vector<F> v(B); vector<hobbit::Fe> hv(B); bool ok = true;
        for (size_t off = 0; off < 4 * cs; off += B) { read_stream(fd, v, (int)B); hobbit::read_stream(hfd, hv, (int)B); ok = ok && !memcmp(v.data(), hv.data(), B * 16); }
        stream_descriptor fw; fw.name = "wiring_consistency_check_opt"; fw.size = 8 * cs; reset_stream(fw);
        hobbit::stream_descriptor hfw; hfw.name = "wiring_consistency_check_opt"; hfw.size = 8 * cs;
        vector<F> w(2 * B); vector<hobbit::Fe> hw(2 * B);
        for (size_t off = 0; off < 8 * cs; off += 2 * B) { read_stream(fw, w, (int)(2 * B)); hobbit::read_stream(hfw, hw, (int)(2 * B)); ok = ok && !memcmp(w.data(), hw.data(), 2 * B * 16); }
        stream_descriptor ft; ft.name = "transcript_stream"; ft.size = cs; reset_stream(ft);
        hobbit::stream_descriptor hft; hft.name = "transcript_stream"; hft.size = cs;
        vector<F> l(B), r(B), o(B); vector<int> sv(B); vector<hobbit::Fe> hl(B), hr(B), ho(B); vector<int> hs(B);
        for (size_t off = 0; off < cs; off += B) {
            read_trace(ft, l, r, o, sv); hobbit::read_trace(hft, hl, hr, ho, hs);
            ok = ok && !memcmp(l.data(), hl.data(), B * 16) && !memcmp(r.data(), hr.data(), B * 16) && !memcmp(o.data(), ho.data(), B * 16) && sv == hs;
        }
        if (lookups) {
            stream_descriptor fl; fl.name = "lookup_basic"; fl.size = 2 * cs; reset_stream(fl);
            hobbit::stream_descriptor hfl; hfl.name = "lookup_basic"; hfl.size = 2 * cs;
            for (size_t off = 0; off < 2 * cs; off += 2 * B) { read_stream(fl, w, (int)(2 * B)); hobbit::read_stream(hfl, hw, (int)(2 * B)); ok = ok && !memcmp(w.data(), hw.data(), 2 * B * 16); }
            stream_descriptor fq; fq.name = "lookup_witness_basic"; fq.size = 2 * cs; reset_stream(fq);
            hobbit::stream_descriptor hfq; hfq.name = "lookup_witness_basic"; hfq.size = 2 * cs;
            for (size_t off = 0; off < 2 * cs; off += B) { read_stream(fq, v, (int)B); hobbit::read_stream(hfq, hv, (int)B); ok = ok && !memcmp(v.data(), hv.data(), B * 16); }
        }
        CHECK(ok, fun == 9 ? "GPU MLP evaluator (8f.4): witness, wiring and transcript streams identical to the reference's (labels, access counters, values)"
                : fun == 8 ? "GPU pruned-MLP evaluator (8f.4): witness, wiring and transcript streams identical to the reference's (sparsity pattern replayed from libc rand())"
                : fun == 6 ? "GPU SQL evaluator (8f.4): witness, wiring, transcript and both lookup streams identical to the reference's"
                           : "GPU AES evaluator (8f.4): witness, wiring, transcript and both lookup streams identical to the reference's");
        printf("      trace on the GPU in %.4f s (producer thread + upload: %.4f s)\n", t_eval, t_trace);
    }
    // ---- commit(witness) ---------------------------------------------------------------------------------------------------------------
    vector<vector<_hash>> MT, MT0; vector<vector<hobbit::_hash>> hMT, hMT0;
    {
        stream_descriptor fd; fd.name = "witness"; fd.size = 4 * cs; reset_stream(fd);
        init_commitment(false);
        _hash comm; t0 = now(); commit(fd, comm, MT); tr_ref[0] = now() - t0;
        hobbit::stream_descriptor hfd; hfd.name = "witness"; hfd.size = 4 * cs;
        hobbit::init_commitment(false);
        hobbit::_hash hcomm; t0 = now(); hobbit::commit(hfd, hcomm, hMT); tr_gpu[0] = now() - t0;
        vector<vector<_hash>> A = MT; vector<vector<hobbit::_hash>> Bm = hMT;
        A[0].back() = _hash(); memset(&Bm[0].back(), 0, 32);              // the reference's last leaf reads past its buffers
        CHECK(same_levels(A, Bm), "commit(witness): every Merkle level");
        if (fun == 8) { MT0 = MT; hMT0 = hMT; }
    }
    vector<vector<_hash>> MTl; vector<vector<hobbit::_hash>> hMTl;
    if (lookups) {          // main.cpp:913: read_stream_PC does not know this name and commits its synthetic default stream; so does the mirror
        stream_descriptor fd; fd.name = "lookup_witness_basic"; fd.size = 2 * cs; reset_stream(fd);
        _hash comm; commit(fd, comm, MTl);
        hobbit::stream_descriptor hfd; hfd.name = "lookup_witness_basic"; hfd.size = 2 * cs;
        hobbit::_hash hcomm; hobbit::commit(hfd, hcomm, hMTl);
        vector<vector<_hash>> A = MTl; vector<vector<hobbit::_hash>> Bm = hMTl;
        A[0].back() = _hash(); memset(&Bm[0].back(), 0, 32);
        CHECK(same_levels(A, Bm), "commit(lookup_witness_basic): every Merkle level");
    }
    // ---- prove_multiplication_tree_stream_shallow(wiring, 8 vectors) ---------------------------------------------------------------------
    {
        stream_descriptor fd; fd.name = "wiring_consistency_check_opt"; fd.size = 8 * cs; reset_stream(fd);
        hobbit::stream_descriptor hfd; hfd.name = "wiring_consistency_check_opt"; hfd.size = 8 * cs;
        double ps = 0, hps = 0; vector<F> px;
        srand(11); t0 = now(); vector<F> o = prove_multiplication_tree_stream_shallow(fd, 8, fd.size / 8, F(32), 5, px, 0, vt, ps); tr_ref[1] = now() - t0; int r1 = rand();
        srand(11); t0 = now(); vector<hobbit::Fe> ho = hobbit::prove_multiplication_tree_stream_shallow(hfd, 8, (int)(hfd.size / 8), hobbit::Fe(32), 5, vector<hobbit::Fe>(), 0, vt, hps);
        tr_gpu[1] = now() - t0; int r2 = rand();
        bool ok = o.size() == ho.size() && ps == hps && r1 == r2;
        for (size_t i = 0; ok && i < o.size(); i++) ok = o[i].real == ho[i].real && o[i].img == ho[i].img;
        CHECK(ok, "wiring consistency product tree: 8 products, ps, RNG state");
        F rd = o[0] * o[1] * o[2] * o[7], wr = o[4] * o[5] * o[6] * o[3];
        CHECK(rd == wr, "memory consistency holds: prod(read set) * prod(final) == prod(write set) * prod(init)");
        printf("      ps %f / %f KB\n", ps, hps);
    }
    if (lookups) {          // main.cpp:916: the lookup argument's product tree, 2 vectors
        stream_descriptor fd; fd.name = "lookup_basic"; fd.size = 2 * cs; reset_stream(fd);
        hobbit::stream_descriptor hfd; hfd.name = "lookup_basic"; hfd.size = 2 * cs;
        double ps = 0, hps = 0; vector<F> px;
        srand(12); t0 = now(); vector<F> o = prove_multiplication_tree_stream_shallow(fd, 2, fd.size / 2, F(32), 5, px, 0, vt, ps); tr_ref[1] += now() - t0; int r1 = rand();
        srand(12); t0 = now(); vector<hobbit::Fe> ho = hobbit::prove_multiplication_tree_stream_shallow(hfd, 2, (int)(hfd.size / 2), hobbit::Fe(32), 5, vector<hobbit::Fe>(), 0, vt, hps);
        tr_gpu[1] += now() - t0; int r2 = rand();
        bool ok = o.size() == ho.size() && ps == hps && r1 == r2;
        for (size_t i = 0; ok && i < o.size(); i++) ok = o[i].real == ho[i].real && o[i].img == ho[i].img;
        CHECK(ok, "lookup product tree: 2 products, ps, RNG state");
        printf("      ps %f / %f KB\n", ps, hps);
    }
    // ---- prove_gate_consistency(transcript) ----------------------------------------------------------------------------------------------
    {
        stream_descriptor fd; fd.name = "transcript_stream"; fd.size = cs; reset_stream(fd);
        hobbit::stream_descriptor hfd; hfd.name = "transcript_stream"; hfd.size = cs;
        double ps = 0, hps = 0;
        srand(13); vector<F> r = generate_randomness((int)log2(cs)); t0 = now();
        if (lookups) prove_gate_consistency_lookups(fd, r, vt, ps); else prove_gate_consistency(fd, r, vt, ps);
        tr_ref[2] = now() - t0; int r1 = rand();
        srand(13); vector<hobbit::Fe> hr = hobbit::generate_randomness((int)log2(cs)); t0 = now();
        if (lookups) hobbit::prove_gate_consistency_lookups(hfd, hr, vt, hps); else hobbit::prove_gate_consistency(hfd, hr, vt, hps);
        tr_gpu[2] = now() - t0; int r2 = rand();
        CHECK(ps == hps && r1 == r2, lookups ? "prove_gate_consistency_lookups: ps, RNG state (the prover's own identities hold on both sides)"
                                             : "prove_gate_consistency: ps, RNG state (the prover's own three identities hold on both sides)");
        printf("      ps %f / %f KB\n", ps, hps);
    }
    // ---- open(witness) ---------------------------------------------------------------------------------------------------------------------
    {
        stream_descriptor fd; fd.name = "witness"; fd.size = 4 * cs; reset_stream(fd);
        hobbit::stream_descriptor hfd; hfd.name = "witness"; hfd.size = 4 * cs;
        double ps = 0, hps = 0;
        srand(17); vector<F> x = generate_randomness((int)log2(fd.size)); t0 = now(); open(fd, x, MT, vt, ps); tr_ref[3] = now() - t0; int r1 = rand();
        srand(17); vector<hobbit::Fe> hx = hobbit::generate_randomness((int)log2(hfd.size)); t0 = now(); hobbit::open(hfd, hx, hMT, vt, hps); tr_gpu[3] = now() - t0; int r2 = rand();
        CHECK(ps == hps && r1 == r2, "open(witness): ps, RNG state");
        printf("      ps %f / %f KB\n", ps, hps);
    }
    if (fun == 8) {         // prove_arbitrary_circuit (main.cpp:853): the circuit description is opened against a COPY of the witness tree.
        // Mirror only (see above: the reference aborts while reading "circuit"); the opening's own identities (Error in fft, recursion checks) run.
        vector<vector<hobbit::_hash>> hCH = hMT0;
        hobbit::stream_descriptor hfd; hfd.name = "circuit"; hfd.size = 16 * cs;
        double hps = 0;
        srand(23); vector<hobbit::Fe> hx = hobbit::generate_randomness((int)log2(hfd.size)); t0 = now(); hobbit::open(hfd, hx, hCH, vt, hps); tr_gpu[3] += now() - t0;
        CHECK(hps > 0 && hCH.empty(), "open(circuit) on the mirror: completes, tree freed");
        printf("      ps %f KB\n", hps);
    }
    if (lookups) {          // main.cpp:919: the second open of the process (the global query vector I keeps the first call's positions)
        stream_descriptor fd; fd.name = "lookup_witness_basic"; fd.size = 2 * cs; reset_stream(fd);
        hobbit::stream_descriptor hfd; hfd.name = "lookup_witness_basic"; hfd.size = 2 * cs;
        double ps = 0, hps = 0;
        srand(19); vector<F> x = generate_randomness((int)log2(fd.size)); t0 = now(); open(fd, x, MTl, vt, ps); tr_ref[3] += now() - t0; int r1 = rand();
        srand(19); vector<hobbit::Fe> hx = hobbit::generate_randomness((int)log2(hfd.size)); t0 = now(); hobbit::open(hfd, hx, hMTl, vt, hps); tr_gpu[3] += now() - t0; int r2 = rand();
        CHECK(ps == hps && r1 == r2, "open(lookup_witness_basic): ps, RNG state");
        printf("      ps %f / %f KB\n", ps, hps);
    }
    printf("{\"workload\": \"%s circuit 2^%d gates, BUFFER_SPACE 2^%d\", \"trace_upload_s\": %.4f, \"ref_s\": {\"commit\": %.4f, \"mul_tree\": %.4f, \"gate\": %.4f, \"open\": %.4f}, "
           "\"gpu_s\": {\"commit\": %.4f, \"mul_tree\": %.4f, \"gate\": %.4f, \"open\": %.4f}, \"identical\": %s}\n",
           fun == 9 ? "MLP" : fun == 8 ? "pruned MLP" : fun == 5 ? "AES" : "SQL", (int)log2(cs), (int)log2(B), t_trace, tr_ref[0], tr_ref[1], tr_ref[2], tr_ref[3], tr_gpu[0], tr_gpu[1], tr_gpu[2], tr_gpu[3], failures ? "false" : "true");
    printf(failures ? "CIRC: %d FAILURES\n" : "CIRC: all identical\n", failures);
    fflush(stdout);
    _exit(failures ? 1 : 0);                                          // the producer thread is still blocked on its mutex
}
