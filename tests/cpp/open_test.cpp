// Opening-recursion drop-in test: the UNMODIFIED reference (global namespace, oracle/_ref/libhobbit_ref.so) and the
// host mirror (namespace hobbit, hobbit_b200/host/hobbit_open.cpp) run the same calls with the same libc RNG state; every proof field,
// every Merkle level, the proof-size counter and the RNG state afterwards must agree.
// Built twice by oracle/Makefile: `open_test` against libhobbit_b200.so (needs a GPU) and `open_test_emul` against the CPU emulation
// of the C ABI (oracle/hb_emul.cpp) so that the host-side logic is covered without a GPU.  Run by tests/test_open_cpp.py.
#include "../../hobbit_b200/host/hobbit_host.hpp"
namespace hobbit { typedef F Fe; }          // the reference #defines F as a macro (config_pc.hpp:10)
#include "config_pc.hpp"
#include "utils.hpp"
#include "mimc.h"
#include "Our_PC.hpp"
#include "witness_stream.h"
#include "Elastic_PC.hpp"
#include "sumcheck.h"
#include "PC_utils.h"
#include "Virgo.h"
#include <cstdio>
#include <cstring>

extern bool linear_time;
extern int tensor_row_size;
extern size_t BUFFER_SPACE;
extern bool __encode_initialized;
extern shockwave_data *C_f, *C_c;

static int failures = 0;
#define CHECK(cond, what) do { if (!(cond)) { printf("FAIL: %s (%s:%d)\n", what, __FILE__, __LINE__); failures++; } else printf("ok:   %s\n", what); fflush(stdout); } while (0)

static bool eqF(const F &a, const hobbit::Fe &b) { return a.real == b.real && a.img == b.img; }
static vector<hobbit::Fe> conv(const vector<F> &v) { vector<hobbit::Fe> o(v.size()); memcpy(o.data(), v.data(), v.size() * 16); return o; }
static vector<vector<hobbit::Fe>> conv2(const vector<vector<F>> &v) { vector<vector<hobbit::Fe>> o; for (auto &r : v) o.push_back(conv(r)); return o; }
static bool same_levels(const vector<vector<_hash>> &a, const vector<vector<hobbit::_hash>> &b) {
    if (a.size() != b.size()) return false;
    for (size_t l = 0; l < a.size(); l++) {
        if (a[l].size() != b[l].size()) return false;
        if (memcmp(a[l].data(), b[l].data(), a[l].size() * 32)) return false;
    }
    return true;
}
static bool same_proof(const proof &P, const hobbit::proof &H) {
    if (P.q_poly.size() != H.q_poly.size() || P.randomness.size() != H.randomness.size() || P.vr.size() != H.vr.size()) return false;
    for (size_t i = 0; i < P.q_poly.size(); i++)
        if (!(eqF(P.q_poly[i].a, H.q_poly[i].a) && eqF(P.q_poly[i].b, H.q_poly[i].b) && eqF(P.q_poly[i].c, H.q_poly[i].c))) return false;
    for (size_t k = 0; k < P.randomness.size(); k++) {
        if (P.randomness[k].size() != H.randomness[k].size()) return false;
        for (size_t i = 0; i < P.randomness[k].size(); i++) if (!eqF(P.randomness[k][i], H.randomness[k][i])) return false;
    }
    for (size_t i = 0; i < P.vr.size(); i++) if (!eqF(P.vr[i], H.vr[i])) return false;
    return eqF(P.final_rand, H.final_rand);
}
// full-width pseudo-random field elements (products of small randoms)
static vector<F> rand_vec(size_t n) {
    vector<F> a = generate_randomness(n), b = generate_randomness(n);
    for (size_t i = 0; i < n; i++) a[i] = a[i] * b[(i * 7 + 3) % n] + F(3, (long long)i) * b[i];
    return a;
}

int main(int argc, char **argv) {
    const bool big = argc > 1 && !strcmp(argv[1], "big");
    init_hash();
    hobbit::init_backend(0);
    double vt = 0;

    if (argc > 1 && !strcmp(argv[1], "deep")) {
        // Runs in its OWN process: open_layers opens with BUFFER_SPACE 2^9, and the reference's global query vector `I` (Elastic_PC.cpp:314)
        // would still hold positions of earlier, larger openings (its own "Error pos,size" exit).
    // ---- deep streaming product tree (layers > distance): batched layers + commit_layers / open_layers (sumcheck.cpp:983-1011, 1871-1911) ----
        for (int cfg = 0; cfg < 2; cfg++) {
            // the reference's top-layer read (read_mul_tree_layer, witness_stream.cpp:2415-2459) needs total >> layers >= 2^layers, else it spins
            BUFFER_SPACE = 1 << 9; hobbit::BUFFER_SPACE = BUFFER_SPACE;
            extern int BUFFER_SPACE_tr; BUFFER_SPACE_tr = BUFFER_SPACE / 8;
            const size_t total = 1 << 20; const int vectors = cfg ? 2 : 8;                      // layers = 10 -> 2 batches of 5
            double ps = 0, hps = 0;
            stream_descriptor fd; fd.name = "test"; fd.size = total; reset_stream(fd);
            srand(33); vector<F> o = prove_multiplication_tree_stream_shallow(fd, vectors, total / vectors, F(32), 5, vector<F>(), 0, vt, ps);
            int r1 = rand();
            hobbit::stream_descriptor hfd; hfd.name = "test"; hfd.size = total;
            srand(33); vector<hobbit::Fe> ho = hobbit::prove_multiplication_tree_stream_shallow(hfd, vectors, total / vectors, hobbit::Fe(32), 5, vector<hobbit::Fe>(), 0, vt, hps);
            int r2 = rand();
            bool ok = o.size() == ho.size() && ps == hps && r1 == r2;
            for (size_t i = 0; ok && i < o.size(); i++) ok = eqF(o[i], ho[i]);
            CHECK(ok, cfg ? "deep product tree (2 x 2^19, BUFFER_SPACE 2^9: 2 batched layers, committed layer): products, ps, RNG state"
                          : "deep product tree (8 x 2^17, BUFFER_SPACE 2^9: 2 batched layers, committed layer): products, ps, RNG state");
            printf("      ps %f / %f KB\n", ps, hps);
        }
        printf(failures ? "OPEN: %d FAILURES\n" : "OPEN: all identical\n", failures);
        return failures ? 1 : 0;
    }

    // ---- shockwave_commit ---------------------------------------------------------------------------------------------------------
    {
        srand(41);
        vector<F> poly = rand_vec(1 << 13);
        for (int i = 0; i < 256; i++) poly[3 * 256 + i] = F(0);       // one all-zero row (the reference skips its _fft)
        vector<hobbit::Fe> hp = conv(poly);
        shockwave_data *d = shockwave_commit(poly, 32);
        hobbit::shockwave_data *h = hobbit::shockwave_commit(hp, 32);
        vector<hobbit::Fe> enc = h->encoded_host();
        bool ok = true;
        for (int i = 0; ok && i < 32; i++) ok = memcmp(d->encoded_matrix[i], enc.data() + (size_t)i * 512, 512 * 16) == 0;
        CHECK(ok, "shockwave_commit: encoded_matrix");
        CHECK(same_levels(d->MT, h->MT_host()), "shockwave_commit: every Merkle level");
        // ---- shockwave_prove on it (N/k = 256: no WHIR) ----
        srand(43);
        vector<F> x = generate_randomness(13);
        double ps = 0, hps = 0;
        srand(44); shockwave_prove(d, x, vt, ps); int r1 = rand();
        srand(44); hobbit::shockwave_prove(h, conv(x), vt, hps); int r2 = rand();
        CHECK(ps == hps && r1 == r2, "shockwave_prove (no WHIR): ps and RNG state");
        printf("      ps %f / %f\n", ps, hps);
    }
    // ---- prove_fft / prove_fft_matrix ----------------------------------------------------------------------------------------------
    {
        srand(51);
        vector<F> m = rand_vec(1 << 10), r = rand_vec(11);
        vector<hobbit::Fe> hm = conv(m);
        double ps = 0, hps = 0;
        // previous_sum = MLE of the zero-extended transform at r
        vector<F> t = m; t.resize(2048, F(0)); _fft(t.data(), 11, false);
        F y = evaluate_vector(t, r);
        proof P = prove_fft(m, r, y, vt, ps);
        hobbit::proof H = hobbit::prove_fft(hm, conv(r), hobbit::Fe(y.real, y.img), vt, hps);
        CHECK(same_proof(P, H) && ps == hps && m.size() == hm.size(), "prove_fft: q_poly, randomness, vr, final_rand, ps");
        CHECK(P.q_poly[0].eval(0) + P.q_poly[0].eval(1) == y, "prove_fft: claimed sum is the transform's MLE");

        const size_t rows = 16, cols = 128;
        vector<vector<F>> M(rows), Mp(rows);
        vector<F> flat;
        for (size_t i = 0; i < rows; i++) {
            M[i] = rand_vec(cols); Mp[i] = M[i]; Mp[i].resize(2 * cols, F(0)); _fft(Mp[i].data(), 8, false);
            flat.insert(flat.end(), Mp[i].begin(), Mp[i].end());
        }
        vector<F> rr = rand_vec(8 + 4 + 1);
        F y1 = evaluate_vector(flat, rr);
        ps = hps = 0;
        P = prove_fft_matrix(M, rr, y1, vt, ps);
        H = hobbit::prove_fft_matrix(conv2(M), conv(rr), hobbit::Fe(y1.real, y1.img), vt, hps);
        CHECK(same_proof(P, H) && ps == hps, "prove_fft_matrix: q_poly, randomness (+r1), vr, final_rand, ps");
    }
    // ---- prove_linear_code ----------------------------------------------------------------------------------------------------------
    for (int n : {16, 128}) {
        srand(61); __encode_initialized = false; expander_init_store(n);
        srand(61); hobbit::expander_init_store(n);
        srand(62);
        vector<F> msg = rand_vec(n), cw(2 * n, F(0));
        encode_monolithic(msg.data(), cw.data(), n);
        vector<hobbit::Fe> hcw = conv(cw);
        double ps = 0, hps = 0;
        srand(63); proof P = prove_linear_code(cw, n, vt, ps); int r1 = rand();
        srand(63); hobbit::proof H = hobbit::prove_linear_code(hcw, n, vt, hps); int r2 = rand();
        CHECK(same_proof(P, H) && ps == hps && r1 == r2, n == 16 ? "prove_linear_code n=16: proof, ps, RNG" : "prove_linear_code n=128: proof, ps, RNG");
        CHECK(P.q_poly[0].eval(0) + P.q_poly[0].eval(1) == F(0), "prove_linear_code: the parity check vanishes on a codeword");
    }
    // ---- whir_commit / _whir_prove -----------------------------------------------------------------------------------------------------
    for (int logn : {10, 13}) {
        srand(71);
        vector<F> poly = rand_vec((size_t)1 << logn), x = rand_vec(logn);
        vector<hobbit::Fe> hp = conv(poly);
        Whir_data D; hobbit::Whir_data H;
        whir_commit(poly, D); hobbit::whir_commit(hp, H);
        CHECK(same_levels(D.MT, H.MT_host()), "whir_commit: every Merkle level");
        double ps = 0, hps = 0;
        srand(72); _whir_prove(D, x, vt, ps); int r1 = rand();
        srand(72); hobbit::_whir_prove(H, conv(x), vt, hps); int r2 = rand();
        bool ok = ps == hps && r1 == r2 && D.FRI_MT.size() == H.FRI_MT.size();
        for (size_t i = 0; ok && i < D.FRI_MT.size(); i++) if (D.FRI_MT[i].size()) ok = H.FRI_MT[i] && same_levels(D.FRI_MT[i], H.FRI_MT_host((int)i));
        vector<hobbit::Fe> fp = H.poly_host();
        size_t rem = (size_t)1 << (logn - 4 * ((logn - 1) / 4));
        for (size_t i = 0; ok && i < rem; i++) ok = eqF(D.poly[i], fp[i]);
        CHECK(ok, logn == 10 ? "_whir_prove 2^10: FRI trees, folded polynomial, ps, RNG" : "_whir_prove 2^13: FRI trees, folded polynomial, ps, RNG");
        printf("      ps %f / %f\n", ps, hps);
    }
    // ---- shockwave_prove with WHIR -------------------------------------------------------------------------------------------------------
    {
        srand(81);
        vector<F> poly = rand_vec(1 << 15), x = rand_vec(15);
        vector<hobbit::Fe> hp = conv(poly);
        shockwave_data *d = shockwave_commit(poly, 32);
        hobbit::shockwave_data *h = hobbit::shockwave_commit(hp, 32);
        double ps = 0, hps = 0;
        srand(82); shockwave_prove(d, x, vt, ps); int r1 = rand();
        srand(82); hobbit::shockwave_prove(h, conv(x), vt, hps); int r2 = rand();
        CHECK(ps == hps && r1 == r2, "shockwave_prove 2^15 (WHIR on 2^10): ps and RNG state");
        printf("      ps %f / %f\n", ps, hps);
    }
    // ---- recursive_prover_RS ------------------------------------------------------------------------------------------------------------
    {
        const size_t B = 1 << 13;
        tensor_row_size = 16; hobbit::tensor_row_size = 16; linear_time = false; hobbit::linear_time = false;
        srand(91);
        vector<F> agg = rand_vec(B);
        vector<hobbit::Fe> hagg = conv(agg);
        vector<vector<size_t>> I(790);
        for (auto &q : I) { q.push_back(rand() % (2 * B / tensor_row_size)); q.push_back(rand() % (2 * tensor_row_size)); }
        vector<F> buff = agg; C_f = shockwave_commit(buff, 32);
        hobbit::C_f = hobbit::shockwave_commit(hagg, 32);
        double ps = 0, hps = 0;
        srand(92); recursive_prover_RS(agg, I, vt, ps); int r1 = rand();
        srand(92); hobbit::recursive_prover_RS(hagg, I, vt, hps); int r2 = rand();
        CHECK(ps == hps && r1 == r2, "recursive_prover_RS (B = 2^13, trs = 16, 790 queries): ps and RNG state");
        printf("      ps %f / %f\n", ps, hps);
    }
    // ---- recursive_prover_Spielman --------------------------------------------------------------------------------------------------------
    {
        const size_t B = 1 << 13;
        tensor_row_size = 16; hobbit::tensor_row_size = 16; linear_time = true; hobbit::linear_time = true;
        srand(101); __encode_initialized = false; expander_init_store(16);
        srand(101); hobbit::expander_init_store(16);
        srand(102);
        vector<F> agg = rand_vec(B);
        vector<hobbit::Fe> hagg = conv(agg);
        vector<vector<F>> T; compute_tensorcode(agg, T);
        T.erase(T.begin(), T.begin() + T.size() / 2);
        vector<size_t> I(5900);
        for (auto &q : I) { size_t c = rand() % (2 * B / tensor_row_size), r = rand() % (2 * tensor_row_size); q = c + (2 * B / tensor_row_size) * r; }
        vector<F> buff = agg; C_f = shockwave_commit(buff, 32);
        buff = convert2vector(T); C_c = shockwave_commit(buff, 32);
        hobbit::C_f = hobbit::shockwave_commit(hagg, 32);
        vector<hobbit::Fe> hbuff = conv(buff); hobbit::C_c = hobbit::shockwave_commit(hbuff, 32);
        vector<vector<hobbit::Fe>> hT = conv2(T);
        double ps = 0, hps = 0;
        srand(103); recursive_prover_Spielman(agg, T, I, vt, ps); int r1 = rand();
        srand(103); hobbit::recursive_prover_Spielman(hagg, hT, I, vt, hps); int r2 = rand();
        CHECK(ps == hps && r1 == r2, "recursive_prover_Spielman (B = 2^13, trs = 16, 5900 queries): ps and RNG state");
        printf("      ps %f / %f\n", ps, hps);
    }
    // ---- commit_standard + open_standard, the test_PC(N, 4, K) flow (Our_PC.cpp:806-826) ---------------------------------------------------
    for (int lin = 1; lin >= 0; lin--) {
        const size_t N = big ? 1 << 20 : 1 << 18; const int K = big ? 32 : 8;
        srand(1);
        vector<F> poly = generate_randomness(N);
        linear_time = lin; tensor_row_size = N / (K * 1ULL << 11);
        if (lin) { __encode_initialized = false; expander_init_store(tensor_row_size); }
        _hash comm; vector<vector<_hash>> MT; vector<vector<vector<F>>> T;
        commit_standard(poly, comm, MT, T, K);
        double ps = 0, hps = 0;
        vector<F> x = generate_randomness((int)log2(poly.size()));
        open_standard(poly, x, MT, T, K, vt, ps);
        int r1 = rand();

        srand(1);
        vector<hobbit::Fe> hpoly = hobbit::generate_randomness(N);
        hobbit::linear_time = lin; hobbit::tensor_row_size = N / (K * 1ULL << 11);
        if (lin) hobbit::expander_init_store(hobbit::tensor_row_size);
        hobbit::_hash hcomm; vector<vector<hobbit::_hash>> hMT; vector<vector<vector<hobbit::Fe>>> hT;
        hobbit::materialize_tensor = false;
        hobbit::commit_standard(hpoly, hcomm, hMT, hT, K);
        vector<hobbit::Fe> hx = hobbit::generate_randomness((int)log2(hpoly.size()));
        hobbit::open_standard(hpoly, hx, hMT, hT, K, vt, hps);
        int r2 = rand();
        CHECK(same_levels(MT, hMT), lin ? "test_PC flow (Orion columns): commit_standard levels" : "test_PC flow (RS columns): commit_standard levels");
        CHECK(ps == hps && r1 == r2, lin ? "test_PC flow (Orion columns): open_standard ps and RNG state" : "test_PC flow (RS columns): open_standard ps and RNG state");
        printf("      ps %f / %f KB\n", ps, hps);
        if (big && lin) CHECK(hps == 3839.078125, "test_PC(2^20, 4, 32): ps == 3839.078125 KB (SURVEY §9 KAT)");
    }
    // ---- Elastic_PC commit + open on the synthetic stream (test_Elastic_PC flow, Elastic_PC.cpp:736-785), both column codes ------------------
    for (int lin = 0; lin < 2; lin++) {
        const size_t N = 1 << 17; BUFFER_SPACE = 1 << 14; hobbit::BUFFER_SPACE = BUFFER_SPACE;
        srand(3);
        init_commitment(lin); tensor_row_size = 16;
        if (lin) { __encode_initialized = false; expander_init_store(tensor_row_size); }
        stream_descriptor fd; fd.name = "test"; fd.size = N;
        _hash comm; vector<vector<_hash>> MT; commit(fd, comm, MT);
        vector<vector<_hash>> MT0 = MT;
        double ps = 0, hps = 0;
        vector<F> x = generate_randomness(17);
        open(fd, x, MT, vt, ps);
        int r1 = rand();
        srand(3);
        hobbit::init_commitment(lin); hobbit::tensor_row_size = 16;
        if (lin) hobbit::expander_init_store(hobbit::tensor_row_size);
        hobbit::stream_descriptor hfd; hfd.name = "test"; hfd.size = N;
        hobbit::_hash hcomm; vector<vector<hobbit::_hash>> hMT; hobbit::commit(hfd, hcomm, hMT);
        vector<vector<hobbit::_hash>> hMT0 = hMT;
        vector<hobbit::Fe> hx = hobbit::generate_randomness(17);
        hobbit::open(hfd, hx, hMT, vt, hps);
        int r2 = rand();
        MT0[0].back() = _hash(); memset(&hMT0[0].back(), 0, 32);          // the reference's last leaf reads past its buffers
        CHECK(same_levels(MT0, hMT0), lin ? "Elastic commit (Orion columns): every level" : "Elastic commit (RS columns): every level");
        CHECK(ps == hps && r1 == r2 && MT.empty() && hMT.empty(), lin ? "Elastic open (Orion columns, Spielman_stream recursion): ps, RNG state, tree freed"
                                                                      : "Elastic open (RS columns): ps, RNG state, tree freed");
        printf("      ps %f / %f KB\n", ps, hps);
    }
    printf(failures ? "OPEN: %d FAILURES\n" : "OPEN: all identical\n", failures);
    return failures ? 1 : 0;
}
