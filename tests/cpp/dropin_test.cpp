// Drop-in test: links the UNMODIFIED reference (global namespace, oracle/_ref/libhobbit_ref.so) next to the host mirror
// (namespace hobbit, libhobbit_host.so -> libhobbit_b200.so) and runs the same calls through both, with the libc RNG
// reset to the same state before each side.  Needs a GPU.  Built by oracle/Makefile (it needs the reference headers),
// run by tests/test_dropin_cpp.py.
#include "../../hobbit_b200/host/hobbit_host.hpp"
namespace hobbit { typedef F Fe; }          // the reference #defines F as a macro (config_pc.hpp:10)
#include "config_pc.hpp"
#include "utils.hpp"
#include "mimc.h"
#include "Our_PC.hpp"
#include "witness_stream.h"
#include "Elastic_PC.hpp"
#include "sumcheck.h"
#include <cstdio>
#include <cstring>

extern bool linear_time;
extern int tensor_row_size;
extern size_t BUFFER_SPACE;
extern bool __encode_initialized;
proof batch_3product_sumcheck(vector<vector<F>> &arr1, vector<vector<F>> &arr2, vector<vector<F>> &arr3, vector<F> a, double &vt, double &ps);

static int failures = 0;
#define CHECK(cond, what) do { if (!(cond)) { printf("FAIL: %s (%s:%d)\n", what, __FILE__, __LINE__); failures++; } else printf("ok:   %s\n", what); } while (0)

static bool eqF(const F &a, const hobbit::Fe &b) { return a.real == b.real && a.img == b.img; }
static vector<hobbit::Fe> conv(const vector<F> &v) { vector<hobbit::Fe> o(v.size()); memcpy(o.data(), v.data(), v.size() * 16); return o; }
static bool same_levels(const vector<vector<_hash>> &a, const vector<vector<hobbit::_hash>> &b) {
    if (a.size() != b.size()) return false;
    for (size_t l = 0; l < a.size(); l++) {
        if (a[l].size() != b[l].size()) return false;
        if (memcmp(a[l].data(), b[l].data(), a[l].size() * 32)) return false;
    }
    return true;
}

int main() {
    init_hash();
    hobbit::init_backend(0);

    // ---- test_PC(2^16, option 4, K=8) commit, then the front half of open_standard --------------------------------
    {
        const size_t N = 1 << 16; const int K = 8;
        srand(1);
        vector<F> poly = generate_randomness(N);
        linear_time = true; tensor_row_size = N / (K * 1ULL << 11);
        __encode_initialized = false;
        expander_init_store(tensor_row_size);
        _hash comm; vector<vector<_hash>> MT; vector<vector<vector<F>>> T;
        commit_standard(poly, comm, MT, T, K);
        int r_ref = rand();

        srand(1);
        vector<hobbit::Fe> hpoly = hobbit::generate_randomness(N);
        hobbit::linear_time = true; hobbit::tensor_row_size = N / (K * 1ULL << 11);
        hobbit::expander_init_store(hobbit::tensor_row_size);
        hobbit::_hash hcomm; vector<vector<hobbit::_hash>> hMT; vector<vector<vector<hobbit::Fe>>> hT;
        hobbit::materialize_tensor = true;
        hobbit::commit_standard(hpoly, hcomm, hMT, hT, K);
        int r_h = rand();
        CHECK(memcmp(poly.data(), hpoly.data(), N * 16) == 0, "generate_randomness identical");
        CHECK(r_ref == r_h, "libc RNG state identical after expander_init_store + commit_standard");
        CHECK(same_levels(MT, hMT), "commit_standard: every Merkle level identical");
        bool tsame = T.size() == hT.size();
        for (size_t i = 0; tsame && i < T.size(); i++) for (size_t r = 0; tsame && r < T[i].size(); r++)
            tsame = T[i][r].size() == hT[i][r].size() && memcmp(T[i][r].data(), hT[i][r].data(), T[i][r].size() * 16) == 0;
        CHECK(tsame, "commit_standard: _tensor identical");

        // open front: same x, same RNG state -> same queries, replies, paths, aggregate
        srand(5);
        vector<F> x = generate_randomness(16);
        srand(9);
        hobbit::open_front o = hobbit::open_standard_front(hpoly, conv(x), hMT, K);
        srand(9);
        vector<F> x1(x.begin(), x.begin() + 3), beta; precompute_beta(x1, beta);
        F rv0 = generate_randomness(1)[0];
        CHECK(eqF(rv0, o.r_v0), "open_standard: r_v[0] drawn at the same RNG position");
        bool ok = beta.size() == o.beta.size();
        for (size_t i = 0; ok && i < beta.size(); i++) ok = eqF(beta[i], o.beta[i]);
        CHECK(ok, "open_standard: beta = eq(x1)");
        size_t B = N / K; vector<F> agg(B, F(0));
        for (int i = 0; i < K; i++) for (size_t j = 0; j < B; j++) agg[j] += beta[i] * poly[i * B + j];
        ok = true; for (size_t j = 0; ok && j < B; j++) ok = eqF(agg[j], o.aggr_vector[j]);
        CHECK(ok, "open_standard: aggregated vector (Our_PC.cpp:265-272)");
        ok = true;
        for (int q = 0; q < 5900 && ok; q++) {
            size_t c0 = rand() % (2 * B / tensor_row_size), c1 = rand() % (2 * tensor_row_size);
            ok = o.I[q][0] == c0 && o.I[q][1] == c1;
            for (int i = 0; ok && i < K; i++) ok = eqF(T[i][c1][c0], o.reply[q][i]);
            vector<size_t> c = {c0, c1};
            vector<_hash> path = merkle_tree::merkle_tree_prover::open_tree_blake(MT, c, 2 * B / tensor_row_size);
            ok = ok && path.size() == o.commitment_paths[q].size() && memcmp(path.data(), o.commitment_paths[q].data(), path.size() * 32) == 0;
        }
        CHECK(ok, "open_standard: 5900 queries, replies and Merkle paths");
    }
    // ---- Elastic_PC commit, both codes ---------------------------------------------------------------------------------
    for (int lin = 0; lin < 2; lin++) {
        const size_t N = 1 << 16; BUFFER_SPACE = 1 << 12; hobbit::BUFFER_SPACE = BUFFER_SPACE;
        srand(3);
        init_commitment(lin);
        if (lin) { __encode_initialized = false; expander_init_store(tensor_row_size); }
        stream_descriptor fd; fd.name = "test"; fd.size = N;
        _hash comm; vector<vector<_hash>> MT; commit(fd, comm, MT);
        srand(3);
        hobbit::init_commitment(lin);
        if (lin) hobbit::expander_init_store(hobbit::tensor_row_size);
        hobbit::stream_descriptor hfd; hfd.name = "test"; hfd.size = N;
        hobbit::_hash hcomm; vector<vector<hobbit::_hash>> hMT; hobbit::commit(hfd, hcomm, hMT);
        // the reference's LAST leaf hashes one element past its parked buffers (see oracle/hobbit_oracle.c): excluded
        MT[0].back() = _hash(); memset(&hMT[0].back(), 0, 32);
        CHECK(same_levels(MT, hMT), lin ? "Elastic commit (Spielman columns): every level" : "Elastic commit (RS columns): every level");
    }
    // ---- sumchecks -------------------------------------------------------------------------------------------------------
    {
        const size_t n = 1 << 12;
        srand(11);
        vector<F> v1 = generate_randomness(n), v2 = generate_randomness(n), v3 = generate_randomness(n);
        for (size_t i = 0; i < n; i++) { v1[i] = v1[i] * v2[(i * 7) % n] + F(3, i); v3[i] = v3[i] * v1[i]; }
        vector<hobbit::Fe> h1 = conv(v1), h2 = conv(v2), h3 = conv(v3);
        double vt = 0, ps = 0, hps = 0;
        proof P = generate_2product_sumcheck_proof(v1, v2, F(9), vt, ps);
        hobbit::proof H = hobbit::generate_2product_sumcheck_proof(h1, h2, hobbit::Fe(9), vt, hps);
        bool ok = P.q_poly.size() == H.q_poly.size() && ps == hps && eqF(P.final_rand, H.final_rand) && eqF(P.vr[0], H.vr[0]) && eqF(P.vr[1], H.vr[1]);
        for (size_t i = 0; ok && i < P.q_poly.size(); i++)
            ok = eqF(P.q_poly[i].a, H.q_poly[i].a) && eqF(P.q_poly[i].b, H.q_poly[i].b) && eqF(P.q_poly[i].c, H.q_poly[i].c) && eqF(P.randomness[0][i], H.randomness[0][i]);
        CHECK(ok, "generate_2product_sumcheck_proof: q_poly, randomness, vr, final_rand, ps");

        ps = hps = 0;
        vector<F> a1 = v1, a2 = v2, a3 = v3;
        P = _generate_3product_sumcheck_proof(a1, a2, a3, F(9), vt, ps);
        H = hobbit::_generate_3product_sumcheck_proof(h1, h2, h3, hobbit::Fe(9), vt, hps);
        ok = P.c_poly.size() == H.c_poly.size() && ps == hps && eqF(P.final_rand, H.final_rand);
        for (int i = 0; ok && i < 3; i++) ok = eqF(P.vr[i], H.vr[i]);
        for (size_t i = 0; ok && i < P.c_poly.size(); i++)
            ok = eqF(P.c_poly[i].a, H.c_poly[i].a) && eqF(P.c_poly[i].b, H.c_poly[i].b) && eqF(P.c_poly[i].c, H.c_poly[i].c) && eqF(P.c_poly[i].d, H.c_poly[i].d) && eqF(P.randomness[0][i], H.randomness[0][i]);
        CHECK(ok, "_generate_3product_sumcheck_proof: c_poly, randomness, vr, final_rand, ps");

        vector<vector<F>> A = {v1, vector<F>(v2.begin(), v2.begin() + 256), vector<F>(v3.begin(), v3.begin() + 4)}, Bv = {v2, vector<F>(v3.begin(), v3.begin() + 256), vector<F>(v1.begin(), v1.begin() + 4)},
                          C = {v3, vector<F>(v1.begin(), v1.begin() + 256), vector<F>(v2.begin(), v2.begin() + 4)};
        vector<vector<hobbit::Fe>> hA, hB, hC;
        for (int i = 0; i < 3; i++) { hA.push_back(conv(A[i])); hB.push_back(conv(Bv[i])); hC.push_back(conv(C[i])); }
        vector<F> a = {F(17, 4), F(23, 6), F(5, 5)};
        ps = hps = 0;
        P = batch_3product_sumcheck(A, Bv, C, a, vt, ps);
        H = hobbit::batch_3product_sumcheck(hA, hB, hC, conv(a), vt, hps);
        ok = P.c_poly.size() == H.c_poly.size() && P.vr.size() == H.vr.size() && ps == hps;
        for (size_t i = 0; ok && i < P.vr.size(); i++) ok = eqF(P.vr[i], H.vr[i]);
        for (size_t i = 0; ok && i < P.c_poly.size(); i++) ok = eqF(P.c_poly[i].a, H.c_poly[i].a) && eqF(P.c_poly[i].d, H.c_poly[i].d) && eqF(P.randomness[0][i], H.randomness[0][i]);
        CHECK(ok, "batch_3product_sumcheck: c_poly, randomness, vr, ps");

        vector<vector<F>> in(8); vector<vector<hobbit::Fe>> hin(8);
        for (int i = 0; i < 8; i++) { in[i].assign(v1.begin() + i * 512, v1.begin() + (i + 1) * 512); hin[i] = conv(in[i]); }
        ps = hps = 0;
        srand(21); mul_tree_proof M = prove_multiplication_tree_new(in, F(32), vector<F>(), vt, ps);
        srand(21); hobbit::mul_tree_proof HM = hobbit::prove_multiplication_tree_new(hin, hobbit::Fe(32), vector<hobbit::Fe>(), vt, hps);
        ok = ps == hps && eqF(M.out_eval, HM.out_eval) && eqF(M.final_eval, HM.final_eval) && M.final_r.size() == HM.final_r.size() && M.proofs.size() == HM.proofs.size();
        for (size_t i = 0; ok && i < M.final_r.size(); i++) ok = eqF(M.final_r[i], HM.final_r[i]);
        for (size_t i = 0; ok && i < M.output.size(); i++) ok = eqF(M.output[i], HM.output[i]);
        for (size_t l = 0; ok && l < M.proofs.size(); l++) {
            ok = M.proofs[l].c_poly.size() == HM.proofs[l].c_poly.size() && eqF(M.proofs[l].final_rand, HM.proofs[l].final_rand);
            for (size_t i = 0; ok && i < M.proofs[l].c_poly.size(); i++) ok = eqF(M.proofs[l].c_poly[i].a, HM.proofs[l].c_poly[i].a) && eqF(M.proofs[l].c_poly[i].c, HM.proofs[l].c_poly[i].c);
        }
        for (size_t i = 0; ok && i < M.individual_randomness.size(); i++) ok = eqF(M.individual_randomness[i], HM.individual_randomness[i]);
        for (size_t i = 0; ok && i < M.global_randomness.size(); i++) ok = eqF(M.global_randomness[i], HM.global_randomness[i]);
        CHECK(ok, "prove_multiplication_tree_new (8 x 512): output, out_eval, layer proofs, final_r, final_eval, ps");
    }
    // ---- streaming product tree on the synthetic stream (MLP-shaped: 8 vectors, 4 streamed layers) ----------------------------
    {
        BUFFER_SPACE = 1 << 10; hobbit::BUFFER_SPACE = BUFFER_SPACE;
        extern int BUFFER_SPACE_tr; BUFFER_SPACE_tr = BUFFER_SPACE / 8;
        const size_t total = 1 << 15;
        double vt = 0, ps = 0, hps = 0;
        stream_descriptor fd; fd.name = "test"; fd.size = total; reset_stream(fd);
        srand(31); vector<F> o = prove_multiplication_tree_stream_shallow(fd, 8, total / 8, F(32), 5, vector<F>(), 0, vt, ps);
        int r1 = rand();
        hobbit::stream_descriptor hfd; hfd.name = "test"; hfd.size = total;
        srand(31); vector<hobbit::Fe> ho = hobbit::prove_multiplication_tree_stream_shallow(hfd, 8, total / 8, hobbit::Fe(32), 5, vector<hobbit::Fe>(), 0, vt, hps);
        int r2 = rand();
        bool ok = o.size() == ho.size() && ps == hps && r1 == r2;
        for (size_t i = 0; ok && i < o.size(); i++) ok = eqF(o[i], ho[i]);
        CHECK(ok, "prove_multiplication_tree_stream_shallow (8 x 4096, 4 streamed layers): products, ps, RNG state");
    }
    printf(failures ? "DROPIN: %d FAILURES\n" : "DROPIN: all identical\n", failures);
    return failures ? 1 : 0;
}
