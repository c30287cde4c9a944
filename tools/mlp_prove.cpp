// Reference-free end-to-end MLP proof on the GPU (BASELINE config 3 without the CPU evaluator): the circuit is evaluated on the GPU
// (hobbit::trace_generate_mlp), every stream is derived there, then the prove_circuit sequence (main.cpp:862-886):
// commit(witness) -> prove_multiplication_tree_stream_shallow(wiring, 8) -> prove_gate_consistency -> open(witness).
// `mlp_prove <b> aes <n>` / `mlp_prove <b> sql <n>` are the same for the lookup circuits (`pigeon 5 b n 1`, 2^n AES blocks / `pigeon 6 b n 1`,
// 2^n database rows; hobbit::trace_generate_aes / trace_generate_sql and the
// fun == 5 sequence main.cpp:887-917: commit(witness), lookup_rand, commit(lookup_witness_basic), the wiring and lookup product trees,
// prove_gate_consistency_lookups, both opens).
// Links only libhobbit_host.so / libhobbit_b200.so.   usage: mlp_prove <log2 BUFFER_SPACE> <layer sizes...> [--reps R] | mlp_prove <b> aes <n> [--reps R]
#include "../hobbit_b200/host/hobbit_host.hpp"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
using namespace hobbit;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    int b = argc > 1 ? atoi(argv[1]) : 18, reps = 2;
    std::vector<int> layers;
    int aes_n = -1; bool sql = false;                        // aes_n: log2 of the blocks (AES) or rows (SQL) of a lookup circuit
    for (int i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "--reps")) { reps = atoi(argv[++i]); continue; }
        if (!strcmp(argv[i], "--resident-levels")) { commit_levels_on_host = false; continue; }      // prover-only run: big Merkle levels stay in HBM
        if (!strcmp(argv[i], "--async-levels")) { commit_levels_async = true; continue; }            // levels reach the host in the background, under the provers
        if (!strcmp(argv[i], "aes") || !strcmp(argv[i], "sql")) { sql = !strcmp(argv[i], "sql"); aes_n = atoi(argv[++i]); continue; }
        layers.push_back(atoi(argv[i]));
    }
    if (layers.empty()) layers = {1024, 256, 256, 16};
    // one process per GPU: RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT from the environment (tools/run_ranks.sh, torchrun);
    // every rank evaluates the circuit (a millisecond on the GPU) and proves its share of every sumcheck and commitment
    dist_init_from_env((size_t)3 << 29);
    int saved = dup(1); FILE *nul = fopen("/dev/null", "w");
    double best[6] = {1e9, 1e9, 1e9, 1e9, 1e9, 1e9}, ps = 0; size_t cs = 0;
    unsigned long long transcript = 0, launches_last = 0;
    for (int rep = 0; rep <= reps; rep++) {                 // rep 0 = warm-up (tables, allocator)
        fflush(stdout); if (!getenv("HOBBIT_KEEP_STDOUT")) dup2(fileno(nul), 1);   // the reference-style printf chatter of the provers
        srand(1);
        hb_transcript_digest(backend(), 1);
        const unsigned long long l0 = hb_launch_count(backend());
        double t0 = now();
        if (aes_n >= 0 && sql) trace_generate_sql(1 << aes_n); else if (aes_n >= 0) trace_generate_aes(1 << aes_n); else trace_generate_mlp(layers);
        cs = trace_end();
        double t1 = now();
        BUFFER_SPACE = (size_t)1 << b;
        if (aes_n >= 0) { if (BUFFER_SPACE > ((size_t)128 << aes_n)) BUFFER_SPACE = (size_t)128 << aes_n; }      // init_stream (main.cpp:1093)
        else if (BUFFER_SPACE > cs) BUFFER_SPACE = cs / 4;
        has_lookups = aes_n >= 0;
        a_w = F(random()); b_w = F(random());                 // main() (main.cpp:1227) ...
        a_w = F(random()); b_w = F(random());                 // ... and again in prove_circuit (:873): same libc position as `pigeon 9 ...`
        double vt = 0; ps = 0;
        stream_descriptor fd1; fd1.name = "transcript_stream"; fd1.size = cs;
        stream_descriptor fd2; fd2.name = "wiring_consistency_check_opt"; fd2.size = 8 * cs;
        stream_descriptor fdw; fdw.name = "witness"; fdw.size = 4 * cs;
        std::vector<std::vector<_hash>> MT; _hash comm;
        stream_descriptor fd3; fd3.name = "lookup_basic"; fd3.size = 2 * cs;
        stream_descriptor fdl; fdl.name = "lookup_witness_basic"; fdl.size = 2 * cs;
        std::vector<std::vector<_hash>> MTl;
        init_commitment(false);
        commit(fdw, comm, MT);
        if (has_lookups) { lookup_rand = generate_randomness(4); commit(fdl, comm, MTl); }
        double t2 = now();
        std::vector<F> prods = prove_multiplication_tree_stream_shallow(fd2, 8, (int)(fd2.size / 8), F(32), 5, std::vector<F>(), 0, vt, ps);
        if (has_lookups) prove_multiplication_tree_stream_shallow(fd3, 2, (int)(fd3.size / 2), F(32), 5, std::vector<F>(), 0, vt, ps);
        double t3 = now();
        if (has_lookups) prove_gate_consistency_lookups(fd1, generate_randomness((int)std::log2((double)fd1.size)), vt, ps);
        else prove_gate_consistency(fd1, generate_randomness((int)std::log2((double)fd1.size)), vt, ps);
        double t4 = now();
        open(fdw, generate_randomness((int)std::log2((double)fdw.size)), MT, vt, ps);
        if (has_lookups) open(fdl, generate_randomness((int)std::log2((double)fdl.size)), MTl, vt, ps);
        double t5 = now();
        wait_levels();
        transcript = hb_transcript_digest(backend(), 0); launches_last = hb_launch_count(backend()) - l0;
        for (auto &lv : MT) if (lv.size() == 1) for (int q = 0; q < 32; q++) transcript = (transcript ^ ((const unsigned char *)&lv[0])[q]) * 0x100000001b3ULL;   // + the commitment root
        fflush(stdout); dup2(saved, 1);
        F rd = prods[0] * prods[1] * prods[2] * prods[7], wr = prods[4] * prods[5] * prods[6] * prods[3];
        if (rd != wr) { printf("memory consistency check FAILED\n"); return 1; }
        if (rep) { const double t[6] = {t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0}; for (int i = 0; i < 6; i++) best[i] = std::min(best[i], t[i]); }
    }
    if (dist_rank() != 0) return 0;                          // every rank holds the same proof; rank 0 reports
    if (aes_n >= 0) printf("{\"workload\": \"%s prove_circuit (lookups), 2^%d %s", sql ? "SQL" : "AES", aes_n, sql ? "rows" : "blocks");
    else { printf("{\"workload\": \"MLP prove_circuit, layers"); for (int l : layers) printf(" %d", l); }
    printf(", circuit_size 2^%d, BUFFER_SPACE 2^%d%s\", \"evaluate_s\": %.5f, \"commit_s\": %.5f, \"mul_tree_s\": %.5f, \"gate_s\": %.5f, \"open_s\": %.5f, \"total_s\": %.5f, "
           "\"ps_kb\": %.6f, \"gates_per_s\": %.1f, \"gpu_launches\": %llu, \"launches_per_proof\": %llu, \"n_gpus\": %d, \"transcript\": \"%016llx\", \"rng_next\": %ld}\n",
           (int)std::log2((double)cs), (int)std::log2((double)BUFFER_SPACE), !commit_levels_on_host ? ", Merkle levels resident in HBM" : commit_levels_async ? ", Merkle levels to the host in the background" : "", best[0], best[1], best[2], best[3], best[4], best[5], ps, cs / best[5],
           (unsigned long long)hb_launch_count(backend()), (unsigned long long)launches_last, dist_world(), (unsigned long long)transcript, random());
    return 0;
}
