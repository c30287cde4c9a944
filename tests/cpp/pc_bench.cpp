// BASELINE config 1 side by side: test_PC(N, option 4 / 1, K) = commit_standard + open_standard (Our_PC.cpp:763-779, 806-826) through the
// unmodified reference (1 host thread) and through the host mirror on the GPU, in one process; prints one JSON line.
// usage: pc_bench logN K lin [reps] [skip_ref]
#include "../../hobbit_b200/host/hobbit_host.hpp"
namespace hobbit { typedef F Fe; }
#include "config_pc.hpp"
#include "utils.hpp"
#include "mimc.h"
#include "Our_PC.hpp"
#include "sumcheck.h"
#include <chrono>
#include <cstdio>
#include <cstring>
extern bool linear_time;
extern int tensor_row_size;
extern bool __encode_initialized;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    int logN = argc > 1 ? atoi(argv[1]) : 20, K = argc > 2 ? atoi(argv[2]) : 32, lin = argc > 3 ? atoi(argv[3]) : 1;
    int reps = argc > 4 ? atoi(argv[4]) : 3; bool skip_ref = argc > 5 && atoi(argv[5]);
    const size_t N = (size_t)1 << logN;
    init_hash();
    hobbit::init_backend(0);
    double vt = 0, ps = 0, hps = 0, ref_commit = 0, ref_open = 0;
    std::vector<std::vector<hobbit::_hash>> hMT; std::vector<std::vector<_hash>> MT;
    FILE *devnull = fopen("/dev/null", "w"); FILE *real_stdout = stdout;
    if (!skip_ref) {
        srand(1);
        vector<F> poly = generate_randomness(N);
        linear_time = lin; tensor_row_size = N / (K * 1ULL << 11);
        if (lin) { __encode_initialized = false; expander_init_store(tensor_row_size); }
        _hash comm; vector<vector<vector<F>>> T;
        stdout = devnull;
        double t0 = now(); commit_standard(poly, comm, MT, T, K); double t1 = now();
        vector<F> x = generate_randomness(logN);
        double t2 = now(); open_standard(poly, x, MT, T, K, vt, ps); double t3 = now();
        stdout = real_stdout;
        ref_commit = t1 - t0; ref_open = t3 - t2;
    }
    double best_c = 1e9, best_o = 1e9;
    for (int rep = 0; rep < reps + 1; rep++) {          // rep 0 = warm-up (context, twiddle tables, allocator)
        srand(1);
        std::vector<hobbit::Fe> hpoly = hobbit::generate_randomness(N);
        hobbit::linear_time = lin; hobbit::tensor_row_size = N / (K * 1ULL << 11);
        if (lin) hobbit::expander_init_store(hobbit::tensor_row_size);
        hobbit::_hash hcomm; std::vector<std::vector<std::vector<hobbit::Fe>>> hT;
        stdout = devnull;
        double t0 = now(); hobbit::commit_standard(hpoly, hcomm, hMT, hT, K); double t1 = now();
        std::vector<hobbit::Fe> hx = hobbit::generate_randomness(logN);
        hps = 0;
        double t2 = now(); hobbit::open_standard(hpoly, hx, hMT, hT, K, vt, hps); double t3 = now();
        stdout = real_stdout;
        if (rep) { best_c = std::min(best_c, t1 - t0); best_o = std::min(best_o, t3 - t2); }
    }
    bool same = skip_ref || (ps == hps && MT.size() == hMT.size() && !memcmp(MT.back().data(), hMT.back().data(), 32) && !memcmp(MT[0].data(), hMT[0].data(), MT[0].size() * 32));
    printf("{\"workload\": \"test_PC(2^%d, %s, K=%d)\", \"gpu_commit_s\": %.6f, \"gpu_open_s\": %.6f, \"ref_commit_s\": %.4f, \"ref_open_s\": %.4f, "
           "\"ref_threads\": 1, \"ps_kb\": %.6f, \"identical\": %s, \"gpu_launches\": %llu}\n",
           logN, lin ? "Orion columns" : "RS columns", K, best_c, best_o, ref_commit, ref_open, hps, same ? "true" : "false",
           (unsigned long long)hb_launch_count(hobbit::backend()));
    return same ? 0 : 1;
}
