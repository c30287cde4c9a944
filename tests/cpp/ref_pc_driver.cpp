// Calls the reference's OWN test_PC (Our_PC.cpp:757-861) / test_Elastic_PC (Elastic_PC.cpp:736-808) — its main.cpp has these calls commented
// out (main.cpp:1176-1179) — in a binary whose commit/open entry points are the GPU adapter's (oracle/build_pigeon_gpu.sh).
//   ref_pc_gpu pc <logN> <option> <K>          ref_pc_gpu elastic <logN> <logB> <option>
#include "config_pc.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
void test_PC(size_t N, int option, int K);
void test_Elastic_PC(size_t N, int option);
void init_hash();
extern size_t BUFFER_SPACE;
extern std::mutex mtx, mtx2;
int main(int argc, char **argv) {
    if (argc < 5) { printf("usage: ref_pc_gpu pc <logN> <option> <K> | ref_pc_gpu elastic <logN> <logB> <option>\n"); return 2; }
    init_hash();
    if (!strcmp(argv[1], "pc")) test_PC((size_t)1 << atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
    else { BUFFER_SPACE = (size_t)1 << atoi(argv[3]); test_Elastic_PC((size_t)1 << atoi(argv[2]), atoi(argv[4])); }
    return 0;
}
