// Development check (CPU only, no GPU needed): the closed-form AES / SQL records of hobbit_b200/csrc/aes_circuit.cuh and sql_circuit.cuh, computed on the host and
// pushed through hb_trace_push into the C-ABI emulation (oracle/libhb_emul.so, test infrastructure), must give the same derived streams as
// the emulation's own gate-by-gate evaluator (hb_trace_generate_aes there restates Seval.cpp:957-1084 one gate at a time).
// Build + run:  nvcc -O2 -o build/aes_records_check tools/aes_records_check.cu -Loracle -lhb_emul -Xlinker -rpath=$PWD/oracle && build/aes_records_check 5
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../include/hobbit_b200.h"
#include "../hobbit_b200/csrc/aes_circuit.cuh"
#include "../hobbit_b200/csrc/sql_circuit.cuh"
using namespace hb;

static void streams(hb_ctx *c, size_t cs, std::vector<hb_F> &all) {
    hb_F a_w = {123456789, 987654321}, b_w = {5555, 7777}, lr[4] = {{11, 12}, {13, 14}, {15, 16}, {17, 18}};
    all.assign(4 * cs + 4 * cs + 8 * cs + 2 * cs + 2 * cs, hb_F{0, 0});
    hb_F *p = all.data();
    if (hb_trace_witness(c, cs, p)) exit(2);
    if (hb_trace_transcript(c, cs, 1, p + 4 * cs, p + 5 * cs, p + 6 * cs, p + 7 * cs)) exit(2);
    if (hb_trace_wiring(c, cs, &a_w, &b_w, p + 8 * cs)) exit(2);
    if (hb_trace_lookup_basic(c, cs, lr, p + 16 * cs)) exit(2);
    if (hb_trace_lookup_witness(c, cs, lr, p + 18 * cs)) exit(2);
}
int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 5;
    const bool sql = argc > 2 && !strcmp(argv[2], "sql");                      // aes_records_check <rows> sql: the SQL range-query circuit instead
    hb_ctx *a, *b;
    if (hb_ctx_create(&a, 0) || hb_ctx_create(&b, 0)) { printf("ctx\n"); return 1; }
    size_t na = 0;
    if (sql ? hb_trace_generate_sql(a, n, &na) : hb_trace_generate_aes(a, n, &na)) return 1;
    std::vector<TrTuple> rec;
    if (sql) for (size_t g = 0; g < sql_records(n); g++) rec.push_back(sql_record(n, g));
    else {
        for (int blk = 0; blk < n; blk++) for (int r = 0; r < kAesRecs; r++) rec.push_back(aes_record(n, blk, r));
        for (int i = 0; i <= 16 * n + 160; i++) rec.push_back(aes_tail_record(n, i));
    }
    TrTuple end; memset(&end, 0, sizeof end); end.type = 255; rec.push_back(end);
    int done = 0;
    if (hb_trace_begin(b, rec.size()) || hb_trace_push(b, rec.data(), rec.size(), &done) || !done) return 1;
    size_t cnt_a[3], cnt_b[3];
    hb_trace_finish(a, &cnt_a[0], &cnt_a[1], &cnt_a[2]); hb_trace_finish(b, &cnt_b[0], &cnt_b[1], &cnt_b[2]);
    printf("records %zu / %zu, ops %zu / %zu, deletes %zu / %zu\n", cnt_a[0], cnt_b[0], cnt_a[1], cnt_b[1], cnt_a[2], cnt_b[2]);
    if (memcmp(cnt_a, cnt_b, sizeof cnt_a) || na != rec.size() - 1) { printf("FAIL: counts\n"); return 1; }
    size_t cs = 1; while (cs < cnt_a[2]) cs *= 2;
    std::vector<hb_F> sa, sb;
    streams(a, cs, sa); streams(b, cs, sb);
    for (size_t i = 0; i < sa.size(); i++)
        if (memcmp(&sa[i], &sb[i], 16)) { printf("FAIL: stream element %zu of %zu (cs %zu): %llu,%llu vs %llu,%llu\n", i, sa.size(), cs,
            (unsigned long long)sa[i].real, (unsigned long long)sa[i].img, (unsigned long long)sb[i].real, (unsigned long long)sb[i].img); return 1; }
    printf("ok: %d %s, every derived stream identical (cs %zu)\n", n, sql ? "rows" : "blocks", cs);
    return 0;
}
