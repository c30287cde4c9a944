// 8f.4: closed-form records of the SQL range-query circuit evaluator (included by trace.cu; host+device so that the closed forms can also
// be checked on a CPU against the gate-by-gate restatement, tools/aes_records_check.cu).
// Reference: Seval.cpp:1085-1166 range_query, :667-687 get_bytes, :223-293 ltu_gate, :190-221 lookup_gate, fun == 6 driver :1398-1416:
// DB[i] = (i + 12) % 16, L = 21, R = 321 (never decomposed: both bound vectors are the bytes of L, :1099), two bytes per word.
// Every row runs the same gate program (34 records for row 0, 60 for every later row, which also folds the row into the running maximum);
// the shared gates (range table, byte weights, zero, minus one, the bound bytes) only differ by access counters that are closed forms of the
// row index.  With the driver's data every row value is below L, so the selected value and the running maximum stay zero and no row
// depends on an earlier one through VALUES; a database with rows inside the range would need the sequential recurrence and is not
// what this entry point generates.
#pragma once
#include "aes_circuit.cuh"

namespace hb {

constexpr int kSqlPrologue = 16, kSqlRow0 = 34, kSqlRow = 60, kSqlRow0Labels = 18, kSqlRowLabels = 30, kSqlTailFixed = 268;
struct SqlGate { int label; u64 val; int acc; };
__host__ __device__ __forceinline__ u64 sql_mul(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) % P61); }
__host__ __device__ __forceinline__ u64 sql_add(u64 a, u64 b) { u64 s = a + b; return s >= P61 ? s - P61 : s; }
__host__ __device__ __forceinline__ void sql_op(TrTuple &t, int type, SqlGate l, SqlGate r, int lo, u64 vo) {
    t.type = (uint8_t)type; t.idx_l = l.label; t.value_l = mkF(l.val, 0); t.access_l = l.acc; t.idx_r = r.label; t.value_r = mkF(r.val, 0); t.access_r = r.acc;
    t.idx_o = lo; t.value_o = mkF(vo, 0); t.access_o = 0;
}
__host__ __device__ __forceinline__ void sql_del(TrTuple &t, int label, u64 val, int acc) { t.type = 0; t.idx_o = label; t.value_o = mkF(val, 0); t.access_o = acc; }

// labels of the shared gates (n rows)
struct SqlLabels {
    int n;
    __host__ __device__ int db(int i) const { return 1 + i; }
    __host__ __device__ int R() const { return n + 1; }
    __host__ __device__ int L() const { return n + 2; }
    __host__ __device__ int range(int e) const { return n + 3 + e; }
    __host__ __device__ int pow(int j) const { return n + 259 + j; }
    __host__ __device__ int zero() const { return n + 261; }
    __host__ __device__ int minus_one() const { return n + 262; }
    __host__ __device__ int lb(int j) const { return n + 263 + j; }            // the bytes of L from the first get_bytes(L) ...
    __host__ __device__ int rb(int j) const { return n + 268 + j; }            // ... and from the second
    __host__ __device__ int row(int i) const { return n + 273 + (i ? kSqlRow0Labels + kSqlRowLabels * (i - 1) : 0); }
};
// earlier rows whose low byte is zero: i' < i with (i' + 12) % 16 == 0
__host__ __device__ __forceinline__ int sql_zero_rows(int i) { return (i + 11) / 16; }
// access counter of range-table entry e at the k-th lookup (k = 0: low byte, 1: high byte) of row i; the prologue read entries 21, 0, 21, 0
__host__ __device__ __forceinline__ int sql_range_access(int i, int k) {
    const int v = (i + 12) % 16;
    const int zero_before = 2 + i + sql_zero_rows(i);                          // reads of entry 0 before this row
    if (k == 0) return v == 0 ? zero_before : i / 16;
    return zero_before + (v == 0 ? 1 : 0);
}
// records of one get_bytes(word) (8 records, 5 labels from `lab`): the two range lookups, the recomposition check and its deletes.
// zacc / pacc: access counters of zero / of the byte weights on entry; racc[k]: of the range entries
__host__ __device__ __forceinline__ void sql_get_bytes(TrTuple &t, int q, const SqlLabels &S, u64 word, int lab, int zacc, int pacc, int racc0, int racc1) {
    const u64 b[2] = {word & 255, (word >> 8) & 255};
    const u64 check = b[0], tp = sql_mul(256, b[1]), tc = sql_add(check, tp);
    if (q < 2) sql_op(t, 3, SqlGate{S.range((int)b[q]), b[q], q ? racc1 : racc0}, SqlGate{S.zero(), 0, zacc + q}, lab + q, b[q]);
    else if (q == 2) sql_op(t, 2, SqlGate{S.pow(0), 1, pacc}, SqlGate{lab, b[0], 1}, lab + 2, check);
    else if (q == 3) sql_op(t, 2, SqlGate{S.pow(1), 256, pacc}, SqlGate{lab + 1, b[1], 1}, lab + 3, tp);
    else if (q == 4) sql_op(t, 1, SqlGate{lab + 2, check, 1}, SqlGate{lab + 3, tp, 1}, lab + 4, tc);
    else if (q == 5) sql_del(t, lab + 2, check, 2);
    else if (q == 6) sql_del(t, lab + 3, tp, 2);
    else sql_del(t, lab + 4, tc, 1);
}
// records of one ltu_gate(x, y) on two-byte operands (9 records, 5 labels from `lab`); returns the value of its output (label lab + 4)
__host__ __device__ __forceinline__ u64 sql_ltu(TrTuple &t, int q, const SqlGate x[2], const SqlGate y[2], int lab) {
    const u64 lt0 = y[0].val < x[0].val, lt1 = y[1].val < x[1].val, eq0 = x[0].val == y[0].val, out = sql_mul(eq0, lt0), to = sql_add(out, lt1);
    if (q == 0) sql_op(t, 4, x[0], y[0], lab, lt0);
    else if (q == 1) sql_op(t, 4, x[1], y[1], lab + 1, lt1);
    else if (q == 2) sql_op(t, 5, SqlGate{x[0].label, x[0].val, x[0].acc + 1}, SqlGate{y[0].label, y[0].val, y[0].acc + 1}, lab + 2, eq0);
    else if (q == 3) sql_op(t, 2, SqlGate{lab + 2, eq0, 1}, SqlGate{lab, lt0, 1}, lab + 3, out);
    else if (q == 4) sql_op(t, 1, SqlGate{lab + 3, out, 1}, SqlGate{lab + 1, lt1, 1}, lab + 4, to);
    else if (q == 5) sql_del(t, lab + 3, out, 2);
    else if (q == 6) sql_del(t, lab, lt0, 2);
    else if (q == 7) sql_del(t, lab + 1, lt1, 2);
    else if (q == 8) sql_del(t, lab + 2, eq0, 2);
    return to;
}
// record q of row i (q < 34 for row 0, < 60 otherwise)
__host__ __device__ __forceinline__ TrTuple sql_row_record(int n, int i, int q) {
    const SqlLabels S{n};
    TrTuple t; memset(&t, 0, sizeof t);
    const int lab = S.row(i);
    const u64 v = (u64)((i + 12) % 16), lv = 21;
    const u64 b[2] = {v & 255, v >> 8}, lbv[2] = {lv & 255, lv >> 8};
    // the row's bytes after get_bytes have access 2; ltu(bytes, Rb) reads them at 2 / 2, ltu(Lb, bytes) at 4 / 3
    const SqlGate by1[2] = {{lab, b[0], 2}, {lab + 1, b[1], 2}}, by2[2] = {{lab, b[0], 4}, {lab + 1, b[1], 3}};
    const SqlGate Rb[2] = {{S.rb(0), lbv[0], 2 + 2 * i}, {S.rb(1), lbv[1], 2 + i}}, Lb[2] = {{S.lb(0), lbv[0], 2 + 2 * i}, {S.lb(1), lbv[1], 2 + i}};
    TrTuple scratch;
    const u64 bit1 = sql_ltu(q >= 8 && q < 17 ? t : scratch, q >= 8 && q < 17 ? q - 8 : 0, by1, Rb, lab + 5);
    const u64 bit2 = sql_ltu(q >= 17 && q < 26 ? t : scratch, q >= 17 && q < 26 ? q - 17 : 0, Lb, by2, lab + 10);
    const u64 bit = sql_mul(bit1, bit2), tm[2] = {sql_mul(b[0], bit), sql_mul(b[1], bit)};
    if (q < 8) { memset(&t, 0, sizeof t); sql_get_bytes(t, q, S, v, lab, 4 + 2 * i, 2 + i, sql_range_access(i, 0), sql_range_access(i, 1)); }
    else if (q < 26) {}
    else if (q == 26) sql_op(t, 2, SqlGate{lab + 9, bit1, 1}, SqlGate{lab + 14, bit2, 1}, lab + 15, bit);
    else if (q == 27) sql_del(t, lab + 9, bit1, 2);
    else if (q == 28) sql_del(t, lab + 14, bit2, 2);
    else if (q == 29) sql_op(t, 2, SqlGate{lab, b[0], 6}, SqlGate{lab + 15, bit, 1}, lab + 16, tm[0]);
    else if (q == 30) sql_op(t, 2, SqlGate{lab + 1, b[1], 4}, SqlGate{lab + 15, bit, 2}, lab + 17, tm[1]);
    else if (q == 31) sql_del(t, lab + 15, bit, 3);
    else if (q == 32) sql_del(t, lab, b[0], 7);
    else if (q == 33) sql_del(t, lab + 1, b[1], 5);
    else {
        // fold into the running maximum (rows >= 1): max is (0, 0) with the driver's data; its gates are row 0's selected bytes or the previous row's sums
        const int ml[2] = {i == 1 ? S.row(0) + 16 : S.row(i - 1) + 26, i == 1 ? S.row(0) + 17 : S.row(i - 1) + 29};
        const SqlGate T[2] = {{lab + 16, tm[0], 1}, {lab + 17, tm[1], 1}}, M[2] = {{ml[0], 0, 1}, {ml[1], 0, 1}};
        const u64 mb = sql_ltu(q < 43 ? t : scratch, q < 43 ? q - 34 : 0, T, M, lab + 18);
        const u64 mbn = sql_add(mb, P61 - 1);
        if (q < 43) {}
        else if (q == 43) sql_op(t, 1, SqlGate{lab + 22, mb, 1}, SqlGate{S.minus_one(), P61 - 1, i - 1}, lab + 23, mbn);
        else if (q < 58) {
            const int j = (q - 44) / 7, e = (q - 44) % 7, pl = lab + 24 + 3 * j;
            const int macc = (j == 0 ? 3 : 2), tacc = (j == 0 ? 3 : 2);                 // after ltu(tm, max): byte 0 was read twice, byte 1 once
            const u64 p1 = sql_mul(M[j].val, mb), p2 = sql_mul(tm[j], mbn), nm = sql_add(p1, p2);
            if (e == 0) sql_op(t, 2, SqlGate{M[j].label, M[j].val, macc}, SqlGate{lab + 22, mb, 2 + j}, pl, p1);
            else if (e == 1) sql_op(t, 2, SqlGate{T[j].label, tm[j], tacc}, SqlGate{lab + 23, mbn, 1 + j}, pl + 1, p2);
            else if (e == 2) sql_del(t, M[j].label, M[j].val, macc + 1);
            else if (e == 3) sql_del(t, T[j].label, tm[j], tacc + 1);
            else if (e == 4) sql_op(t, 1, SqlGate{pl, p1, 1}, SqlGate{pl + 1, p2, 1}, pl + 2, nm);
            else if (e == 5) sql_del(t, pl, p1, 2);
            else sql_del(t, pl + 1, p2, 2);
        }
        else if (q == 58) sql_del(t, lab + 22, mb, 4);
        else sql_del(t, lab + 23, mbn, 3);
    }
    return t;
}
// the 16 records before the first row: get_bytes(L) twice
__host__ __device__ __forceinline__ TrTuple sql_prologue_record(int n, int q) {
    const SqlLabels S{n};
    TrTuple t; memset(&t, 0, sizeof t);
    const int second = q >= 8;
    // range entries 21 and 0 are read at access `second` each; zero at 2 * second (+1); the byte weights at `second`
    sql_get_bytes(t, q - 8 * second, S, 21, second ? S.rb(0) : S.lb(0), 2 * second, second, second, second);
    return t;
}
// the records after the last row: delete the running maximum, zero, the byte weights, the range table, the bound bytes, the rows, R, L, minus one
__host__ __device__ __forceinline__ TrTuple sql_tail_record(int n, int q) {
    const SqlLabels S{n};
    TrTuple t; memset(&t, 0, sizeof t);
    if (q < 2) sql_del(t, n == 1 ? S.row(0) + 16 + q : S.row(n - 1) + 26 + 3 * q, 0, 1);
    else if (q == 2) sql_del(t, S.zero(), 0, 4 + 2 * n);
    else if (q < 5) sql_del(t, S.pow(q - 3), q == 3 ? 1 : 256, 2 + n);
    else if (q < 261) {
        const int e = q - 5;
        int acc = 0;
        if (e == 0) acc = 2 + n + sql_zero_rows(n);
        else if (e < 16) { const int r = (e + 4) % 16; acc = n > r ? (n - r + 15) / 16 : 0; }
        if (e == 21) acc += 2;
        sql_del(t, S.range(e), (u64)e, acc);
    }
    else if (q < 265) { const int j = (q - 261) >> 1, right = (q - 261) & 1; sql_del(t, right ? S.rb(j) : S.lb(j), j ? 0 : 21, 2 + (j ? n : 2 * n)); }
    else if (q < 265 + n) sql_del(t, S.db(q - 265), (u64)((q - 265 + 12) % 16), 0);
    else if (q == 265 + n) sql_del(t, S.R(), 321, 0);
    else if (q == 266 + n) sql_del(t, S.L(), 21, 0);
    else sql_del(t, S.minus_one(), P61 - 1, n - 1);
    return t;
}
__host__ __device__ __forceinline__ size_t sql_records(int n) { return (size_t)kSqlPrologue + kSqlRow0 + (size_t)kSqlRow * (n - 1) + kSqlTailFixed + n; }
// record g of the whole trace
__host__ __device__ __forceinline__ TrTuple sql_record(int n, size_t g) {
    if (g < kSqlPrologue) return sql_prologue_record(n, (int)g);
    g -= kSqlPrologue;
    if (g < kSqlRow0) return sql_row_record(n, 0, (int)g);
    g -= kSqlRow0;
    if (g < (size_t)kSqlRow * (n - 1)) return sql_row_record(n, 1 + (int)(g / kSqlRow), (int)(g % kSqlRow));
    return sql_tail_record(n, (int)(g - (size_t)kSqlRow * (n - 1)));
}

}  // namespace hb
