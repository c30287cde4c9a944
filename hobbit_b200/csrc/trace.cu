// W1/W2 — the circuit witness streams derived on the GPU from the evaluator's trace.
// Reference: tr_tuple (src/Seval.h:4-9), the stateful readers read_witness / _read_witness (src/witness_stream.cpp:768-874, 1260-1338),
// read_trace (:1701-1807), read_memory_opt / read_memory_trancript / read_final_memory_trancript (:1055-1258, 1620-1698) and the
// "wiring_consistency_check_opt" branch of read_stream (:2276-2311).
//
// The reference re-executes the circuit once per pass over a stream (a producer thread refills an 80-byte-per-gate ring buffer and
// the readers above pull from it under a mutex hand-off: "streaming time", about 20 % of its prover time).  Here ONE pass of the
// trace is uploaded and kept in HBM (80 B per tuple: 160 MiB for the 2^20-gate MLP) and every named stream is a pure function of it:
// a stream position is the rank of an op tuple (type 1..254) or of a delete tuple (type 0), so two exclusive scans over the type byte
// give every element its destination and the streams are written by one scatter kernel each — HBM-bound, 80 B read per tuple.
#include "common.cuh"
#include "aes_circuit.cuh"        // TrTuple + the AES gate program in closed form
#include "sql_circuit.cuh"        // the SQL range-query gate program in closed form
#include <algorithm>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

namespace hb {


__device__ __forceinline__ F f_int(int x) { return mkF(x >= 0 ? (u64)x : P61 - (u64)(-(long long)x), 0); }

__global__ void __launch_bounds__(256) trace_flags_kernel(const TrTuple *__restrict__ tr, size_t n, unsigned *__restrict__ is_op, unsigned *__restrict__ is_del) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t t = tr[i].type;
    is_op[i] = (t > 0 && t != 255) ? 1u : 0u;
    is_del[i] = (t == 0) ? 1u : 0u;
}
// "witness" (4cs): [ (l, r, o) of op tuple p at 3p.. | value_o of delete tuple q at 3cs + q ], zero padded (memset by the caller)
__global__ void __launch_bounds__(256)
trace_witness_kernel(const TrTuple *__restrict__ tr, size_t n, const unsigned *__restrict__ op_pos, const unsigned *__restrict__ del_pos, size_t cs, F *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TrTuple &t = tr[i];
    if (t.type == 0) out[3 * cs + del_pos[i]] = t.value_o;
    else if (t.type != 255) { size_t p = 3 * (size_t)op_pos[i]; out[p] = t.value_l; out[p + 1] = t.value_r; out[p + 2] = t.value_o; }
}
// "transcript_stream" (read_trace): L, R, O and the selector S (no lookups: add 1 / mul 0; lookups: add 0 / mul 1 / table 2)
__global__ void __launch_bounds__(256)
trace_transcript_kernel(const TrTuple *__restrict__ tr, size_t n, const unsigned *__restrict__ op_pos, int has_lookups,
                        F *__restrict__ L, F *__restrict__ R, F *__restrict__ O, F *__restrict__ S) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TrTuple &t = tr[i];
    if (t.type == 0 || t.type == 255) return;
    size_t p = op_pos[i];
    L[p] = t.value_l; R[p] = t.value_r; O[p] = t.value_o;
    int s = t.type == 1 ? (has_lookups ? 0 : 1) : t.type == 2 ? (has_lookups ? 1 : 0) : 2;
    S[p] = mkF((u64)s, 0);
}
// "wiring_consistency_check_opt" in its logical two-half form [X | Y] (4cs each):
//   X[3p+c] = idx+1 + a_w value + b_w access (read set), Y = X + b_w (write set), or 1 where X == 1
//   X[3cs+q] = idx_o+1 + a_w value_o (init set), Y = X + b_w access_o (final set); padding positions are 1 in both halves.
__device__ __forceinline__ F wire_entry(int idx, F value, int access, F a_w, F b_w) {
    return fadd(fadd(fadd(f_int(idx), mkF(1, 0)), fmul(a_w, value)), fmul(b_w, f_int(access)));
}
__global__ void __launch_bounds__(256)
trace_wiring_kernel(const TrTuple *__restrict__ tr, size_t n, const unsigned *__restrict__ op_pos, const unsigned *__restrict__ del_pos, size_t cs,
                    F a_w, F b_w, F *__restrict__ X, F *__restrict__ Y) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TrTuple &t = tr[i];
    const F one = mkF(1, 0);
    if (t.type == 0) {
        size_t q = 3 * cs + del_pos[i];
        F x = fadd(fadd(f_int(t.idx_o), one), fmul(a_w, t.value_o));
        X[q] = x; Y[q] = fadd(x, fmul(b_w, f_int(t.access_o)));
    } else if (t.type != 255) {
        size_t p = 3 * (size_t)op_pos[i];
        F x0 = wire_entry(t.idx_l, t.value_l, t.access_l, a_w, b_w), x1 = wire_entry(t.idx_r, t.value_r, t.access_r, a_w, b_w),
          x2 = wire_entry(t.idx_o, t.value_o, t.access_o, a_w, b_w);
        X[p] = x0; X[p + 1] = x1; X[p + 2] = x2;
        Y[p] = feq(x0, one) ? x0 : fadd(x0, b_w); Y[p + 1] = feq(x1, one) ? x1 : fadd(x1, b_w); Y[p + 2] = feq(x2, one) ? x2 : fadd(x2, b_w);
    }
}
// "circuit" (16cs, no lookups; witness_stream.cpp:2123-2162 with read_circuit_trace :2021-2104 and read_memory_circuit :1812-2019):
//   [ selector per op record: 1 add / 0 mul (cs) | (idx, access) pairs of (l, r, o) per op record (6cs) | (idx_o, access_o) pairs per
//     delete record (2cs) | zeros (7cs) ]   — zero padded inside each section (memset by the caller)
__global__ void __launch_bounds__(256)
trace_circuit_kernel(const TrTuple *__restrict__ tr, size_t n, const unsigned *__restrict__ op_pos, const unsigned *__restrict__ del_pos, size_t cs, F *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TrTuple &t = tr[i];
    if (t.type == 255) return;
    if (t.type == 0) {
        F *m = out + cs + 6 * cs + 2 * (size_t)del_pos[i];
        m[0] = f_int(t.idx_o); m[1] = f_int(t.access_o);
    } else {
        const size_t p = op_pos[i];
        out[p] = mkF(t.type == 1 ? 1 : 0, 0);
        F *m = out + cs + 6 * p;
        m[0] = f_int(t.idx_l); m[1] = f_int(t.access_l); m[2] = f_int(t.idx_r); m[3] = f_int(t.access_r); m[4] = f_int(t.idx_o); m[5] = f_int(t.access_o);
    }
}
// ---- lookup streams (witness_stream.cpp:920-1053, 2198-2247) -------------------------------------------------------------------------
// The reference threads a mutable access_table through the pass: a lookup record carries the number of EARLIER lookups of the same table
// entry.  On the GPU that is the rank of the record among equal (table, entry) keys in trace order: a stable radix sort of the keys,
// a max-scan of the run starts, and a scatter of (position - run start) back to trace order.
__global__ void __launch_bounds__(256)
lookup_keys_kernel(const TrTuple *__restrict__ tr, size_t n, unsigned long long *__restrict__ keys, unsigned *__restrict__ idx, unsigned *__restrict__ is_lkp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TrTuple &t = tr[i];
    const bool lk = t.type >= 3 && t.type != 255;
    unsigned long long entry = t.type == 3 ? t.value_l.re : t.value_l.re + 256ull * t.value_r.re;
    keys[i] = lk ? (((unsigned long long)t.type << 40) | (entry & ((1ull << 40) - 1))) : ~0ull;
    idx[i] = (unsigned)i;
    is_lkp[i] = lk ? 1u : 0u;
}
__global__ void __launch_bounds__(256) run_start_kernel(const unsigned long long *__restrict__ keys, size_t n, unsigned *__restrict__ start) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    start[i] = (i == 0 || keys[i] != keys[i - 1]) ? (unsigned)i : 0u;
}
__global__ void __launch_bounds__(256)
access_scatter_kernel(const unsigned *__restrict__ idx_sorted, const unsigned *__restrict__ start, size_t n, unsigned *__restrict__ access) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    access[idx_sorted[i]] = (unsigned)i - start[i];
}
struct MaxOp { __device__ __forceinline__ unsigned operator()(unsigned a, unsigned b) const { return a > b ? a : b; } };
// "lookup_basic": X = 1 + value_l + lr0 value_r + lr1 value_o + lr2 access + lr3 type at the op rank of a lookup record (1 elsewhere), Y = X + lr2
__global__ void __launch_bounds__(256)
lookup_basic_kernel(const TrTuple *__restrict__ tr, size_t n, const unsigned *__restrict__ op_pos, const unsigned *__restrict__ access,
                    F lr0, F lr1, F lr2, F lr3, F *__restrict__ X, F *__restrict__ Y) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TrTuple &t = tr[i];
    if (t.type < 3 || t.type == 255) return;
    F v = fadd(mkF(1, 0), t.value_l);
    v = fadd(v, fmul(lr0, t.value_r)); v = fadd(v, fmul(lr1, t.value_o));
    v = fadd(v, fmul(lr2, mkF(access[i], 0))); v = fadd(v, fmul(lr3, mkF(t.type, 0)));
    const size_t p = op_pos[i];
    X[p] = v; Y[p] = feq(v, mkF(1, 0)) ? v : fadd(v, lr2);
}
// "lookup_witness_basic": (value_o + lr0 value_l + lr1 value_r, access) at the rank of the record among the lookup records
__global__ void __launch_bounds__(256)
lookup_witness_kernel(const TrTuple *__restrict__ tr, size_t n, const unsigned *__restrict__ lkp_pos, const unsigned *__restrict__ access,
                      F lr0, F lr1, F *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TrTuple &t = tr[i];
    if (t.type < 3 || t.type == 255) return;
    const size_t q = lkp_pos[i];
    out[2 * q] = fadd(fadd(t.value_o, fmul(lr0, t.value_l)), fmul(lr1, t.value_r));
    out[2 * q + 1] = mkF(access[i], 0);
}
__global__ void __launch_bounds__(256) fill_kernel(F *__restrict__ v, size_t n, F x) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) v[i] = x;
}

// ---- 8f.4: the MLP circuit evaluator on the GPU (Seval.cpp:1238-1286 MLP_inference, :1462-1489 the fun == 9 driver) --------------------
// The evaluator is sequential only in its bookkeeping: every field of every trace record is a closed-form function of (layer, neuron, input)
// except the running sums, which are a per-neuron prefix scan.  One CTA per neuron; the records land exactly where the CPU evaluator's
// single pass would put them, labels and access counters included:
//   labels   : 1.. weights (layer, neuron, input order) | inputs | zero | then gates in creation order (1 per first product, 2 per later input)
//   neuron j : mul(w_0, x_0) ; then per k >= 1: mul(w_k, x_k), add(sum, product), delete(product), delete(old sum)
//   access   : a weight is read once; input k of a layer is read by every neuron (access_r = a0 + j, a0 = 0 for the network inputs, 1 for
//              hidden values, which are born with access 1); sums and products are read once (deleted with access 2)
struct MlpLayer {
    int n, m, a0;                 // inputs, neurons, initial access counter of the inputs
    int layer;                    // i (the weight value is (j + i + k) % 256)
    long long w_label0;           // label of weight (i, 0, 0)
    long long g_label0;           // label of the first gate of this layer
    size_t rec0;                  // first record of this layer
};
__global__ void __launch_bounds__(256)
mlp_layer_kernel(TrTuple *__restrict__ tr, MlpLayer L, const F *__restrict__ in_val, const int *__restrict__ in_idx, F *__restrict__ out_val, int *__restrict__ out_idx) {
    __shared__ F sc[256];
    __shared__ F carry;
    const int j = blockIdx.x, n = L.n;
    const size_t per = 1 + 4 * (size_t)(n - 1);
    TrTuple *base = tr + L.rec0 + (size_t)j * per;
    const long long gl = L.g_label0 + (long long)j * (2 * n - 1);
    if (threadIdx.x == 0) carry = mkF(0, 0);
    __syncthreads();
    for (int k0 = 0; k0 < n; k0 += blockDim.x) {
        const int k = k0 + threadIdx.x;
        F w = mkF(0, 0), x = mkF(0, 0), p = mkF(0, 0);
        if (k < n) { w = mkF((u64)((j + L.layer + k) % 256), 0); x = in_val[k]; p = fmul(w, x); }
        sc[threadIdx.x] = p;
        __syncthreads();
        for (int d = 1; d < (int)blockDim.x; d <<= 1) {                    // inclusive scan of the products
            F v = sc[threadIdx.x];
            if ((int)threadIdx.x >= d) v = fadd(v, sc[threadIdx.x - d]);
            __syncthreads();
            sc[threadIdx.x] = v;
            __syncthreads();
        }
        const F c = carry;
        const F sum = fadd(c, sc[threadIdx.x]);                             // w_0 x_0 + .. + w_k x_k
        const F prev = (threadIdx.x == 0) ? c : fadd(c, sc[threadIdx.x - 1]);
        if (k < n) {
            const int widx = (int)(L.w_label0 + (long long)j * n + k);
            TrTuple t;
            t.type = 2; t.value_l = w; t.value_r = x; t.value_o = p;
            t.idx_l = widx; t.idx_r = in_idx[k]; t.access_l = 0; t.access_r = L.a0 + j; t.access_o = 0;
            for (int q = 0; q < 7; q++) t.pad_[q] = 0;
            if (k == 0) { t.idx_o = (int)gl; base[0] = t; }
            else {
                const int ml = (int)(gl + 1 + 2 * (long long)(k - 1)), hl = (k == 1) ? (int)gl : ml - 1;     // product label, label of the old sum
                TrTuple *r = base + 1 + 4 * (size_t)(k - 1);
                t.idx_o = ml; r[0] = t;
                TrTuple a = t;
                a.type = 1; a.value_l = prev; a.value_r = p; a.value_o = sum; a.idx_l = hl; a.idx_r = ml; a.idx_o = ml + 1; a.access_l = 1; a.access_r = 1; a.access_o = 0;
                r[1] = a;
                TrTuple d1 = a; d1.type = 0; d1.idx_o = ml; d1.value_o = p; d1.access_o = 2; r[2] = d1;
                TrTuple d2 = a; d2.type = 0; d2.idx_o = hl; d2.value_o = prev; d2.access_o = 2; r[3] = d2;
            }
            if (k == n - 1) { out_val[j] = sum; out_idx[j] = (n == 1) ? (int)gl : (int)(gl + 2 * (long long)n - 2); }
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = sum;
        __syncthreads();
    }
}
// delete records: idx/value from arrays (or the weight formula), constant access counter
__global__ void __launch_bounds__(256)
mlp_delete_kernel(TrTuple *__restrict__ tr, size_t n, const F *__restrict__ val, const int *__restrict__ idx, int access) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    TrTuple t; memset(&t, 0, sizeof t);
    t.type = 0; t.idx_o = idx[i]; t.value_o = val[i]; t.access_o = access;
    tr[i] = t;
}
__global__ void __launch_bounds__(256) mlp_delete_weights_kernel(TrTuple *__restrict__ tr, int layer, int m, int n, long long w_label0) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)m * n) return;
    const int j = (int)(i / n), k = (int)(i % n);
    TrTuple t; memset(&t, 0, sizeof t);
    t.type = 0; t.idx_o = (int)(w_label0 + (long long)i); t.value_o = mkF((u64)((j + layer + k) % 256), 0); t.access_o = 1;
    tr[i] = t;
}
__global__ void __launch_bounds__(256) mlp_inputs_kernel(F *__restrict__ val, int *__restrict__ idx, int n, long long label0) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { val[k] = mkF((u64)((k + 1) % 256), 0); idx[k] = (int)(label0 + k); }
}

__global__ void __launch_bounds__(256) aes_block_kernel(TrTuple *__restrict__ tr, int n) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < (size_t)n * kAesRecs) tr[gid] = aes_record(n, (int)(gid / kAesRecs), (int)(gid % kAesRecs));
}
__global__ void __launch_bounds__(256) aes_tail_kernel(TrTuple *__restrict__ tr, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= 16 * n + 160) tr[i] = aes_tail_record(n, i);
}

// ---- 8f.4: the pruned two-layer MLP (Seval.cpp:1170-1236 `inference`, fun == 8 driver :1424-1461) ------------------------------------------
// Same per-neuron program as the dense MLP over a CSR pattern (the reference's indexes[l][i], host data drawn from libc rand()); what is
// new is that the access counter of an input is the number of EARLIER reads of it in evaluation order — a stable rank of the position among
// equal column indices (radix sort + run starts, as for the lookup counters) — and that a neuron without inputs is a copy of the gate `zero`.
struct PrunedLayer {
    int m;                               // neurons
    const int *rowptr, *cols;            // CSR pattern (device)
    const long long *recoff, *laboff;    // exclusive prefix sums over neurons of the records / labels they emit
    const unsigned *rank;                // per CSR position: earlier reads of the same input
    long long w_label0, g_label0;        // label of weight (0, 0) of this layer, of the layer's first gate
    size_t rec0;                         // first record of this layer
    int zero_label;
};
__global__ void __launch_bounds__(256)
pruned_layer_kernel(TrTuple *__restrict__ tr, PrunedLayer L, const F *__restrict__ in_val, const int *__restrict__ in_idx, const int *__restrict__ in_acc0,
                    F *__restrict__ out_val, int *__restrict__ out_idx, int *__restrict__ out_acc) {
    __shared__ F sc[256];
    __shared__ F carry;
    const int i = blockIdx.x, r0 = L.rowptr[i], f = L.rowptr[i + 1] - r0;
    if (f == 0) {                                                              // hidden_layer[i] = zero: a copy (label of zero, access 0)
        if (threadIdx.x == 0) { out_val[i] = mkF(0, 0); out_idx[i] = L.zero_label; out_acc[i] = 0; }
        return;
    }
    TrTuple *base = tr + L.rec0 + L.recoff[i];
    const long long gl = L.g_label0 + L.laboff[i];
    if (threadIdx.x == 0) carry = mkF(0, 0);
    __syncthreads();
    for (int k0 = 0; k0 < f; k0 += blockDim.x) {
        const int k = k0 + threadIdx.x;
        F w = mkF(0, 0), x = mkF(0, 0), p = mkF(0, 0);
        int col = 0;
        if (k < f) { w = mkF((u64)((k + i) % 256), 0); col = L.cols[r0 + k]; x = in_val[col]; p = fmul(w, x); }
        sc[threadIdx.x] = p;
        __syncthreads();
        for (int d = 1; d < (int)blockDim.x; d <<= 1) {                    // inclusive scan of the products
            F v = sc[threadIdx.x];
            if ((int)threadIdx.x >= d) v = fadd(v, sc[threadIdx.x - d]);
            __syncthreads();
            sc[threadIdx.x] = v;
            __syncthreads();
        }
        const F c = carry;
        const F sum = fadd(c, sc[threadIdx.x]);
        const F prev = (threadIdx.x == 0) ? c : fadd(c, sc[threadIdx.x - 1]);
        if (k < f) {
            TrTuple t;
            t.type = 2; t.value_l = w; t.value_r = x; t.value_o = p;
            t.idx_l = (int)(L.w_label0 + r0 + k); t.idx_r = in_idx[col]; t.access_l = 0; t.access_r = in_acc0[col] + (int)L.rank[r0 + k]; t.access_o = 0;
            for (int q = 0; q < 7; q++) t.pad_[q] = 0;
            if (k == 0) { t.idx_o = (int)gl; base[0] = t; }
            else {
                const int ml = (int)(gl + 1 + 2 * (long long)(k - 1)), hl = (k == 1) ? (int)gl : ml - 1;
                TrTuple *r = base + 1 + 4 * (size_t)(k - 1);
                t.idx_o = ml; r[0] = t;
                TrTuple a = t;
                a.type = 1; a.value_l = prev; a.value_r = p; a.value_o = sum; a.idx_l = hl; a.idx_r = ml; a.idx_o = ml + 1; a.access_l = 1; a.access_r = 1; a.access_o = 0;
                r[1] = a;
                TrTuple d1 = a; d1.type = 0; d1.idx_o = ml; d1.value_o = p; d1.access_o = 2; r[2] = d1;
                TrTuple d2 = a; d2.type = 0; d2.idx_o = hl; d2.value_o = prev; d2.access_o = 2; r[3] = d2;
            }
            if (k == f - 1) { out_val[i] = sum; out_idx[i] = (f == 1) ? (int)gl : (int)(gl + 2 * (long long)f - 2); out_acc[i] = 1; }
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = sum;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) pruned_keys_kernel(const int *__restrict__ cols, size_t n, unsigned long long *__restrict__ keys, unsigned *__restrict__ idx, unsigned *__restrict__ count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = (unsigned long long)cols[i]; idx[i] = (unsigned)i;
    atomicAdd(&count[cols[i]], 1u);
}
// delete records of a layer's weights: one CTA per neuron
__global__ void __launch_bounds__(256) pruned_delete_weights_kernel(TrTuple *__restrict__ tr, const int *__restrict__ rowptr, long long w_label0) {
    const int i = blockIdx.x, r0 = rowptr[i], f = rowptr[i + 1] - r0;
    for (int k = threadIdx.x; k < f; k += blockDim.x) {
        TrTuple t; memset(&t, 0, sizeof t);
        t.type = 0; t.idx_o = (int)(w_label0 + r0 + k); t.value_o = mkF((u64)((k + i) % 256), 0); t.access_o = 1;
        tr[r0 + k] = t;
    }
}
// delete records of inputs / hidden values / outputs: access = initial counter + number of reads
__global__ void __launch_bounds__(256)
pruned_delete_kernel(TrTuple *__restrict__ tr, size_t n, const F *__restrict__ val, const int *__restrict__ idx, const int *__restrict__ acc0, const unsigned *__restrict__ reads) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    TrTuple t; memset(&t, 0, sizeof t);
    t.type = 0; t.idx_o = idx[i]; t.value_o = val[i]; t.access_o = (acc0 ? acc0[i] : 0) + (reads ? (int)reads[i] : 0);
    tr[i] = t;
}

__global__ void __launch_bounds__(256) sql_kernel(TrTuple *__restrict__ tr, int n, size_t total) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < total) tr[gid] = sql_record(n, gid);
}

}  // namespace hb

using namespace hb;

// MLP_inference on the GPU: fills the context's resident trace as if the producer had been drained (hb_trace_begin/push), one pass.
extern "C" int hb_trace_generate_mlp(hb_ctx *ctx, const int *layer_size, int nsizes, size_t *n_records) { HB_DEV(ctx);
    if (nsizes < 2) HB_FAIL(ctx, "hb_trace_generate_mlp: need at least an input and an output layer");
    size_t recs = 0, wt = 0; int maxw = 0;
    for (int i = 0; i + 1 < nsizes; i++) {
        const size_t n = (size_t)layer_size[i], m = (size_t)layer_size[i + 1];
        if (n < 1 || m < 1) HB_FAIL(ctx, "hb_trace_generate_mlp: layer sizes must be positive");
        recs += m * (1 + 4 * (n - 1)) + n; wt += m * n;
    }
    for (int i = 0; i < nsizes; i++) maxw = std::max(maxw, layer_size[i]);
    recs += wt + (size_t)layer_size[nsizes - 1] + 1;                               // weight deletes, output deletes, zero
    if (wt + recs >= ((size_t)1 << 31)) HB_FAIL(ctx, "hb_trace_generate_mlp: labels do not fit the reference's int");
    HB_TRY(hb_trace_begin(ctx, recs));
    TraceState &t = ctx->trace;
    TrTuple *tr = (TrTuple *)t.tuples;
    F *val; HB_CHECK(ctx, cudaMallocAsync(&val, 2 * (size_t)maxw * (sizeof(F) + sizeof(int)), ctx->stream));
    F *va = val, *vb = val + maxw; int *ia = (int *)(vb + maxw), *ib = ia + maxw;
    const long long in_label0 = (long long)wt + 1, zero_label = in_label0 + layer_size[0];
    long long g_label = zero_label + 1, w_label = 1;
    HB_LAUNCH(ctx, mlp_inputs_kernel, (unsigned)((layer_size[0] + 255) / 256), 256, 0, va, ia, layer_size[0], in_label0);
    size_t rec = 0;
    for (int i = 0; i + 1 < nsizes; i++) {
        MlpLayer L; L.n = layer_size[i]; L.m = layer_size[i + 1]; L.a0 = i ? 1 : 0; L.layer = i; L.w_label0 = w_label; L.g_label0 = g_label; L.rec0 = rec;
        HB_LAUNCH(ctx, mlp_layer_kernel, (unsigned)L.m, 256, 0, tr, L, va, ia, vb, ib);
        rec += (size_t)L.m * (1 + 4 * (size_t)(L.n - 1));
        HB_LAUNCH(ctx, mlp_delete_kernel, (unsigned)((L.n + 255) / 256), 256, 0, tr + rec, (size_t)L.n, va, ia, L.a0 + L.m);     // the layer's inputs
        rec += (size_t)L.n;
        w_label += (long long)L.m * L.n; g_label += (long long)L.m * (2 * (long long)L.n - 1);
        std::swap(va, vb); std::swap(ia, ib);
    }
    w_label = 1;
    for (int i = 0; i + 1 < nsizes; i++) {
        const int n = layer_size[i], m = layer_size[i + 1];
        HB_LAUNCH(ctx, mlp_delete_weights_kernel, (unsigned)(((size_t)m * n + 255) / 256), 256, 0, tr + rec, i, m, n, w_label);
        rec += (size_t)m * n; w_label += (long long)m * n;
    }
    const int nout = layer_size[nsizes - 1];
    HB_LAUNCH(ctx, mlp_delete_kernel, (unsigned)((nout + 255) / 256), 256, 0, tr + rec, (size_t)nout, va, ia, 1);                 // the outputs (never read)
    rec += (size_t)nout;
    TrTuple z; memset(&z, 0, sizeof z); z.type = 0; z.idx_o = (int)zero_label; z.access_o = 0;
    HB_CHECK(ctx, cudaMemcpyAsync(tr + rec, &z, sizeof z, cudaMemcpyHostToDevice, ctx->stream));
    rec += 1;
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(val, ctx->stream);
    if (rec != recs) HB_FAIL(ctx, "hb_trace_generate_mlp: record count mismatch");
    t.n = rec; t.done = true;
    if (n_records) *n_records = rec;
    return 0;
}

// stable rank of every position among equal keys + reads per key (device arrays; rank and count are outputs)
static int rank_positions(hb_ctx *ctx, const int *cols, size_t n, size_t nkeys, unsigned *rank, unsigned *count) {
    HB_CHECK(ctx, cudaMemsetAsync(count, 0, nkeys * sizeof(unsigned), ctx->stream));
    if (n == 0) return 0;
    char *buf; HB_CHECK(ctx, cudaMallocAsync(&buf, n * (2 * 8 + 3 * 4), ctx->stream));
    unsigned long long *keys = (unsigned long long *)buf, *keys_s = keys + n;
    unsigned *idx = (unsigned *)(keys_s + n), *idx_s = idx + n, *start = idx_s + n;
    const unsigned g = (unsigned)((n + 255) / 256);
    HB_LAUNCH(ctx, pruned_keys_kernel, g, 256, 0, cols, n, keys, idx, count);
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, keys, keys_s, idx, idx_s, (int)n, 0, 32, ctx->stream);
    cub::DeviceScan::InclusiveScan(nullptr, b2, start, start, MaxOp(), (int)n, ctx->stream);
    void *tmp; HB_CHECK(ctx, cudaMallocAsync(&tmp, std::max(b1, b2), ctx->stream));
    HB_CHECK(ctx, cub::DeviceRadixSort::SortPairs(tmp, b1, keys, keys_s, idx, idx_s, (int)n, 0, 32, ctx->stream));         // stable: evaluation order inside a run
    HB_LAUNCH(ctx, run_start_kernel, g, 256, 0, keys_s, n, start);
    HB_CHECK(ctx, cub::DeviceScan::InclusiveScan(tmp, b2, start, start, MaxOp(), (int)n, ctx->stream));
    HB_LAUNCH(ctx, access_scatter_kernel, g, 256, 0, idx_s, start, n, rank);
    ctx->launches += 2;
    cudaFreeAsync(tmp, ctx->stream); cudaFreeAsync(buf, ctx->stream);
    return 0;
}

// the pruned MLP on the GPU (fun == 8): rowptr / cols are HOST arrays (the reference driver's sparsity pattern)
extern "C" int hb_trace_generate_pruned_mlp(hb_ctx *ctx, int n_inputs, int n_hidden, int n_out, const int *rowptr0, const int *cols0, const int *rowptr1,
                                            const int *cols1, size_t *n_records) { HB_DEV(ctx);
    if (n_inputs < 1 || n_hidden < 1 || n_out < 1) HB_FAIL(ctx, "hb_trace_generate_pruned_mlp: layer sizes must be positive");
    const int m[2] = {n_hidden, n_out}, nin[2] = {n_inputs, n_hidden};
    const int *rp[2] = {rowptr0, rowptr1}, *cl[2] = {cols0, cols1};
    size_t W[2], R[2], Lb[2];
    std::vector<long long> recoff[2], laboff[2];
    for (int l = 0; l < 2; l++) {
        if (rp[l][0] != 0) HB_FAIL(ctx, "hb_trace_generate_pruned_mlp: rowptr must start at 0");
        recoff[l].assign(m[l] + 1, 0); laboff[l].assign(m[l] + 1, 0);
        for (int i = 0; i < m[l]; i++) {
            const long long f = (long long)rp[l][i + 1] - rp[l][i];
            if (f < 0) HB_FAIL(ctx, "hb_trace_generate_pruned_mlp: rowptr must be non-decreasing");
            recoff[l][i + 1] = recoff[l][i] + (f ? 1 + 4 * (f - 1) : 0);
            laboff[l][i + 1] = laboff[l][i] + (f ? 2 * f - 1 : 0);
        }
        W[l] = (size_t)rp[l][m[l]]; R[l] = (size_t)recoff[l][m[l]]; Lb[l] = (size_t)laboff[l][m[l]];
        for (size_t p = 0; p < W[l]; p++) if (cl[l][p] < 0 || cl[l][p] >= nin[l]) HB_FAIL(ctx, "hb_trace_generate_pruned_mlp: column index out of range");
    }
    const size_t recs = R[0] + R[1] + W[0] + W[1] + (size_t)n_inputs + n_hidden + n_out + 1;
    const size_t labels = (size_t)n_inputs + W[0] + W[1] + 1 + Lb[0] + Lb[1];
    if (labels >= ((size_t)1 << 31) || recs >= ((size_t)1 << 31)) HB_FAIL(ctx, "hb_trace_generate_pruned_mlp: labels do not fit the reference's int");
    HB_TRY(hb_trace_begin(ctx, recs));
    TraceState &t = ctx->trace;
    TrTuple *tr = (TrTuple *)t.tuples;
    // device copies of the pattern and the prefix sums
    int *d_rp[2], *d_cl[2]; long long *d_ro[2], *d_lo[2]; unsigned *d_rank[2], *d_cnt[2];
    for (int l = 0; l < 2; l++) {
        HB_CHECK(ctx, cudaMallocAsync(&d_rp[l], (m[l] + 1) * sizeof(int), ctx->stream));
        HB_CHECK(ctx, cudaMallocAsync(&d_cl[l], std::max<size_t>(W[l], 1) * sizeof(int), ctx->stream));
        HB_CHECK(ctx, cudaMallocAsync(&d_ro[l], (m[l] + 1) * sizeof(long long), ctx->stream));
        HB_CHECK(ctx, cudaMallocAsync(&d_lo[l], (m[l] + 1) * sizeof(long long), ctx->stream));
        HB_CHECK(ctx, cudaMallocAsync(&d_rank[l], std::max<size_t>(W[l], 1) * sizeof(unsigned), ctx->stream));
        HB_CHECK(ctx, cudaMallocAsync(&d_cnt[l], nin[l] * sizeof(unsigned), ctx->stream));
        HB_CHECK(ctx, cudaMemcpyAsync(d_rp[l], rp[l], (m[l] + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        if (W[l]) HB_TRY(copy_from_host(ctx, d_cl[l], cl[l], W[l] * sizeof(int), ctx->stream));
        HB_CHECK(ctx, cudaMemcpyAsync(d_ro[l], recoff[l].data(), (m[l] + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        HB_CHECK(ctx, cudaMemcpyAsync(d_lo[l], laboff[l].data(), (m[l] + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        HB_TRY(rank_positions(ctx, d_cl[l], W[l], (size_t)nin[l], d_rank[l], d_cnt[l]));
    }
    // values, labels and initial access counters of the inputs / hidden values / outputs
    const size_t nv = (size_t)n_inputs + n_hidden + n_out;
    F *val; int *idx, *acc;
    HB_CHECK(ctx, cudaMallocAsync(&val, nv * sizeof(F), ctx->stream));
    HB_CHECK(ctx, cudaMallocAsync(&idx, 2 * nv * sizeof(int), ctx->stream));
    acc = idx + nv;
    HB_CHECK(ctx, cudaMemsetAsync(acc, 0, nv * sizeof(int), ctx->stream));
    F *v_in = val, *v_h = val + n_inputs, *v_o = v_h + n_hidden;
    int *i_in = idx, *i_h = idx + n_inputs, *i_o = i_h + n_hidden, *a_in = acc, *a_h = acc + n_inputs, *a_o = a_h + n_hidden;
    HB_LAUNCH(ctx, mlp_inputs_kernel, (unsigned)((n_inputs + 255) / 256), 256, 0, v_in, i_in, n_inputs, (long long)1);
    const long long w_label[2] = {(long long)n_inputs + 1, (long long)n_inputs + 1 + (long long)W[0]};
    const long long zero_label = w_label[1] + (long long)W[1];
    PrunedLayer L0{m[0], d_rp[0], d_cl[0], d_ro[0], d_lo[0], d_rank[0], w_label[0], zero_label + 1, 0, (int)zero_label};
    HB_LAUNCH(ctx, pruned_layer_kernel, (unsigned)m[0], 256, 0, tr, L0, v_in, i_in, a_in, v_h, i_h, a_h);
    PrunedLayer L1{m[1], d_rp[1], d_cl[1], d_ro[1], d_lo[1], d_rank[1], w_label[1], zero_label + 1 + (long long)Lb[0], R[0], (int)zero_label};
    HB_LAUNCH(ctx, pruned_layer_kernel, (unsigned)m[1], 256, 0, tr, L1, v_h, i_h, a_h, v_o, i_o, a_o);
    size_t rec = R[0] + R[1];
    for (int l = 0; l < 2; l++) { HB_LAUNCH(ctx, pruned_delete_weights_kernel, (unsigned)m[l], 256, 0, tr + rec, d_rp[l], w_label[l]); rec += W[l]; }
    HB_LAUNCH(ctx, pruned_delete_kernel, (unsigned)((n_inputs + 255) / 256), 256, 0, tr + rec, (size_t)n_inputs, v_in, i_in, a_in, d_cnt[0]); rec += n_inputs;
    HB_LAUNCH(ctx, pruned_delete_kernel, (unsigned)((n_hidden + 255) / 256), 256, 0, tr + rec, (size_t)n_hidden, v_h, i_h, a_h, d_cnt[1]); rec += n_hidden;
    HB_LAUNCH(ctx, pruned_delete_kernel, (unsigned)((n_out + 255) / 256), 256, 0, tr + rec, (size_t)n_out, v_o, i_o, a_o, (const unsigned *)nullptr); rec += n_out;
    TrTuple z; memset(&z, 0, sizeof z); z.type = 0; z.idx_o = (int)zero_label; z.access_o = 0;
    HB_CHECK(ctx, cudaMemcpyAsync(tr + rec, &z, sizeof z, cudaMemcpyHostToDevice, ctx->stream));
    rec += 1;
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int l = 0; l < 2; l++) { cudaFreeAsync(d_rp[l], ctx->stream); cudaFreeAsync(d_cl[l], ctx->stream); cudaFreeAsync(d_ro[l], ctx->stream); cudaFreeAsync(d_lo[l], ctx->stream); cudaFreeAsync(d_rank[l], ctx->stream); cudaFreeAsync(d_cnt[l], ctx->stream); }
    cudaFreeAsync(val, ctx->stream); cudaFreeAsync(idx, ctx->stream);
    if (rec != recs) HB_FAIL(ctx, "hb_trace_generate_pruned_mlp: record count mismatch");
    t.n = rec; t.done = true;
    if (n_records) *n_records = rec;
    return 0;
}

// the SQL range query on the GPU (fun == 6: `pigeon 6 b n d`, input_size = 2^n rows)
extern "C" int hb_trace_generate_sql(hb_ctx *ctx, int input_size, size_t *n_records) { HB_DEV(ctx);
    if (input_size < 1) HB_FAIL(ctx, "hb_trace_generate_sql: need at least one row");
    const size_t recs = sql_records(input_size);
    if ((size_t)input_size * (kSqlRowLabels + 1) + 300 >= ((size_t)1 << 31)) HB_FAIL(ctx, "hb_trace_generate_sql: labels do not fit the reference's int");
    HB_TRY(hb_trace_begin(ctx, recs));
    TraceState &t = ctx->trace;
    HB_LAUNCH(ctx, sql_kernel, (unsigned)((recs + 255) / 256), 256, 0, (TrTuple *)t.tuples, input_size, recs);
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    t.n = recs; t.done = true;
    if (n_records) *n_records = recs;
    return 0;
}

// AES on the GPU (fun == 5: `pigeon 5 b n d`, input_size = 2^n blocks): the resident trace of one pass of the CPU evaluator
extern "C" int hb_trace_generate_aes(hb_ctx *ctx, int input_size, size_t *n_records) { HB_DEV(ctx);
    if (input_size < 1) HB_FAIL(ctx, "hb_trace_generate_aes: need at least one block");
    const size_t n = (size_t)input_size, recs = n * kAesRecs + 16 * n + 161;
    if (n * (16 + kAesLabels) + 162 >= ((size_t)1 << 31) || n * kAesLookups >= ((size_t)1 << 31)) HB_FAIL(ctx, "hb_trace_generate_aes: labels do not fit the reference's int");
    HB_TRY(hb_trace_begin(ctx, recs));
    TraceState &t = ctx->trace;
    TrTuple *tr = (TrTuple *)t.tuples;
    HB_LAUNCH(ctx, aes_block_kernel, (unsigned)((n * kAesRecs + 255) / 256), 256, 0, tr, input_size);
    HB_LAUNCH(ctx, aes_tail_kernel, (unsigned)((16 * n + 161 + 255) / 256), 256, 0, tr + n * kAesRecs, input_size);
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    t.n = recs; t.done = true;
    if (n_records) *n_records = recs;
    return 0;
}

extern "C" int hb_trace_begin(hb_ctx *ctx, size_t capacity) { HB_DEV(ctx);
    TraceState &t = ctx->trace;
    if (t.tuples && t.capacity < capacity) { cudaFree(t.tuples); t.tuples = nullptr; }
    if (!t.tuples) { HB_CHECK(ctx, cudaMalloc(&t.tuples, std::max<size_t>(capacity, 1) * 80)); t.capacity = std::max<size_t>(capacity, 1); }
    t.n = 0; t.n_ops = t.n_del = 0; t.done = false; t.indexed = false;
    return 0;
}

extern "C" int hb_trace_push(hb_ctx *ctx, const void *tuples, size_t n, int *done) { HB_DEV(ctx);
    TraceState &t = ctx->trace;
    if (!t.tuples) HB_FAIL(ctx, "hb_trace_push: call hb_trace_begin first");
    if (done) *done = t.done ? 1 : 0;
    if (t.done || n == 0) return 0;
    if (is_device_ptr(tuples)) HB_FAIL(ctx, "hb_trace_push: the producer's buffer is host memory");
    const TrTuple *src = reinterpret_cast<const TrTuple *>(tuples);
    size_t take = n;
    for (size_t i = 0; i < n; i++) if (src[i].type == 255) { take = i; t.done = true; break; }      // end-of-circuit marker (Seval.cpp:1273-1283)
    if (t.n + take > t.capacity) {                                                                   // grow geometrically
        size_t cap = std::max(t.capacity * 2, t.n + take);
        void *p; HB_CHECK(ctx, cudaMalloc(&p, cap * 80));
        HB_CHECK(ctx, cudaMemcpyAsync(p, t.tuples, t.n * 80, cudaMemcpyDeviceToDevice, ctx->stream));
        HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(t.tuples); t.tuples = p; t.capacity = cap;
    }
    // the producer reuses its buffer as soon as we return: the copy must have left the host buffer before that
    HB_CHECK(ctx, cudaMemcpyAsync((char *)t.tuples + t.n * 80, src, take * 80, cudaMemcpyHostToDevice, ctx->stream));
    HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    t.n += take;
    if (done) *done = t.done ? 1 : 0;
    return 0;
}

static int trace_index(hb_ctx *ctx) {
    TraceState &t = ctx->trace;
    if (t.indexed) return 0;
    const size_t n = std::max<size_t>(t.n, 1);
    if (t.pos_capacity < 4 * n) {
        if (t.pos) cudaFree(t.pos);
        t.pos = nullptr; t.pos_capacity = 0;
        HB_CHECK(ctx, cudaMalloc(&t.pos, 4 * n * sizeof(unsigned)));
        t.pos_capacity = 4 * n;
    }
    unsigned *is_op = t.pos, *is_del = t.pos + n, *op_pos = t.pos + 2 * n, *del_pos = t.pos + 3 * n;
    if (t.n) {
        HB_LAUNCH(ctx, trace_flags_kernel, (unsigned)((t.n + 255) / 256), 256, 0, (const TrTuple *)t.tuples, t.n, is_op, is_del);
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, is_op, op_pos, (int)t.n, ctx->stream);
        void *tmp; HB_CHECK(ctx, cudaMallocAsync(&tmp, tmp_bytes, ctx->stream));
        HB_CHECK(ctx, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, is_op, op_pos, (int)t.n, ctx->stream));
        HB_CHECK(ctx, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, is_del, del_pos, (int)t.n, ctx->stream));
        ctx->launches += 2;
        cudaFreeAsync(tmp, ctx->stream);
        unsigned last[4];
        HB_CHECK(ctx, cudaMemcpyAsync(&last[0], is_op + t.n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
        HB_CHECK(ctx, cudaMemcpyAsync(&last[1], op_pos + t.n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
        HB_CHECK(ctx, cudaMemcpyAsync(&last[2], is_del + t.n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
        HB_CHECK(ctx, cudaMemcpyAsync(&last[3], del_pos + t.n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
        HB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        t.n_ops = (size_t)last[0] + last[1]; t.n_del = (size_t)last[2] + last[3];
    }
    t.indexed = true;
    return 0;
}

extern "C" int hb_trace_finish(hb_ctx *ctx, size_t *n_tuples, size_t *n_ops, size_t *n_deletes) { HB_DEV(ctx);
    if (!ctx->trace.tuples) HB_FAIL(ctx, "hb_trace_finish: no trace");
    HB_TRY(trace_index(ctx));
    if (n_tuples) *n_tuples = ctx->trace.n;
    if (n_ops) *n_ops = ctx->trace.n_ops;
    if (n_deletes) *n_deletes = ctx->trace.n_del;
    return 0;
}

static int trace_check(hb_ctx *ctx, size_t cs, const char *who) {
    HB_TRY(trace_index(ctx));
    if (cs == 0 || (cs & (cs - 1))) HB_FAIL(ctx, std::string(who) + ": circuit_size must be a power of two");
    if (ctx->trace.n_ops > cs || ctx->trace.n_del > cs) HB_FAIL(ctx, std::string(who) + ": the trace has more op / delete tuples than circuit_size");
    return 0;
}

extern "C" int hb_trace_witness(hb_ctx *ctx, size_t cs, hb_F *out) { HB_DEV(ctx);
    HB_TRY(trace_check(ctx, cs, "hb_trace_witness"));
    TraceState &t = ctx->trace;
    Staged so(ctx);
    HB_TRY(so.outbuf(out, 4 * cs * sizeof(F)));
    HB_CHECK(ctx, cudaMemsetAsync(so.dev, 0, 4 * cs * sizeof(F), ctx->stream));
    if (t.n) HB_LAUNCH(ctx, trace_witness_kernel, (unsigned)((t.n + 255) / 256), 256, 0, (const TrTuple *)t.tuples, t.n, t.pos + 2 * t.n, t.pos + 3 * t.n, cs, so.as<F>());
    HB_TRY(so.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_trace_circuit(hb_ctx *ctx, size_t cs, hb_F *out) { HB_DEV(ctx);
    HB_TRY(trace_check(ctx, cs, "hb_trace_circuit"));
    TraceState &t = ctx->trace;
    Staged so(ctx);
    HB_TRY(so.outbuf(out, 16 * cs * sizeof(F)));
    HB_CHECK(ctx, cudaMemsetAsync(so.dev, 0, 16 * cs * sizeof(F), ctx->stream));
    if (t.n) HB_LAUNCH(ctx, trace_circuit_kernel, (unsigned)((t.n + 255) / 256), 256, 0, (const TrTuple *)t.tuples, t.n, t.pos + 2 * t.n, t.pos + 3 * t.n, cs, so.as<F>());
    HB_TRY(so.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_trace_transcript(hb_ctx *ctx, size_t cs, int has_lookups, hb_F *L, hb_F *R, hb_F *O, hb_F *S) { HB_DEV(ctx);
    HB_TRY(trace_check(ctx, cs, "hb_trace_transcript"));
    TraceState &t = ctx->trace;
    Staged sl(ctx), sr(ctx), so(ctx), ss(ctx);
    HB_TRY(sl.outbuf(L, cs * sizeof(F))); HB_TRY(sr.outbuf(R, cs * sizeof(F))); HB_TRY(so.outbuf(O, cs * sizeof(F))); HB_TRY(ss.outbuf(S, cs * sizeof(F)));
    for (Staged *s : {&sl, &sr, &so, &ss}) HB_CHECK(ctx, cudaMemsetAsync(s->dev, 0, cs * sizeof(F), ctx->stream));
    if (t.n) HB_LAUNCH(ctx, trace_transcript_kernel, (unsigned)((t.n + 255) / 256), 256, 0, (const TrTuple *)t.tuples, t.n, t.pos + 2 * t.n, has_lookups,
                       sl.as<F>(), sr.as<F>(), so.as<F>(), ss.as<F>());
    for (Staged *s : {&sl, &sr, &so, &ss}) HB_TRY(s->finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_trace_wiring(hb_ctx *ctx, size_t cs, const hb_F *a_w, const hb_F *b_w, hb_F *xy) { HB_DEV(ctx);
    HB_TRY(trace_check(ctx, cs, "hb_trace_wiring"));
    TraceState &t = ctx->trace;
    Staged so(ctx);
    HB_TRY(so.outbuf(xy, 8 * cs * sizeof(F)));
    F *X = so.as<F>(), *Y = X + 4 * cs;
    HB_LAUNCH(ctx, fill_kernel, (unsigned)std::min<size_t>((8 * cs + 255) / 256, (size_t)ctx->sm_count * 16), 256, 0, X, 8 * cs, mkF(1, 0));
    if (t.n) HB_LAUNCH(ctx, trace_wiring_kernel, (unsigned)((t.n + 255) / 256), 256, 0, (const TrTuple *)t.tuples, t.n, t.pos + 2 * t.n, t.pos + 3 * t.n, cs,
                       mkF(a_w->real, a_w->img), mkF(b_w->real, b_w->img), X, Y);
    HB_TRY(so.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

// access counters + rank among lookup records; caller frees *scratch with cudaFreeAsync
static int lookup_index(hb_ctx *ctx, unsigned **access, unsigned **lkp_pos, void **scratch) {
    TraceState &t = ctx->trace;
    const size_t n = std::max<size_t>(t.n, 1);
    // layout: keys[n] keys_sorted[n] (u64) | idx[n] idx_sorted[n] start[n] access[n] is_lkp[n] lkp_pos[n] (u32)
    char *buf; HB_CHECK(ctx, cudaMallocAsync(&buf, n * (2 * 8 + 6 * 4), ctx->stream));
    unsigned long long *keys = (unsigned long long *)buf, *keys_s = keys + n;
    unsigned *idx = (unsigned *)(keys_s + n), *idx_s = idx + n, *start = idx_s + n, *acc = start + n, *is_lkp = acc + n, *lpos = is_lkp + n;
    *scratch = buf; *access = acc; *lkp_pos = lpos;
    if (!t.n) return 0;
    const unsigned g = (unsigned)((t.n + 255) / 256);
    HB_LAUNCH(ctx, lookup_keys_kernel, g, 256, 0, (const TrTuple *)t.tuples, t.n, keys, idx, is_lkp);
    size_t b1 = 0, b2 = 0, b3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, keys, keys_s, idx, idx_s, (int)t.n, 0, 48, ctx->stream);
    cub::DeviceScan::InclusiveScan(nullptr, b2, start, start, MaxOp(), (int)t.n, ctx->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, b3, is_lkp, lpos, (int)t.n, ctx->stream);
    void *tmp; HB_CHECK(ctx, cudaMallocAsync(&tmp, std::max(b1, std::max(b2, b3)), ctx->stream));
    HB_CHECK(ctx, cub::DeviceRadixSort::SortPairs(tmp, b1, keys, keys_s, idx, idx_s, (int)t.n, 0, 48, ctx->stream));      // stable: trace order inside a run
    HB_LAUNCH(ctx, run_start_kernel, g, 256, 0, keys_s, t.n, start);
    HB_CHECK(ctx, cub::DeviceScan::InclusiveScan(tmp, b2, start, start, MaxOp(), (int)t.n, ctx->stream));
    HB_LAUNCH(ctx, access_scatter_kernel, g, 256, 0, idx_s, start, t.n, acc);
    HB_CHECK(ctx, cub::DeviceScan::ExclusiveSum(tmp, b3, is_lkp, lpos, (int)t.n, ctx->stream));
    ctx->launches += 3;
    cudaFreeAsync(tmp, ctx->stream);
    return 0;
}

extern "C" int hb_trace_lookup_basic(hb_ctx *ctx, size_t cs, const hb_F *lookup_rand4, hb_F *xy) { HB_DEV(ctx);
    HB_TRY(trace_check(ctx, cs, "hb_trace_lookup_basic"));
    TraceState &t = ctx->trace;
    Staged so(ctx);
    HB_TRY(so.outbuf(xy, 2 * cs * sizeof(F)));
    F *X = so.as<F>(), *Y = X + cs;
    HB_LAUNCH(ctx, fill_kernel, (unsigned)std::min<size_t>((2 * cs + 255) / 256, (size_t)ctx->sm_count * 16), 256, 0, X, 2 * cs, mkF(1, 0));
    unsigned *access, *lpos; void *scratch;
    HB_TRY(lookup_index(ctx, &access, &lpos, &scratch));
    const hb_F *lr = lookup_rand4;
    if (t.n) HB_LAUNCH(ctx, lookup_basic_kernel, (unsigned)((t.n + 255) / 256), 256, 0, (const TrTuple *)t.tuples, t.n, t.pos + 2 * t.n, access,
                       mkF(lr[0].real, lr[0].img), mkF(lr[1].real, lr[1].img), mkF(lr[2].real, lr[2].img), mkF(lr[3].real, lr[3].img), X, Y);
    cudaFreeAsync(scratch, ctx->stream);
    HB_TRY(so.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_trace_lookup_witness(hb_ctx *ctx, size_t cs, const hb_F *lookup_rand2, hb_F *out) { HB_DEV(ctx);
    HB_TRY(trace_check(ctx, cs, "hb_trace_lookup_witness"));
    TraceState &t = ctx->trace;
    Staged so(ctx);
    HB_TRY(so.outbuf(out, 2 * cs * sizeof(F)));
    HB_CHECK(ctx, cudaMemsetAsync(so.dev, 0, 2 * cs * sizeof(F), ctx->stream));
    unsigned *access, *lpos; void *scratch;
    HB_TRY(lookup_index(ctx, &access, &lpos, &scratch));
    const hb_F *lr = lookup_rand2;
    if (t.n) HB_LAUNCH(ctx, lookup_witness_kernel, (unsigned)((t.n + 255) / 256), 256, 0, (const TrTuple *)t.tuples, t.n, lpos, access,
                       mkF(lr[0].real, lr[0].img), mkF(lr[1].real, lr[1].img), so.as<F>());
    cudaFreeAsync(scratch, ctx->stream);
    HB_TRY(so.finish());
    HB_TRY(end_call(ctx));
    return 0;
}
