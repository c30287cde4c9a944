/* hobbit_b200 — C ABI of the B200-native backend for HOBBIT's data-parallel prover hot path.
 *
 * The reference (ChristodoulosPappas/HOBBIT-…, `/root/reference`) has no FFI layer: its "API" is a set of free
 * C++ functions over STL containers.  This header is the flat boundary those functions forward
 * to; the C++ shims with the reference's exact signatures live in hobbit_b200/host/hobbit_host.hpp and
 * INTEGRATION.md shows the binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns 0 on success, non-zero on error
 *    (hb_last_error(ctx) gives the text).  No exceptions cross the boundary.
 *  - hb_F is the reference's 16-byte POD `virgo::fieldElement {u64 real; u64 img}` (fieldElement.hpp:96-97),
 *    canonical limbs in [0, 2^61-1).  Digests are 32 raw bytes == reference `_hash` (Blake3_hash.h:3-5).
 *  - every DATA pointer may be host memory (pageable or pinned) or device memory; the library detects which
 *    (cudaPointerGetAttributes) and stages copies on the context's stream.  HOST outputs are complete when the call
 *    returns.  The building-block calls (field, NTT, encode, Merkle, eq tables, the 8f.1 and trace primitives) whose data pointers
 *    are ALL device memory are stream-ordered on the context's stream and return without synchronising it, so that a chain of
 *    them on HBM-resident tables costs launches only; use hb_sync (or hb_stream + events) before touching such results from
 *    another stream.  The provers and the commit entry points always synchronise.
 *  - a context is bound to one CUDA device and is thread-compatible (one caller at a time), like the reference's
 *    non-reentrant entry points.
 *  - there is NO CPU fallback: without a CUDA device hb_ctx_create fails.
 */
#ifndef HOBBIT_B200_H
#define HOBBIT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hb_ctx hb_ctx;
typedef struct { uint64_t real, img; } hb_F;

/* ---- context / memory ------------------------------------------------------------------------------------ */
int  hb_ctx_create(hb_ctx **out, int device);
void hb_ctx_destroy(hb_ctx *ctx);
const char *hb_last_error(hb_ctx *ctx);
int  hb_sync(hb_ctx *ctx);
/* number of kernels this context has launched so far (bench.py reports the delta as gpu_launches) */
uint64_t hb_launch_count(hb_ctx *ctx);
/* FNV-1a digest of every value the provers of this context have read back from the GPU so far (round-polynomial sums, final table
 * values): two runs with the same digest produced the same Fiat–Shamir transcript.  reset != 0 restarts it. */
uint64_t hb_transcript_digest(hb_ctx *ctx, int reset);
/* the CUDA stream (cudaStream_t) all work of this context is enqueued on — for event timing by the caller */
void *hb_stream(hb_ctx *ctx);
/* per-kernel timing: CUDA events around every launch on the context's stream.  enable(1) clears the records;
 * report writes JSON {"kernel": {"launches": n, "total_ms": t}, ...} and returns the bytes needed. */
/* measured integer-pipe roofs of this GPU, warp-instructions per SECOND chip-wide (CUDA-event timed): out3[0] IMAD.WIDE.U32 (the only
 * wide multiplier of sm_100a), out3[1] 32-bit ALU work (SHF / LOP3 / IADD), out3[2] a 1:3 mix of the two — the binding roofs of the
 * field and hash kernels, reported by bench.py next to the HBM roofline */
int  hb_ubench_pipes(hb_ctx *ctx, double *out3);
int  hb_profile_enable(hb_ctx *ctx, int on);
size_t hb_profile_report(hb_ctx *ctx, char *buf, size_t cap);
int  hb_malloc_device(hb_ctx *ctx, void **p, size_t bytes);
int  hb_free_device(hb_ctx *ctx, void *p);
/* stream-ordered scratch for tables that live between calls of one proof (the context's pool; valid for use by later hb_* calls) */
int  hb_malloc_stream(hb_ctx *ctx, void **p, size_t bytes);
int  hb_free_stream(hb_ctx *ctx, void *p);
int  hb_malloc_pinned(hb_ctx *ctx, void **p, size_t bytes);
int  hb_free_pinned(hb_ctx *ctx, void *p);
int  hb_memcpy(hb_ctx *ctx, void *dst, const void *src, size_t bytes);   /* any direction; synchronous unless device-to-device */

/* ---- F1/F2: field (reference fieldElement.cpp:34-104, 206-209) ------------------------------------------- */
/* op: 0 a+b, 1 a-b, 2 a*b, 3 -a, 4 a^-1 (b ignored for 3,4) */
int hb_field_binop(hb_ctx *ctx, int op, const hb_F *a, const hb_F *b, hb_F *c, size_t n);
/* getRootOfUnity (utils.cpp:452-463), mimc_hash (mimc.cpp:95-107): scalar, evaluated on the host */
void hb_root_of_unity(int logn, hb_F *out);
void hb_mimc_hash(const hb_F *input, const hb_F *k, hb_F *out);

/* ---- N1: NTT == _fft(arr, logn, false) (utils.cpp:605-673), batched over rows ------------------------------ */
/* `batch` rows of 2^logn elements each, row r starts at data + r*stride (elements); in place. */
int hb_ntt_batch(hb_ctx *ctx, hb_F *data, int logn, size_t batch, size_t stride);

/* ---- E1/E2: Orion/Spielman expander code (expanders.h:20-47,78-92; linear_code_encode.h:62-119) ------------ */
/* Upload the graphs the HOST generated with libc rand()/random() (RNG order stays on the host).
 * levels = recursion depth; for dep < levels: C graph (L_C[dep] x R_C[dep], degree deg_C) and D graph
 * (L_D[dep] x R_D[dep], degree deg_D); nbr_X[dep][i*deg+j] = target, w_X[dep][i*deg+j] = weight (real, < 2^32). */
int hb_expander_set(hb_ctx *ctx, long long n, int levels, int deg_C, int deg_D,
                    const long long *L_C, const long long *R_C, const uint32_t *const *nbr_C, const uint64_t *const *w_C,
                    const long long *L_D, const long long *R_D, const uint32_t *const *nbr_D, const uint64_t *const *w_D);
/* codeword length of the installed code (n + L + R), 0 if none */
long long hb_expander_codeword_len(hb_ctx *ctx);
/* encode_monolithic for `ncols` messages at once.  src: n x ncols row-major (message c = column c);
 * dst: 2n x ncols row-major; rows >= codeword length are zero (the reference's caller buffer is 2n, zero tail).
 * Message entries must be canonical field elements (both limbs < p, as every fieldElement the reference produces is): with weights
 * < 2^31 (the reference draws them with random()) the kernel sums four edge products on the 64-bit addend of IMAD.WIDE before it touches
 * the wide accumulators, which has no headroom for limbs >= 2^61.  Graphs with a weight >= 2^31 take the one-edge-at-a-time path. */
int hb_encode_batch(hb_ctx *ctx, const hb_F *src, hb_F *dst, long long n, size_t ncols);

/* ---- H1..H4: BLAKE3 leaves and the Merkle tree (Blake3_hash.cpp:5-10; merkle_tree.cpp:62-87,193-287) ------ */
int hb_blake3_64(hb_ctx *ctx, const uint8_t *src, uint8_t *dst, size_t count);       /* count x (64 B -> 32 B) */
/* create_tree_blake: `levels` holds nleaves digests at offset 0 and room for (2*nleaves-1)*32 bytes; the upper
 * levels are appended level by level.  Parent = H1(left || left), as in the reference (merkle_tree.cpp:275-280). */
int hb_merkle_tree(hb_ctx *ctx, uint8_t *levels, size_t nleaves);
/* MT_commit_Blake: leaf i = H1(leafs[4i..4i+3]); then the tree.  levels: (2*(N/4)-1)*32 bytes. */
int hb_mt_commit(hb_ctx *ctx, const hb_F *leafs, size_t N, uint8_t *levels);

/* ---- T1: tensor code (PC_utils.cpp:9-123) ------------------------------------------------------------------ */
/* msg: n elements -> tensor: (2*trs) x (2n/trs) row-major.  linear_time != 0: rows RS (NTT), columns expander
 * (needs hb_expander_set for n = trs); == 0: RS x RS. */
int hb_tensorcode(hb_ctx *ctx, const hb_F *msg, size_t n, int trs, int linear_time, hb_F *tensor);

/* ---- C1: commit_standard (Our_PC.cpp:146-171) -------------------------------------------------------------- */
/* poly: N elements, K chunks.  levels_out: (2*(N/K)-1)*32 bytes (every level, leaves first).
 * tensor_out: NULL, or K*4*(N/K) elements (chunk-major, then row-major) == the reference's `_tensor`.
 * The encoded tensor always stays resident in the context for hb_open_standard_* (see hb_tensor_device). */
int hb_commit_standard(hb_ctx *ctx, const hb_F *poly, size_t N, int K, int trs, int linear_time,
                       uint8_t *levels_out, hb_F *tensor_out);
/* commit_standard split for sharding across GPUs: the leaf of position p is the Merkle–Damgård chain
 *   leaf[p] <- H1( H1(4-row quad of chunk c at p) | leaf[p] ),  c = 0..K-1 in order, starting from 32 zero bytes.
 * hb_commit_encode_chunks: the chunk-independent part for `nchunks` consecutive chunks of B coefficients — tensor code (kept
 *   resident, like hb_commit_standard: this call fills chunks [first_chunk, first_chunk+nchunks) of a tensor of total_chunks;
 *   total_chunks = 0 means "just these") and the INNER digests (32 bytes each) in the layout
 *   inner_out[(part * nchunks + c) * (B/leaf_parts) + off] for leaf position p = part * (B/leaf_parts) + off:
 *   leaf_parts = 1 is the plain [chunk][leaf] order; leaf_parts = G makes the slice for each destination rank contiguous, so
 *   the all_to_all needs no permute pass.
 * hb_md_chain: leaves[p] <- chain over inner[c*nleaves + p], c = 0..nchunks-1, continuing from the current leaves[p].
 * A rank that owns a chunk range calls the first; after exchanging inner digests by leaf range, the owner of a leaf range calls the
 * second and hb_merkle_tree on its subtree (hobbit_b200/dist.py). */
int hb_commit_encode_chunks(hb_ctx *ctx, const hb_F *poly, size_t nchunks, size_t B, int trs, int linear_time, uint8_t *inner_out,
                            size_t leaf_parts, size_t first_chunk, size_t total_chunks);
int hb_md_chain(hb_ctx *ctx, const uint8_t *inner, size_t nchunks, size_t nleaves, uint8_t *leaves);
/* The same split for Elastic_PC commit (Elastic_PC.cpp:174-285): `ngroups` groups of 4 consecutive chunks of B coefficients ->
 * inner[g][p] = H1(c0[p+1] | c1[p+1] | c2[p] | T[p]) for the 4B positions of each group (same exchange layouts as above with 4B leaves);
 * the leaves are then hb_md_chain over the groups in order, and hb_merkle_tree over 4B leaves. */
int hb_elastic_encode_groups(hb_ctx *ctx, const hb_F *chunks, size_t ngroups, size_t B, int trs, int linear_time, uint8_t *inner_out,
                             size_t leaf_parts, size_t first_group, size_t total_groups);
/* device pointer to the resident `_tensor` of the last hb_commit_standard (K*4*(N/K) elements), or NULL */
const hb_F *hb_tensor_device(hb_ctx *ctx);
/* _compute_aggregation_reply (Our_PC.cpp:291-305): reply[q*K + i] = _tensor[i][row[q]][col[q]] */
int hb_tensor_gather(hb_ctx *ctx, const uint32_t *col, const uint32_t *row, size_t queries, hb_F *reply);
/* _aggregate's axpy (Our_PC.cpp:258-273): agg[j] = sum_i beta[i] * poly[i*(N/K) + j] */
int hb_aggregate(hb_ctx *ctx, const hb_F *poly, size_t N, int K, const hb_F *beta, hb_F *agg);

/* ---- C2: Elastic_PC commit (Elastic_PC.cpp:174-285), chunk at a time --------------------------------------- */
/* begin: B = BUFFER_SPACE, trs = tensor_row_size.  push: one chunk of B elements (call N/B times, in order);
 * all-zero chunks skip the encode like the reference (:206-222).  finish: levels_out (2*4B-1)*32 bytes. */
int hb_elastic_begin(hb_ctx *ctx, size_t B, int trs, int linear_time);
int hb_elastic_push(hb_ctx *ctx, const hb_F *chunk);
int hb_elastic_finish(hb_ctx *ctx, uint8_t *levels_out);
/* the same, level l (4B >> l digests) written straight to level_ptrs[l] — the reference's MT_hashes[l] — nlevels = log2(4B)+1;
 * a NULL level_ptrs[l] skips that level (a prover that only needs the root and the level SIZES, see hobbit::commit_levels_on_host) */
int hb_elastic_finish_levels(hb_ctx *ctx, uint8_t *const *level_ptrs, int nlevels);
/* hb_elastic_finish_levels with the host copies in the BACKGROUND (worker thread, own stream, pinned double buffer): returns once the tree
 * kernels are queued, so the copy of 64-256 MiB of levels into fresh pageable pages overlaps whatever the context proves next.  The
 * destination arrays must not be read or freed before hb_levels_wait returned. */
int hb_elastic_finish_levels_async(hb_ctx *ctx, uint8_t *const *level_ptrs, int nlevels);
int hb_levels_copy_async(hb_ctx *ctx, uint8_t *dev_levels, int take_ownership, uint8_t *const *level_ptrs, int nlevels, size_t nleaves);
int hb_levels_wait(hb_ctx *ctx);
/* ---- O2, data-parallel front half of Elastic_PC open (Elastic_PC.cpp:316-333 aggregate, 487-533 compute_aggregation_reply) --------
 * begin: the `queries` cells (col[q], row[q]) drawn by the host (Elastic_PC.cpp:649-655), nchunks = N/B.  push chunk i with beta[i]:
 *   agg[j] += beta[i] * chunk[j]  and  reply[q*nchunks + i] = tensorcode(chunk)[row[q]][col[q]].  finish: agg (B), reply (queries*nchunks).
 * RS columns (linear_time = 0) match update_reply (:59-110) exactly.  For linear_time != 0 the reference's update_reply_spielman
 * (:431-486) groups replies by sorted column and reads past its buffer for parity rows (undefined behaviour); here every reply is
 * simply the encoded tensor cell, in query order. */
int hb_elastic_open_begin(hb_ctx *ctx, size_t B, int trs, int linear_time, const uint32_t *col, const uint32_t *row, size_t queries, size_t nchunks);
int hb_elastic_open_push(hb_ctx *ctx, const hb_F *chunk, const hb_F *beta_i);
int hb_elastic_open_finish(hb_ctx *ctx, hb_F *agg_out, hb_F *reply_out);
/* read_stream_PC's synthetic default stream (witness_stream.cpp:2405-2411): v[0]=322322, v[i+1]=v[i]^2+i */
int hb_stream_pc_test(hb_ctx *ctx, hb_F *out, size_t n);

/* ---- S9: eq table and MLE evaluation (utils.cpp:251-296, 789-802) ------------------------------------------ */
int hb_precompute_beta(hb_ctx *ctx, const hb_F *r, int nr, hb_F *out);          /* out: 2^nr */
int hb_evaluate_vector(hb_ctx *ctx, const hb_F *v, size_t n, const hb_F *r, hb_F *out);

/* ---- S1/S2/S3/S5: sumcheck provers (sumcheck.cpp:2391-2460, 1974-2058, 275-372, 35-257) --------------------- */
/* Flat proof layouts (hb_F units), rounds = log2(n):
 *   S1: (a,b,c) x rounds | randomness x rounds | vr[2] | final_rand                       -> 4*rounds+3
 *   S2: (a,b,c,d) x rounds | randomness x rounds | vr[3] | final_rand                     -> 5*rounds+4
 *   S3: (a,b,c,d) x rounds | randomness x rounds | vr[3*batches]                          -> 5*rounds+3*batches
 * *ps accumulates the reference's proof-size counter (KB).  Input tables are not modified. */
int hb_sumcheck2(hb_ctx *ctx, const hb_F *v1, const hb_F *v2, size_t n, const hb_F *prev_r, hb_F *proof, double *ps);
int hb_sumcheck3(hb_ctx *ctx, const hb_F *v1, const hb_F *v2, const hb_F *v3, size_t n, const hb_F *prev_r,
                 hb_F *proof, double *ps);
int hb_batch_sumcheck3(hb_ctx *ctx, const hb_F *t1, const hb_F *t2, const hb_F *t3, const size_t *sizes, int batches,
                       const hb_F *a, hb_F *proof, double *ps);
/* prove_multiplication_tree_new for `vectors` tables of n elements (powers of two).  x_rand: the log2(vectors)
 * libc-drawn points (generate_randomness) when vectors > 1, ignored otherwise.  Output layout:
 *   output[vectors] | out_eval | final_r[nfr] | final_eval | per layer: (a,b,c,d) x rounds | vr[3] | final_rand
 * Returns the number of hb_F written in *written. */
int hb_mul_tree(hb_ctx *ctx, const hb_F *input, int vectors, size_t n, const hb_F *prev_r, const hb_F *x_rand,
                hb_F *out, size_t *written, int *nfr, double *ps);

/* One S2 round on device tables, for callers that own the round loop (sumcheck sharded across GPUs by hypercube prefix:
 * every rank runs this on its contiguous slice and the 4 coefficients are all-reduced, hobbit_b200/dist.py).
 * coeffs4 = cubic coefficients (a,b,c,d) over the L pairs of the slice; out_t[j] = in_t[2j] + rand*(in_t[2j+1]-in_t[2j]). */
int hb_sc3_round(hb_ctx *ctx, const hb_F *in1, const hb_F *in2, const hb_F *in3, hb_F *out1, hb_F *out2, hb_F *out3, size_t L,
                 const hb_F *rand, hb_F *coeffs4);

/* ---- gate consistency, in-memory form: prove_gate_consistency_standard (sumcheck.cpp:434-501) --------------------------------- */
/* Degree-4 sumcheck of beta(x) (mul(x) L(x) R(x) + add(x) (L(x)+R(x)) - O(x)), mul = 1 - add, beta = eq(r), rand_0 = F(213).
 * The reference folds its arguments in place and returns nothing; here the inputs are untouched and
 * out = (a,b,c,d,e,rand) per round [6*log2 n] | final add, L, R, O, mul, beta [6]. */
int hb_gate_consistency_standard(hb_ctx *ctx, const hb_F *L, const hb_F *R, const hb_F *O, const hb_F *add_gate, size_t n,
                                 const hb_F *r, hb_F *out);

/* S7: prove_gate_consistency (sumcheck.cpp:796-981, no lookups) with the transcript resident in HBM: L, R, O (gate wire values) and
 * S (F(1) = add gate, F(0) = mul gate), `cs` entries as read_trace emits them (witness_stream.cpp:1701-1807), processed in chunks of
 * B = BUFFER_SPACE.  r: log2 B points; rnd10 = generate_randomness(4) | generate_randomness(6), drawn by the HOST in that order.
 * The reference returns nothing; out = R[cs/B] | (a,b,c,d,e,rand) x log2 B | final L,R,O,add,mul,beta | Peval[6][cs/B] |
 * flat 2-product proof (4*log2(cs/B)+3).  The reference's self-checks ("Error in gate consistency 1/2/3") fail the call. */
int hb_gate_consistency_stream(hb_ctx *ctx, const hb_F *L, const hb_F *R, const hb_F *O, const hb_F *S, size_t cs, size_t B,
                               const hb_F *r, const hb_F *rnd10, hb_F *out, double *ps);

/* S8: prove_gate_consistency_lookups (sumcheck.cpp:503-794): as above with S = F(0) add / F(1) mul / F(2) lookup, lookup_rand2 = lookup_rand[0..1],
 * rnd13 = generate_randomness(5) | generate_randomness(8).  out = R[cs/B] | (a,b,c,d,e,rand) x log2 B | final L,R,O,add_L,add_R,mul,lkp,lkp_O,beta |
 * Peval[8][cs/B] | flat 2-product proof (4*log2(cs/B)+3). */
int hb_gate_consistency_lookups_stream(hb_ctx *ctx, const hb_F *L, const hb_F *R, const hb_F *O, const hb_F *S, size_t cs, size_t B,
                                       const hb_F *r, const hb_F *lookup_rand2, const hb_F *rnd13, hb_F *out, double *ps);

/* ---- S4/S6: streaming folding sumcheck with the witness stream resident in HBM (sumcheck.cpp:1093-1392, 1746-1915) --------- */
/* xy: the stream in its logical two-half form [X | Y] (`total` elements; what read_stream emits as X-block | Y-block per read,
 * witness_stream.cpp:2276-2311).  Layer l is [seg_l(X) | seg_l(Y)] with 2^l-element segment products (read_mul_tree_data).
 * hb_stream_sumcheck_layer = generate_3product_sumcheck_beta_stream_batch_optimized with batches = 1, distance = 1:
 *   r: log2((total>>layer_id)/2) points; rnd4 = (a, b0, b1, pad): generate_randomness(1), generate_randomness(2), random() drawn by
 *   the HOST in that order; new_r gets 1 + log2(total>>layer_id) - 1 points.  Error checks behave like the reference
 *   ("Error in sumcheck 0" warns, 1 and 2 fail).
 * hb_mul_tree_stream = prove_multiplication_tree_stream_shallow for layers <= distance (or naive): out = the `vectors` products;
 *   x_rand: log2(vectors) points; rnd: 4 per streamed layer (top layer first); *layers_out = number of streamed layers. */
int hb_stream_sumcheck_layer(hb_ctx *ctx, const hb_F *xy, size_t total, size_t B, int layer_id, const hb_F *r, int nr,
                             const hb_F *old_claim, const hb_F *rnd4, hb_F *new_claim, hb_F *new_r, int *n_new_r, double *ps);
int hb_mul_tree_stream(hb_ctx *ctx, const hb_F *xy, size_t total, int vectors, size_t B, int distance, int naive,
                       const hb_F *prev_r, const hb_F *x_rand, const hb_F *rnd, hb_F *out, int *layers_out, double *ps);

/* ---- 8f.1: building blocks of the opening recursion behind open_standard / Elastic_PC open ----------------------------------------
 * (shockwave_commit / shockwave_prove Virgo.cpp:120-157,435-517; whir_commit / _whir_prove :160-178,519-686; recursive_prover_Spielman
 *  / _RS PC_utils.cpp:290-512; prove_fft / prove_fft_matrix sumcheck.cpp:2975-3027).  The host mirror (hobbit_b200/host/hobbit_open.cpp)
 * owns the libc RNG order and the Fiat–Shamir scalars and calls these on tables that stay resident in HBM (device pointers). */
int hb_vec_zero(hb_ctx *ctx, hb_F *v, size_t n);
/* `rows` rows of in_len elements (contiguous), each zero-extended to 2^logn and transformed (_fft) into dst (rows x 2^logn) */
int hb_rs_encode_rows(hb_ctx *ctx, const hb_F *src, size_t in_len, size_t rows, hb_F *dst, int logn);
/* out[j] = sum_i w[i] M[i*stride + j] (j < cols)  — shockwave_prove's aggregation (:446-456), prepare_matrix(transpose(M), r) with
 * w = eq(r) (utils.cpp:758-775), the column evaluations of the recursive provers (PC_utils.cpp:332-340) */
int hb_matvec_cols(hb_ctx *ctx, const hb_F *M, size_t rows, size_t cols, size_t stride, const hb_F *w, hb_F *out);
/* out[i] = sum_j s[j] M[i*stride + j] (i < rows)  — aggr_c of recursive_prover_Spielman (PC_utils.cpp:318-328), dot products */
int hb_matvec_rows(hb_ctx *ctx, const hb_F *M, size_t rows, size_t cols, size_t stride, const hb_F *s, hb_F *out);
int hb_axpy(hb_ctx *ctx, hb_F *y, const hb_F *x, const hb_F *a, size_t n);                     /* y += a * x */
/* out[0..n) = 0, then out[idx[k]] = val[k]; idx must be unique (the host resolves the reference's sequential duplicate semantics) */
int hb_scatter(hb_ctx *ctx, hb_F *out, size_t n, const uint64_t *idx, const hb_F *val, size_t m);
/* out[q*rows + j] = M[j*stride + col[q]]  — column replies (Virgo.cpp:468-472) */
int hb_gather_cols(hb_ctx *ctx, const hb_F *M, size_t rows, size_t cols, size_t stride, const uint64_t *col, size_t m, hb_F *out);
/* out[j*m + q] = M[j*stride + col[q]] (rows x m: message q = column q, the layout hb_encode_batch takes); out (cols x rows) = in^T;
 * *out = 1 iff some element of v is non-zero (Elastic_PC's all-zero chunk test, Elastic_PC.cpp:206-213, 509-516) */
int hb_select_cols(hb_ctx *ctx, const hb_F *M, size_t rows, size_t cols, size_t stride, const uint64_t *col, size_t m, hb_F *out);
int hb_transpose(hb_ctx *ctx, const hb_F *in, size_t rows, size_t cols, hb_F *out);
int hb_any_nonzero(hb_ctx *ctx, const hb_F *v, size_t n, int *out);
/* phiGInit(phi_g, r, F(1), n, false) (utils.cpp:694-755): out has 2^n entries, the upper half stays zero like the reference's */
int hb_phi_g_init(hb_ctx *ctx, const hb_F *r, int n, hb_F *out);
/* shockwave_commit's leaves: leaves[c] = root of MT_commit_Blake over the k cells enc[0..k)[c] of column c (Virgo.cpp:140-152) */
int hb_shockwave_leaves(hb_ctx *ctx, const hb_F *enc, int k, size_t cols, uint8_t *leaves);
int hb_change_form(hb_ctx *ctx, hb_F *poly, int logn);                                          /* Virgo.cpp:104-118, in place */
int hb_regroup(hb_ctx *ctx, const hb_F *in, size_t n, int k, hb_F *out);                        /* out[j*2^k + t] = in[j + t*(n>>k)] (:168-175) */
/* one WHIR sumcheck round over the pairs (j, j+L) of (poly, beta) (Virgo.cpp:548-568): coefficients (a,b,c), then the in-place fold */
int hb_whir_poly(hb_ctx *ctx, const hb_F *poly, const hb_F *beta, size_t L, hb_F *coeffs3);
int hb_whir_fold(hb_ctx *ctx, hb_F *poly, hb_F *beta, size_t L, const hb_F *a);
/* out-of-domain / shift queries of one WHIR iteration (:597-623): zetas = repeats x v points, y[i] = sum_j eq(z_i)[j] poly[j],
 * beta[j] += sum_i pows[i] eq(z_i)[j]   (poly, beta: 2^v entries) */
int hb_whir_zeta(hb_ctx *ctx, const hb_F *poly, hb_F *beta, int v, const hb_F *zetas, int repeats, const hb_F *pows, hb_F *y);

/* ---- W1/W2 (8f.2): circuit witness streams derived on the GPU from ONE pass of the evaluator's trace ------------------------------------
 * A trace record is the reference's 80-byte `tr_tuple` (Seval.h:4-9): F value_o, value_l, value_r; int idx_o, idx_l, idx_r;
 * int access_o, access_l, access_r; uint8_t type (0 delete/output, 1 add, 2 mul, >= 3 lookup, 255 end of circuit).
 * begin / push (the producer's host buffers, in order; everything from the first type-255 record on is ignored and *done becomes 1) /
 * finish (counts).  The trace stays resident in the context; each stream below is what the reference's stateful readers emit when the
 * stream is read front to back (witness_stream.cpp:768-874, 1055-1338, 1620-1807, 2276-2311), `cs` = circuit_size = ops = gates:
 *   witness     4cs : (value_l, value_r, value_o) per op record, zero padded to 3cs | value_o per delete record, zero padded to cs
 *   transcript  cs  : L, R, O = values per op record; S = F(1) add / F(0) mul  (has_lookups: 0 add / 1 mul / 2 table), zero padded
 *   wiring      8cs : "wiring_consistency_check_opt" in its logical two-half form [X | Y] (what hb_mul_tree_stream takes):
 *                     X = idx+1 + a_w value + b_w access over (l, r, o) of op records (3cs) | idx_o+1 + a_w value_o over deletes (cs);
 *                     Y = X + b_w (or 1 where X == 1) | X + b_w access_o; padding = 1 */
int hb_trace_begin(hb_ctx *ctx, size_t capacity_records);
int hb_trace_push(hb_ctx *ctx, const void *records, size_t n, int *done);
int hb_trace_finish(hb_ctx *ctx, size_t *n_records, size_t *n_ops, size_t *n_deletes);
/* 8f.4: the MLP circuit evaluated on the GPU instead of by the producer thread (Seval.cpp:1238-1286 MLP_inference under the fun == 9
 * driver, :1462-1489: weights F((j+i+k)%256), inputs F((k+1)%256)): fills the resident trace with exactly the records, labels and access
 * counters one pass of the CPU evaluator emits.  layer_size: the network shape (`pigeon 9 ... n l0 l1 ...`). */
int hb_trace_generate_mlp(hb_ctx *ctx, const int *layer_size, int nsizes, size_t *n_records);
/* 8f.4: the AES circuit (Seval.cpp:957-1084 lookup_box / encrypt / AES under the fun == 5 driver, :1353-1396: input_size blocks of 16 bytes
 * F((i*122+j)%256), 10 round keys F((i+j+1)%256), toy tables (x+21)%256 / (x+3)%256 / (x+4)%256 / xor) evaluated on the GPU: 1824 records
 * per block + the final deletes, exactly the records, labels and access counters of one pass of the CPU evaluator. */
int hb_trace_generate_aes(hb_ctx *ctx, int input_size, size_t *n_records);
/* 8f.4: the pruned two-layer MLP (Seval.cpp:1170-1236 `inference` under the fun == 8 driver, :1424-1461): n_inputs inputs F((i+1)%256);
 * neuron i of layer l reads cols_l[rowptr_l[i] .. rowptr_l[i+1]) (the reference's indexes[l][i], drawn from libc rand() by its driver: HOST
 * arrays) with weights F((j+i)%256); layer 1 reads hidden values; a neuron without inputs is a copy of the gate `zero`.  The access counter
 * of an input is the number of earlier reads of it in evaluation order (a stable rank, computed on the GPU). */
int hb_trace_generate_pruned_mlp(hb_ctx *ctx, int n_inputs, int n_hidden, int n_out, const int *rowptr0, const int *cols0, const int *rowptr1,
                                 const int *cols1, size_t *n_records);
/* 8f.4: the SQL range-query circuit (Seval.cpp:1085-1166 range_query with get_bytes :667-687, ltu_gate :223-293 and lookup_gate :190-221
 * under the fun == 6 driver, :1398-1416: DB[i] = F((i+12)%16), L = 21, R = 321, two bytes per word) evaluated on the GPU. */
int hb_trace_generate_sql(hb_ctx *ctx, int input_size, size_t *n_records);
int hb_trace_witness(hb_ctx *ctx, size_t cs, hb_F *out);
int hb_trace_transcript(hb_ctx *ctx, size_t cs, int has_lookups, hb_F *L, hb_F *R, hb_F *O, hb_F *S);
int hb_trace_wiring(hb_ctx *ctx, size_t cs, const hb_F *a_w, const hb_F *b_w, hb_F *xy);
/* "circuit" (16cs, has_lookups = false; witness_stream.cpp:2123-2162): selector 1 add / 0 mul per op record (cs) | (idx, access) pairs of
 * (l, r, o) per op record (6cs) | (idx_o, access_o) pairs per delete record (2cs) | zeros (7cs) — what prove_arbitrary_circuit opens */
int hb_trace_circuit(hb_ctx *ctx, size_t cs, hb_F *out);
/* lookup streams (witness_stream.cpp:920-1053, 2198-2247); access = number of EARLIER lookups of the same table entry in the pass
 * (table = type-3, entry = value_l for the range table (type 3), value_l + 256 value_r otherwise):
 *   lookup_basic    2cs [X | Y]: per op record X = 1 + value_l + lr0 value_r + lr1 value_o + lr2 access + lr3 type on lookup records, 1 elsewhere;
 *                                Y = X + lr2, or 1 where X == 1
 *   lookup_witness  2cs        : per LOOKUP record (value_o + lr0 value_l + lr1 value_r, access), zero padded */
int hb_trace_lookup_basic(hb_ctx *ctx, size_t cs, const hb_F *lookup_rand4, hb_F *xy);
int hb_trace_lookup_witness(hb_ctx *ctx, size_t cs, const hb_F *lookup_rand2, hb_F *out);

/* ---- multi-GPU: one process per GPU, the GPUs of one NVSwitch box ------------------------------------------------------
 * The reference is single-process; these entry points implement the partitioning §8e derives from it: commit_standard / Elastic_PC commit
 * are chunk-parallel up to the Merkle–Damgård chain of each leaf (Our_PC.cpp:146-171, Elastic_PC.cpp:174-285), sumchecks are sharded by
 * hypercube prefix (sumcheck.cpp:1974-2058 and the streaming provers).  Ranks exchange data with peer stores over NVLink through a
 * WINDOW of device memory each rank exports with CUDA IPC; there is no host round trip and no separate collective on the data path.
 * Bootstrap: every rank calls hb_dist_local_info (allocates its window: 64 KiB of control + data_bytes) and gets a 256-byte blob; the
 * caller all-gathers the blobs (torch.distributed, MPI, the TCP store of the C++ host mirror ...) and passes all `world` blobs, in rank
 * order, to hb_dist_connect.  world must be 1, 2, 4 or 8.  All ranks must then make the same sequence of hb_dist_* / prover calls. */
int hb_dist_local_info(hb_ctx *ctx, size_t data_bytes, void *blob256);
int hb_dist_connect(hb_ctx *ctx, int rank, int world, const void *blobs);
int hb_dist_disconnect(hb_ctx *ctx);
int hb_dist_rank(hb_ctx *ctx);
int hb_dist_world(hb_ctx *ctx);
int hb_dist_barrier(hb_ctx *ctx);                                   /* device-side barrier over the ranks, then stream sync */
int hb_dist_allreduce(hb_ctx *ctx, hb_F *vec, size_t n);           /* field sum over the ranks, in place (needs n*16*world data bytes) */
/* commit_standard of K_total chunks of B coefficients; this rank passes its K_total/world consecutive chunks (chunk range
 * [rank*K_total/world, ...)), host or device memory.  The inner leaf digests are stored by the encode kernel straight into the window
 * of the rank that owns the leaf range, every rank chains its B/world leaves over all chunks and builds its subtree, subtrees are
 * scattered to every rank and the top log2(world) levels rebuilt: levels_out receives ALL (2B-1)*32 bytes on every rank (NULL: skip
 * the copy; the tree stays in the window).  Needs data_bytes >= K_total*(B/world)*32 + (2B-1)*32 + 256.  world == 1: hb_commit_standard. */
int hb_dist_commit_standard(hb_ctx *ctx, const hb_F *poly_local, size_t K_total, size_t B, int trs, int linear_time, uint8_t *levels_out);
/* Elastic_PC commit of groups_total groups of 4 chunks of B coefficients, sharded the same way (4B leaves): levels_out (8B-1)*32 bytes */
int hb_dist_elastic_commit(hb_ctx *ctx, const hb_F *chunks_local, size_t groups_total, size_t B, int trs, int linear_time, uint8_t *levels_out);
/* the streaming form: begin, hb_elastic_push for this rank's 4*groups_total/world chunks (its consecutive groups, in order), then
 * hb_elastic_finish / hb_elastic_finish_levels: every rank receives the whole tree */
int hb_dist_elastic_begin(hb_ctx *ctx, size_t B, int trs, int linear_time, size_t groups_total);
/* Elastic_PC open, stream pass of a chunk RANGE (call right after hb_elastic_open_begin(.., nchunks = the range length)): the pushes are
 * chunks [first, first+nchunks) of `total`; reply_out of hb_elastic_open_finish then has queries*total entries, this range filled and
 * zeros elsewhere, so hb_dist_allreduce over the ranks assembles aggregate and replies (Elastic_PC.cpp:316-333, 487-533 are sums /
 * independent cells over the chunks) */
int hb_elastic_open_range(hb_ctx *ctx, size_t first, size_t total);
/* on != 0: the provers (hb_sumcheck3, hb_batch_sumcheck3, hb_mul_tree, hb_stream_sumcheck_layer, hb_mul_tree_stream,
 * hb_gate_consistency_stream, hb_gate_consistency_lookups_stream) shard their work over the ranks: every rank passes the SAME full
 * tables (replicated in HBM) and works on its contiguous part of each table / of each BUFFER_SPACE chunk; the round sums are added
 * across ranks inside the round kernel, so every rank derives the same Fiat–Shamir challenges and returns the same proof, bit-identical
 * to the single-GPU proof. */
int hb_dist_shard(hb_ctx *ctx, int on);
/* counters since hb_dist_connect: cross-rank reductions done inside round kernels, small all-gathers, device barriers */
void hb_dist_stats(hb_ctx *ctx, uint64_t *out3);

#ifdef __cplusplus
}
#endif
#endif /* HOBBIT_B200_H */
