/* TEST INFRASTRUCTURE ONLY.
 * CPU restatement (plain C) of the reference's hot-path algorithms.  It is the
 * checker for the CUDA path: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load it.  The product library
 * (hobbit_b200/libhobbit_b200.so) never links or calls anything in oracle/.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against
 * the unmodified reference compiled in place (oracle/_ref/libhobbit_ref.so,
 * built by oracle/Makefile from /root/reference) in tests/test_oracle_vs_ref.py,
 * and against the committed golden vectors in tests/golden/ (generated from
 * the reference by tests/golden/make_golden.py).  The reference ships no test
 * vectors of its own.
 *
 * All reference citations are relative to /root/reference/src/.
 */
#ifndef HOBBIT_ORACLE_H
#define HOBBIT_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t re, im; } orc_F;   /* fieldElement.hpp:96-97 */

/* F1/F2 */
void orc_field_binop(int op, const orc_F *a, const orc_F *b, orc_F *c, size_t n); /* 0 add 1 sub 2 mul 3 neg 4 inv */
void orc_root_of_unity(int n, orc_F *out);
void orc_mimc_hash(const orc_F *in, const orc_F *k, orc_F *out);
/* N1 */
void orc_fft(orc_F *arr, int logn);
/* RNG-driven (libc rand()/random(), same call order as the reference) */
void orc_generate_randomness(int n, orc_F *out);
void orc_inject_randomness(const orc_F *q, size_t n);   /* NULL: back to libc */
/* E1/E2 */
long long orc_expander_init_store(long long n);
int  orc_expander_levels(long long n);
long long orc_expander_dims(int which, int dep, long long *R, int *deg);
void orc_expander_dump(int which, int dep, uint32_t *nbr, uint64_t *w);
void orc_expander_install(int which, int dep, long long L, long long R, int deg, const uint32_t *nbr, const uint64_t *w);
int  orc_encode_monolithic(const orc_F *src, orc_F *dst, long long n);
int  orc_encode_reseed(const orc_F *src, orc_F *dst, long long n);       /* E3: encode(), graph re-drawn from fixed seeds per call */
/* H1..H4 */
void orc_blake3_hash(const uint8_t *src64, uint8_t *dst32);
void orc_md_leaf(const orc_F *xyzw, const uint8_t *prev, uint8_t *out);
int  orc_mt_commit_blake(const orc_F *leafs, int N, uint8_t *out);
int  orc_create_tree_blake(const uint8_t *leaves, int n, uint8_t *out);
/* T1, C1, C2 */
void orc_compute_tensorcode(const orc_F *msg, size_t n, int trs, int lin, orc_F *tensor_out);
void orc_commit_standard(const orc_F *poly, size_t N, int K, int trs, int lin, uint8_t *levels_out, orc_F *tensor_out);
void orc_read_stream_pc_test(orc_F *out, size_t n);
void orc_elastic_commit(size_t N, size_t B, int trs, int lin, uint8_t *levels_out);
void orc_elastic_commit_stream(const orc_F *stream, size_t N, size_t B, int trs, int lin, uint8_t *levels_out);
/* S9 */
void orc_precompute_beta(const orc_F *r, int nr, orc_F *out);
void orc_evaluate_vector(const orc_F *v, size_t n, const orc_F *r, int nr, orc_F *out);
/* S1/S2/S3/S5 — flat proof layouts identical to oracle/ref_shim.cpp */
double orc_sumcheck2(const orc_F *v1, const orc_F *v2, size_t n, const orc_F *prev_r, orc_F *out);
double orc_sumcheck3(const orc_F *v1, const orc_F *v2, const orc_F *v3, size_t n, const orc_F *prev_r, orc_F *out);
double orc_batch_sumcheck3(const orc_F *t1, const orc_F *t2, const orc_F *t3, const size_t *sizes, int batches,
                           const orc_F *a, orc_F *out);
size_t orc_mul_tree(const orc_F *input, int vectors, size_t n, const orc_F *prev_r, orc_F *out, int *nfr_out, double *ps_out);

/* S4/S6 — streaming folding sumcheck over one product-tree layer and the shallow streaming product tree.
 * xy: the witness stream in its logical two-half form [X | Y] (total elements). */
int orc_stream_sumcheck_layer(const orc_F *xy, size_t total, size_t B, int layer_id, const orc_F *r, int nr, const orc_F *old_claim,
                              orc_F *new_claim, orc_F *new_r, double *ps_out);
int orc_stream_sumcheck_batch(const orc_F *xy, size_t total, size_t B, int layer_id, int distance, int batches, const orc_F *r, int rs, const int *rlen,
                              const orc_F *old_claims, orc_F *new_claims, orc_F *new_r, double *ps_out);
double orc_mul_tree_stream(const orc_F *xy, size_t total, int vectors, size_t B, int distance, int naive, const orc_F *prev_r, orc_F *out);

/* prove_gate_consistency_standard (sumcheck.cpp:434-501); out: (a,b,c,d,e,rand) x rounds | final add, L, R, O, mul, beta */
void orc_gate_consistency_standard(const orc_F *L, const orc_F *R, const orc_F *O, const orc_F *add_gate, size_t n, const orc_F *r, orc_F *out);
/* S7: prove_gate_consistency (sumcheck.cpp:796-981) on a resident transcript (L, R, O, S); out layout in hobbit_oracle.c */
double orc_gate_consistency_stream(const orc_F *L, const orc_F *R, const orc_F *O, const orc_F *S, size_t cs, size_t B, const orc_F *r, orc_F *out);
/* S8: prove_gate_consistency_lookups (sumcheck.cpp:503-794) on a resident transcript; S = F(0) add / F(1) mul / F(2) lookup; lr = lookup_rand[0..1] */
double orc_gate_consistency_lookups_stream(const orc_F *L, const orc_F *R, const orc_F *O, const orc_F *S, size_t cs, size_t B, const orc_F *r,
                                           const orc_F *lr, orc_F *out);
/* O2 front: streaming aggregate + query replies */
void orc_elastic_open_front(const orc_F *stream, size_t nchunks, size_t B, int trs, int lin, const orc_F *beta, const uint32_t *col,
                            const uint32_t *row, size_t queries, orc_F *agg, orc_F *reply);
/* one S2 round on a slice */
void orc_sc3_round(const orc_F *v1, const orc_F *v2, const orc_F *v3, orc_F *o1, orc_F *o2, orc_F *o3, size_t L, const orc_F *rand, orc_F *coeffs4);
/* C1 split for sharding: inner leaf digests of chunks, and the Merkle–Damgård chain over chunks */
void orc_commit_encode_chunks(const orc_F *poly, size_t nchunks, size_t B, int trs, int lin, uint8_t *inner_out);
void orc_elastic_encode_groups(const orc_F *chunks, size_t ngroups, size_t B, int trs, int lin, uint8_t *inner_out);
void orc_md_chain(const uint8_t *inner, size_t nchunks, size_t nleaves, uint8_t *leaves);

#ifdef __cplusplus
}
#endif
#endif
