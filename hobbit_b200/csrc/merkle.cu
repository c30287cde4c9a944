// H1..H4 — BLAKE3 leaf hashing and the Merkle tree of the commitment.
// Reference: blake3_hash (src/Blake3_hash.cpp:5-10), hash_double_field_element_merkle_damgard_blake
// (src/merkle_tree.cpp:62-87), MT_commit_Blake (:193-221), create_tree_blake (:255-287).
//
// One leaf (or one parent) per thread; state and message in registers; 32-byte digests move as two uint4.
// Adjacent threads own adjacent columns, so the four 16-byte cells of a leaf are four fully coalesced row reads.
// The tree is built level by level; the reference's parent rule H1(left || LEFT) (merkle_tree.cpp:275-280 reads
// hashes[lvl-1][2i] twice) is reproduced on purpose — every level is an output (it serves authentication paths).
#include "common.cuh"
#include "blake3.cuh"

namespace hb {

__device__ __forceinline__ void load_digest(const uint8_t *p, uint32_t (&w)[8]) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
__device__ __forceinline__ void store_digest(uint8_t *p, const uint32_t (&w)[8]) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(w[0], w[1], w[2], w[3]);
    q[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
__device__ __forceinline__ void cell_words(F x, uint32_t *m) {
    m[0] = (uint32_t)x.re; m[1] = (uint32_t)(x.re >> 32); m[2] = (uint32_t)x.im; m[3] = (uint32_t)(x.im >> 32);
}

__global__ void __launch_bounds__(256) blake3_64_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t m[16], out[8];
    const uint4 *q = reinterpret_cast<const uint4 *>(src + i * 64);
#pragma unroll
    for (int k = 0; k < 4; k++) { uint4 v = q[k]; m[4 * k] = v.x; m[4 * k + 1] = v.y; m[4 * k + 2] = v.z; m[4 * k + 3] = v.w; }
    blake3_compress64(m, out);
    store_digest(dst + i * 32, out);
}

// commit_standard leaves (Our_PC.cpp:160-166): leaf[j*cols+k] <- H1( H1(T[4j][k]|T[4j+1][k]|T[4j+2][k]|T[4j+3][k]) | leaf[j*cols+k] ),
// chained over the chunks in order.  Split in two kernels: the inner digests of all chunks are independent ...
__global__ void __launch_bounds__(256) md_inner_standard_kernel(const F *__restrict__ Tbase, size_t chunk_stride, size_t rows, size_t cols,
                                                                uint8_t *__restrict__ inner_base, InnerLayout lay) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t nl = (rows / 4) * cols;
    if (i >= nl) return;
    const F *T = Tbase + (size_t)blockIdx.y * chunk_stride;
    size_t j = i / cols, k = i % cols;
    uint32_t m[16], out[8];
#pragma unroll
    for (int q = 0; q < 4; q++) cell_words(T[(4 * j + q) * cols + k], m + 4 * q);
    blake3_compress64(m, out);
    store_digest(lay.addr(inner_base, blockIdx.y, i), out);
}
// ... and only the outer compression is sequential in the chunk index (one thread walks one leaf position).
__global__ void __launch_bounds__(256) md_chain_kernel(const uint8_t *__restrict__ inner, size_t nchunks, size_t nleaves, uint8_t *__restrict__ leaves) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nleaves) return;
    uint32_t m[16], h[8];
    load_digest(leaves + p * 32, h);
    for (size_t c = 0; c < nchunks; c++) {
        uint32_t in[8];
        load_digest(inner + (c * nleaves + p) * 32, in);
#pragma unroll
        for (int k = 0; k < 8; k++) { m[k] = in[k]; m[8 + k] = h[k]; }
        blake3_compress64(m, h);
    }
    store_digest(leaves + p * 32, h);
}

// Elastic commit leaves (Elastic_PC.cpp:230-239).  The reference passes
// (c0[counter], c1[counter], c2[counter++], T[j][k]) by value and GCC evaluates right-to-left, so the tuple
// hashed at position p is (c0[p+1], c1[p+1], c2[p], T[p]); one past the end reads as zero (see oracle/hobbit_oracle.c).
__global__ void __launch_bounds__(256) md_leaves_stream4_kernel(const F *__restrict__ c0, const F *__restrict__ c1, const F *__restrict__ c2,
                                                                const F *__restrict__ T, size_t cells, uint8_t *__restrict__ leaves) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cells) return;
    uint32_t m[16], prev[8], out[8];
    F z = mkF(0, 0);
    cell_words(p + 1 < cells ? c0[p + 1] : z, m);
    cell_words(p + 1 < cells ? c1[p + 1] : z, m + 4);
    cell_words(c2[p], m + 8);
    cell_words(T[p], m + 12);
    load_digest(leaves + p * 32, prev);
    md_leaf(m, prev, out);
    store_digest(leaves + p * 32, out);
}

// The same tuple WITHOUT the chunk-order-dependent half: inner[p] = H1(c0[p+1] | c1[p+1] | c2[p] | T[p]) of one group of 4 encoded chunks
// (tensors contiguous at T4, 4B cells each).  The chain leaf <- H1(inner | leaf) over the groups is md_chain_kernel's job, so groups can be
// encoded on different GPUs and only 32 B per position cross NVLink (hobbit_b200/dist.py).  grid.y = group.
__global__ void __launch_bounds__(256) md_inner_stream4_kernel(const F *__restrict__ T4, size_t cells, uint8_t *__restrict__ inner, InnerLayout lay) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cells) return;
    const F *c0 = T4 + (size_t)blockIdx.y * 4 * cells, *c1 = c0 + cells, *c2 = c1 + cells, *T = c2 + cells;
    uint32_t m[16], out[8];
    F z = mkF(0, 0);
    cell_words(p + 1 < cells ? c0[p + 1] : z, m);
    cell_words(p + 1 < cells ? c1[p + 1] : z, m + 4);
    cell_words(c2[p], m + 8);
    cell_words(T[p], m + 12);
    blake3_compress64(m, out);
    store_digest(lay.addr(inner, blockIdx.y, p), out);
}

// MT_commit_Blake leaves: leaf i = H1(leafs[4i..4i+3])
__global__ void __launch_bounds__(256) mt_leaves_kernel(const F *__restrict__ x, size_t nleaves, uint8_t *__restrict__ leaves) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nleaves) return;
    uint32_t m[16], out[8];
#pragma unroll
    for (int q = 0; q < 4; q++) cell_words(x[4 * i + q], m + 4 * q);
    blake3_compress64(m, out);
    store_digest(leaves + i * 32, out);
}

// one tree level: out[i] = H1(in[2i] || in[2i])
__global__ void __launch_bounds__(256) merkle_level_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, size_t n_out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    uint32_t l[8], m[16], o[8];
    load_digest(in + 2 * i * 32, l);
#pragma unroll
    for (int k = 0; k < 8; k++) { m[k] = l[k]; m[8 + k] = l[k]; }
    blake3_compress64(m, o);
    store_digest(out + i * 32, o);
}
// the last levels (<= 1024 nodes in) in one CTA: level after level through shared memory, every level also stored
__global__ void __launch_bounds__(512) merkle_top_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, unsigned n_in) {
    __shared__ uint32_t sh[2][512][8];
    int cur = 0;
    unsigned n = n_in / 2;               // nodes of the level being produced
    size_t off = 0;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        uint32_t l[8], m[16], o[8];
        load_digest(in + 2 * (size_t)i * 32, l);
#pragma unroll
        for (int k = 0; k < 8; k++) { m[k] = l[k]; m[8 + k] = l[k]; }
        blake3_compress64(m, o);
#pragma unroll
        for (int k = 0; k < 8; k++) sh[cur][i][k] = o[k];
        store_digest(out + (off + i) * 32, o);
    }
    __syncthreads();
    while (n > 1) {
        off += n; n >>= 1;
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
            uint32_t m[16], o[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { m[k] = sh[cur][2 * i][k]; m[8 + k] = m[k]; }
            blake3_compress64(m, o);
#pragma unroll
            for (int k = 0; k < 8; k++) sh[cur ^ 1][i][k] = o[k];
            store_digest(out + (off + i) * 32, o);
        }
        cur ^= 1;
        __syncthreads();
    }
}

static inline unsigned blocks_for(size_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

int blake3_64_dev(hb_ctx *ctx, const uint8_t *src, uint8_t *dst, size_t count) {
    if (count) HB_LAUNCH(ctx, blake3_64_kernel, blocks_for(count, 256), 256, 0, src, dst, count);
    return 0;
}
int md_inner_standard_dev(hb_ctx *ctx, const F *T, size_t rows, size_t cols, size_t nchunks, size_t chunk_stride, uint8_t *inner, InnerLayout lay) {
    size_t n = (rows / 4) * cols;
    if (lay.part_leaves == 0) lay = InnerLayout::plain(n, nchunks);
    if (n && nchunks) HB_LAUNCH(ctx, md_inner_standard_kernel, dim3(blocks_for(n, 256), (unsigned)nchunks), 256, 0, T, chunk_stride, rows, cols, inner, lay);
    return 0;
}
int md_chain_dev(hb_ctx *ctx, const uint8_t *inner, size_t nchunks, size_t nleaves, uint8_t *leaves) {
    if (nleaves && nchunks) HB_LAUNCH(ctx, md_chain_kernel, blocks_for(nleaves, 256), 256, 0, inner, nchunks, nleaves, leaves);
    return 0;
}
int md_leaves_stream4_dev(hb_ctx *ctx, const F *c0, const F *c1, const F *c2, const F *T, size_t cells, uint8_t *leaves) {
    if (cells) HB_LAUNCH(ctx, md_leaves_stream4_kernel, blocks_for(cells, 256), 256, 0, c0, c1, c2, T, cells, leaves);
    return 0;
}
int md_inner_stream4_dev(hb_ctx *ctx, const F *T4, size_t cells, size_t ngroups, uint8_t *inner, InnerLayout lay) {
    if (cells && ngroups) HB_LAUNCH(ctx, md_inner_stream4_kernel, dim3(blocks_for(cells, 256), (unsigned)ngroups), 256, 0, T4, cells, inner, lay);
    return 0;
}
int merkle_tree_dev(hb_ctx *ctx, uint8_t *levels, size_t nleaves) {
    size_t in_off = 0, n = nleaves;
    while (n > 1024) {
        size_t out_off = in_off + n;
        HB_LAUNCH(ctx, merkle_level_kernel, blocks_for(n / 2, 256), 256, 0, levels + in_off * 32, levels + out_off * 32, n / 2);
        in_off = out_off; n /= 2;
    }
    if (n >= 2) HB_LAUNCH(ctx, merkle_top_kernel, 1, 512, 0, levels + in_off * 32, levels + (in_off + n) * 32, (unsigned)n);
    return 0;
}

}  // namespace hb

// ---------------------------------------------------------------------------------------------------------
extern "C" int hb_blake3_64(hb_ctx *ctx, const uint8_t *src, uint8_t *dst, size_t count) { HB_DEV(ctx);
    using namespace hb;
    Staged s(ctx), d(ctx);
    HB_TRY(s.in(src, count * 64)); HB_TRY(d.outbuf(dst, count * 32));
    HB_TRY(blake3_64_dev(ctx, s.as<uint8_t>(), d.as<uint8_t>(), count));
    HB_TRY(d.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_merkle_tree(hb_ctx *ctx, uint8_t *levels, size_t nleaves) { HB_DEV(ctx);
    using namespace hb;
    if (nleaves == 0 || (nleaves & (nleaves - 1))) HB_FAIL(ctx, "hb_merkle_tree: nleaves must be a power of two");
    Staged d(ctx);
    HB_TRY(d.outbuf(levels, (2 * nleaves - 1) * 32, true));
    HB_TRY(merkle_tree_dev(ctx, d.as<uint8_t>(), nleaves));
    HB_TRY(d.finish());
    HB_TRY(end_call(ctx));
    return 0;
}

extern "C" int hb_mt_commit(hb_ctx *ctx, const hb_F *leafs, size_t N, uint8_t *levels) { HB_DEV(ctx);
    using namespace hb;
    size_t nl = N / 4;
    if (nl == 0 || (nl & (nl - 1))) HB_FAIL(ctx, "hb_mt_commit: N/4 must be a power of two");
    Staged s(ctx), d(ctx);
    HB_TRY(s.in(leafs, N * sizeof(F))); HB_TRY(d.outbuf(levels, (2 * nl - 1) * 32));
    HB_LAUNCH(ctx, mt_leaves_kernel, blocks_for(nl, 256), 256, 0, s.as<F>(), nl, d.as<uint8_t>());
    HB_TRY(merkle_tree_dev(ctx, d.as<uint8_t>(), nl));
    HB_TRY(d.finish());
    HB_TRY(end_call(ctx));
    return 0;
}
