// Single-block BLAKE3 (64 bytes in, 32 bytes out) for the Merkle layer.
// Replaces blake3_hash() (reference src/Blake3_hash.cpp:5-10): hashing exactly 64 bytes is ONE compression with
// cv = IV, counter = 0, block_len = 64, flags = CHUNK_START|CHUNK_END|ROOT (Blake/blake3_impl.h:18-21).
// The whole state lives in registers; the message permutation between rounds is resolved at compile time
// (fully unrolled schedule), so there is no data movement for it at all.
#pragma once
#include <cstdint>

namespace hb {

struct Digest { uint32_t w[8]; };   // 32 bytes, little-endian words == reference `_hash::arr`

__device__ __forceinline__ uint32_t rotr(uint32_t x, int c) { return __funnelshift_r(x, x, c); }

// Pipe balancing (measured on B200, tools/ubench_blake.cu): XOR and rotate can only issue on the ALU pipe (8 per G,
// 2 cycles each per SM sub-partition), so the additions must stay off it.  ptxas fuses `a + b + m` into one 3-input
// IADD3 on the ALU pipe (variant 0: 32.2 G compress/s, ALU pipe 96 % busy, FMA pipe 10 %).  Making `a + b` an opaque
// IMAD (`a * one + b`, `one` read from %nctaid.z == 1, unknown to ptxas) leaves only 2-input adds, which ptxas emits
// as IMAD.IADD on the FMA pipe: variant 3 reaches 38.9 G compress/s = the ALU floor of 466 LOP3/SHF per compression.
// (Forcing every add through the register-form IMAD — variant 1 — is slower: 25.8 G/s.)
__device__ __forceinline__ uint32_t hb_one() { uint32_t v; asm("mov.u32 %0, %%nctaid.z;" : "=r"(v)); return v; }
__device__ __forceinline__ uint32_t add_fma(uint32_t a, uint32_t b, uint32_t one) {
    uint32_t d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b)); return d;
}

#ifndef HB_BLAKE_VARIANT
#define HB_BLAKE_VARIANT 3
#endif
#if HB_BLAKE_VARIANT == 0      // plain adds (ptxas picks IADD3 on the ALU pipe)
#define HB_ADD3(a, b, x) ((a) + (b) + (x))
#define HB_ADD2(c, d) ((c) + (d))
#elif HB_BLAKE_VARIANT == 1    // every add on the FMA pipe
#define HB_ADD3(a, b, x) add_fma(add_fma(a, b, one), (x), one)
#define HB_ADD2(c, d) add_fma(c, d, one)
#elif HB_BLAKE_VARIANT == 2    // a+b+x stays one IADD3, c+d goes to the FMA pipe
#define HB_ADD3(a, b, x) ((a) + (b) + (x))
#define HB_ADD2(c, d) add_fma(c, d, one)
#elif HB_BLAKE_VARIANT == 3    // a+b on the FMA pipe, the rest plain
#define HB_ADD3(a, b, x) (add_fma(a, b, one) + (x))
#define HB_ADD2(c, d) ((c) + (d))
#else                          // two 2-input adds kept apart by an empty asm so that ptxas cannot fuse them into IADD3
__device__ __forceinline__ uint32_t opaque(uint32_t v) { asm("" : "+r"(v)); return v; }
#define HB_ADD3(a, b, x) (opaque((a) + (b)) + (x))
#define HB_ADD2(c, d) ((c) + (d))
#endif
#define HB_G(a, b, c, d, x, y)                                                                          \
    do {                                                                                                \
        a = HB_ADD3(a, b, x); d = rotr(d ^ a, 16); c = HB_ADD2(c, d); b = rotr(b ^ c, 12); \
        a = HB_ADD3(a, b, y); d = rotr(d ^ a, 8);  c = HB_ADD2(c, d); b = rotr(b ^ c, 7);  \
    } while (0)

// schedule[r][i] = index into the ORIGINAL message words used at position i of round r
// (MSG_PERMUTATION = {2,6,3,10,7,0,4,13,1,11,12,5,9,14,15,8} applied r times).
#define HB_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    HB_G(v0, v4, v8,  v12, m[s0],  m[s1]);  HB_G(v1, v5, v9,  v13, m[s2],  m[s3]);      \
    HB_G(v2, v6, v10, v14, m[s4],  m[s5]);  HB_G(v3, v7, v11, v15, m[s6],  m[s7]);      \
    HB_G(v0, v5, v10, v15, m[s8],  m[s9]);  HB_G(v1, v6, v11, v12, m[s10], m[s11]);     \
    HB_G(v2, v7, v8,  v13, m[s12], m[s13]); HB_G(v3, v4, v9,  v14, m[s14], m[s15]);

__device__ __forceinline__ void blake3_compress64(const uint32_t (&m)[16], uint32_t (&out)[8]) {
    const uint32_t one = hb_one();
    uint32_t v0 = 0x6A09E667u, v1 = 0xBB67AE85u, v2 = 0x3C6EF372u, v3 = 0xA54FF53Au;
    uint32_t v4 = 0x510E527Fu, v5 = 0x9B05688Cu, v6 = 0x1F83D9ABu, v7 = 0x5BE0CD19u;
    uint32_t v8 = 0x6A09E667u, v9 = 0xBB67AE85u, v10 = 0x3C6EF372u, v11 = 0xA54FF53Au;
    uint32_t v12 = 0u, v13 = 0u, v14 = 64u, v15 = 11u;   // counter lo/hi, block_len, CHUNK_START|CHUNK_END|ROOT
    HB_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    HB_ROUND(2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
    HB_ROUND(3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1)
    HB_ROUND(10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6)
    HB_ROUND(12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4)
    HB_ROUND(9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7)
    HB_ROUND(11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13)
    out[0] = v0 ^ v8;  out[1] = v1 ^ v9;  out[2] = v2 ^ v10; out[3] = v3 ^ v11;
    out[4] = v4 ^ v12; out[5] = v5 ^ v13; out[6] = v6 ^ v14; out[7] = v7 ^ v15;
}

// H2 of the reference (merkle_tree.cpp:62-87): h = H1(x|y|z|w) ; return H1(h | prev)
__device__ __forceinline__ void md_leaf(const uint32_t (&cells)[16], const uint32_t (&prev)[8], uint32_t (&out)[8]) {
    uint32_t m[16];
    uint32_t h[8];
    blake3_compress64(cells, h);
#pragma unroll
    for (int i = 0; i < 8; i++) { m[i] = h[i]; m[8 + i] = prev[i]; }
    blake3_compress64(m, out);
}

}  // namespace hb
