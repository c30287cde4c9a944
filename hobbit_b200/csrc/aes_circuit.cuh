// 8f.4: closed-form records of the AES circuit evaluator (included by trace.cu; host+device so that the closed forms can also be checked
// on a CPU against the gate-by-gate restatement, tools/aes_records_check.cu).
#pragma once
#include "field.cuh"
#include <cstring>

namespace hb {

struct TrTuple {                       // == reference tr_tuple: 3 F, 3 idx, 3 access counters, type; 80 bytes
    F value_o, value_l, value_r;
    int idx_o, idx_l, idx_r;
    int access_o, access_l, access_r;
    uint8_t type; uint8_t pad_[7];
};
static_assert(sizeof(TrTuple) == 80, "tr_tuple layout");

// ---- 8f.4: the AES circuit evaluator (Seval.cpp:957-1084 lookup_box / encrypt / AES under the fun == 5 driver, :1353-1396) ---------------
// Every block runs the same 1824-record gate program (an xor with key 0, then 8 rounds of: S-box lookup, two mix lookups, the mixing xors of
// each 4-byte group, add-round-key, and the deletes in between) and — there is no row shift in this circuit — the four bytes of a group
// never meet another group, so every field of every record is a closed-form function of (block, record position) plus at most 8 rounds of
// byte arithmetic on 4 bytes.  One thread per record, records written coalesced in the order of the CPU evaluator's single pass:
//   labels   : inputs 1..16n | keys (10 x 16) | zero | then 912 per block: the 16 initial xors, then per round 16 S-box outputs, 32 mix
//              outputs (2 per byte), 48 mixing xors (xor1, xor2, xor per output byte) and 16 add-round-key outputs
//   access   : zero is the right operand of every lookup (384 per block, counted through); key bytes are read once per block; a gate is
//              born with access 1; the S-box output is read by both mix lookups and — as the copy temp2[.][0] — twice more by the mixing
//   tables   : 2 -> (x + 21) % 256, 3 -> (x + 3) % 256, 4 -> (x + 4) % 256, 1 -> xor (create_Sbox / create_MixBox / create_xor_table)
constexpr int kAesRecs = 1824, kAesLabels = 912, kAesLookups = 384, kAesRound = 224, kAesRoundLabels = 112;
__host__ __device__ __forceinline__ int aes_in(int b, int j) { return (b * 122 + j) & 255; }
__host__ __device__ __forceinline__ int aes_key(int i, int j) { return (i + j + 1) & 255; }
// which of temp2[.][0..2] of the group's bytes (a, b, c, d) feeds output byte m (xor1 = (a, b), xor2 = (c, d)): a -> 1,0,0,2  b -> 2,1,0,0
// c -> 0,2,1,0  d -> 0,0,2,1 for m = 0..3: the first mix output on the diagonal, the second one place before it, the S-box output elsewhere
__host__ __device__ __forceinline__ int aes_pick(int p, int m) { return m == p ? 1 : (((m + 1) & 3) == p ? 2 : 0); }
struct AesGroup { int out[4], v[4][3], t3[4]; };
__host__ __device__ __forceinline__ void aes_mix(AesGroup &g) {
    for (int p = 0; p < 4; p++) { const int t1 = (g.out[p] + 21) & 255; g.v[p][0] = t1; g.v[p][1] = (t1 + 3) & 255; g.v[p][2] = (t1 + 4) & 255; }
    for (int m = 0; m < 4; m++)
        g.t3[m] = (g.v[0][aes_pick(0, m)] ^ g.v[1][aes_pick(1, m)]) ^ (g.v[2][aes_pick(2, m)] ^ g.v[3][aes_pick(3, m)]);
}
// state of group k of block b at the start of round `round` (1-based), mixed for that round
__host__ __device__ __forceinline__ AesGroup aes_group(int b, int k, int round) {
    AesGroup g;
    for (int p = 0; p < 4; p++) g.out[p] = aes_in(b, 4 * k + p) ^ aes_key(0, 4 * k + p);
    aes_mix(g);
    for (int i = 1; i < round; i++) {
        for (int p = 0; p < 4; p++) g.out[p] = g.t3[p] ^ aes_key(i, 4 * k + p);
        aes_mix(g);
    }
    return g;
}
__host__ __device__ __forceinline__ void aes_op(TrTuple &t, int type, int il, int vl, int al, int ir, int vr, int ar, int io, int vo) {
    t.type = (uint8_t)type; t.idx_l = il; t.value_l = mkF((u64)vl, 0); t.access_l = al; t.idx_r = ir; t.value_r = mkF((u64)vr, 0); t.access_r = ar;
    t.idx_o = io; t.value_o = mkF((u64)vo, 0); t.access_o = 0;
}
__host__ __device__ __forceinline__ void aes_del(TrTuple &t, int io, int vo, int ao) { t.type = 0; t.idx_o = io; t.value_o = mkF((u64)vo, 0); t.access_o = ao; }
// record r (0..1823) of block b of n
__host__ __device__ __forceinline__ TrTuple aes_record(int n, int b, int r) {
    const int key0 = 16 * n + 1, zero = 16 * n + 161, Lb = 16 * n + 162 + kAesLabels * b;
    TrTuple t; memset(&t, 0, sizeof t);
    if (r < 16) {                                                                     // out[j] = element[j] ^ key[0][j]
        const int vi = aes_in(b, r), vk = aes_key(0, r);
        aes_op(t, 4, 1 + 16 * b + r, vi, 0, key0 + r, vk, b, Lb + r, vi ^ vk);
    } else if (r >= 16 + 8 * kAesRound) {                                             // the last round's outputs are never read
        const int j = r - (16 + 8 * kAesRound);
        const AesGroup g = aes_group(b, j >> 2, 8);
        aes_del(t, Lb + 16 + 7 * kAesRoundLabels + 96 + j, g.t3[j & 3] ^ aes_key(8, j), 1);
    } else {
        const int i = (r - 16) / kAesRound + 1, q = (r - 16) % kAesRound;
        const int Li = Lb + 16 + kAesRoundLabels * (i - 1);                           // first label of this round
        const int Lo = i == 1 ? Lb : Li - kAesRoundLabels + 96;                       // labels of the round's input bytes
        const int z0 = kAesLookups * b + 48 * (i - 1);                                // access counter of zero at the round's first lookup
        if (q < 16) {                                                                 // temp1[j] = Sbox(out[j])
            const AesGroup g = aes_group(b, q >> 2, i);
            aes_op(t, 5, Lo + q, g.out[q & 3], 1, zero, 0, z0 + q, Li + q, g.v[q & 3][0]);
        } else if (q < 48) {                                                          // temp2[j][1 + s] = Mix_s(temp1[j])
            const int j = (q - 16) >> 1, s = (q - 16) & 1;
            const AesGroup g = aes_group(b, j >> 2, i);
            aes_op(t, 6 + s, Li + j, g.v[j & 3][0], 1 + s, zero, 0, z0 + 16 + 2 * j + s, Li + 16 + 2 * j + s, g.v[j & 3][1 + s]);
        } else if (q < 128) {                                                         // mixing: per group, per output byte: xor1, xor2, xor, 2 deletes
            const int u = q - 48, k = u / 20, m = (u % 20) / 5, e = u % 5;
            const AesGroup g = aes_group(b, k, i);
            const int x1 = Li + 48 + 12 * k + 3 * m;
            int val[4], lab[4], acc[4];
            for (int p = 0; p < 4; p++) {
                const int vsel = aes_pick(p, m), j = 4 * k + p;
                val[p] = g.v[p][vsel]; lab[p] = vsel == 0 ? Li + j : Li + 16 + 2 * j + (vsel - 1);
                int earlier = 0;
                for (int mm = 0; mm < m; mm++) earlier += aes_pick(p, mm) == 0;
                acc[p] = vsel == 0 ? 3 + earlier : 1;
            }
            const int v1 = val[0] ^ val[1], v2 = val[2] ^ val[3];
            if (e == 0) aes_op(t, 4, lab[0], val[0], acc[0], lab[1], val[1], acc[1], x1, v1);
            else if (e == 1) aes_op(t, 4, lab[2], val[2], acc[2], lab[3], val[3], acc[3], x1 + 1, v2);
            else if (e == 2) aes_op(t, 4, x1, v1, 1, x1 + 1, v2, 1, x1 + 2, v1 ^ v2);
            else if (e == 3) aes_del(t, x1, v1, 2);
            else aes_del(t, x1 + 1, v2, 2);
        } else if (q < 176) {                                                         // delete temp2[j][0..2]
            const int j = (q - 128) / 3, vsel = (q - 128) % 3;
            const AesGroup g = aes_group(b, j >> 2, i);
            aes_del(t, vsel == 0 ? Li + j : Li + 16 + 2 * j + (vsel - 1), g.v[j & 3][vsel], vsel == 0 ? 5 : 2);
        } else {
            const int j = (q - 176) & 15, sec = (q - 176) >> 4;
            const AesGroup g = aes_group(b, j >> 2, i);
            const int l3 = Li + 48 + 12 * (j >> 2) + 3 * (j & 3) + 2, v3 = g.t3[j & 3];
            if (sec == 0) aes_op(t, 4, l3, v3, 1, key0 + 16 * i + j, aes_key(i, j), b, Li + 96 + j, v3 ^ aes_key(i, j));   // add round key
            else if (sec == 1) aes_del(t, l3, v3, 2);                                 // delete temp3[j]
            else aes_del(t, Lo + j, g.out[j & 3], 2);                                 // delete the round's input byte
        }
    }
    return t;
}
// record i (0..16n+160) after the last block: delete the inputs (read once), the keys (read once per block) and zero
__host__ __device__ __forceinline__ TrTuple aes_tail_record(int n, int i) {
    TrTuple t; memset(&t, 0, sizeof t);
    if (i < 16 * n) aes_del(t, 1 + i, aes_in(i >> 4, i & 15), 1);
    else if (i < 16 * n + 160) { const int u = i - 16 * n; aes_del(t, 1 + i, aes_key(u >> 4, u & 15), (u >> 4) < 9 ? n : 0); }   // key 9 is never read
    else aes_del(t, 1 + i, 0, kAesLookups * n);
    return t;
}
}  // namespace hb
