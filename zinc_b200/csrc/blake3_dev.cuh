// zinc_b200/csrc/blake3_dev.cuh -- register-resident BLAKE3 compression for the Merkle kernels.
//
// Every hash on zinc's commit path is ONE compression:
//   leaf  = blake3::hash(Int<K>::to_bytes())            zip/pcs/utils.rs:87-93, field/int.rs:201-210
//   node  = Hasher::update(l); update(r); finalize()    zip/pcs/utils.rs:95-118   (64-byte input, NOT parent mode)
// both with cv = IV, counter = 0, flags = CHUNK_START|CHUNK_END|ROOT, block_len = input length.
// Leaves wider than 64 bytes (out_limbs > 8) chain several blocks of one chunk.
//
// The 16-word state and the 16 message words live in registers; the 7-round message schedule is resolved at
// compile time (full unroll), so no word is ever moved.
#pragma once
#include <stdint.h>

namespace zipgpu {
namespace b3 {

constexpr uint32_t IV0 = 0x6A09E667u, IV1 = 0xBB67AE85u, IV2 = 0x3C6EF372u, IV3 = 0xA54FF53Au;
constexpr uint32_t IV4 = 0x510E527Fu, IV5 = 0x9B05688Cu, IV6 = 0x1F83D9ABu, IV7 = 0x5BE0CD19u;
constexpr uint32_t CHUNK_START = 1, CHUNK_END = 2, ROOT = 8;

struct Schedule {
    unsigned char s[7][16];
};
constexpr Schedule make_schedule() {
    constexpr unsigned char perm[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
    Schedule r{};
    for (int i = 0; i < 16; i++) r.s[0][i] = (unsigned char)i;
    for (int k = 1; k < 7; k++)
        for (int i = 0; i < 16; i++) r.s[k][i] = r.s[k - 1][perm[i]];
    return r;
}

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

// Pipe balancing (measured on B200 with scratch/hashbench.cu and scratch/pipes.cu, see profiles/r1_int32_pipes.md):
// LOP3/SHF/PRMT/IADD3 all issue on the "alu" pipe (64 lanes/clk/SM); a plain C++ G compiles to 10 alu ops
// (4 xor, 4 rotate, 2 IADD3) + 2 IMAD.IADD and is bound by the alu pipe at 20.8 clk per G per warp.  IMAD issues
// on the "fma" pipe, so the a+b of every half-G is issued as IMAD with a run-time multiplier `one` (== 1, a
// kernel argument, so ptxas cannot fold it back into a 3-input IADD3); ptxas then turns the remaining 2-input
// adds (+m, c+d) into IMAD.IADD by itself: 8 alu + 6 fma per G -> 17.6 clk per G measured (alu floor: 16).
// Issuing every add as IMAD-with-register-multiplier is slower again (21.3 clk): the immediate form is cheaper.
__device__ __forceinline__ uint32_t add_fma(uint32_t a, uint32_t b, uint32_t one) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
    return r;
}

// (Round 2, measured slower: issuing the reg + constant adds of round 0 -- "+ x" onto a folded a + b, and c + d with c
// still a constant -- as IMAD with the run-time multiplier instead of VIADD removed 8 alu-pipe instructions per
// compression, but the fused commit kernel went from 1.849 to 1.923 ms: the three-register IMAD form is the expensive
// one, as the 21.3 clk figure above already said.)
#define ZIPGPU_B3_G(a, b, c, d, x, y, HX, HY)         \
    a = add_fma(a, b, one);                           \
    if (HX) a = a + (x);                              \
    d = rotr(d ^ a, 16);                              \
    c = c + d;                                        \
    b = rotr(b ^ c, 12);                              \
    a = add_fma(a, b, one);                           \
    if (HY) a = a + (y);                              \
    d = rotr(d ^ a, 8);                               \
    c = c + d;                                        \
    b = rotr(b ^ c, 7);

// Generic single compression.  `m` has 16 words; words >= NW are known zero at compile time.
template <int NW>
__device__ __forceinline__ void compress(const uint32_t (&cv)[8], const uint32_t (&m_in)[16], uint32_t counter_lo,
                                         uint32_t counter_hi, uint32_t block_len, uint32_t flags,
                                         uint32_t (&out)[8], uint32_t one) {
    constexpr Schedule S = make_schedule();
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = (i < NW) ? m_in[i] : 0u;
    uint32_t s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    uint32_t s8 = IV0, s9 = IV1, s10 = IV2, s11 = IV3, s12 = counter_lo, s13 = counter_hi, s14 = block_len,
             s15 = flags;
#pragma unroll
    for (int r = 0; r < 7; r++) {
        ZIPGPU_B3_G(s0, s4, s8, s12, m[S.s[r][0]], m[S.s[r][1]], S.s[r][0] < NW, S.s[r][1] < NW)
        ZIPGPU_B3_G(s1, s5, s9, s13, m[S.s[r][2]], m[S.s[r][3]], S.s[r][2] < NW, S.s[r][3] < NW)
        ZIPGPU_B3_G(s2, s6, s10, s14, m[S.s[r][4]], m[S.s[r][5]], S.s[r][4] < NW, S.s[r][5] < NW)
        ZIPGPU_B3_G(s3, s7, s11, s15, m[S.s[r][6]], m[S.s[r][7]], S.s[r][6] < NW, S.s[r][7] < NW)
        ZIPGPU_B3_G(s0, s5, s10, s15, m[S.s[r][8]], m[S.s[r][9]], S.s[r][8] < NW, S.s[r][9] < NW)
        ZIPGPU_B3_G(s1, s6, s11, s12, m[S.s[r][10]], m[S.s[r][11]], S.s[r][10] < NW, S.s[r][11] < NW)
        ZIPGPU_B3_G(s2, s7, s8, s13, m[S.s[r][12]], m[S.s[r][13]], S.s[r][12] < NW, S.s[r][13] < NW)
        ZIPGPU_B3_G(s3, s4, s9, s14, m[S.s[r][14]], m[S.s[r][15]], S.s[r][14] < NW, S.s[r][15] < NW)
    }
    out[0] = s0 ^ s8;
    out[1] = s1 ^ s9;
    out[2] = s2 ^ s10;
    out[3] = s3 ^ s11;
    out[4] = s4 ^ s12;
    out[5] = s5 ^ s13;
    out[6] = s6 ^ s14;
    out[7] = s7 ^ s15;
}

struct Digest {
    uint32_t w[8];
};

// hash of a 64-byte node: left digest || right digest (digest words are little-endian, no byte shuffling)
__device__ __forceinline__ void hash_node(const uint32_t (&l)[8], const uint32_t (&r)[8], uint32_t (&out)[8],
                                          uint32_t one) {
    const uint32_t cv[8] = {IV0, IV1, IV2, IV3, IV4, IV5, IV6, IV7};
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        m[i] = l[i];
        m[8 + i] = r[i];
    }
    compress<16>(cv, m, 0u, 0u, 64u, CHUNK_START | CHUNK_END | ROOT, out, one);
}

// Out-of-line copy: the subtree kernels call this from several merge levels; keeping ONE instance of the ~700
// instruction compression keeps the kernel inside the instruction cache (the fully inlined version stalled on
// instruction fetch: smsp no_instruction 2.9 warps per issue, profiles/r1_hash_leaf_after_imad.txt).
static __device__ __noinline__ Digest hash_node_call(Digest l, Digest r, uint32_t one) {
    Digest o;
    hash_node(l.w, r.w, o.w, one);
    return o;
}

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0u, 0x0123); }

// hash of one Int<K> leaf given as LEAF32 = 2*K little-endian u32 words (lo32, hi32 of each u64 limb).
// to_bytes() writes each u64 limb big-endian (int.rs:201-210) and BLAKE3 loads message words little-endian,
// so m[2k] = bswap32(hi32(limb k)), m[2k+1] = bswap32(lo32(limb k)).
template <int LEAF32>
__device__ __forceinline__ void hash_leaf(const uint32_t (&x)[LEAF32], uint32_t (&out)[8], uint32_t one) {
    static_assert(LEAF32 % 2 == 0 && LEAF32 >= 2, "whole u64 limbs");
    const uint32_t iv[8] = {IV0, IV1, IV2, IV3, IV4, IV5, IV6, IV7};
    if constexpr (LEAF32 <= 16) {
        uint32_t m[16];
#pragma unroll
        for (int i = 0; i < 16; i++) m[i] = 0u;
#pragma unroll
        for (int k = 0; k < LEAF32 / 2; k++) {
            m[2 * k] = bswap32(x[2 * k + 1]);
            m[2 * k + 1] = bswap32(x[2 * k]);
        }
        compress<LEAF32>(iv, m, 0u, 0u, LEAF32 * 4u, CHUNK_START | CHUNK_END | ROOT, out, one);
    } else {
        // one chunk (<= 1024 bytes) of several 64-byte blocks
        static_assert(LEAF32 <= 256, "leaf must fit one BLAKE3 chunk");
        constexpr int NB = (LEAF32 + 15) / 16;
        uint32_t cv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) cv[i] = iv[i];
#pragma unroll 1
        for (int b = 0; b < NB; b++) {
            uint32_t m[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int w = 16 * b + i;           // message word index
                const int src = (w & ~1) | ((w & 1) ^ 1);  // swap lo/hi of the limb
                m[i] = (w < LEAF32) ? bswap32(x[src]) : 0u;
            }
            const uint32_t blen = (b == NB - 1) ? (uint32_t)(LEAF32 * 4 - 64 * b) : 64u;
            const uint32_t flags = (b == 0 ? CHUNK_START : 0u) | (b == NB - 1 ? (CHUNK_END | ROOT) : 0u);
            uint32_t o[8];
            compress<16>(cv, m, 0u, 0u, blen, flags, o, one);
#pragma unroll
            for (int i = 0; i < 8; i++) cv[i] = o[i];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) out[i] = cv[i];
    }
}

}  // namespace b3
}  // namespace zipgpu
