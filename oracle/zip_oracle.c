/*
 * oracle/zip_oracle.c -- TEST INFRASTRUCTURE ONLY (see zip_oracle.h for scope, citations and pin status).
 *
 * Plain-C CPU restatement of zinc's Zip commit: RAA encode over Int<N> limbs, BLAKE3 leaves, one Merkle
 * tree per encoded row.  Used only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
 */
#define _GNU_SOURCE
#include "zip_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * Int<N> arithmetic (field/int.rs over crypto_bigint::Int<N>)
 * ---------------------------------------------------------------------------------------------- */

/* int.rs:194-199 `From<&Int<M>> for Int<N>` -> crypto_bigint resize: sign-extending (pinned by
 * zip/utils.rs:208-234), zero-padding for non-negative values (zip/utils.rs:163-206). */
void zo_widen(const uint64_t *in, int in_limbs, uint64_t *out, int out_limbs) {
    uint64_t fill = (in[in_limbs - 1] >> 63) ? ~(uint64_t)0 : 0;
    for (int i = 0; i < out_limbs; i++) out[i] = i < in_limbs ? in[i] : fill;
}

/* int.rs:122-134 `AddAssign`: limb-wise add with carry, LSW first.  crypto-bigint's `+=` is an
 * overflow-checked add; the width assert in code_raa.rs:53-72 makes overflow unreachable on the commit
 * path.  We wrap and report the signed-overflow condition so tests can assert it never fires. */
int zo_add_assign(uint64_t *acc, const uint64_t *rhs, int limbs) {
    unsigned __int128 c = 0;
    uint64_t sa = acc[limbs - 1] >> 63, sb = rhs[limbs - 1] >> 63;
    for (int i = 0; i < limbs; i++) {
        c += (unsigned __int128)acc[i] + rhs[i];
        acc[i] = (uint64_t)c;
        c >>= 64;
    }
    uint64_t sr = acc[limbs - 1] >> 63;
    return (sa == sb) && (sr != sa);
}

/* int.rs:201-210 `ToBytes`: words in LSW->MSW order, each word big-endian. */
void zo_int_to_bytes(const uint64_t *v, int limbs, uint8_t *out) {
    for (int i = 0; i < limbs; i++)
        for (int b = 0; b < 8; b++) out[i * 8 + b] = (uint8_t)(v[i] >> (56 - 8 * b));
}

/* ------------------------------------------------------------------------------------------------
 * RAA pieces (zip/code_raa.rs)
 * ---------------------------------------------------------------------------------------------- */

/* code_raa.rs:142-152: out[j] = Out::from(&in[j mod row_len]) for j < row_len*rep */
void zo_repeat(const uint64_t *in, size_t row_len, int in_limbs, size_t rep, uint64_t *out, int out_limbs) {
    for (size_t j = 0; j < row_len * rep; j++)
        zo_widen(in + (j % row_len) * in_limbs, in_limbs, out + j * out_limbs, out_limbs);
}

/* code_raa.rs:164-171: for i in 1..n { v[i] += v[i-1] } */
int zo_accumulate(uint64_t *v, size_t n, int limbs) {
    int ov = 0;
    for (size_t i = 1; i < n; i++) ov |= zo_add_assign(v + i * limbs, v + (i - 1) * limbs, limbs);
    return ov;
}

static size_t ilog2_sz(size_t x) { size_t l = 0; while (x >>= 1) l++; return l; }
static size_t next_pow2_sz(size_t x) { size_t p = 1; while (p < x) p <<= 1; return p; }
static uint64_t isqrt_u64(uint64_t x) {
    uint64_t r = 0;
    for (uint64_t bit = (uint64_t)1 << 31; bit; bit >>= 1) { uint64_t t = r | bit; if (t * t <= x) r = t; }
    return r;
}

/* code_raa.rs:42-43: row_len = isqrt(1 << num_vars).next_power_of_two() */
size_t zo_raa_row_len(size_t poly_size) {
    size_t num_vars = ilog2_sz(poly_size);
    return next_pow2_sz((size_t)isqrt_u64((uint64_t)1 << num_vars));
}

/* pcs/structs.rs:82: num_rows = ((1 << num_vars) / row_len).next_power_of_two() */
size_t zo_num_rows(size_t poly_size, size_t row_len) {
    size_t num_vars = ilog2_sz(poly_size);
    return next_pow2_sz(((size_t)1 << num_vars) / row_len);
}

/* code_raa.rs:53-72: K::num_bits() >= N::num_bits() + num_vars_even + 2*log2(next_pow2(rep)) */
int zo_raa_width_ok(int in_limbs, int out_limbs, size_t poly_size, size_t rep) {
    size_t num_vars = ilog2_sz(poly_size);
    size_t nv_even = (num_vars % 2 == 0) ? num_vars : num_vars + 1;
    size_t need = (size_t)in_limbs * 64 + nv_even + 2 * ilog2_sz(next_pow2_sz(rep));
    return (size_t)out_limbs * 64 >= need;
}

/* ------------------------------------------------------------------------------------------------
 * shuffle_seeded (zip/utils.rs:139-142): StdRng::seed_from_u64(seed); slice.shuffle(&mut rng)
 * Third-party, absent from /root/reference: rand 0.9.2 / rand_chacha 0.9 / rand_core 0.9 (Cargo.toml:34).
 * PARITY UNPINNED against the real crate (no Rust toolchain, no vectors in the reference).
 * ---------------------------------------------------------------------------------------------- */

/* rand_core 0.9 SeedableRng::seed_from_u64: PCG32 stream fills the 32-byte seed, 4 bytes at a time. */
void zo_seed_from_u64(uint64_t state, uint32_t key[8]) {
    for (int i = 0; i < 8; i++) {
        state = state * 6364136223846793005ULL + 11634580027462260723ULL;
        uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31)); /* LE bytes -> LE word */
    }
}

#define ROTL32(x, n) (((x) << (n)) | ((x) >> (32 - (n))))
#define CHACHA_QR(a, b, c, d) \
    a += b; d ^= a; d = ROTL32(d, 16); c += d; b ^= c; b = ROTL32(b, 12); \
    a += b; d ^= a; d = ROTL32(d, 8);  c += d; b ^= c; b = ROTL32(b, 7);

/* rand_chacha 0.9: constants "expand 32-byte k", 256-bit key, 64-bit block counter (words 12,13),
 * 64-bit stream id (words 14,15). */
void zo_chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                      key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t x[16];
    memcpy(x, s, sizeof x);
    for (int r = 0; r < rounds; r += 2) {
        CHACHA_QR(x[0], x[4], x[8], x[12]) CHACHA_QR(x[1], x[5], x[9], x[13])
        CHACHA_QR(x[2], x[6], x[10], x[14]) CHACHA_QR(x[3], x[7], x[11], x[15])
        CHACHA_QR(x[0], x[5], x[10], x[15]) CHACHA_QR(x[1], x[6], x[11], x[12])
        CHACHA_QR(x[2], x[7], x[8], x[13]) CHACHA_QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

/* StdRng = ChaCha12Rng; BlockRng buffers 4 blocks (64 words) and hands words out in order. */
typedef struct {
    uint32_t key[8];
    uint64_t counter;
    uint32_t buf[64];
    int idx;
} zo_stdrng;

static void stdrng_init(zo_stdrng *r, uint64_t seed) {
    zo_seed_from_u64(seed, r->key);
    r->counter = 0;
    r->idx = 64;
}
static inline uint32_t stdrng_next_u32(zo_stdrng *r) {
    if (r->idx >= 64) {
        for (int b = 0; b < 4; b++) zo_chacha_block(r->key, r->counter + b, 0, 12, r->buf + 16 * b);
        r->counter += 4;
        r->idx = 0;
    }
    return r->buf[r->idx++];
}
void zo_stdrng_words(uint64_t seed, uint32_t *out, size_t n) {
    zo_stdrng r;
    stdrng_init(&r, seed);
    for (size_t i = 0; i < n; i++) out[i] = stdrng_next_u32(&r);
}

/* rand 0.9 UniformInt<u32>::sample_single_inclusive (Canon's method, biased variant) for `..bound`. */
static inline uint32_t random_below_u32(zo_stdrng *r, uint32_t bound) {
    uint64_t m = (uint64_t)stdrng_next_u32(r) * bound;
    uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
    if (lo > (uint32_t)(0u - bound)) {
        uint32_t new_hi = (uint32_t)(((uint64_t)stdrng_next_u32(r) * bound) >> 32);
        if ((uint64_t)lo + new_hi > 0xffffffffULL) hi += 1;
    }
    return hi;
}

/* rand 0.9 seq::increasing_uniform::IncreasingUniform: draws one u32 per "chunk" of consecutive indices */
typedef struct {
    zo_stdrng *rng;
    uint32_t n, chunk;
    uint8_t chunk_remaining;
} zo_incr_uniform;

static void calc_bound_u32(uint32_t m, uint32_t *bound, uint8_t *count) {
    uint32_t product = m, current = m + 1;
    for (;;) {
        uint64_t p = (uint64_t)product * current;
        if (p > 0xffffffffULL) { *bound = product; *count = (uint8_t)(current - m); return; }
        product = (uint32_t)p;
        current++;
    }
}
static inline size_t incr_next_index(zo_incr_uniform *u) {
    uint32_t next_n = u->n + 1;
    uint8_t next_remaining;
    if (u->chunk_remaining > 0) {
        next_remaining = u->chunk_remaining - 1;
    } else {
        uint32_t bound; uint8_t remaining;
        calc_bound_u32(next_n, &bound, &remaining);
        u->chunk = random_below_u32(u->rng, bound);
        next_remaining = remaining - 1;
    }
    size_t result;
    if (next_remaining == 0) {
        result = u->chunk;
    } else {
        result = u->chunk % next_n;
        u->chunk /= next_n;
    }
    u->chunk_remaining = next_remaining;
    u->n = next_n;
    return result;
}

/* rand 0.9 SliceRandom::shuffle -> partial_shuffle(len): for i in 0..len { swap(i, chooser.next_index()) } */
void zo_shuffle_seeded(void *slice, size_t n, size_t elem_bytes, uint64_t seed) {
    if (n <= 1) return;
    zo_stdrng rng;
    stdrng_init(&rng, seed);
    zo_incr_uniform u = {&rng, 0, 0, 1}; /* n = 0 -> the first index is 0 without a draw */
    uint8_t *p = (uint8_t *)slice;
    uint8_t tmp[512];
    for (size_t i = 0; i < n; i++) {
        size_t j = incr_next_index(&u);
        if (i != j) {
            if (elem_bytes == 4) {
                uint32_t t = ((uint32_t *)p)[i]; ((uint32_t *)p)[i] = ((uint32_t *)p)[j]; ((uint32_t *)p)[j] = t;
            } else {
                memcpy(tmp, p + i * elem_bytes, elem_bytes);
                memcpy(p + i * elem_bytes, p + j * elem_bytes, elem_bytes);
                memcpy(p + j * elem_bytes, tmp, elem_bytes);
            }
        }
    }
}

void zo_perm_from_seed(uint32_t *idx, size_t n, uint64_t seed) {
    for (size_t i = 0; i < n; i++) idx[i] = (uint32_t)i;
    zo_shuffle_seeded(idx, n, 4, seed);
}

/* ------------------------------------------------------------------------------------------------
 * encode (code_raa.rs:89-105)
 * ---------------------------------------------------------------------------------------------- */

int zo_encode_row_seeded(const uint64_t *row, size_t row_len, int in_limbs, size_t rep,
                         uint64_t seed1, uint64_t seed2, uint64_t *out, int out_limbs) {
    size_t cw = row_len * rep;
    int ov = 0;
    zo_repeat(row, row_len, in_limbs, rep, out, out_limbs);               /* code_raa.rs:98 */
    zo_shuffle_seeded(out, cw, (size_t)out_limbs * 8, seed1);             /* :99 */
    ov |= zo_accumulate(out, cw, out_limbs);                              /* :100 */
    zo_shuffle_seeded(out, cw, (size_t)out_limbs * 8, seed2);             /* :101 */
    ov |= zo_accumulate(out, cw, out_limbs);                              /* :102 */
    return ov ? -1 : 0;
}

int zo_encode_row_perm(const uint64_t *row, size_t row_len, int in_limbs, size_t rep,
                       const uint32_t *perm1, const uint32_t *perm2, uint64_t *out, int out_limbs,
                       uint64_t *scratch) {
    size_t cw = row_len * rep;
    int ov = 0;
    /* repeat then permute == gather from the row at perm1[i] mod row_len */
    for (size_t i = 0; i < cw; i++)
        zo_widen(row + (perm1[i] % row_len) * in_limbs, in_limbs, scratch + i * out_limbs, out_limbs);
    ov |= zo_accumulate(scratch, cw, out_limbs);
    for (size_t i = 0; i < cw; i++)
        memcpy(out + i * out_limbs, scratch + (size_t)perm2[i] * out_limbs, (size_t)out_limbs * 8);
    ov |= zo_accumulate(out, cw, out_limbs);
    return ov ? -1 : 0;
}

/* commit.rs:158-183 (semantic definition pinned by commit.rs:356-380: row i = encode_wide(evals[i*row_len..])) */
int zo_encode_rows_perm(const uint64_t *evals, size_t num_rows, size_t row_len, int in_limbs, size_t rep,
                        const uint32_t *perm1, const uint32_t *perm2, uint64_t *rows_out, int out_limbs) {
    size_t cw = row_len * rep;
    uint64_t *scratch = (uint64_t *)malloc(cw * out_limbs * 8);
    int rc = 0;
    for (size_t r = 0; r < num_rows; r++)
        if (zo_encode_row_perm(evals + r * row_len * in_limbs, row_len, in_limbs, rep, perm1, perm2,
                               rows_out + r * cw * out_limbs, out_limbs, scratch))
            rc = -1;
    free(scratch);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * BLAKE3 (third-party crate blake3 1.8.2; algorithm per the BLAKE3 specification)
 * ---------------------------------------------------------------------------------------------- */

static const uint32_t B3_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                  0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t B3_PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum { B3_CHUNK_START = 1, B3_CHUNK_END = 2, B3_PARENT = 4, B3_ROOT = 8 };

#define ROTR32(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
#define B3_G(a, b, c, d, mx, my) \
    a = a + b + (mx); d = ROTR32(d ^ a, 16); c = c + d; b = ROTR32(b ^ c, 12); \
    a = a + b + (my); d = ROTR32(d ^ a, 8);  c = c + d; b = ROTR32(b ^ c, 7);

static void b3_compress_portable(const uint32_t cv[8], const uint32_t block[16], uint64_t counter,
                                 uint32_t block_len, uint32_t flags, uint32_t out_cv[8]) {
    uint32_t s[16] = {cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], cv[6], cv[7],
                      B3_IV[0], B3_IV[1], B3_IV[2], B3_IV[3],
                      (uint32_t)counter, (uint32_t)(counter >> 32), block_len, flags};
    uint32_t m[16], t[16];
    memcpy(m, block, sizeof m);
    for (int r = 0; r < 7; r++) {
        B3_G(s[0], s[4], s[8], s[12], m[0], m[1])   B3_G(s[1], s[5], s[9], s[13], m[2], m[3])
        B3_G(s[2], s[6], s[10], s[14], m[4], m[5])  B3_G(s[3], s[7], s[11], s[15], m[6], m[7])
        B3_G(s[0], s[5], s[10], s[15], m[8], m[9])  B3_G(s[1], s[6], s[11], s[12], m[10], m[11])
        B3_G(s[2], s[7], s[8], s[13], m[12], m[13]) B3_G(s[3], s[4], s[9], s[14], m[14], m[15])
        for (int i = 0; i < 16; i++) t[i] = m[B3_PERM[i]];
        memcpy(m, t, sizeof m);
    }
    for (int i = 0; i < 8; i++) out_cv[i] = s[i] ^ s[i + 8];
}

#if defined(__x86_64__)
/* Row-vectorised single-block compression (state rows in 4 xmm registers), so that the CPU baseline
 * is not handicapped relative to the blake3 crate's SIMD single-block path. */
#define B3_VEC_BODY(ROT16, ROT12, ROT8, ROT7)                                                         \
    __m128i r0 = _mm_loadu_si128((const __m128i *)cv), r1 = _mm_loadu_si128((const __m128i *)(cv + 4)); \
    __m128i r2 = _mm_loadu_si128((const __m128i *)B3_IV);                                              \
    __m128i r3 = _mm_set_epi32((int)flags, (int)block_len, (int)(counter >> 32), (int)counter);        \
    uint32_t m[16], t[16];                                                                             \
    memcpy(m, block, sizeof m);                                                                        \
    for (int r = 0; r < 7; r++) {                                                                      \
        __m128i mx = _mm_set_epi32((int)m[6], (int)m[4], (int)m[2], (int)m[0]);                        \
        __m128i my = _mm_set_epi32((int)m[7], (int)m[5], (int)m[3], (int)m[1]);                        \
        r0 = _mm_add_epi32(_mm_add_epi32(r0, r1), mx); r3 = _mm_xor_si128(r3, r0); r3 = ROT16(r3);     \
        r2 = _mm_add_epi32(r2, r3); r1 = _mm_xor_si128(r1, r2); r1 = ROT12(r1);                        \
        r0 = _mm_add_epi32(_mm_add_epi32(r0, r1), my); r3 = _mm_xor_si128(r3, r0); r3 = ROT8(r3);      \
        r2 = _mm_add_epi32(r2, r3); r1 = _mm_xor_si128(r1, r2); r1 = ROT7(r1);                         \
        /* diagonalise: rotate rows 1,2,3 left by 1,2,3 lanes */                                       \
        r1 = _mm_shuffle_epi32(r1, _MM_SHUFFLE(0, 3, 2, 1));                                           \
        r2 = _mm_shuffle_epi32(r2, _MM_SHUFFLE(1, 0, 3, 2));                                           \
        r3 = _mm_shuffle_epi32(r3, _MM_SHUFFLE(2, 1, 0, 3));                                           \
        mx = _mm_set_epi32((int)m[14], (int)m[12], (int)m[10], (int)m[8]);                             \
        my = _mm_set_epi32((int)m[15], (int)m[13], (int)m[11], (int)m[9]);                             \
        r0 = _mm_add_epi32(_mm_add_epi32(r0, r1), mx); r3 = _mm_xor_si128(r3, r0); r3 = ROT16(r3);     \
        r2 = _mm_add_epi32(r2, r3); r1 = _mm_xor_si128(r1, r2); r1 = ROT12(r1);                        \
        r0 = _mm_add_epi32(_mm_add_epi32(r0, r1), my); r3 = _mm_xor_si128(r3, r0); r3 = ROT8(r3);      \
        r2 = _mm_add_epi32(r2, r3); r1 = _mm_xor_si128(r1, r2); r1 = ROT7(r1);                         \
        r1 = _mm_shuffle_epi32(r1, _MM_SHUFFLE(2, 1, 0, 3));                                           \
        r2 = _mm_shuffle_epi32(r2, _MM_SHUFFLE(1, 0, 3, 2));                                           \
        r3 = _mm_shuffle_epi32(r3, _MM_SHUFFLE(0, 3, 2, 1));                                           \
        for (int i = 0; i < 16; i++) t[i] = m[B3_PERM[i]];                                             \
        memcpy(m, t, sizeof m);                                                                        \
    }                                                                                                  \
    _mm_storeu_si128((__m128i *)out_cv, _mm_xor_si128(r0, r2));                                        \
    _mm_storeu_si128((__m128i *)(out_cv + 4), _mm_xor_si128(r1, r3));

#define SSE_ROT(x, n) _mm_or_si128(_mm_srli_epi32(x, n), _mm_slli_epi32(x, 32 - (n)))
#define SSE_ROT16(x) _mm_shuffle_epi8(x, _mm_set_epi8(13, 12, 15, 14, 9, 8, 11, 10, 5, 4, 7, 6, 1, 0, 3, 2))
#define SSE_ROT8(x) _mm_shuffle_epi8(x, _mm_set_epi8(12, 15, 14, 13, 8, 11, 10, 9, 4, 7, 6, 5, 0, 3, 2, 1))
#define SSE_ROT12(x) SSE_ROT(x, 12)
#define SSE_ROT7(x) SSE_ROT(x, 7)
__attribute__((target("sse4.1,ssse3"))) static void b3_compress_sse41(
    const uint32_t cv[8], const uint32_t block[16], uint64_t counter, uint32_t block_len, uint32_t flags,
    uint32_t out_cv[8]) {
    B3_VEC_BODY(SSE_ROT16, SSE_ROT12, SSE_ROT8, SSE_ROT7)
}
#define AVX_ROT16(x) _mm_ror_epi32(x, 16)
#define AVX_ROT12(x) _mm_ror_epi32(x, 12)
#define AVX_ROT8(x) _mm_ror_epi32(x, 8)
#define AVX_ROT7(x) _mm_ror_epi32(x, 7)
__attribute__((target("avx512f,avx512vl"))) static void b3_compress_avx512(
    const uint32_t cv[8], const uint32_t block[16], uint64_t counter, uint32_t block_len, uint32_t flags,
    uint32_t out_cv[8]) {
    B3_VEC_BODY(AVX_ROT16, AVX_ROT12, AVX_ROT8, AVX_ROT7)
}
#endif

typedef void (*b3_compress_fn)(const uint32_t *, const uint32_t *, uint64_t, uint32_t, uint32_t, uint32_t *);
static b3_compress_fn b3_compress_impl = 0;
static int b3_force_portable = 0;
void zo_blake3_force_portable(int on) { b3_force_portable = on; b3_compress_impl = 0; }

static inline void b3_compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter,
                               uint32_t block_len, uint32_t flags, uint32_t out_cv[8]) {
    b3_compress_fn f = b3_compress_impl;
    if (!f) {
        f = b3_compress_portable;
#if defined(__x86_64__)
        if (!b3_force_portable) {
            __builtin_cpu_init();
            if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl")) f = b3_compress_avx512;
            else if (__builtin_cpu_supports("sse4.1") && __builtin_cpu_supports("ssse3")) f = b3_compress_sse41;
        }
#endif
        b3_compress_impl = f;
    }
    f(cv, block, counter, block_len, flags, out_cv);
}

static void b3_load_block(const uint8_t *in, size_t len, uint32_t block[16]) {
    uint8_t buf[64] = {0};
    memcpy(buf, in, len);
    for (int i = 0; i < 16; i++)
        block[i] = (uint32_t)buf[4 * i] | ((uint32_t)buf[4 * i + 1] << 8) | ((uint32_t)buf[4 * i + 2] << 16) |
                   ((uint32_t)buf[4 * i + 3] << 24);
}

/* chaining value of one chunk (<= 1024 bytes); root_flag is B3_ROOT when the chunk is the whole input */
static void b3_chunk_cv(const uint8_t *in, size_t len, uint64_t chunk_counter, uint32_t root_flag, uint32_t cv[8]) {
    memcpy(cv, B3_IV, 32);
    size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
    for (size_t b = 0; b < nblocks; b++) {
        size_t blen = (b + 1 == nblocks) ? len - 64 * b : 64;
        uint32_t block[16], flags = 0;
        b3_load_block(in + 64 * b, blen, block);
        if (b == 0) flags |= B3_CHUNK_START;
        if (b + 1 == nblocks) flags |= B3_CHUNK_END | root_flag;
        b3_compress(cv, block, chunk_counter, (uint32_t)blen, flags, cv);
    }
}

static void b3_subtree_cv(const uint8_t *in, size_t len, uint64_t chunk_counter, uint32_t root_flag, uint32_t cv[8]) {
    if (len <= 1024) { b3_chunk_cv(in, len, chunk_counter, root_flag, cv); return; }
    /* left subtree: the largest power-of-two number of chunks that leaves at least one byte on the right */
    size_t left_chunks = 1;
    while (left_chunks * 2 * 1024 < len) left_chunks *= 2;
    size_t left_len = left_chunks * 1024;
    uint32_t block[16];
    b3_subtree_cv(in, left_len, chunk_counter, 0, block);
    b3_subtree_cv(in + left_len, len - left_len, chunk_counter + left_chunks, 0, block + 8);
    b3_compress(B3_IV, block, 0, 64, B3_PARENT | root_flag, cv);
}

void zo_blake3(const uint8_t *in, size_t len, uint8_t out[32]) {
    uint32_t cv[8];
    b3_subtree_cv(in, len, 0, B3_ROOT, cv);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)cv[i]; out[4 * i + 1] = (uint8_t)(cv[i] >> 8);
        out[4 * i + 2] = (uint8_t)(cv[i] >> 16); out[4 * i + 3] = (uint8_t)(cv[i] >> 24);
    }
}

/* ------------------------------------------------------------------------------------------------
 * MerkleTree (zip/pcs/utils.rs:66-118) and MerkleProof (:152-211)
 * ---------------------------------------------------------------------------------------------- */

int zo_merkle_tree_new(size_t depth, const uint64_t *leaves, size_t num_leaves, int leaf_limbs,
                       uint8_t *layers_out, uint8_t *root_out) {
    /* utils.rs:75-76: assert power of two and == 1 << depth (reference panics) */
    if (num_leaves == 0 || (num_leaves & (num_leaves - 1)) || num_leaves != ((size_t)1 << depth)) return -2;
    size_t total = ((size_t)2 << depth) - 1; /* utils.rs:77 */
    uint8_t *layers = (uint8_t *)malloc(total * 32);
    uint8_t bytes[1024];
    if ((size_t)leaf_limbs * 8 > sizeof bytes) { free(layers); return -2; }
    /* compute_leaves_hashes, utils.rs:87-93: hash(leaf.to_bytes()) */
    for (size_t j = 0; j < num_leaves; j++) {
        zo_int_to_bytes(leaves + j * leaf_limbs, leaf_limbs, bytes);
        zo_blake3(bytes, (size_t)leaf_limbs * 8, layers + 32 * j);
    }
    /* merklize_leaves_hashes, utils.rs:95-118: next[k] = hash(cur[2k] || cur[2k+1]) */
    size_t offset = 0;
    for (size_t d = depth; d >= 1; d--) {
        size_t width = (size_t)1 << d;
        const uint8_t *cur = layers + 32 * offset;
        uint8_t *next = layers + 32 * (offset + width);
        for (size_t k = 0; k < width / 2; k++) zo_blake3(cur + 64 * k, 64, next + 32 * k);
        offset += width;
    }
    /* utils.rs:80-84: root = layers.pop() */
    memcpy(root_out, layers + 32 * (total - 1), 32);
    if (layers_out) memcpy(layers_out, layers, 32 * (total - 1));
    free(layers);
    return 0;
}

/* utils.rs:163-176 */
void zo_merkle_create_proof(size_t depth, const uint8_t *layers, size_t leaf, uint8_t *path_out) {
    size_t offset = 0, n = 0;
    for (size_t d = depth; d >= 1; d--) {
        size_t width = (size_t)1 << d;
        size_t idx = (leaf >> (depth - d)) ^ 1;
        memcpy(path_out + 32 * n++, layers + 32 * (offset + idx), 32);
        offset += width;
    }
}

/* utils.rs:178-210.  NOTE: for depth >= 1 the last path element of create_proof is a level-1 sibling, and
 * the final hash is compared with the root. */
int zo_merkle_verify(size_t depth, const uint8_t *path, const uint8_t root[32],
                     const uint64_t *leaf_value, int leaf_limbs, size_t leaf_index) {
    uint8_t cur[32], buf[64], bytes[1024];
    zo_int_to_bytes(leaf_value, leaf_limbs, bytes);
    zo_blake3(bytes, (size_t)leaf_limbs * 8, cur);
    size_t index = leaf_index;
    for (size_t i = 0; i < depth; i++) {
        if ((index & 1) == 0) { memcpy(buf, cur, 32); memcpy(buf + 32, path + 32 * i, 32); }
        else { memcpy(buf, path + 32 * i, 32); memcpy(buf + 32, cur, 32); }
        zo_blake3(buf, 64, cur);
        index /= 2;
    }
    return memcmp(cur, root, 32) == 0 ? 0 : -1;
}

/* ------------------------------------------------------------------------------------------------
 * commit (zip/pcs/commit.rs:50-87)
 * ---------------------------------------------------------------------------------------------- */

int zo_commit_perm(const uint64_t *evals, size_t num_rows, size_t row_len, int in_limbs, size_t rep,
                   const uint32_t *perm1, const uint32_t *perm2, int out_limbs,
                   uint64_t *rows_out, uint8_t *layers_out, uint8_t *roots_out) {
    size_t cw = row_len * rep;
    size_t depth = ilog2_sz(next_pow2_sz(cw)); /* commit.rs:67 */
    uint64_t *rows = rows_out ? rows_out : (uint64_t *)malloc(num_rows * cw * out_limbs * 8);
    int rc = zo_encode_rows_perm(evals, num_rows, row_len, in_limbs, rep, perm1, perm2, rows, out_limbs); /* :69 */
    size_t per_row = ((size_t)2 << depth) - 2;
    for (size_t r = 0; r < num_rows && rc == 0; r++) /* :71-74 */
        if (zo_merkle_tree_new(depth, rows + r * cw * out_limbs, cw, out_limbs,
                               layers_out ? layers_out + r * per_row * 32 : 0, roots_out + 32 * r))
            rc = -2;
    if (!rows_out) free(rows);
    return rc;
}

/* ---- multithreaded baseline ---- */
typedef struct {
    const uint64_t *evals;
    size_t row_begin, row_end, row_len, rep;
    int in_limbs, out_limbs, faithful, rc;
    uint64_t seed1, seed2;
    const uint32_t *perm1, *perm2;
    uint64_t *rows_out;
    uint8_t *layers_out, *roots_out;
} zo_mt_job;

static void *zo_mt_worker(void *arg) {
    zo_mt_job *j = (zo_mt_job *)arg;
    size_t cw = j->row_len * j->rep, depth = ilog2_sz(next_pow2_sz(cw));
    size_t per_row = ((size_t)2 << depth) - 2;
    uint64_t *scratch = (uint64_t *)malloc(cw * j->out_limbs * 8);
    uint64_t *rowbuf = (uint64_t *)malloc(cw * j->out_limbs * 8);
    for (size_t r = j->row_begin; r < j->row_end; r++) {
        const uint64_t *row = j->evals + r * j->row_len * j->in_limbs;
        uint64_t *out = j->rows_out ? j->rows_out + r * cw * j->out_limbs : rowbuf;
        int rc = j->faithful
                     ? zo_encode_row_seeded(row, j->row_len, j->in_limbs, j->rep, j->seed1, j->seed2, out, j->out_limbs)
                     : zo_encode_row_perm(row, j->row_len, j->in_limbs, j->rep, j->perm1, j->perm2, out, j->out_limbs, scratch);
        if (rc) j->rc = rc;
        if (j->roots_out &&
            zo_merkle_tree_new(depth, out, cw, j->out_limbs,
                               j->layers_out ? j->layers_out + r * per_row * 32 : 0, j->roots_out + 32 * r))
            j->rc = -2;
    }
    free(scratch);
    free(rowbuf);
    return 0;
}

int zo_commit_mt(const uint64_t *evals, size_t num_rows, size_t row_len, int in_limbs, size_t rep,
                 uint64_t seed1, uint64_t seed2, const uint32_t *perm1, const uint32_t *perm2, int out_limbs,
                 uint64_t *rows_out, uint8_t *layers_out, uint8_t *roots_out, int threads, int faithful) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > num_rows) threads = (int)num_rows;
    size_t rows_per_thread = (num_rows + threads - 1) / threads; /* commit.rs:164 */
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    zo_mt_job *jobs = (zo_mt_job *)calloc(threads, sizeof(zo_mt_job));
    int rc = 0, started = 0;
    for (int t = 0; t < threads; t++) {
        size_t b = (size_t)t * rows_per_thread, e = b + rows_per_thread;
        if (b >= num_rows) break;
        if (e > num_rows) e = num_rows;
        zo_mt_job j = {evals, b, e, row_len, rep, in_limbs, out_limbs, faithful, 0, seed1, seed2,
                       perm1, perm2, rows_out, layers_out, roots_out};
        jobs[t] = j;
        pthread_create(&tid[t], 0, zo_mt_worker, &jobs[t]);
        started++;
    }
    for (int t = 0; t < started; t++) {
        pthread_join(tid[t], 0);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    free(tid);
    free(jobs);
    return rc;
}

/* ---- ZipLinearCode (zip/code.rs:77-215): the sparse code.  The two sampled matrices are INPUTS here, like the RAA
 * permutations: `cols`/`coef` are SparseMatrixZ::cells (code.rs:271-296) in order, d cells per matrix row. ---- */

/* acc += coef * sext(v)   (mod 2^(64*out_limbs)); expand::<L,M>(coeff) * expand::<N,M>(value), code.rs:314 */
static void zo_mul_add(uint64_t *acc, int out_limbs, const uint64_t *v, int in_limbs, int64_t coef) {
    uint64_t w[16], prod[16];
    if (coef == 0) return; /* the product is zero; the reference still multiplies (code.rs:314) */
    zo_widen(v, in_limbs, w, out_limbs);
    if (coef == 1) { /* the only non-zero value KeccakTranscript draws: a plain widening add ("tuned" CPU baseline) */
        uint64_t c1 = 0;
        for (int i = 0; i < out_limbs; i++) {
            unsigned __int128 t = (unsigned __int128)acc[i] + w[i] + c1;
            acc[i] = (uint64_t)t;
            c1 = (uint64_t)(t >> 64);
        }
        return;
    }
    uint64_t mag = coef < 0 ? (uint64_t)0 - (uint64_t)coef : (uint64_t)coef;
    unsigned __int128 carry = 0;
    for (int i = 0; i < out_limbs; i++) {
        unsigned __int128 t = (unsigned __int128)w[i] * mag + carry;
        prod[i] = (uint64_t)t;
        carry = t >> 64;
    }
    if (coef < 0) { /* two's complement negate */
        uint64_t c = 1;
        for (int i = 0; i < out_limbs; i++) {
            uint64_t x = ~prod[i] + c;
            c = (c && x == 0) ? 1 : 0;
            prod[i] = x;
        }
    }
    uint64_t c = 0;
    for (int i = 0; i < out_limbs; i++) {
        unsigned __int128 t = (unsigned __int128)acc[i] + prod[i] + c;
        acc[i] = (uint64_t)t;
        c = (uint64_t)(t >> 64);
    }
}

/* SparseMatrixZ::mat_vec_mul (code.rs:299-321) */
int zo_sparse_mat_vec(size_t n, size_t m, size_t d, const uint32_t *cols, const int64_t *coef,
                      const uint64_t *vec, int in_limbs, uint64_t *out, int out_limbs) {
    if (out_limbs > 16 || in_limbs > out_limbs) return -2;
    for (size_t i = 0; i < n; i++) {
        uint64_t *acc = out + i * out_limbs;
        memset(acc, 0, 8 * (size_t)out_limbs);
        for (size_t k = 0; k < d; k++) {
            uint32_t c = cols[i * d + k];
            if (c >= m) return -2;
            zo_mul_add(acc, out_limbs, vec + (size_t)c * in_limbs, in_limbs, coef[i * d + k]);
        }
    }
    return 0;
}

typedef struct {
    const uint64_t *evals;
    size_t row_begin, row_end, row_len, n, d;
    int in_limbs, out_limbs, rc;
    const uint32_t *cols_a, *cols_b;
    const int64_t *coef_a, *coef_b;
    uint64_t *rows_out;
    uint8_t *layers_out, *roots_out;
} zo_sparse_job;

static void *zo_sparse_worker(void *arg) {
    zo_sparse_job *j = (zo_sparse_job *)arg;
    size_t cw = 2 * j->n, depth = ilog2_sz(next_pow2_sz(cw));
    size_t per_row = ((size_t)2 << depth) - 2;
    uint64_t *rowbuf = j->rows_out ? 0 : (uint64_t *)malloc(cw * j->out_limbs * 8);
    for (size_t r = j->row_begin; r < j->row_end; r++) {
        const uint64_t *row = j->evals + r * j->row_len * j->in_limbs;
        uint64_t *out = j->rows_out ? j->rows_out + r * cw * j->out_limbs : rowbuf;
        /* encode_wide code.rs:186-201: a.mat_vec_mul(row) then b.mat_vec_mul(row) */
        if (zo_sparse_mat_vec(j->n, j->row_len, j->d, j->cols_a, j->coef_a, row, j->in_limbs, out, j->out_limbs) ||
            zo_sparse_mat_vec(j->n, j->row_len, j->d, j->cols_b, j->coef_b, row, j->in_limbs,
                              out + j->n * j->out_limbs, j->out_limbs))
            j->rc = -2;
        if (j->roots_out && cw == ((size_t)1 << depth) &&
            zo_merkle_tree_new(depth, out, cw, j->out_limbs,
                               j->layers_out ? j->layers_out + r * per_row * 32 : 0, j->roots_out + 32 * r))
            j->rc = -2;
    }
    free(rowbuf);
    return 0;
}

/* encode_rows (commit.rs:158-183) + one MerkleTree per row (commit.rs:71-74) with ZipLinearCode as the code.
 * n = codeword_len/2 matrix rows, d cells per row; roots_out NULL = encode only. */
int zo_sparse_commit_mt(const uint64_t *evals, size_t num_rows, size_t row_len, int in_limbs, size_t n, size_t d,
                        const uint32_t *cols_a, const int64_t *coef_a, const uint32_t *cols_b, const int64_t *coef_b,
                        int out_limbs, uint64_t *rows_out, uint8_t *layers_out, uint8_t *roots_out, int threads) {
    if (num_rows == 0) return 0;
    if (threads < 1) threads = 1;
    if ((size_t)threads > num_rows) threads = (int)num_rows;
    size_t rows_per_thread = (num_rows + threads - 1) / threads;
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    zo_sparse_job *jobs = (zo_sparse_job *)calloc(threads, sizeof(zo_sparse_job));
    int rc = 0, started = 0;
    for (int t = 0; t < threads; t++) {
        size_t b = (size_t)t * rows_per_thread, e = b + rows_per_thread;
        if (b >= num_rows) break;
        if (e > num_rows) e = num_rows;
        zo_sparse_job j = {evals, b, e, row_len, n, d, in_limbs, out_limbs, 0, cols_a, cols_b, coef_a, coef_b,
                           rows_out, layers_out, roots_out};
        jobs[t] = j;
        pthread_create(&tid[t], 0, zo_sparse_worker, &jobs[t]);
        started++;
    }
    for (int t = 0; t < started; t++) {
        pthread_join(tid[t], 0);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    free(tid);
    free(jobs);
    return rc;
}
