/*
 * oracle/zip_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of NethermindEth/zinc's Zip PCS *commit* path.  It exists to CHECK the
 * CUDA product in zinc_b200/ and to serve as the timed CPU baseline of bench.py; nothing under
 * zinc_b200/ may include, link or call it.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Parity pin status (see DESIGN.md "Oracle"):
 *   - BLAKE3 + Int<N>::to_bytes + tree layout: PINNED against the `blake3` crate (Python binding 1.0.8 of
 *     the same Rust crate the reference depends on) through tests/golden/ fixtures and the KATs in
 *     SURVEY.md 8(c); the reference itself ships no digest known-answer vectors for this path.
 *   - repeat / accumulate / widening: PINNED against the reference's own unit-test examples
 *     (code_raa.rs:199-244, zip/utils.rs:163-234).
 *   - shuffle_seeded (rand 0.9.2 StdRng + SliceRandom::shuffle): PARITY UNPINNED.  The crate is a
 *     Cargo dependency absent from /root/reference and no Rust toolchain exists here; the algorithm is
 *     restated from the published crate sources (rand_core 0.9 seed_from_u64, rand_chacha 0.9 ChaCha12,
 *     rand 0.9 IncreasingUniform + Canon's method).  The ChaCha core is validated at 20 rounds against
 *     OpenSSL's ChaCha20.  Because of this the permutations are *inputs* at the product boundary.
 */
#ifndef ZIP_ORACLE_H
#define ZIP_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Int<N>: N u64 limbs, least-significant first, two's complement (field/int.rs:23-25,230-232) ---- */
void zo_widen(const uint64_t *in, int in_limbs, uint64_t *out, int out_limbs);      /* int.rs:194-199 */
int  zo_add_assign(uint64_t *acc, const uint64_t *rhs, int limbs);                  /* int.rs:122-134; returns 1 on signed overflow (reference would panic) */
void zo_int_to_bytes(const uint64_t *v, int limbs, uint8_t *out);                   /* int.rs:201-210 */

/* ---- RAA code pieces (zip/code_raa.rs) ---- */
void zo_repeat(const uint64_t *in, size_t row_len, int in_limbs, size_t rep, uint64_t *out, int out_limbs); /* code_raa.rs:142-152 */
int  zo_accumulate(uint64_t *v, size_t n, int limbs);                                                        /* code_raa.rs:164-171 */
int  zo_raa_width_ok(int in_limbs, int out_limbs, size_t poly_size, size_t rep);                             /* code_raa.rs:53-72 */
size_t zo_raa_row_len(size_t poly_size);                                                                     /* code_raa.rs:42-43 */
size_t zo_num_rows(size_t poly_size, size_t row_len);                                                        /* pcs/structs.rs:79-90 */

/* ---- shuffle_seeded (zip/utils.rs:139-142) = rand 0.9.2 StdRng::seed_from_u64 + SliceRandom::shuffle ---- */
void zo_shuffle_seeded(void *slice, size_t n, size_t elem_bytes, uint64_t seed);
void zo_perm_from_seed(uint32_t *idx, size_t n, uint64_t seed);   /* the same shuffle applied to 0..n: shuffled[i] = original[idx[i]] */
void zo_chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]);
void zo_seed_from_u64(uint64_t state, uint32_t key[8]);
/* raw StdRng stream, for tests */
void zo_stdrng_words(uint64_t seed, uint32_t *out, size_t n);

/* ---- encode (code_raa.rs:89-105, commit.rs:158-183) ---- */
/* literal restatement: repeat -> shuffle_seeded(seed1) -> accumulate -> shuffle_seeded(seed2) -> accumulate */
int zo_encode_row_seeded(const uint64_t *row, size_t row_len, int in_limbs, size_t rep,
                         uint64_t seed1, uint64_t seed2, uint64_t *out, int out_limbs);
/* same with the two permutations given as gather index arrays (perm[i] = source index) */
int zo_encode_row_perm(const uint64_t *row, size_t row_len, int in_limbs, size_t rep,
                       const uint32_t *perm1, const uint32_t *perm2, uint64_t *out, int out_limbs,
                       uint64_t *scratch /* cw*out_limbs u64 */);
int zo_encode_rows_perm(const uint64_t *evals, size_t num_rows, size_t row_len, int in_limbs, size_t rep,
                        const uint32_t *perm1, const uint32_t *perm2, uint64_t *rows_out, int out_limbs);

/* ---- BLAKE3 (crate blake3 1.8.2: pcs/utils.rs:68-70,90,107-111) ---- */
void zo_blake3(const uint8_t *in, size_t len, uint8_t out[32]);

/* ---- MerkleTree::new (pcs/utils.rs:74-118) ----
 * layers_out: (2<<depth)-2 digests: [leaf hashes | level depth-1 | ... | level 1]; root_out: 32 bytes. */
int zo_merkle_tree_new(size_t depth, const uint64_t *leaves, size_t num_leaves, int leaf_limbs,
                       uint8_t *layers_out, uint8_t *root_out);
/* MerkleProof::create_proof / verify (pcs/utils.rs:163-210) */
void zo_merkle_create_proof(size_t depth, const uint8_t *layers, size_t leaf, uint8_t *path_out /* depth*32 */);
int  zo_merkle_verify(size_t depth, const uint8_t *path, const uint8_t root[32],
                      const uint64_t *leaf_value, int leaf_limbs, size_t leaf_index);

/* ---- commit (commit.rs:50-87) ----
 * rows_out: num_rows*cw*out_limbs u64 (nullable -> internal scratch); layers_out nullable;
 * roots_out: num_rows*32 bytes.  returns 0 ok, -1 overflow, -2 bad shape. */
int zo_commit_perm(const uint64_t *evals, size_t num_rows, size_t row_len, int in_limbs, size_t rep,
                   const uint32_t *perm1, const uint32_t *perm2, int out_limbs,
                   uint64_t *rows_out, uint8_t *layers_out, uint8_t *roots_out);

/* Multithreaded commit used as the CPU baseline (bench.py cpu_baseline / --impl reference).
 * Row chunks over `threads` threads like commit.rs:164-180 / zip/utils.rs:36-52.
 * faithful=1 regenerates both permutations from the seeds for every row with swaps of out_limbs-wide
 * elements, as the reference does (code_raa.rs:98-102); faithful=0 uses the precomputed gather arrays.
 * Returns 0 ok. */
int zo_commit_mt(const uint64_t *evals, size_t num_rows, size_t row_len, int in_limbs, size_t rep,
                 uint64_t seed1, uint64_t seed2, const uint32_t *perm1, const uint32_t *perm2, int out_limbs,
                 uint64_t *rows_out, uint8_t *layers_out, uint8_t *roots_out, int threads, int faithful);

/* ---- ZipLinearCode, the sparse code (zip/code.rs:77-215).  The two sampled matrices are inputs:
 * cols/coef = SparseMatrixZ::cells (code.rs:271-296) in order, d cells per matrix row, n = codeword_len/2 rows.
 * Coefficients are passed as i64 (KeccakTranscript draws 0/1, transcript.rs:176-181; the tests' MockTranscript
 * a counter, pcs/tests.rs:30-33).  Arithmetic is mod 2^(64*out_limbs); the reference's checked ops never wrap
 * for the sizes it is used at. ---- */
int zo_sparse_mat_vec(size_t n, size_t m, size_t d, const uint32_t *cols, const int64_t *coef,
                      const uint64_t *vec, int in_limbs, uint64_t *out, int out_limbs);          /* code.rs:299-321 */
int zo_sparse_commit_mt(const uint64_t *evals, size_t num_rows, size_t row_len, int in_limbs, size_t n, size_t d,
                        const uint32_t *cols_a, const int64_t *coef_a, const uint32_t *cols_b, const int64_t *coef_b,
                        int out_limbs, uint64_t *rows_out, uint8_t *layers_out, uint8_t *roots_out, int threads);

#ifdef __cplusplus
}
#endif
#endif
