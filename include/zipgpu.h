/*
 * include/zipgpu.h -- C ABI of libzipgpu: the B200 (sm_100a) drop-in for the *commit* path of zinc's Zip PCS.
 *
 * The reference (NethermindEth/zinc, pure Rust) has no FFI seam; the seam this library is shaped for is the
 * generic Rust API itself (citations relative to /root/reference):
 *
 *   MultilinearZip::commit            src/zip/pcs/commit.rs:50-87     -> zipgpu_commit / zipgpu_commit_device / zipgpu_commit_resident
 *   MultilinearZip::commit_no_merkle  src/zip/pcs/commit.rs:104-119   -> zipgpu_encode_rows / zipgpu_encode_rows_device
 *   MultilinearZip::batch_commit      src/zip/pcs/commit.rs:134-142   -> zipgpu_batch_commit
 *   MultilinearZip::encode_rows       src/zip/pcs/commit.rs:158-183   -> zipgpu_encode_rows
 *   RaaCode::{new,encode_inner}       src/zip/code_raa.rs:35-105      -> zipgpu_code_create (+ zipgpu_perm_from_seed)
 *   ZipLinearCode::{new,encode_wide}  src/zip/code.rs:100-147,186-201,299-321 -> zipgpu_sparse_code_create
 *   MerkleTree::new                   src/zip/pcs/utils.rs:74-118     -> zipgpu_merkle_rows / zipgpu_merkle_rows_device
 *   MerkleProof::create_proof,
 *   ColumnOpening::open_at_column     src/zip/pcs/utils.rs:163-176,221-233, open_z.rs:124-143 -> zipgpu_data_open_columns
 *   PcsTranscript::write_integers /
 *   write_merkle_proof (wire bytes)   src/zip/pcs_transcript.rs:115-135,198-211           -> zipgpu_data_open_columns_wire
 *   combine_rows (proximity test)     src/zip/utils.rs:94-127, open_z.rs:100-113        -> zipgpu_data_combine_rows / zipgpu_combine_rows_device
 *
 * Conventions
 *   - Integers are the in-memory layout of `[Int<n>]`: n little-endian u64 limbs per value, least significant
 *     limb first (field/int.rs:23-25,230-232).  `evals` is row-major: row i = evals[i*row_len .. (i+1)*row_len).
 *   - `rows` (u-hat) is num_rows*cw values of out_limbs limbs, row-major (structs.rs:33-38).
 *   - `layers` is, per row, (2<<depth)-2 BLAKE3 digests of 32 bytes laid out as MerkleTree::layers after the
 *     root is popped (pcs/utils.rs:77-85): [leaf hashes (cw) | level depth-1 (cw/2) | ... | level 1 (2)];
 *     rows are concatenated.  `roots` is num_rows*32 bytes (structs.rs:40-45).
 *   - The two RAA permutations are inputs: perm[i] = index the shuffled slice takes element i from, i.e. the
 *     array obtained by applying `shuffle_seeded(&mut v, seed)` (zip/utils.rs:139-142) to v = [0,1,..,cw).
 *     A Rust host passes exactly that; zipgpu_perm_from_seed is a restatement of rand 0.9.2 for hosts
 *     without the crate (parity with the real crate unpinned).
 *   - Every function returns 0 on success or a negative zipgpu_status; it never aborts or unwinds.
 *     zipgpu_last_error() gives a thread-local message for the last failure.
 *   - `*_device` entry points take device pointers and enqueue on `stream` (a cudaStream_t passed as void*,
 *     NULL = the context's own stream) without synchronising.  The context's stream is NON-BLOCKING: it does not
 *     order itself against the legacy default stream, so a caller that prepares buffers on another stream passes
 *     that stream here or synchronises first (zipgpu_ctx_sync waits for the context's streams).  The others take HOST pointers, perform the
 *     host<->device copies themselves (pipelined with the kernels) and return when the outputs are valid.
 *   - A context (zipgpu_ctx) is bound to one GPU.  Multi-GPU comes in two forms, both sharding a commit by contiguous
 *     row range and a batch_commit by polynomial (rows are independent through encode AND hash, commit.rs:71-81):
 *       zipgpu_mgpu_*                ONE process drives n GPUs (what `ZincProver` needs, zinc/prover.rs:313-315): a worker
 *                                    thread and a context per device, slices copied over all PCIe links at once;
 *       zipgpu_*_sharded + zipgpu_peer_roots_*   one process (or thread) per GPU, e.g. under torch.distributed.
 *     In both the only exchange step -- every GPU receives every other GPU's 32-byte roots -- is part of the kernel
 *     that produces the roots (P2P stores over NVLink), not a separate collective.  See INTEGRATION.md 4.
 */
#ifndef ZIPGPU_H
#define ZIPGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zipgpu_ctx zipgpu_ctx;
typedef struct zipgpu_code zipgpu_code;
typedef struct zipgpu_data zipgpu_data;

typedef enum {
    ZIPGPU_OK = 0,
    ZIPGPU_ERR_INVALID = -1,     /* bad argument / shape (maps to Error::InvalidPcsParam or the reference's asserts) */
    ZIPGPU_ERR_CUDA = -2,        /* CUDA runtime failure */
    ZIPGPU_ERR_NOMEM = -3,       /* device or pinned-host allocation failed */
    ZIPGPU_ERR_UNSUPPORTED = -4, /* shape outside what the kernels implement */
    ZIPGPU_ERR_NO_DEVICE = -5,   /* no CUDA device / wrong architecture */
    ZIPGPU_ERR_WIDTH = -6,       /* code_raa.rs:68-72: out type too narrow for the codeword entries */
    ZIPGPU_ERR_PEER_TIMEOUT = -7 /* multi-GPU roots exchange: a peer GPU did not publish its roots in time */
} zipgpu_status;

/* ---- library / context ------------------------------------------------------------------------------- */
const char *zipgpu_version(void);
const char *zipgpu_last_error(void);
int zipgpu_device_count(int *count);

int zipgpu_ctx_create(int device, zipgpu_ctx **out);
void zipgpu_ctx_destroy(zipgpu_ctx *ctx);
int zipgpu_ctx_device(const zipgpu_ctx *ctx);
int zipgpu_ctx_sync(zipgpu_ctx *ctx);
/* number of zipgpu kernels launched by this context so far (bench.py's gpu_launches) */
uint64_t zipgpu_ctx_launch_count(const zipgpu_ctx *ctx);

/* pinned host memory for the host-pointer entry points (pageable pointers also work, through a staging copy).
 * zipgpu_host_alloc: from 2 MiB up, a 2 MiB-aligned anonymous mapping with transparent huge pages requested,
 * pre-faulted and page-locked (cudaHostRegister, portable) -- the DMA rate of a pinned buffer depends on what backs
 * it (measured: 20-27 GB/s from a buffer cudaHostAlloc carved out of a fragmented process, 46-55 GB/s from 2 MiB
 * pages); smaller requests and failures fall back to cudaHostAlloc. */
int zipgpu_host_alloc(size_t bytes, void **out);
int zipgpu_host_free(void *p);
int zipgpu_host_register(void *p, size_t bytes);
int zipgpu_host_unregister(void *p);

/* ---- code (RaaCode, per-pp state) -------------------------------------------------------------------- */
/* shuffle_seeded(&mut [0..n), seed) restated (zip/utils.rs:139-142; rand 0.9.2).  Host-side, no GPU needed. */
int zipgpu_perm_from_seed(uint64_t seed, uint32_t n, uint32_t *perm_out);
/* The ChaCha block function underneath it (rand_chacha 0.9 state layout: 8 key words, 64-bit counter, stream 0), so
 * that the core can be checked against published ChaCha vectors (StdRng uses rounds = 12).  out: 16 words. */
int zipgpu_chacha_block(const uint32_t *key, uint64_t counter, int rounds, uint32_t *out);
/* RaaCode::new geometry (code_raa.rs:42-43), MultilinearZip::setup (structs.rs:79-90). Host-side. */
size_t zipgpu_raa_row_len(size_t poly_size);
size_t zipgpu_num_rows(size_t poly_size, size_t row_len);
/* bits a codeword entry needs (code_raa.rs:53-67) */
int zipgpu_raa_codeword_width_bits(int in_limbs, size_t poly_size, size_t repetition_factor);

/* Uploads the permutations once per pp.  perm1/perm2: cw = row_len*repetition_factor entries each, validated
 * to be permutations of [0,cw).  in_limbs/out_limbs: limbs of ZT::N and ZT::K (1 and 4 for INT_LIMBS=1).
 * ZIPGPU_ERR_WIDTH if 64*out_limbs < 64*in_limbs + 2*ceil(log2(cw)) (the reference's width assert can then
 * not hold for any poly_size with this cw). */
int zipgpu_code_create(zipgpu_ctx *ctx, size_t row_len, size_t repetition_factor, int in_limbs, int out_limbs,
                       const uint32_t *perm1, const uint32_t *perm2, zipgpu_code **out);
/* ZipLinearCode, the sparse code: `ZipLinearCode::new` / `new_multilinear` src/zip/code.rs:100-147 with the two
 * sampled matrices as INPUTS (like the RAA permutations): cols_x / coef_x are `SparseMatrixZ::cells`
 * (code.rs:271-296) of matrix a and b in order -- codeword_len/2 matrix rows of cells_per_row (column, coefficient)
 * cells each.  Coefficients are i64: KeccakTranscript::get_encoding_element draws 0 or 1 (src/transcript.rs:176-181),
 * the reference tests' MockTranscript a counter (src/zip/pcs/tests.rs:30-33).  The handle is a zipgpu_code: every
 * entry point below (encode_rows = `ZipLinearCode::encode_wide` code.rs:186-201 per row, commit, batch_commit,
 * commit_resident, ...) takes it unchanged.  Sums are exact modulo 2^(64*out_limbs) (the reference's checked
 * arithmetic never wraps at the sizes it is used at). */
int zipgpu_sparse_code_create(zipgpu_ctx *ctx, size_t row_len, size_t codeword_len, size_t cells_per_row, int in_limbs,
                              int out_limbs, const uint32_t *cols_a, const int64_t *coef_a, const uint32_t *cols_b,
                              const int64_t *coef_b, zipgpu_code **out);
/* -1: an RAA code; 0: sparse code on the generic kernel; 1: sparse code on the tensor-core kernel (all coefficients
 * in 0..255 and row_len, codeword_len multiples of 128) */
int zipgpu_code_sparse_kind(const zipgpu_code *code);
void zipgpu_code_destroy(zipgpu_code *code);
size_t zipgpu_code_row_len(const zipgpu_code *code);
size_t zipgpu_code_codeword_len(const zipgpu_code *code);
/* merkle depth = codeword_len.next_power_of_two().ilog2()  (commit.rs:67) */
int zipgpu_code_merkle_depth(const zipgpu_code *code);

/* ---- encode_rows / commit_no_merkle (commit.rs:104-119,158-183) -------------------------------------- */
int zipgpu_encode_rows(zipgpu_code *code, size_t num_rows, const uint64_t *evals, uint64_t *rows_out);
int zipgpu_encode_rows_device(zipgpu_code *code, size_t num_rows, const uint64_t *d_evals, uint64_t *d_rows_out,
                              void *stream);

/* ---- encode_f (code_raa.rs:133-138): the RAA code over field elements, as the verifier re-encodes the combined row
 * (verify_z.rs:141-142).  Same permutations, every addition is the field's (RandomField AddAssign: add the stored
 * residues, subtract the modulus once if the sum overflowed or is >= modulus; field/arithmetic.rs:66-77,
 * field/config.rs:53-76).  Elements are `limbs` u64 words, least significant first -- the stored BigInt<N> of a
 * RandomField<N> value (Montgomery form or not: addition does not care) -- and must be < modulus (checked).
 *   rows: num_rows * row_len * limbs u64 (host);  out: num_rows * codeword_len * limbs u64 (host). */
int zipgpu_encode_f(zipgpu_code *code, size_t num_rows, int limbs, const uint64_t *modulus, const uint64_t *rows,
                    uint64_t *out);

/* ---- encode_wide (code_raa.rs:125-131) for any In = Int<in_limbs>, Out = Int<out_limbs>, 1 <= in <= out <= 8: entries
 * sign-extended (Out::from(&In), int.rs:194-199), two's-complement adds at the output width.  The verifier runs it on
 * the combined row of every proximity test with In = Out = M = Int<8> (verify_z.rs:74-78); a latency kernel (one CTA per
 * row), not the prover's encoder.  Uses the permutations of `code`; its own in/out limbs do not matter.
 *   rows: num_rows * row_len * in_limbs u64 (host);  out: num_rows * codeword_len * out_limbs u64 (host). */
int zipgpu_encode_wide(zipgpu_code *code, size_t num_rows, int in_limbs, int out_limbs, const uint64_t *rows,
                       uint64_t *out);

/* ---- MerkleTree::new batched over rows (pcs/utils.rs:74-118) ----------------------------------------- */
/* leaves: num_rows * (1<<depth) values of leaf_limbs limbs.  layers_out nullable.  roots_out: num_rows*32. */
int zipgpu_merkle_rows(zipgpu_ctx *ctx, size_t num_rows, int depth, int leaf_limbs, const uint64_t *leaves,
                       uint8_t *layers_out, uint8_t *roots_out);
int zipgpu_merkle_rows_device(zipgpu_ctx *ctx, size_t num_rows, int depth, int leaf_limbs, const uint64_t *d_leaves,
                              uint8_t *d_layers_out, uint8_t *d_roots_out, void *stream);

/* ---- commit (commit.rs:50-87) ------------------------------------------------------------------------ */
/* host buffers; rows_out / layers_out nullable (then they are not copied back; see zipgpu_commit_resident) */
int zipgpu_commit(zipgpu_code *code, size_t num_rows, const uint64_t *evals, uint64_t *rows_out, uint8_t *layers_out,
                  uint8_t *roots_out);
/* device buffers; d_rows_out / d_layers_out nullable (internal scratch is used) */
int zipgpu_commit_device(zipgpu_code *code, size_t num_rows, const uint64_t *d_evals, uint64_t *d_rows_out,
                         uint8_t *d_layers_out, uint8_t *d_roots_out, void *stream);
/* batch_commit (commit.rs:134-142): num_polys polynomials sharing one pp; evals[p] -> outputs[p] (arrays of
 * host pointers; rows_out / layers_out may be NULL, or hold NULL entries) */
int zipgpu_batch_commit(zipgpu_code *code, size_t num_polys, size_t num_rows, const uint64_t *const *evals,
                        uint64_t *const *rows_out, uint8_t *const *layers_out, uint8_t *const *roots_out);

/* ---- device-resident prover data (MultilinearZipData, structs.rs:33-38) ------------------------------ */
/* host evals in, host roots out; rows + layers stay on the GPU behind *handle for the opening phase */
int zipgpu_commit_resident(zipgpu_code *code, size_t num_rows, const uint64_t *evals, uint8_t *roots_out,
                           zipgpu_data **handle);
void zipgpu_data_free(zipgpu_data *data);
size_t zipgpu_data_num_rows(const zipgpu_data *data);
const uint64_t *zipgpu_data_rows_device(const zipgpu_data *data);
const uint8_t *zipgpu_data_layers_device(const zipgpu_data *data);
const uint8_t *zipgpu_data_roots_device(const zipgpu_data *data);
/* row-sharded commits (zipgpu_commit_resident_sharded): ALL roots of the commitment, total_rows*32 bytes; else NULL */
const uint8_t *zipgpu_data_all_roots_device(const zipgpu_data *data);
/* copy a row range of `rows` / `layers` back to the host */
int zipgpu_data_read_rows(const zipgpu_data *data, size_t row_begin, size_t row_count, uint64_t *rows_out);
int zipgpu_data_read_layers(const zipgpu_data *data, size_t row_begin, size_t row_count, uint8_t *layers_out);
/* Column openings (open_z.rs:124-143 + pcs/utils.rs:163-176,221-233): for each requested column j, the
 * num_rows column entries rows[i*cw + j] (out_limbs limbs each) and, per row, the Merkle path of `depth`
 * digests in create_proof order (sibling at the leaf level first).
 *   col_values_out: num_cols * num_rows * out_limbs u64;  paths_out: num_cols * num_rows * depth * 32 bytes */
int zipgpu_data_open_columns(const zipgpu_data *data, size_t num_cols, const uint32_t *columns,
                             uint64_t *col_values_out, uint8_t *paths_out);
/* The same for a shard of a larger commitment: the outputs are laid out for `total_rows` rows per column and this
 * data's rows land at rows [row_offset, row_offset + num_rows) of every column (multi-GPU openings in row order). */
int zipgpu_data_open_columns_strided(const zipgpu_data *data, size_t num_cols, const uint32_t *columns,
                                     uint64_t *col_values_out, uint8_t *paths_out, size_t total_rows, size_t row_offset);

/* The same openings as the exact byte stream `open` appends to the proof for these columns (open_z.rs:124-143 over
 * PcsTranscript::write_integers / write_merkle_proof, pcs_transcript.rs:115-135,198-211): per column the num_rows
 * entries as little-endian u64 limbs, then per row be64(depth) followed by the depth path digests.
 *   stream_out: num_cols * zipgpu_data_open_columns_wire_bytes(data) bytes (host). */
size_t zipgpu_data_open_columns_wire_bytes(const zipgpu_data *data);
int zipgpu_data_open_columns_wire(const zipgpu_data *data, size_t num_cols, const uint32_t *columns, uint8_t *stream_out);

/* Proximity-test row combination (open_z.rs:100-113; zip/utils.rs:94-127): combined[col] = sum_i coeffs[i] *
 * evals[i*row_len + col], exact signed integer arithmetic, every operand expanded to out_limbs limbs as
 * `expand::<N, M>` does (zip/utils.rs:129-137).  Int<1> operands only (in_limbs == 1); out_limbs >= 3.
 *   coeffs: num_rows u64 (two's complement);  combined_out: row_len * out_limbs u64.
 * zipgpu_data_combine_rows works on the evaluations kept behind a zipgpu_commit_resident handle (host coeffs in,
 * host row out); the _device form takes device pointers (scratch is taken from the context's allocator). */
int zipgpu_data_combine_rows(const zipgpu_data *data, const uint64_t *coeffs, int out_limbs, uint64_t *combined_out);
int zipgpu_combine_rows_device(zipgpu_ctx *ctx, size_t num_rows, size_t row_len, const uint64_t *d_evals,
                               const uint64_t *d_coeffs, int out_limbs, uint64_t *d_combined_out, void *stream);

/* ---- multi-GPU, one process (or thread) per GPU: the roots exchange of a row-sharded commit ------------------
 * A commit sharded by row range over G GPUs of one node has one exchange step: the list of roots,
 * MultilinearZipCommitment::roots (structs.rs:40-45), must be complete on every GPU.  A zipgpu_peer_roots object holds
 * this rank's result buffers (2 x total_rows*32 bytes, double-buffered by step) and flag words, and the peers' as
 * mapped over NVLink.  Every exchange is one "step"; all ranks must run the same number of steps.
 *   create : allocates the buffers; fills ipc_out (ZIPGPU_IPC_BYTES bytes, nullable) for ranks in other processes;
 *   connect: ipc_all = the ZIPGPU_IPC_BYTES-byte blocks of ALL ranks in rank order (exchanged by the host, e.g.
 *            torch.distributed.all_gather); opens the peers' buffers (cudaIpcOpenMemHandle);
 *   connect_local: all = the objects of ranks 0..n-1, created in THIS process on different contexts; uses direct peer
 *            access (cudaDeviceEnablePeerAccess), no IPC;
 *   zipgpu_commit_device_sharded: commit of rows [row_begin, row_begin+count) of the commitment from device-resident
 *            evaluations; the kernel that produces the roots stores each one into every peer's result buffer and its
 *            last CTA publishes this rank's step counter and waits for the peers' (st.release.sys / ld.acquire.sys).
 *            Enqueued on `stream`; when the stream has passed this point *d_all_roots_out (returned immediately, valid
 *            until the call after next) holds all total_rows roots.  d_rows_out / d_layers_out nullable (scratch);
 *   zipgpu_commit_resident_sharded: the same from HOST evaluations of the row range (pinned or pageable), prover data
 *            kept behind *handle (nullable; NULL is returned for count == 0), all roots copied to roots_all_out
 *            (host, total_rows*32 bytes, nullable).  Synchronises and reports ZIPGPU_ERR_PEER_TIMEOUT;
 *   allgather: the stand-alone exchange for roots that already sit in device memory (d_local_roots: count*32 bytes,
 *            16-byte aligned);
 *   status : 0, or ZIPGPU_ERR_PEER_TIMEOUT if a kernel gave up waiting for a peer (bounded wait, default 20 s,
 *            ZIPGPU_PEER_TIMEOUT_MS); check after synchronising.  After a timeout recreate the object on all ranks. */
#define ZIPGPU_IPC_BYTES 64
typedef struct zipgpu_peer_roots zipgpu_peer_roots;
int zipgpu_peer_roots_create(zipgpu_ctx *ctx, size_t total_rows, int rank, int world, zipgpu_peer_roots **out,
                             uint8_t *ipc_out);
int zipgpu_peer_roots_connect(zipgpu_peer_roots *pr, const uint8_t *ipc_all);
int zipgpu_peer_roots_connect_local(zipgpu_peer_roots *const *all, int n);
int zipgpu_peer_roots_allgather(zipgpu_peer_roots *pr, size_t row_begin, size_t count, const uint8_t *d_local_roots,
                                void *stream, uint8_t **d_all_out);
int zipgpu_peer_roots_status(zipgpu_peer_roots *pr);
void zipgpu_peer_roots_destroy(zipgpu_peer_roots *pr);
int zipgpu_commit_device_sharded(zipgpu_code *code, zipgpu_peer_roots *pr, size_t row_begin, size_t count,
                                 const uint64_t *d_evals, uint64_t *d_rows_out, uint8_t *d_layers_out, void *stream,
                                 uint8_t **d_all_roots_out);
int zipgpu_commit_resident_sharded(zipgpu_code *code, zipgpu_peer_roots *pr, size_t row_begin, size_t count,
                                   const uint64_t *evals, uint8_t *roots_all_out, zipgpu_data **handle);

/* ---- multi-GPU, ONE process: a multi-device context behind the same commit API ---------------------------------
 * What the reference's caller needs (one process: zinc/prover.rs:313-315 calls RaaCode::new, setup, commit): the
 * device list is given once, every call below shards its work over those GPUs (a worker thread and a zipgpu_ctx per
 * device; host slices go over all PCIe links concurrently) and returns results in the reference's order.
 *   commit / commit_resident / encode_rows: GPU g of n takes rows [g*R/n, (g+1)*R/n) (balanced);
 *   batch_commit: polynomial p goes to device p mod n (commit.rs:134-142 has no cross-polynomial dependency);
 *   roots: exchanged between the GPUs inside the roots-producing kernel (see zipgpu_peer_roots above), so every
 *          device holds the complete commitment; the host copy is ONE D2H from device 0;
 *   data_open_columns / open_columns_wire: every GPU extracts its rows' entries and paths, output in row order
 *          (the row -> column redistribution of `open`, open_z.rs:124-143: only the opened columns ever move);
 *   data_combine_rows: per-GPU partial sums over its rows, added exactly on the host. */
typedef struct zipgpu_mgpu zipgpu_mgpu;
typedef struct zipgpu_mgpu_code zipgpu_mgpu_code;
typedef struct zipgpu_mgpu_data zipgpu_mgpu_data;
/* devices == NULL: the first n visible devices (n <= 0: all of them) */
int zipgpu_mgpu_create(const int *devices, int n, zipgpu_mgpu **out);
void zipgpu_mgpu_destroy(zipgpu_mgpu *m);
int zipgpu_mgpu_num_devices(const zipgpu_mgpu *m);
zipgpu_ctx *zipgpu_mgpu_ctx(zipgpu_mgpu *m, int index);
/* kernels launched so far by all of its contexts */
uint64_t zipgpu_mgpu_launch_count(const zipgpu_mgpu *m);
/* same arguments and errors as zipgpu_code_create; the tables are uploaded to every device */
int zipgpu_mgpu_code_create(zipgpu_mgpu *m, size_t row_len, size_t repetition_factor, int in_limbs, int out_limbs,
                            const uint32_t *perm1, const uint32_t *perm2, zipgpu_mgpu_code **out);
void zipgpu_mgpu_code_destroy(zipgpu_mgpu_code *code);
zipgpu_code *zipgpu_mgpu_code_device(zipgpu_mgpu_code *code, int index);
/* commit_no_merkle / encode_rows (commit.rs:104-119,158-183), host buffers */
int zipgpu_mgpu_encode_rows(zipgpu_mgpu_code *code, size_t num_rows, const uint64_t *evals, uint64_t *rows_out);
/* commit (commit.rs:50-87), host buffers; rows_out / layers_out nullable */
int zipgpu_mgpu_commit(zipgpu_mgpu_code *code, size_t num_rows, const uint64_t *evals, uint64_t *rows_out,
                       uint8_t *layers_out, uint8_t *roots_out);
/* batch_commit (commit.rs:134-142); arrays as in zipgpu_batch_commit */
int zipgpu_mgpu_batch_commit(zipgpu_mgpu_code *code, size_t num_polys, size_t num_rows, const uint64_t *const *evals,
                             uint64_t *const *rows_out, uint8_t *const *layers_out, uint8_t *const *roots_out);
/* commit with the prover data left on the GPUs that produced it */
int zipgpu_mgpu_commit_resident(zipgpu_mgpu_code *code, size_t num_rows, const uint64_t *evals, uint8_t *roots_out,
                                zipgpu_mgpu_data **handle);
void zipgpu_mgpu_data_free(zipgpu_mgpu_data *data);
size_t zipgpu_mgpu_data_num_rows(const zipgpu_mgpu_data *data);
/* the shard held by device `index`: its zipgpu_data (NULL if it owns no rows) and row range */
zipgpu_data *zipgpu_mgpu_data_shard(const zipgpu_mgpu_data *data, int index, size_t *row_begin, size_t *row_count);
/* ALL roots of the commitment as they sit on device `index` after the exchange (num_rows*32 bytes; NULL if that
 * device owns no rows) */
const uint8_t *zipgpu_mgpu_data_roots_device(const zipgpu_mgpu_data *data, int index);
/* as zipgpu_data_open_columns / _wire / _combine_rows, over all shards, outputs in row order */
int zipgpu_mgpu_data_open_columns(const zipgpu_mgpu_data *data, size_t num_cols, const uint32_t *columns,
                                  uint64_t *col_values_out, uint8_t *paths_out);
size_t zipgpu_mgpu_data_open_columns_wire_bytes(const zipgpu_mgpu_data *data);
int zipgpu_mgpu_data_open_columns_wire(const zipgpu_mgpu_data *data, size_t num_cols, const uint32_t *columns,
                                       uint8_t *stream_out);
int zipgpu_mgpu_data_combine_rows(const zipgpu_mgpu_data *data, const uint64_t *coeffs, int out_limbs,
                                  uint64_t *combined_out);

/* ---- measurement helpers (used by bench.py; no reference counterpart) -------------------------------- */
/* When enabled, every commit/encode/merkle call records CUDA events around its kernels on the launch stream. */
int zipgpu_profile_enable(zipgpu_ctx *ctx, int on);
/* Sums since the last reset, milliseconds of device time: encoder kernel, hash kernels; and call count. */
int zipgpu_profile_read(zipgpu_ctx *ctx, double *encode_ms, double *hash_ms, uint64_t *calls, int reset);
/* Dependency-free INT32 micro-benchmark: runs `iters` rounds of `kind` ops per thread on the whole GPU and
 * returns achieved warp-level lane-ops/s.  kind 0 = 3 alu + 1 IMAD.IADD, 1 = 8 alu : 6 IMAD (a BLAKE3 G), 2 = LOP3/SHF
 * only: the alu pipe's own peak (the roofline of the hash kernels). */
int zipgpu_microbench_int32(zipgpu_ctx *ctx, int kind, int iters, double *lane_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* ZIPGPU_H */
