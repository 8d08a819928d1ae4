// zinc_b200/csrc/kernels.h -- host-visible launchers of the sm_100a kernels (internal to libzipgpu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace zipgpu {

struct RootsFanout;

// ---- K1: RAA encoder (raa_encode.cu) ----
struct EncodeArgs {
    const uint32_t *evals;   // [num_rows][row_len][2*in_limbs]
    uint32_t *rows_out;      // [num_rows][cw][out32]
    const uint16_t *tab1;    // device tables from build_encode_tables, encode_perm_padded_len(cw) entries each
    const uint16_t *tab2;
    const uint8_t *colw;
    const uint32_t *perm1_raw = nullptr;  // device u32[cw]: perm1 as uploaded (the cw = 16384 commit kernels gather through it)
    const uint32_t *perm2_raw = nullptr;  // device u32[cw]: perm2 as uploaded (commit_wsc.cu)
    void *wsc_scratch = nullptr;          // commit_wsc_scratch_bytes(num_sms) of device memory (commit_wsc.cu: s1 of the rows in flight)
    uint32_t num_rows, row_len, cw, out32;
    int in_limbs;
    int num_sms;
    uint32_t *evals_copy = nullptr;  // non-NULL (exact shapes only): `evals` is mapped pinned HOST memory read in place
                                     // (zero-copy over PCIe); every staged row is also written here, in HBM
    uint32_t *row_counter = nullptr; // device word for dynamic row claiming (armed by the launcher); NULL = static rows
    uint8_t *fuse_layers = nullptr;  // non-NULL: the fused commit kernel also writes the lowest Merkle levels
    int *fused_levels_out = nullptr; // receives the level up to which that launch builds the trees (<= encode_fused_levels():
                                     // the warp-specialised kernel stops 1 or 2 levels lower when it hashes sub-row units)
    // Whole trees in the fused launch (commit_ws.cu, "tops" epilogue): non-NULL tops_roots asks the warp-specialised
    // kernel to finish the trees of the rows it hashed and write their roots here; if it does, *fused_levels_out = depth.
    // With `fan` the same launch also carries the multi-GPU roots exchange (see RootsFanout), *fan_fused reports it.
    uint8_t *tops_roots = nullptr;
    uint32_t *pair_flags = nullptr;  // zeroed device words, one per CTA boundary (rows hashed as two units by two CTAs)
    const RootsFanout *fan = nullptr;
    unsigned long long fan_step = 0;
    uint32_t fan_row_begin = 0;
    bool *fan_fused = nullptr;
    cudaStream_t stream;
};
// the warp-specialised commit kernel (commit_ws.cu) for the encoder configuration (E, T) of an exact Int<1> -> Int<4> shape
// the cw = 16384 form (commit_ws16k.cu): one plane set, the hash warps read the codeword back from global memory
// the cw = 16384 form as a 2-CTA cluster (commit_wsc.cu): two plane sets split over the shared memories of an SM pair
bool commit_wsc_supported(uint32_t row_len, uint32_t cw);
int commit_wsc_levels();
size_t commit_wsc_scratch_bytes(int num_sms);
cudaError_t launch_commit_wsc(const EncodeArgs &a);
bool commit_ws16k_supported(uint32_t row_len, uint32_t cw);
int commit_ws16k_levels();
cudaError_t launch_commit_ws16k(const EncodeArgs &a);
bool commit_ws_supported(int E, int T);
bool commit_ws_whole_trees(uint32_t cw, uint32_t rows);
cudaError_t launch_commit_ws(const EncodeArgs &a, int E, int T, int *fused_levels);
int encode_fused_levels(int in_limbs, int out_limbs, uint32_t row_len, uint32_t cw);
size_t encode_perm_padded_len(uint32_t cw);
void build_encode_tables(const uint32_t *perm1, const uint32_t *perm2, uint32_t row_len, uint32_t cw, int in_limbs,
                         int out_limbs, uint16_t *tab1, uint16_t *tab2, uint8_t *colw);
int encode_compute_limbs(int in_limbs, uint32_t cw);
bool encode_supported(int in_limbs, uint32_t cw, uint32_t row_len);
cudaError_t launch_raa_encode(const EncodeArgs &a);

// ---- K1b: RAA encoder for codewords longer than one SM's shared memory holds (raa_big.cu): chunked scans through
// global scratch; the permutations are used as uploaded ----
struct BigEncodeArgs {
    const uint32_t *evals;    // [num_rows][row_len][2*in_limbs]
    uint32_t *rows_out;       // [num_rows][cw][out32]
    const uint32_t *perm1, *perm2;  // device u32[cw]
    uint32_t *scratch;        // raa_big_plan() bytes
    uint32_t num_rows, row_len, cw, out32, batch_rows;
    int in_limbs;
    uint8_t *fuse_layers = nullptr;  // non-NULL: the fused commit form (leaves + tree levels 1..raa_big_fused_levels())
    int num_sms = 0;          // > 0 allows the row-per-CTA form (one persistent CTA per SM)
    size_t scratch_bytes = 0; // size of `scratch`
    cudaStream_t stream;
};
bool raa_big_supported(int in_limbs, uint32_t cw);
int raa_big_fused_levels(int in_limbs, int out_limbs, uint32_t cw);
void raa_big_plan(int in_limbs, uint32_t cw, uint32_t num_rows, uint32_t *batch_rows, size_t *scratch_bytes);
cudaError_t launch_raa_encode_big(const BigEncodeArgs &a, int *launches);

// ---- encode_f: the RAA code over field elements (encode_f.cu) ----
struct EncodeFArgs {
    const uint32_t *rows_in;  // device [num_rows][row_len][2*limbs] residues < modulus
    uint32_t *out;            // device [num_rows][cw][2*limbs]
    const uint32_t *perm1, *perm2, *modulus;  // device u32[cw], u32[cw], u32[2*limbs]
    uint32_t *scratch;        // device num_rows * cw * 2*limbs words
    uint32_t num_rows, row_len, cw;
    int limbs;                // u64 limbs per field element (1..6) / per output integer (1..8)
    int in_limbs = 0;         // > 0: encode_wide instead -- rows_in entries are Int<in_limbs>, sign-extended to Int<limbs>,
                              // wrap-around adds, `modulus` unused
    cudaStream_t stream;
};
cudaError_t launch_encode_f(const EncodeFArgs &a);

// ---- multi-GPU: the roots exchange of a row-sharded commit (peer_roots.cu, merkle.cu) ----
// One per zipgpu_peer_roots object, in device memory, written once when the peers are connected.  The kernel that
// produces the roots of this GPU's row range stores every root straight into every peer's result buffer (P2P stores
// over NVLink); its last CTA then publishes this rank's step counter in every peer's flag words and waits for theirs.
constexpr int PEER_MAX = 16;
struct RootsFanout {
    uint8_t *bufs[2][PEER_MAX];            // [step parity][rank]: that rank's result buffer (total_rows * 32 bytes)
    unsigned long long *flags[PEER_MAX];   // rank p's PEER_MAX flag words: flags[p][r] = last step rank r published to p
    unsigned int *done;                    // local: CTAs of the publishing kernel that have finished their stores
    unsigned int *status;                  // mapped host word: set to 1 + rank of a peer that did not show up in time
    unsigned long long timeout_ns;         // bound on the wait (the kernel gives up and reports instead of hanging)
    int rank, world;
};

// ---- K2/K3: BLAKE3 leaves + per-row Merkle levels (merkle.cu) ----
struct MerkleArgs {
    const uint32_t *leaves;  // [num_rows][1<<depth][leaf32]
    uint8_t *layers;         // [num_rows][(2<<depth)-2][32]  (may be internal scratch)
    uint8_t *roots;          // [num_rows][32]
    uint32_t num_rows;
    int depth;
    int leaf32;
    cudaStream_t stream;
    int num_sms = 148;
    // row-sharded commit: the launch that produces the roots also stores them at row `fan_row_begin + row` of every
    // peer's result buffer and its last CTA runs the publish/wait handshake of step `fan_step` (see RootsFanout).
    // Only honoured when one launch_merkle_levels call reaches the roots; *fan_fused reports whether it was.
    const RootsFanout *fan = nullptr;
    unsigned long long fan_step = 0;
    uint32_t fan_row_begin = 0;
    bool *fan_fused = nullptr;
};
bool merkle_supported(int leaf32);
// returns the number of kernel launches through *launches
cudaError_t launch_merkle_rows(const MerkleArgs &a, int *launches);
// Runs tree passes starting from level `from_level` (0 = from the raw leaves; > 0 = that level is already in `layers`)
// until a level >= `until_level` is reached (until_level < 0: up to the roots).  *reached = the level produced last.
// Splitting lets a chunked host pipeline run the wide bottom passes per chunk and the narrow, latency-bound top passes
// once over all rows, and lets the fused commit kernel (raa_encode.cu) hand over at level log2(E).
cudaError_t launch_merkle_levels(const MerkleArgs &a, int from_level, int until_level, int *reached, int *launches);

// ---- K4: column openings (open_columns.cu) ----
struct OpenArgs {
    const uint32_t *rows;    // [num_rows][cw][out32]
    const uint8_t *layers;   // [num_rows][(2<<depth)-2][32]
    const uint32_t *columns; // device [num_cols]
    uint32_t *col_values;    // [num_cols][num_rows][out32]
    uint8_t *paths;          // [num_cols][num_rows][depth][32]
    uint32_t num_rows, cw, out32, num_cols;
    int depth;
    cudaStream_t stream;
};
cudaError_t launch_open_columns(const OpenArgs &a);
// the same openings as proof-stream bytes (col_values / paths of OpenArgs unused)
size_t open_columns_wire_bytes(uint32_t num_rows, uint32_t out_limbs, int depth);  // per column
cudaError_t launch_open_columns_wire(const OpenArgs &a, uint8_t *stream_out);

// ---- K5: proximity-test row combination (combine_rows.cu) ----
struct CombineArgs {
    const uint64_t *evals;   // [num_rows][row_len] Int<1> (two's complement u64)
    const uint64_t *coeffs;  // device [num_rows] Int<1>
    uint64_t *scratch;       // combine_rows_scratch_bytes()
    uint64_t *out;           // [row_len][out_limbs]
    uint32_t num_rows, row_len, out_limbs;
    cudaStream_t stream;
};
size_t combine_rows_scratch_bytes(uint32_t num_rows, uint32_t row_len);
cudaError_t launch_combine_rows(const CombineArgs &a, int *launches);

// ---- K6: ZipLinearCode (sparse code) encoder (sparse_encode.cu) ----
struct SparseEncodeArgs {
    const uint64_t *evals;   // [num_rows][row_len] Int<in_limbs>
    uint64_t *rows_out;      // [num_rows][cw] Int<out_limbs>
    uint32_t num_rows, row_len, cw;
    int in_limbs, out_limbs;
    int num_sms;
    // general coefficients: ELL tables transposed to [d][cw] (cell k of codeword entry j at k*cw + j)
    const uint32_t *cols_t = nullptr;
    const int64_t *coef_t = nullptr;
    uint32_t d = 0;
    // all coefficients 0/1 and a shape sparse_gemm_supported() accepts: dense [cw][row_len] bytes, cells per entry,
    // scratch of sparse_planes_bytes()
    const uint8_t *dense = nullptr;
    const uint32_t *nnz = nullptr;
    uint8_t *planes = nullptr;
    bool use_umma = true;    // tcgen05 kernel (sparse_umma.cu); false: the mma.sync kernel
    cudaStream_t stream;
};
cudaError_t launch_sparse_umma(const SparseEncodeArgs &a);  // GEMM stage only, planes already split
bool sparse_gemm_supported(int in_limbs, int out_limbs, uint32_t row_len, uint32_t cw);
size_t sparse_planes_bytes(uint32_t num_rows, uint32_t row_len, int in_limbs);
cudaError_t launch_sparse_encode(const SparseEncodeArgs &a, int *launches);

// ---- multi-GPU: stand-alone form of the roots exchange (peer_roots.cu): copies `nbytes` of local roots into every
// peer's buffer at `offset` and runs the same handshake; used when the roots were not produced by one launch (chunked
// host pipelines that return the layers, ranks without rows) ----
cudaError_t launch_peer_roots_allgather(const RootsFanout *fan, unsigned long long step, const uint8_t *src, size_t offset,
                                        size_t nbytes, cudaStream_t stream);

// ---- INT32 micro-benchmark (microbench.cu) ----
cudaError_t launch_microbench_int32(int kind, int iters, int num_sms, cudaStream_t stream, uint32_t *d_sink,
                                    double *lane_ops);

}  // namespace zipgpu
