// zinc_b200/csrc/open_columns.cu -- K4: column openings from device-resident prover data.
//
// Replaces, for the columns squeezed by `open` (open_z.rs:124-143):
//   column values   rows.iter().skip(column).step_by(cw)            open_z.rs:130-135
//   Merkle paths    MerkleProof::create_proof(tree_i, column)       pcs/utils.rs:163-176, 221-233
// Output order matches the reference's transcript order: per column, all rows' values, then per row a path
// of `depth` digests starting with the sibling at the leaf level.
#include "common.cuh"
#include "kernels.h"

namespace zipgpu {

// one thread moves one 16-byte half of one 32-byte item; items per (col,row): `depth` digests + the value
__global__ void __launch_bounds__(256)
    open_columns_kernel(const uint32_t *__restrict__ rows, const uint8_t *__restrict__ layers,
                        const uint32_t *__restrict__ columns, uint32_t *__restrict__ col_values,
                        uint8_t *__restrict__ paths, uint32_t num_rows, uint32_t cw, uint32_t out32, uint32_t num_cols,
                        uint32_t depth) {
    const size_t row_stride = 2 * (size_t)cw - 2;
    const size_t pairs = (size_t)num_cols * num_rows;
    // ---- paths: 2 threads per digest ----
    const size_t n_path_halves = pairs * depth * 2;
    const size_t n_val_words = pairs * out32;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_path_halves + n_val_words;
         g += (size_t)gridDim.x * blockDim.x) {
        if (g < n_path_halves) {
            const uint32_t half = (uint32_t)(g & 1);
            const size_t item = g >> 1;  // (pair, level)
            const uint32_t level = (uint32_t)(item % depth);
            const size_t pair = item / depth;
            const uint32_t row = (uint32_t)(pair % num_rows);
            const uint32_t col = __ldg(columns + pair / num_rows);
            const uint32_t idx = (col >> level) ^ 1u;
            const size_t off = 2 * (size_t)cw - ((2 * (size_t)cw) >> level);
            const uint4 *src = reinterpret_cast<const uint4 *>(layers + ((size_t)row * row_stride + off + idx) * 32);
            reinterpret_cast<uint4 *>(paths + item * 32)[half] = src[half];
        } else {
            const size_t w = g - n_path_halves;
            const uint32_t q = (uint32_t)(w % out32);
            const size_t pair = w / out32;
            const uint32_t row = (uint32_t)(pair % num_rows);
            const uint32_t col = __ldg(columns + pair / num_rows);
            col_values[w] = rows[((size_t)row * cw + col) * out32 + q];
        }
    }
}

// The same openings as the byte stream `open` appends to the proof (open_z.rs:124-143): per column, write_integers of
// the num_rows column entries (each u64 limb little-endian, pcs_transcript.rs:115-135) followed, per row, by
// write_merkle_proof = be64(path length) || the path digests (pcs_transcript.rs:198-211; utils.rs:221-233).
// One thread moves one 8-byte word of the stream.
__global__ void __launch_bounds__(256)
    open_columns_wire_kernel(const unsigned long long *__restrict__ rows, const unsigned long long *__restrict__ layers,
                             const uint32_t *__restrict__ columns, unsigned long long *__restrict__ out,
                             uint32_t num_rows, uint32_t cw, uint32_t out_limbs, uint32_t num_cols, uint32_t depth) {
    const size_t row_stride = 2 * (size_t)cw - 2;                    // digests per row in `layers`
    const size_t val_words = (size_t)num_rows * out_limbs;           // words of the values part of one column
    const size_t proof_words = 1 + (size_t)depth * 4;                // words of one row's proof
    const size_t col_words = val_words + (size_t)num_rows * proof_words;
    const size_t total = (size_t)num_cols * col_words;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const size_t c = g / col_words, off = g % col_words;
        const uint32_t col = __ldg(columns + c);
        unsigned long long v;
        if (off < val_words) {
            const uint32_t row = (uint32_t)(off / out_limbs), limb = (uint32_t)(off % out_limbs);
            v = rows[((size_t)row * cw + col) * out_limbs + limb];
        } else {
            const size_t po = off - val_words;
            const uint32_t row = (uint32_t)(po / proof_words), r = (uint32_t)(po % proof_words);
            if (r == 0) {
                const unsigned long long d = depth;  // big-endian u64 length prefix
                v = ((d & 0xffull) << 56) | ((d & 0xff00ull) << 40) | ((d & 0xff0000ull) << 24) | ((d & 0xff000000ull) << 8);
            } else {
                const uint32_t level = (r - 1) / 4, word = (r - 1) % 4;
                const uint32_t idx = (col >> level) ^ 1u;
                const size_t lvl_off = 2 * (size_t)cw - ((2 * (size_t)cw) >> level);
                v = layers[((size_t)row * row_stride + lvl_off + idx) * 4 + word];
            }
        }
        out[g] = v;
    }
}

size_t open_columns_wire_bytes(uint32_t num_rows, uint32_t out_limbs, int depth) {
    return ((size_t)num_rows * out_limbs + (size_t)num_rows * (1 + (size_t)depth * 4)) * 8;
}

cudaError_t launch_open_columns_wire(const OpenArgs &a, uint8_t *stream_out) {
    const size_t words = (size_t)a.num_cols * (open_columns_wire_bytes(a.num_rows, a.out32 / 2, a.depth) / 8);
    if (words == 0) return cudaSuccess;
    size_t grid = (words + 255) / 256;
    if (grid > 148 * 64) grid = 148 * 64;
    open_columns_wire_kernel<<<(uint32_t)grid, 256, 0, a.stream>>>(
        reinterpret_cast<const unsigned long long *>(a.rows), reinterpret_cast<const unsigned long long *>(a.layers),
        a.columns, reinterpret_cast<unsigned long long *>(stream_out), a.num_rows, a.cw, a.out32 / 2, a.num_cols,
        (uint32_t)a.depth);
    return cudaGetLastError();
}

cudaError_t launch_open_columns(const OpenArgs &a) {
    const size_t work = (size_t)a.num_cols * a.num_rows * ((size_t)a.depth * 2 + a.out32);
    if (work == 0) return cudaSuccess;
    size_t grid = (work + 255) / 256;
    if (grid > 148 * 64) grid = 148 * 64;
    open_columns_kernel<<<(uint32_t)grid, 256, 0, a.stream>>>(a.rows, a.layers, a.columns, a.col_values, a.paths,
                                                               a.num_rows, a.cw, a.out32, a.num_cols, (uint32_t)a.depth);
    return cudaGetLastError();
}

}  // namespace zipgpu
