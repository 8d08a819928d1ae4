// zinc_b200/csrc/sparse_encode.cu -- K6: ZipLinearCode::encode_wide (the sparse code) for a batch of rows, sm_100a.
//
// Reference: src/zip/code.rs:186-201 (encode_wide = a.mat_vec_mul(row) ‖ b.mat_vec_mul(row)) and
// SparseMatrixZ::mat_vec_mul code.rs:299-321; called per row from MultilinearZip::encode_rows (commit.rs:158-183).
//
// Over a whole matrix of evaluation rows the encoder is   rows_out[r][j] = Σ_c  M[j][c] · evals[r][c]   with M the
// (cw × row_len) stack of the two sampled matrices: a genuine integer matrix product (unlike the RAA code, which has no
// contraction).  Two kernels:
//
//  * sparse_gemm_kernel -- every coefficient is 0 or 1 (what KeccakTranscript::get_encoding_element draws,
//    transcript.rs:176-181).  M is kept dense as bytes; the evaluations are split into byte planes
//    (x + 2^(64·in−1), so all bytes are unsigned); the product runs on the tensor cores as an exact u8×u8→s32 GEMM
//    (mma.sync.m16n8k32; 255·row_len < 2^31), and the epilogue recombines the 8·in_limbs planes of an entry into the
//    multi-limb integer, removes the bias (nnz[j]·2^(64·in−1)) and sign-extends to Int<out_limbs>.
//    Default: the tcgen05 kernel of sparse_umma.cu (TMA → shared memory → tcgen05.mma.kind::i8 → TMEM).  The kernel in
//    this file is the mma.sync version of the same product (ZIPGPU_SPARSE_MMA_SYNC=1): tiles 128×128×128 bytes,
//    3-stage cp.async pipeline, XOR-swizzled 128-byte rows read with ldmatrix.
//  * sparse_generic_kernel -- arbitrary i64 coefficients (the reference tests' MockTranscript draws a counter,
//    pcs/tests.rs:30-33) and the shapes the GEMM tiling does not cover (row_len or cw below 128): one thread per
//    codeword entry walking an ELL table transposed so that a warp reads it coalesced.
#include "common.cuh"
#include "kernels.h"

namespace zipgpu {

// ------------------------------------------------------------------------------------------------------
// generic path
// ------------------------------------------------------------------------------------------------------
template <int OUT>
__global__ void __launch_bounds__(128) sparse_generic_kernel(const uint64_t *__restrict__ evals, uint64_t *__restrict__ rows_out,
                                                             const uint32_t *__restrict__ cols_t,
                                                             const int64_t *__restrict__ coef_t, uint32_t num_rows,
                                                             uint32_t row_len, uint32_t cw, uint32_t d, int in_limbs) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cw) return;
    for (uint32_t r = blockIdx.y; r < num_rows; r += gridDim.y) {
        const uint64_t *row = evals + (size_t)r * row_len * in_limbs;
        uint64_t acc[OUT];
#pragma unroll
        for (int i = 0; i < OUT; i++) acc[i] = 0;
        for (uint32_t k = 0; k < d; k++) {
            const int64_t coef = coef_t[(size_t)k * cw + j];
            if (coef == 0) continue;
            const uint64_t *v = row + (size_t)cols_t[(size_t)k * cw + j] * in_limbs;
            // w = sext(v) to OUT limbs; acc += coef * w  (mod 2^(64*OUT)), code.rs:314
            uint64_t w[OUT];
            const uint64_t fill = (uint64_t)((int64_t)v[in_limbs - 1] >> 63);
#pragma unroll
            for (int i = 0; i < OUT; i++) w[i] = i < in_limbs ? v[i] : fill;
            const uint64_t mag = coef < 0 ? 0ull - (uint64_t)coef : (uint64_t)coef;
            uint64_t carry = 0, neg_c = 1, add_c = 0;
#pragma unroll
            for (int i = 0; i < OUT; i++) {
                const uint64_t lo = w[i] * mag, hi = __umul64hi(w[i], mag);
                uint64_t p = lo + carry;
                carry = hi + (p < lo);
                if (coef < 0) {
                    p = ~p + neg_c;
                    neg_c = (neg_c && p == 0) ? 1 : 0;
                }
                const uint64_t s = acc[i] + p;
                const uint64_t c1 = s < p;
                const uint64_t s2 = s + add_c;
                add_c = c1 | (s2 < s);
                acc[i] = s2;
            }
        }
        uint64_t *o = rows_out + ((size_t)r * cw + j) * OUT;
#pragma unroll
        for (int i = 0; i < OUT; i++) o[i] = acc[i];
    }
}

// ------------------------------------------------------------------------------------------------------
// 0/1 path: byte planes + u8 tensor-core GEMM
// ------------------------------------------------------------------------------------------------------
// planes[(r*P + p)][c] = byte p of (evals[r][c] + 2^(64*IN-1)); 4 consecutive c per thread
template <int IN>
__global__ void __launch_bounds__(256) split_planes_kernel(const uint64_t *__restrict__ evals, uint8_t *__restrict__ planes,
                                                           uint32_t num_rows, uint32_t row_len) {
    constexpr int P = 8 * IN;
    const uint32_t quads = row_len / 4;
    const size_t total = (size_t)num_rows * quads;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(t / quads), q = (uint32_t)(t % quads);
        const uint64_t *src = evals + ((size_t)r * row_len + 4 * q) * IN;
        uint64_t x[4][IN];
#pragma unroll
        for (int e = 0; e < 4; e++)
#pragma unroll
            for (int i = 0; i < IN; i++) x[e][i] = src[e * IN + i];
#pragma unroll
        for (int e = 0; e < 4; e++) x[e][IN - 1] ^= 0x8000000000000000ull;
        uint8_t *dst = planes + (size_t)r * P * row_len + 4 * q;
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int limb = p / 8, sh = 8 * (p % 8);
            const uint32_t w = (uint32_t)((x[0][limb] >> sh) & 0xff) | ((uint32_t)((x[1][limb] >> sh) & 0xff) << 8) |
                               ((uint32_t)((x[2][limb] >> sh) & 0xff) << 16) | ((uint32_t)((x[3][limb] >> sh) & 0xff) << 24);
            *reinterpret_cast<uint32_t *>(dst + (size_t)p * row_len) = w;
        }
    }
}

constexpr int GM = 128, GN = 128, GK = 128, GSTAGES = 3;
constexpr int GEMM_THREADS = 256;
constexpr size_t GEMM_SMEM = (size_t)GSTAGES * (GM + GN) * GK;

__device__ __forceinline__ void cp_async16(uint32_t smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A: dense [cw][K] bytes (0/1);  B: planes [num_rows*P][K] bytes;  out[r][j] = Int<out_limbs>.
// CTA tile: 128 codeword entries × 128 plane rows (= 128/P evaluation rows); 8 warps as 2 (entries) × 4 (plane rows).
template <int IN>
__global__ void __launch_bounds__(GEMM_THREADS, 2) sparse_gemm_kernel(const uint8_t *__restrict__ A, const uint8_t *__restrict__ B,
                                                                      const uint32_t *__restrict__ nnz,
                                                                      uint64_t *__restrict__ rows_out, uint32_t num_rows,
                                                                      uint32_t K, uint32_t cw, int out_limbs) {
    constexpr int P = 8 * IN;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_m = warp & 1, warp_n = warp >> 1;
    const uint32_t m0 = blockIdx.x * GM;
    const size_t n0 = (size_t)blockIdx.y * GN;
    const size_t n_total = (size_t)num_rows * P;

    // stage loads: 4 + 4 16-byte chunks per thread; a tile row is 128 bytes = 8 chunks, chunk c of row r lives at
    // r*128 + ((c ^ (r & 7)) << 4) so that ldmatrix's 8 row addresses fall in 8 different 16-byte bank groups
    // thread t moves chunk (t & 7) of tile rows (t >> 3) + 32*i, i < 4
    const int ld_row = tid >> 3, ld_ch = tid & 7;
    const uint32_t soff = ld_row * GK + ((ld_ch ^ (ld_row & 7)) << 4);
    const uint8_t *ga = A + (size_t)(m0 + ld_row) * K + ld_ch * 16;
    const uint8_t *gb = B + ld_ch * 16;
    auto load_stage = [&](int stage, uint32_t kt) {
        const uint32_t sa = smem_base + stage * (GM + GN) * GK, sb = sa + GM * GK;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            cp_async16(sa + soff + i * 32 * GK, ga + ((size_t)i * 32 * K + (size_t)kt * GK));
            size_t n = n0 + ld_row + 32 * i;
            if (n >= n_total) n = n_total - 1;  // clamped; never stored
            cp_async16(sb + soff + i * 32 * GK, gb + (n * K + (size_t)kt * GK));
        }
    };

    int acc[4][4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[a][b][c] = 0;

    const uint32_t KT = K / GK;
#pragma unroll
    for (int s = 0; s < GSTAGES - 1; s++) {
        if ((uint32_t)s < KT) load_stage(s, s);
        cp_async_commit();
    }
    // per-lane ldmatrix row/chunk selectors
    const int a_row = warp_m * 64 + (lane & 7) + ((lane >> 3) & 1) * 8, a_ch = lane >> 4;
    const int b_row = warp_n * 32 + (lane & 7) + (lane >> 4) * 8, b_ch = (lane >> 3) & 1;

    for (uint32_t kt = 0; kt < KT; kt++) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        if (kt + GSTAGES - 1 < KT) load_stage((kt + GSTAGES - 1) % GSTAGES, kt + GSTAGES - 1);
        cp_async_commit();
        const uint32_t sa = smem_base + (kt % GSTAGES) * (GM + GN) * GK, sb = sa + GM * GK;
#pragma unroll
        for (int ks = 0; ks < GK / 32; ks++) {
            uint32_t bf[4][2];
#pragma unroll
            for (int np = 0; np < 2; np++) {
                const int row = b_row + np * 16, ch = ks * 2 + b_ch;
                ldmatrix_x4(bf[2 * np][0], bf[2 * np][1], bf[2 * np + 1][0], bf[2 * np + 1][1],
                            sb + row * GK + ((ch ^ (row & 7)) << 4));
            }
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                uint32_t af[4];
                const int row = a_row + mt * 16, ch = ks * 2 + a_ch;
                ldmatrix_x4(af[0], af[1], af[2], af[3], sa + row * GK + ((ch ^ (row & 7)) << 4));
#pragma unroll
                for (int nt = 0; nt < 4; nt++) mma_u8(acc[mt][nt], af, bf[nt][0], bf[nt][1]);
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: an evaluation row owns IN consecutive n-tiles (P planes); inside an n-tile a thread holds planes
    // 2*tig and 2*tig+1 of entries g and g+8.  Recombine, reduce over the quad, unbias, sign-extend, store.
    const int g = lane >> 2, tig = lane & 3;
    const int sh = 16 * tig;
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const uint32_t j = m0 + warp_m * 64 + mt * 16 + g + half * 8;
            const uint32_t cnt = nnz[j];
#pragma unroll
            for (int er = 0; er < 4 / IN; er++) {
                uint64_t limb[IN + 1];
#pragma unroll
                for (int i = 0; i <= IN; i++) limb[i] = 0;
#pragma unroll
                for (int q = 0; q < IN; q++) {
                    const int nt = er * IN + q;
                    const uint64_t part = (uint64_t)(uint32_t)acc[mt][nt][2 * half] +
                                          ((uint64_t)(uint32_t)acc[mt][nt][2 * half + 1] << 8);
                    const uint64_t lo = part << sh, hi = sh ? part >> (64 - sh) : 0ull;
                    const uint64_t s0 = limb[q] + lo;
                    const uint64_t c0 = s0 < lo;
                    limb[q] = s0;
                    limb[q + 1] += hi + c0;  // cannot overflow: the limb holds < 2^41 so far
                }
#pragma unroll
                for (int x = 1; x <= 2; x <<= 1) {
                    uint64_t carry = 0;
#pragma unroll
                    for (int i = 0; i <= IN; i++) {
                        const uint64_t o = __shfl_xor_sync(0xffffffffu, limb[i], x);
                        const uint64_t s = limb[i] + o;
                        const uint64_t c1 = s < o;
                        const uint64_t s2 = s + carry;
                        carry = c1 | (s2 < s);
                        limb[i] = s2;
                    }
                }
                // subtract cnt * 2^(64*IN-1)
                {
                    const uint64_t b_lo = (uint64_t)(cnt & 1) << 63, b_hi = (uint64_t)(cnt >> 1);
                    const uint64_t d0 = limb[IN - 1] - b_lo;
                    const uint64_t borrow = limb[IN - 1] < b_lo;
                    limb[IN - 1] = d0;
                    limb[IN] = limb[IN] - b_hi - borrow;
                }
                const uint64_t sign = (uint64_t)((int64_t)limb[IN] >> 63);
                const size_t n = n0 + warp_n * 32 + (size_t)er * P;
                const size_t r = n / P;
                if (r < num_rows) {
                    uint64_t *o = rows_out + (r * cw + j) * out_limbs;
                    // thread tig of the quad writes limbs tig, tig+4, ...
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        if (i < out_limbs && (i & 3) == tig) {
                            uint64_t v = sign;
#pragma unroll
                            for (int l = 0; l <= IN; l++)
                                if (l == i) v = limb[l];
                            o[i] = v;
                        }
                    }
                }
            }
        }
    }
}

bool sparse_gemm_supported(int in_limbs, int out_limbs, uint32_t row_len, uint32_t cw) {
    return (in_limbs == 1 || in_limbs == 2) && out_limbs >= in_limbs && out_limbs <= 8 && row_len % GK == 0 &&
           cw % GM == 0 && row_len <= (1u << 23);
}
size_t sparse_planes_bytes(uint32_t num_rows, uint32_t row_len, int in_limbs) {
    return (size_t)num_rows * row_len * 8 * in_limbs;
}

template <int OUT>
static cudaError_t launch_generic(const SparseEncodeArgs &a) {
    dim3 grid((a.cw + 127) / 128, a.num_rows < 32768 ? a.num_rows : 32768);
    sparse_generic_kernel<OUT><<<grid, 128, 0, a.stream>>>(a.evals, a.rows_out, a.cols_t, a.coef_t, a.num_rows, a.row_len,
                                                          a.cw, a.d, a.in_limbs);
    return cudaGetLastError();
}

template <int IN>
static cudaError_t launch_gemm(const SparseEncodeArgs &a) {
    cudaError_t e = cudaFuncSetAttribute(sparse_gemm_kernel<IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    if (e != cudaSuccess) return e;
    const size_t total = (size_t)a.num_rows * (a.row_len / 4);
    const uint32_t blocks = (uint32_t)((total + 255) / 256 < (size_t)a.num_sms * 16 ? (total + 255) / 256 : (size_t)a.num_sms * 16);
    split_planes_kernel<IN><<<blocks, 256, 0, a.stream>>>(a.evals, a.planes, a.num_rows, a.row_len);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (a.use_umma) return launch_sparse_umma(a);  // tcgen05 + TMA + TMEM (sparse_umma.cu)
    // grid.y is limited to 65535 tiles of GN plane rows: walk very tall matrices in row batches
    const uint32_t rows_per_tile = GN / (8 * IN), max_rows = 65535u * rows_per_tile;
    for (uint32_t r0 = 0; r0 < a.num_rows; r0 += max_rows) {
        const uint32_t nr = a.num_rows - r0 < max_rows ? a.num_rows - r0 : max_rows;
        dim3 grid(a.cw / GM, (nr + rows_per_tile - 1) / rows_per_tile);
        sparse_gemm_kernel<IN><<<grid, GEMM_THREADS, GEMM_SMEM, a.stream>>>(
            a.dense, a.planes + (size_t)r0 * 8 * IN * a.row_len, a.nnz, a.rows_out + (size_t)r0 * a.cw * a.out_limbs, nr,
            a.row_len, a.cw, a.out_limbs);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_sparse_encode(const SparseEncodeArgs &a, int *launches) {
    if (launches) *launches = 0;
    if (a.num_rows == 0) return cudaSuccess;
    if (a.dense) {
        if (launches) *launches = 2;
        return a.in_limbs == 1 ? launch_gemm<1>(a) : launch_gemm<2>(a);
    }
    if (launches) *launches = 1;
    switch (a.out_limbs) {
        case 1: return launch_generic<1>(a);
        case 2: return launch_generic<2>(a);
        case 3: return launch_generic<3>(a);
        case 4: return launch_generic<4>(a);
        case 5: return launch_generic<5>(a);
        case 6: return launch_generic<6>(a);
        case 7: return launch_generic<7>(a);
        case 8: return launch_generic<8>(a);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace zipgpu
