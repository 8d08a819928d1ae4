// zinc_b200/csrc/sparse_umma.cu -- K6, the tcgen05 version of the sparse-code GEMM (see sparse_encode.cu for the
// arithmetic: 0..255 coefficients × unsigned byte planes of the evaluations, exact s32 accumulation).
//
// Reference: ZipLinearCode::encode_wide src/zip/code.rs:186-201, SparseMatrixZ::mat_vec_mul code.rs:299-321.
//
// Persistent: one CTA per SM walks 128 (codeword entries) × 256 (plane rows = 256/(8·in_limbs) evaluation rows) tiles:
//   warp 0   one lane drives TMA: per 128-byte K block, the 128×128 B tile of the coefficient matrix and the 256×128 B
//            tile of the planes land in a 4-stage shared-memory ring (SWIZZLE_128B, K-major), completion on mbarriers;
//   warp 1   one lane issues tcgen05.mma.kind::i8 (M128 × N256 × K32, u8·u8→s32) into one of two 256-column TMEM
//            accumulators, tcgen05.commit releases ring slots and signals the epilogue at the end of a tile;
//   warps 2-5 epilogue: TMEM lane = codeword entry, 8·in_limbs consecutive columns = the byte planes of one
//            evaluation row, so a thread reads its planes with tcgen05.ld, recombines them into the multi-limb sum,
//            removes the bias, sign-extends and stores Int<out_limbs> -- a warp writes 32 consecutive entries.  The
//            epilogue of tile i runs while the MMAs of tile i+1 fill the other accumulator.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "kernels.h"

namespace zipgpu {

namespace umma {
// K block per ring stage.  64 (SWIZZLE_64B, 9 stages of 24 KB: more loads in flight) was measured slower: 0.99 ms
// against 0.75 ms at nv = 24, so the feed limit is not the number of outstanding TMA loads.
#ifndef ZIPGPU_UMMA_TK
#define ZIPGPU_UMMA_TK 128
#endif
constexpr int TM = 128, TN = 256, TK = ZIPGPU_UMMA_TK, STAGES = TK == 128 ? 4 : 9;
static_assert(TK == 128 || TK == 64, "K block = one swizzle atom row: 128 (SWIZZLE_128B) or 64 bytes (SWIZZLE_64B)");
constexpr int A_BYTES = TM * TK, B_BYTES = TN * TK, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;  // two TN-column accumulators
constexpr int EPI_THREADS = 128;
constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded spin: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; spins++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    // stride between 8-row groups = 8 * TK bytes; layout type 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(8 * TK / 16) << 32) | (1ull << 46) |
           ((TK == 128 ? 2ull : 4ull) << 61);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
}  // namespace umma

// instruction descriptor (cute::UMMA::InstrDescriptor): D = s32, A = B = u8, both K-major, N = 256, M = 128
constexpr uint32_t UMMA_IDESC = (2u << 4) | ((uint32_t)(umma::TN >> 3) << 17) | ((uint32_t)(umma::TM >> 4) << 24);

// Tile order: bands of BAND m-blocks; inside a band the m-block runs fastest, then the n-block.  The ~148 tiles in
// flight then cover BAND slices of the coefficient matrix and ~148/BAND plane tiles, both L2-resident, instead of one
// sweep over the whole matrix per plane tile (which at nv = 26 is 128 MiB > L2 and made the kernel DRAM-bound).
// A matrix that fits L2 with room to spare is not banded (BAND = m_blocks): all its m-blocks then share each plane tile.
__device__ __forceinline__ void tile_coords(uint32_t t, uint32_t m_blocks, uint32_t n_blocks, uint32_t BAND, uint32_t &m_blk,
                                            uint32_t &n_blk) {
    const uint32_t per_band = BAND * n_blocks;
    const uint32_t band = t / per_band, r = t - band * per_band;
    const uint32_t width = m_blocks - band * BAND < BAND ? m_blocks - band * BAND : BAND;
    n_blk = r / width;
    m_blk = band * BAND + (r - n_blk * width);
}

template <int IN>
__global__ void __launch_bounds__(umma::THREADS, 1)
    sparse_umma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                       const uint32_t *__restrict__ bias, uint64_t *__restrict__ rows_out, uint32_t num_rows, uint32_t K,
                       uint32_t cw, int out_limbs, uint32_t m_blocks, uint32_t n_blocks, uint32_t band) {
    using namespace umma;
    constexpr int P = 8 * IN;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t bars = tiles + STAGES * STAGE_BYTES;  // full[STAGES], empty[STAGES], acc_full[2], acc_empty[2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + (bars - raw) + 8 * (2 * STAGES + 4));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t KT = K / TK;
    const uint32_t num_tiles = m_blocks * n_blocks;
    auto full = [&](int s) { return bars + 8 * s; };
    auto empty = [&](int s) { return bars + 8 * (STAGES + s); };
    auto acc_full = [&](int b) { return bars + 8 * (2 * STAGES + b); };
    auto acc_empty = [&](int b) { return bars + 8 * (2 * STAGES + 2 + b); };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_b)) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; s++) {
                mbar_init(full(s), 1);
                mbar_init(empty(s), 1);
            }
            for (int b = 0; b < 2; b++) {
                mbar_init(acc_full(b), 1);
                mbar_init(acc_empty(b), EPI_THREADS);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // persistent: this CTA takes tiles blockIdx.x, blockIdx.x + gridDim.x, ... in the banded order of tile_coords()
    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                uint32_t m_blk, n_blk;
                tile_coords(t, m_blocks, n_blocks, band, m_blk, n_blk);
                const int m0 = (int)(m_blk * TM), n0 = (int)(n_blk * TN);
                for (uint32_t kb = 0; kb < KT; kb++, it++) {
                    const int s = it % STAGES;
                    mbar_wait(empty(s), ((it / STAGES) & 1) ^ 1);
                    mbar_expect_tx(full(s), STAGE_BYTES);
                    tma_load_2d(tiles + s * STAGE_BYTES, &tm_a, full(s), (int)(kb * TK), m0);
                    tma_load_2d(tiles + s * STAGE_BYTES + A_BYTES, &tm_b, full(s), (int)(kb * TK), n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, i = 0;
            for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x, i++) {
                const uint32_t ab = i & 1;  // two accumulators of TN columns: the epilogue of a tile overlaps the next tile
                mbar_wait(acc_empty(ab), ((i >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + ab * TN;
                for (uint32_t kb = 0; kb < KT; kb++, it++) {
                    const int s = it % STAGES;
                    mbar_wait(full(s), (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a = tiles + s * STAGE_BYTES, b = a + A_BYTES;
#pragma unroll
                    for (int k = 0; k < TK / 32; k++)
                        mma_i8(d, smem_desc(a + k * 32), smem_desc(b + k * 32), UMMA_IDESC, (kb | (uint32_t)k) != 0);
                    tc_commit(empty(s));  // the slot is free once these MMAs have read it
                }
                tc_commit(acc_full(ab));
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;  // a warp reads TMEM lanes 32*(warp % 4) ..
        constexpr int ROWS_PER_LD = 32 / P;
        uint32_t i = 0;
        for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x, i++) {
            uint32_t m_blk, n_blk;
            tile_coords(t, m_blocks, n_blocks, band, m_blk, n_blk);
            const uint32_t m0 = m_blk * TM, n0 = n_blk * TN;
            const uint32_t ab = i & 1;
            const uint32_t j = m0 + q * 32 + lane;
            const uint32_t cnt = bias[j];
            const uint64_t b_lo = (uint64_t)(cnt & 1) << 63, b_hi = (uint64_t)(cnt >> 1);
            mbar_wait(acc_full(ab), (i >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int g = 0; g < TN / 32; g++) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ab * TN + g * 32, v);
                if (g == TN / 32 - 1) {  // everything of this accumulator is in registers: hand it back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(acc_empty(ab));
                }
#pragma unroll
                for (int e = 0; e < ROWS_PER_LD; e++) {
                    uint64_t limb[IN + 1];
#pragma unroll
                    for (int l = 0; l <= IN; l++) limb[l] = 0;
#pragma unroll
                    for (int ql = 0; ql < IN; ql++) {
                        uint64_t lo = 0, hi = 0;
#pragma unroll
                        for (int p = 0; p < 8; p++) {
                            const uint64_t x = v[e * P + ql * 8 + p];  // < 2^31
                            const uint64_t add = x << (8 * p);
                            const uint64_t s = lo + add;
                            hi += s < add;
                            lo = s;
                            if (8 * p > 32) hi += x >> (64 - 8 * p);
                        }
                        const uint64_t s = limb[ql] + lo;
                        const uint64_t c = s < lo;
                        limb[ql] = s;
                        limb[ql + 1] += hi + c;
                    }
                    {  // minus bias * 2^(64*IN - 1)
                        const uint64_t d0 = limb[IN - 1] - b_lo;
                        const uint64_t borrow = limb[IN - 1] < b_lo;
                        limb[IN - 1] = d0;
                        limb[IN] = limb[IN] - b_hi - borrow;
                    }
                    const uint64_t sign = (uint64_t)((int64_t)limb[IN] >> 63);
                    const size_t r = (size_t)n0 / P + (size_t)g * ROWS_PER_LD + e;
                    if (r < num_rows) {
                        uint64_t *o = rows_out + (r * cw + j) * out_limbs;
                        if (out_limbs == 4) {
                            const uint64_t l2 = IN >= 2 ? limb[IN >= 2 ? 2 : 0] : sign;
                            const uint32_t w[8] = {(uint32_t)limb[0], (uint32_t)(limb[0] >> 32), (uint32_t)limb[1],
                                                   (uint32_t)(limb[1] >> 32), (uint32_t)l2, (uint32_t)(l2 >> 32),
                                                   (uint32_t)sign, (uint32_t)sign};
                            st_stream_v8(o, w);
                        } else {
#pragma unroll
                            for (int l8 = 0; l8 < 8; l8++) {
                                if (l8 < out_limbs) {
                                    uint64_t x = sign;
#pragma unroll
                                    for (int l = 0; l <= IN; l++)
                                        if (l == l8) x = limb[l];
                                    o[l8] = x;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

static PFN_cuTensorMapEncodeTiled encode_fn() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    }
    return fn;
}
// rows × K bytes, row-major; box = box_rows × 128 bytes, SWIZZLE_128B; rows beyond `rows` read as zero
static bool make_map(CUtensorMap *tm, const void *base, uint64_t rows, uint64_t K, uint32_t box_rows) {
    PFN_cuTensorMapEncodeTiled fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {K, rows};
    cuuint64_t strides[1] = {K};
    cuuint32_t box[2] = {(cuuint32_t)umma::TK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, umma::TK == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int IN>
static cudaError_t launch_umma_t(const SparseEncodeArgs &a) {
    cudaError_t e = cudaFuncSetAttribute(sparse_umma_kernel<IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)umma::SMEM);
    if (e != cudaSuccess) return e;
    const uint32_t rows_per_tile = umma::TN / (8 * IN), max_rows = 8192u * rows_per_tile;  // keeps num_tiles far below 2^32
    CUtensorMap tm_a;
    if (!make_map(&tm_a, a.dense, a.cw, a.row_len, umma::TM)) return cudaErrorNotSupported;
    for (uint32_t r0 = 0; r0 < a.num_rows; r0 += max_rows) {
        const uint32_t nr = a.num_rows - r0 < max_rows ? a.num_rows - r0 : max_rows;
        CUtensorMap tm_b;
        if (!make_map(&tm_b, a.planes + (size_t)r0 * 8 * IN * a.row_len, (uint64_t)nr * 8 * IN, a.row_len, umma::TN))
            return cudaErrorNotSupported;
        const uint32_t m_blocks = a.cw / umma::TM, n_blocks = (nr + rows_per_tile - 1) / rows_per_tile;
        const uint32_t num_tiles = m_blocks * n_blocks;
        const uint32_t grid = num_tiles < (uint32_t)a.num_sms ? num_tiles : (uint32_t)a.num_sms;
        sparse_umma_kernel<IN><<<grid, umma::THREADS, umma::SMEM, a.stream>>>(
            tm_a, tm_b, a.nnz, a.rows_out + (size_t)r0 * a.cw * a.out_limbs, nr, a.row_len, a.cw, a.out_limbs, m_blocks,
            n_blocks, (size_t)a.cw * a.row_len <= ((size_t)48 << 20) ? m_blocks : 16u);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// the GEMM stage only (the planes have been split already)
cudaError_t launch_sparse_umma(const SparseEncodeArgs &a) {
    return a.in_limbs == 1 ? launch_umma_t<1>(a) : launch_umma_t<2>(a);
}

}  // namespace zipgpu
