// zinc_b200/csrc/raa_encode.cu -- K1: row-batched RAA encoder for sm_100a.
//
// Replaces, for every row of the evaluation matrix (commit.rs:158-183):
//     repeat -> shuffle_seeded(perm_1_seed) -> accumulate -> shuffle_seeded(perm_2_seed) -> accumulate
// (code_raa.rs:89-105,142-171; zip/utils.rs:139-142) over Int<N> -> Int<K> (field/int.rs).
//
// Design (B200-first, not a translation of the Rust loops):
//   * One CTA owns one row at a time and keeps the whole codeword on chip: the only HBM traffic is the
//     compulsory 8 B read + 64 B written per evaluation (INT_LIMBS=1).  The CTA is persistent over rows so
//     the two permutations (identical for every row of a pp) are loaded ONCE into registers, already
//     translated to shared-memory slots.
//   * repeat o perm1 folds into a gather from the staged input row: y1[i] = widen(row[perm1[i] mod row_len]).
//   * Values are held in W 32-bit limbs (W=3, 96 bit, is exact for Int<1> inputs up to cw = 2^16 because
//     |s2| < 2^63 * cw^2); the stored Int<K> is the sign extension, produced on the way out.
//   * Each thread owns E consecutive codeword positions: a serial carry-chain scan in registers, a
//     warp-shuffle scan of the per-thread totals and one cross-warp step give the row prefix sum.
//   * s1 lives in shared memory as W planes of u32, in a thread-striped XOR-swizzled layout
//     slot(t,k) = k*T + (t ^ (k << log2(32/E))) which is bank-conflict free both for the owner-thread
//     writes (fixed k, consecutive t) and for the coalesced read-out (consecutive i = t*E + k).  The perm2
//     gather is the only randomly-banked access.
//   * Output goes back through the planes so that consecutive lanes store consecutive 32-byte Int<4> values with
//     one 256-bit store each (a scattered store costs the L1 data pipe one wavefront per 32-byte sector).
#include "common.cuh"
#include "kernels.h"

namespace zipgpu {

template <int E>
struct Swz {
    static constexpr int SH = (E >= 32) ? 0 : (E == 16 ? 1 : E == 8 ? 2 : E == 4 ? 3 : E == 2 ? 4 : 5);
};

template <int E>
__device__ __forceinline__ uint32_t slot_of(uint32_t t, uint32_t k, uint32_t T) {
    return k * T + (t ^ ((k << Swz<E>::SH) & 31u));
}

// Scan of W-limb values across the CTA in the logical order i = t*E + k.  On return v[k] holds the inclusive scan
// WITHIN the thread and pre[] the sum of everything owned by lower threads; the caller adds pre to each v[k] at
// the point where it consumes it (keeps the live register set small).  aux: 2 * 32 * W words of shared memory.
template <int W, int E>
__device__ __forceinline__ void block_scan(uint32_t (&v)[E][W], uint32_t (&pre)[W], uint32_t *aux, uint32_t t,
                                           uint32_t nwarps) {
    const uint32_t lane = t & 31u, warp = t >> 5;
#pragma unroll
    for (int k = 1; k < E; k++) add_limbs<W>(v[k], v[k - 1]);
    uint32_t inc[W];
#pragma unroll
    for (int w = 0; w < W; w++) inc[w] = v[E - 1][w];
    // warp inclusive scan of thread totals
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t o[W];
#pragma unroll
        for (int w = 0; w < W; w++) o[w] = __shfl_up_sync(0xffffffffu, inc[w], off);
        if (lane >= (uint32_t)off) add_limbs<W>(inc, o);
    }
    if (lane == 31) {
#pragma unroll
        for (int w = 0; w < W; w++) aux[warp * W + w] = inc[w];
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t wt[W];
#pragma unroll
        for (int w = 0; w < W; w++) wt[w] = (lane < nwarps) ? aux[lane * W + w] : 0u;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t o[W];
#pragma unroll
            for (int w = 0; w < W; w++) o[w] = __shfl_up_sync(0xffffffffu, wt[w], off);
            if (lane >= (uint32_t)off) add_limbs<W>(wt, o);
        }
        // exclusive warp prefix
#pragma unroll
        for (int w = 0; w < W; w++) {
            uint32_t e = __shfl_up_sync(0xffffffffu, wt[w], 1);
            aux[32 * W + lane * W + w] = lane ? e : 0u;
        }
    }
    __syncthreads();
    // exclusive prefix of this thread = warp prefix + inclusive scan of the lower lanes of the warp
#pragma unroll
    for (int w = 0; w < W; w++) {
        uint32_t e = __shfl_up_sync(0xffffffffu, inc[w], 1);
        pre[w] = lane ? e : 0u;
    }
    uint32_t wp[W];
#pragma unroll
    for (int w = 0; w < W; w++) wp[w] = aux[32 * W + warp * W + w];
    add_limbs<W>(pre, wp);
}

// Pre-translated permutation tables (built once per pp by build_encode_tables, below):
//   tab1[i'] = word offset into the staged input row of the element that codeword position i gathers in pass 1
//              ( (perm1[i] mod row_len) * IN32 )
//   tab2[i'] = shared-memory slot of the s1 entry that codeword position i gathers in pass 2 ( slot(perm2[i]) )
// stored "lane-major": position i = t*E + k lives at i' = (k / G) * (T * G) + t * G + (k % G) with G = min(E, 4), so
// the G entries a thread needs for one group are one 16-byte load and consecutive lanes read consecutive
// addresses (512 B per warp request, L2-resident: the tables are shared by every row).
template <int E>
__device__ __forceinline__ uint32_t tab_index(uint32_t t, uint32_t k, uint32_t T) {
    constexpr uint32_t G = E < 4 ? E : 4;
    return (k / G) * (T * G) + t * G + (k % G);
}

template <int IN32, int W, int E, bool CACHE_PERM, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
    raa_encode_kernel(const uint32_t *__restrict__ evals, uint32_t *__restrict__ rows_out,
                      const uint32_t *__restrict__ tab1, const uint32_t *__restrict__ tab2, uint32_t num_rows,
                      uint32_t row_len, uint32_t cw, uint32_t out32) {
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t T = blockDim.x, t = threadIdx.x;
    const uint32_t P = T * E;   // plane size in words (>= cw)
    uint32_t *planes = smem;    // [W][P]
    uint32_t *stage = smem;     // input row, aliases the planes (dead before s1 is written)
    uint32_t *aux = smem + (size_t)W * P;
    const uint32_t nwarps = T >> 5;
    const uint32_t in_words = row_len * IN32;
    constexpr int G = E < 4 ? E : 4;

    // table entries of this thread, group g = entries k in [g*G, g*G+G)
    auto load_group = [&](const uint32_t *tab, int g, uint32_t (&q)[G]) {
        const uint32_t *p = tab + (size_t)g * (T * G) + t * G;
        if constexpr (G == 4) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
            q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < G; j++) q[j] = __ldg(p + j);
        }
    };

    // CACHE_PERM (E <= 8): both tables stay in registers for every row this persistent CTA processes.
    // Otherwise (E = 16 runs at a 64-register budget) the entries are re-read from L2 in each pass.
    uint32_t src1[CACHE_PERM ? E : 1], slot2[CACHE_PERM ? E : 1];
    if constexpr (CACHE_PERM) {
#pragma unroll
        for (int g = 0; g < E / G; g++) {
            uint32_t q1[G], q2[G];
            load_group(tab1, g, q1);
            load_group(tab2, g, q2);
#pragma unroll
            for (int j = 0; j < G; j++) {
                src1[CACHE_PERM ? g * G + j : 0] = q1[j];
                slot2[CACHE_PERM ? g * G + j : 0] = q2[j];
            }
        }
    }

    for (uint32_t row = blockIdx.x; row < num_rows; row += gridDim.x) {
        // ---- 1. stage the input row (coalesced, read-once) ----
        const uint32_t *src = evals + (size_t)row * in_words;
        if ((in_words & 3u) == 0) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            uint4 *d4 = reinterpret_cast<uint4 *>(stage);
            for (uint32_t i = t; i < (in_words >> 2); i += T) d4[i] = ld_stream_v4(s4 + i);
        } else {
            for (uint32_t i = t; i < in_words; i += T) stage[i] = src[i];
        }
        __syncthreads();

        // ---- 2. y1 = widen(row[perm1[i] mod row_len]) ----
        uint32_t v[E][W];
#pragma unroll
        for (int g = 0; g < E / G; g++) {
            uint32_t q[G];
            if constexpr (!CACHE_PERM) load_group(tab1, g, q);
#pragma unroll
            for (int j = 0; j < G; j++) {
                const int k = g * G + j;
                const uint32_t so = CACHE_PERM ? src1[CACHE_PERM ? k : 0] : q[j];
                if (IN32 == 2) {
                    const uint2 x = *reinterpret_cast<const uint2 *>(stage + so);
                    v[k][0] = x.x;
                    v[k][1] = x.y;
                } else {
#pragma unroll
                    for (int w = 0; w < IN32; w++) v[k][w] = stage[so + w];
                }
                const uint32_t sign = (uint32_t)((int32_t)v[k][IN32 - 1] >> 31);
#pragma unroll
                for (int w = IN32; w < W; w++) v[k][w] = sign;
                if ((t * E + k) >= cw) {  // padding positions (cw not a multiple of 32*E) contribute nothing
#pragma unroll
                    for (int w = 0; w < W; w++) v[k][w] = 0u;
                }
            }
        }

        // ---- 3. s1 = prefix sum(y1), parked in the swizzled planes ----
        // (the barriers inside block_scan also order every thread's stage reads before the plane writes)
        uint32_t pre[W];
        block_scan<W, E>(v, pre, aux, t, nwarps);
#pragma unroll
        for (int k = 0; k < E; k++) {
            add_limbs<W>(v[k], pre);
            const uint32_t s = slot_of<E>(t, k, T);
#pragma unroll
            for (int w = 0; w < W; w++) planes[w * P + s] = v[k][w];
        }
        __syncthreads();

        // ---- 4. y2 = s1[perm2[i]] ----
#pragma unroll
        for (int g = 0; g < E / G; g++) {
            uint32_t q[G];
            if constexpr (!CACHE_PERM) load_group(tab2, g, q);
#pragma unroll
            for (int j = 0; j < G; j++) {
                const int k = g * G + j;
                const uint32_t sl = CACHE_PERM ? slot2[CACHE_PERM ? k : 0] : q[j];
                const bool valid = (t * E + k) < cw;
#pragma unroll
                for (int w = 0; w < W; w++) v[k][w] = valid ? planes[w * P + sl] : 0u;
            }
        }

        // ---- 5. s2 = prefix sum(y2), parked in the planes again (same conflict-free layout) ----
        block_scan<W, E>(v, pre, aux, t, nwarps);
#pragma unroll
        for (int k = 0; k < E; k++) {
            add_limbs<W>(v[k], pre);
            const uint32_t s = slot_of<E>(t, k, T);
#pragma unroll
            for (int w = 0; w < W; w++) planes[w * P + s] = v[k][w];
        }
        __syncthreads();

        // ---- 6. coalesced write-out with sign extension to out32 words: consecutive lanes read consecutive
        //         codeword entries back (conflict-free by the XOR swizzle) and each 32-byte Int<4> leaves as ONE
        //         256-bit store, so a warp request covers 1 KiB of contiguous output (8 full 128-byte lines). ----
        uint32_t *dst_row = rows_out + (size_t)row * cw * out32;
#pragma unroll
        for (int it = 0; it < E; it++) {
            const uint32_t i = it * T + t;
            if (i < cw) {
                const uint32_t s = slot_of<E>(i / E, i % E, T);
                uint32_t val[W];
#pragma unroll
                for (int w = 0; w < W; w++) val[w] = planes[w * P + s];
                const uint32_t sign = (uint32_t)((int32_t)val[W - 1] >> 31);
                uint32_t *d = dst_row + (size_t)i * out32;
                if ((out32 & 7u) == 0) {
                    constexpr int QV = (W + 7) / 8;  // 32-byte vectors that still carry value words
#pragma unroll
                    for (int qv = 0; qv < QV; qv++) {
                        uint32_t o[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) o[j] = (8 * qv + j < W) ? val[(8 * qv + j < W) ? 8 * qv + j : 0] : sign;
                        st_global_v8(d + 8 * qv, o);
                    }
                    const uint32_t sg[8] = {sign, sign, sign, sign, sign, sign, sign, sign};
                    for (uint32_t q = 8 * QV; q < out32; q += 8) st_global_v8(d + q, sg);
                } else if ((out32 & 3u) == 0) {
                    constexpr int QV = (W + 3) / 4;
#pragma unroll
                    for (int qv = 0; qv < QV; qv++) {
                        uint4 o;
                        o.x = (4 * qv + 0 < W) ? val[(4 * qv + 0 < W) ? 4 * qv + 0 : 0] : sign;
                        o.y = (4 * qv + 1 < W) ? val[(4 * qv + 1 < W) ? 4 * qv + 1 : 0] : sign;
                        o.z = (4 * qv + 2 < W) ? val[(4 * qv + 2 < W) ? 4 * qv + 2 : 0] : sign;
                        o.w = (4 * qv + 3 < W) ? val[(4 * qv + 3 < W) ? 4 * qv + 3 : 0] : sign;
                        st_stream_v4(reinterpret_cast<uint4 *>(d + 4 * qv), o);
                    }
                    const uint4 sg = make_uint4(sign, sign, sign, sign);
                    for (uint32_t q = 4 * QV; q < out32; q += 4) st_stream_v4(reinterpret_cast<uint4 *>(d + q), sg);
                } else {
#pragma unroll
                    for (int w = 0; w < W; w++) d[w] = val[w];
                    for (uint32_t q = W; q < out32; q++) d[q] = sign;
                }
            }
        }
        __syncthreads();  // the planes are reused as the next row's stage
    }
}

// ------------------------------------------------------------------------------------------------------
// host-side launcher
// ------------------------------------------------------------------------------------------------------
namespace {

struct EncodeCfg {
    int E, T;
    bool cache_perm;
};

EncodeCfg pick_cfg(uint32_t cw) {
    int E;
    if (cw >= 8192) E = 16;
    else if (cw >= 2048) E = 8;
    else if (cw >= 512) E = 4;
    else if (cw >= 64) E = 2;
    else E = 1;
    uint32_t T = (cw + E - 1) / E;
    T = (T + 31) / 32 * 32;
    EncodeCfg c{E, (int)T, E <= 8};
    return c;
}

template <int IN32, int W, int E, bool CP, int MAXT, int MINB>
cudaError_t launch_one(const EncodeArgs &a, int T, size_t smem, int grid_cap_per_sm) {
    auto kern = raa_encode_kernel<IN32, W, E, CP, MAXT, MINB>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int occ = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem);
    if (err != cudaSuccess) return err;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    if (grid_cap_per_sm > 0 && occ > grid_cap_per_sm) occ = grid_cap_per_sm;
    uint32_t grid = (uint32_t)a.num_sms * (uint32_t)occ;
    if (grid > a.num_rows) grid = a.num_rows;
    kern<<<grid, T, smem, a.stream>>>(a.evals, a.rows_out, a.tab1, a.tab2, a.num_rows, a.row_len, a.cw, a.out32);
    return cudaGetLastError();
}

template <int IN32, int W>
cudaError_t launch_w(const EncodeArgs &a) {
    const EncodeCfg c = pick_cfg(a.cw);
    const size_t P = (size_t)c.T * c.E;
    const size_t smem = (W * P + 64 * W) * sizeof(uint32_t);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    switch (c.E) {
        case 16:
            if (c.T <= 512) return launch_one<IN32, W, 16, false, 512, 2>(a, c.T, smem, 0);
            return launch_one<IN32, W, 16, false, 1024, 1>(a, c.T, smem, 0);
        case 8: return launch_one<IN32, W, 8, true, 512, 2>(a, c.T, smem, 0);
        case 4: return launch_one<IN32, W, 4, true, 512, 2>(a, c.T, smem, 0);
        case 2: return launch_one<IN32, W, 2, true, 512, 2>(a, c.T, smem, 0);
        default: return launch_one<IN32, W, 1, true, 512, 2>(a, c.T, smem, 0);
    }
}

}  // namespace

size_t encode_perm_padded_len(uint32_t cw) {
    const EncodeCfg c = pick_cfg(cw);
    return (size_t)c.T * c.E;
}

// host side, once per pp: translate the two gather permutations into the kernel's table layout
void build_encode_tables(const uint32_t *perm1, const uint32_t *perm2, uint32_t row_len, uint32_t cw, int in_limbs,
                         uint32_t *tab1, uint32_t *tab2) {
    const EncodeCfg c = pick_cfg(cw);
    const uint32_t E = (uint32_t)c.E, T = (uint32_t)c.T, G = E < 4 ? E : 4;
    const uint32_t sh = E >= 32 ? 0 : (E == 16 ? 1 : E == 8 ? 2 : E == 4 ? 3 : E == 2 ? 4 : 5);
    const uint32_t in32 = (uint32_t)in_limbs * 2;
    for (uint32_t t = 0; t < T; t++) {
        for (uint32_t k = 0; k < E; k++) {
            const uint32_t i = t * E + k;
            const size_t at = (size_t)(k / G) * (T * G) + (size_t)t * G + (k % G);
            if (i < cw) {
                const uint32_t p2 = perm2[i], t2 = p2 / E, k2 = p2 % E;
                tab1[at] = (perm1[i] % row_len) * in32;
                tab2[at] = k2 * T + (t2 ^ ((k2 << sh) & 31u));
            } else {
                tab1[at] = 0;
                tab2[at] = 0;
            }
        }
    }
}

int encode_compute_limbs(int in_limbs, uint32_t cw) {
    int lg = 0;
    while ((1ull << lg) < cw) lg++;
    const int bits = 64 * in_limbs + 2 * lg;
    return (bits + 31) / 32;
}

bool encode_supported(int in_limbs, uint32_t cw) {
    const int W = encode_compute_limbs(in_limbs, cw);
    if (!((in_limbs == 1 && (W == 3 || W == 4)) || (in_limbs == 2 && (W == 5 || W == 6)))) return false;
    const EncodeCfg c = pick_cfg(cw);
    if (c.T > 1024) return false;
    const size_t P = (size_t)c.T * c.E;
    return (W * P + 64 * W) * sizeof(uint32_t) <= 227 * 1024;
}

cudaError_t launch_raa_encode(const EncodeArgs &a) {
    const int W = encode_compute_limbs(a.in_limbs, a.cw);
    if (a.in_limbs == 1 && W <= 3) return launch_w<2, 3>(a);
    if (a.in_limbs == 1 && W == 4) return launch_w<2, 4>(a);
    if (a.in_limbs == 2 && W <= 5) return launch_w<4, 5>(a);
    if (a.in_limbs == 2 && W == 6) return launch_w<4, 6>(a);
    return cudaErrorInvalidConfiguration;
}

}  // namespace zipgpu
