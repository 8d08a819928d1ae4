// zinc_b200/csrc/raa_common.cuh -- device building blocks shared by the RAA encoder (raa_encode.cu) and the
// warp-specialised commit kernel (commit_ws.cu): the swizzled s2 layout, the multi-limb block scan, the lane-major
// per-pp tables and the warp-local staging of an input row.  See raa_encode.cu for the design notes.
#pragma once
#include "blake3_dev.cuh"
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
struct CtaBarrier {  // all threads of the CTA
    __device__ __forceinline__ static void sync() { __syncthreads(); }
};
template <int ID, int COUNT>
struct NamedBarrier {  // a subset of the CTA's warps (warp-specialised kernels)
    __device__ __forceinline__ static void sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
    __device__ __forceinline__ static void arrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
};

template <int W, int E, class BAR = CtaBarrier>
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
    BAR::sync();
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
    BAR::sync();
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

// ---- per-pp tables (built once by build_encode_tables, below) ---------------------------------------------
//   tab1 (u16): element index of the staged input row that codeword position i gathers in pass 1
//               ( perm1[i] mod row_len )
//   tab2 (u16): shared-memory address (within a plane) of the s1 entry that position i gathers in pass 2
//   colw (u8) : bank (edge colour) at which the owner of s1 entry i parks it; address = write_group*32 + colour
// stored "lane-major" in groups of 16 bytes so that the entries a thread needs are whole vector loads and
// consecutive lanes read consecutive addresses (the tables are shared by every row: L2-resident).
template <int E>
struct Tab16 {  // u16 entries
    static constexpr int G = E < 8 ? E : 8;          // entries per group
    static constexpr int NG = E / G;                 // groups per thread
    static constexpr int NR = (E + 1) / 2;           // packed registers
    __host__ __device__ static size_t at(uint32_t t, uint32_t k, uint32_t T) {
        return ((size_t)(k / G) * T + t) * G + (k % G);
    }
    __device__ __forceinline__ static void load(const uint16_t *tab, uint32_t t, uint32_t T, uint32_t (&r)[NR]) {
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const uint16_t *p = tab + ((size_t)g * T + t) * G;
            if constexpr (G == 8) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
                r[4 * g] = v.x; r[4 * g + 1] = v.y; r[4 * g + 2] = v.z; r[4 * g + 3] = v.w;
            } else if constexpr (G == 4) {
                const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
                r[0] = v.x; r[1] = v.y;
            } else if constexpr (G == 2) {
                r[0] = __ldg(reinterpret_cast<const uint32_t *>(p));
            } else {
                r[0] = __ldg(p);
            }
        }
    }
    __device__ __forceinline__ static uint32_t get(const uint32_t (&r)[NR], int k) {
        return (k & 1) ? (r[k >> 1] >> 16) : (r[k >> 1] & 0xffffu);
    }
};
template <int E>
struct Tab8 {  // u8 entries
    static constexpr int G = E < 16 ? E : 16;
    static constexpr int NG = E / G;
    static constexpr int NR = (E + 3) / 4;
    __host__ __device__ static size_t at(uint32_t t, uint32_t k, uint32_t T) {
        return ((size_t)(k / G) * T + t) * G + (k % G);
    }
    __device__ __forceinline__ static void load(const uint8_t *tab, uint32_t t, uint32_t T, uint32_t (&r)[NR]) {
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const uint8_t *p = tab + ((size_t)g * T + t) * G;
            if constexpr (G == 16) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
                r[4 * g] = v.x; r[4 * g + 1] = v.y; r[4 * g + 2] = v.z; r[4 * g + 3] = v.w;
            } else if constexpr (G == 8) {
                const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
                r[0] = v.x; r[1] = v.y;
            } else if constexpr (G == 4) {
                r[0] = __ldg(reinterpret_cast<const uint32_t *>(p));
            } else if constexpr (G == 2) {
                r[0] = __ldg(reinterpret_cast<const uint16_t *>(p));
            } else {
                r[0] = __ldg(p);
            }
        }
    }
    __device__ __forceinline__ static uint32_t get(const uint32_t (&r)[NR], int k) {
        return (r[k >> 2] >> (8 * (k & 3))) & 0xffu;
    }
};

// coalesced, read-once copy of one input row into shared memory
__device__ __forceinline__ void stage_row(const uint32_t *src, uint32_t *stage, uint32_t in_words, uint32_t t, uint32_t T) {
    if ((in_words & 3u) == 0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
        uint4 *d4 = reinterpret_cast<uint4 *>(stage);
        for (uint32_t i = t; i < (in_words >> 2); i += T) d4[i] = ld_stream_v4(s4 + i);
    } else {
        for (uint32_t i = t; i < in_words; i += T) stage[i] = src[i];
    }
}

// EXACT shapes: every warp stages ITS OWN 1/nwarps of the input row into ITS OWN plane slots (the slots whose s2
// entries it alone reads back in the write-out), so staging the next row needs no CTA-wide barrier after the
// write-out: linear word x of the warp's chunk -> row r = x / 32 of the warp's 32-word slot rows, column x % 32;
// slot row r lives in plane r / E at k = r % E.  build_encode_tables emits tab1 in the same layout.
template <int IN32, int E>
struct WarpStage {
    static constexpr int WPL = E * IN32 / 2;  // words per lane
    static constexpr int NV = WPL / 4 > 0 ? WPL / 4 : 1;
    uint4 v[NV];
    __device__ __forceinline__ void load(const uint32_t *row_src, uint32_t t) {
        static_assert(WPL % 4 == 0, "warp staging moves 16-byte vectors");
        const uint32_t w = t >> 5, L = t & 31u;
        const uint4 *src = reinterpret_cast<const uint4 *>(row_src + (size_t)w * (32 * WPL));
#pragma unroll
        for (int j = 0; j < WPL / 4; j++) v[j] = ld_stream_v4(src + j * 32 + L);
    }
    __device__ __forceinline__ void store(uint32_t *stage, uint32_t P, uint32_t T, uint32_t t) const {
        const uint32_t w = t >> 5, L = t & 31u;
#pragma unroll
        for (int j = 0; j < WPL / 4; j++) {
            const uint32_t r = 4 * j + (L >> 3), col = (L & 7u) * 4;
            *reinterpret_cast<uint4 *>(stage + (r / E) * P + (r % E) * T + 32 * w + col) = v[j];
        }
    }
    // zero-copy input (the row was read straight from pinned host memory): keep a copy in HBM for the opening phase
    __device__ __forceinline__ void copy_out(uint32_t *row_dst, uint32_t t) const {
        const uint32_t w = t >> 5, L = t & 31u;
        uint4 *dst = reinterpret_cast<uint4 *>(row_dst + (size_t)w * (32 * WPL));
#pragma unroll
        for (int j = 0; j < WPL / 4; j++) st_stream_v4(dst + j * 32 + L, v[j]);
    }
};

}  // namespace zipgpu
