// zinc_b200/csrc/raa_big.cu -- RAA encoder for codewords that do not fit one SM's shared memory (cw > 16384: nv >= 27).
//
// Same function as raa_encode.cu -- repeat -> shuffle(perm1) -> accumulate -> shuffle(perm2) -> accumulate
// (code_raa.rs:89-105,142-171) -- with the row cut into chunks of 8192 codeword positions, one CTA per chunk, and the two
// intermediate vectors in global scratch instead of shared-memory planes:
//   pass 1: y1[i] = widen(row[perm1[i] mod row_len]); scan inside the chunk -> s1_local, chunk total
//   pass 2: y2[i] = s1_local[perm2[i]] + (sum of the totals of the chunks before perm2[i]'s); scan inside the chunk ->
//           s2_local, chunk total
//   pass 3: out[i] = sign_extend(s2_local[i] + sum of the totals of the chunks before i's)
// The prefix over the chunk totals is recomputed by every consumer CTA (a row has cw / 8192 chunks: 4 at nv = 27/28), so
// no pass waits for another CTA and there is no look-back chain.  Rows are processed in batches whose scratch stays
// around 1 GiB.  Traffic is ~3x the compulsory 72 B per evaluation (scratch written and re-read, random 16-byte
// gathers); at these sizes the BLAKE3 passes that follow take 4x longer than this encoder, so the commit is still
// alu-bound.  The permutations are used as uploaded (u32[cw]); no per-pp table translation.
#include <algorithm>
#include <cstdlib>

#include "raa_common.cuh"

namespace zipgpu {

namespace {

constexpr int kBigE = 16, kBigT = 512, kBigChunk = kBigE * kBigT;  // 8192 positions per CTA

template <int W, int SW>
__device__ __forceinline__ void load_entry(const uint32_t *p, uint32_t (&x)[W]) {
    const uint4 a = *reinterpret_cast<const uint4 *>(p);
    x[0] = a.x; x[1] = a.y; x[2] = a.z;
    if constexpr (W > 3) x[3] = a.w;
    if constexpr (SW == 8) {
        const uint4 b = *reinterpret_cast<const uint4 *>(p + 4);
        if constexpr (W > 4) x[4] = b.x;
        if constexpr (W > 5) x[5] = b.y;
    }
}
template <int W, int SW>
__device__ __forceinline__ void store_entry(uint32_t *p, const uint32_t (&x)[W]) {
    uint4 a;
    a.x = x[0]; a.y = x[1]; a.z = x[2]; a.w = W > 3 ? x[W > 3 ? 3 : 0] : 0u;
    *reinterpret_cast<uint4 *>(p) = a;
    if constexpr (SW == 8) {
        uint4 b;
        b.x = W > 4 ? x[W > 4 ? 4 : 0] : 0u; b.y = W > 5 ? x[W > 5 ? 5 : 0] : 0u; b.z = 0u; b.w = 0u;
        *reinterpret_cast<uint4 *>(p + 4) = b;
    }
}

// exclusive prefix over the chunk totals of one row into shared memory: pref[c] = sum_{c' < c} tot[c']
template <int W>
__device__ __forceinline__ void chunk_prefixes(const uint32_t *tot_row, uint32_t nchunks, uint32_t *pref, uint32_t t) {
    if (t == 0) {
        uint32_t acc[W];
#pragma unroll
        for (int w = 0; w < W; w++) acc[w] = 0u;
        for (uint32_t c = 0; c < nchunks; c++) {
#pragma unroll
            for (int w = 0; w < W; w++) pref[c * W + w] = acc[w];
            uint32_t x[W];
#pragma unroll
            for (int w = 0; w < W; w++) x[w] = tot_row[c * W + w];
            add_limbs<W>(acc, x);
        }
    }
    __syncthreads();
}

// PASS 1 / PASS 2 (GATHER2): one CTA = one chunk of one row
template <int IN32, int W, int SW, bool GATHER2>
__global__ void __launch_bounds__(kBigT)
    raa_big_scan_kernel(const uint32_t *__restrict__ evals, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ src_local,
                        const uint32_t *__restrict__ src_tot, uint32_t *__restrict__ dst_local, uint32_t *__restrict__ dst_tot,
                        uint32_t row0, uint32_t row_len, uint32_t cw, uint32_t nchunks) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *aux = smem;                 // block_scan scratch: 64 * W words
    uint32_t *pref = smem + 64 * W;       // GATHER2: nchunks * W words
    const uint32_t t = threadIdx.x, chunk = blockIdx.x, lrow = blockIdx.y;
    const size_t row = (size_t)row0 + lrow;
    if constexpr (GATHER2) chunk_prefixes<W>(src_tot + (size_t)lrow * nchunks * W, nchunks, pref, t);
    uint32_t v[kBigE][W];
    const uint32_t i0 = chunk * kBigChunk + t * kBigE;
#pragma unroll
    for (int k = 0; k < kBigE; k++) {
        const uint32_t i = i0 + k;
        if (i < cw) {
            const uint32_t j = __ldg(perm + i);
            if constexpr (!GATHER2) {
                const uint32_t *e = evals + (row * row_len + (j % row_len)) * IN32;
#pragma unroll
                for (int w = 0; w < IN32; w++) v[k][w] = __ldg(e + w);
                const uint32_t sign = (uint32_t)((int32_t)v[k][IN32 - 1] >> 31);
#pragma unroll
                for (int w = IN32; w < W; w++) v[k][w] = sign;
            } else {
                load_entry<W, SW>(src_local + ((size_t)lrow * cw + j) * SW, v[k]);
                uint32_t p[W];
#pragma unroll
                for (int w = 0; w < W; w++) p[w] = pref[(j / kBigChunk) * W + w];
                add_limbs<W>(v[k], p);
            }
        } else {
#pragma unroll
            for (int w = 0; w < W; w++) v[k][w] = 0u;
        }
    }
    uint32_t pre[W];
    block_scan<W, kBigE>(v, pre, aux, t, kBigT / 32);
#pragma unroll
    for (int k = 0; k < kBigE; k++) {
        add_limbs<W>(v[k], pre);
        if (i0 + k < cw) store_entry<W, SW>(dst_local + ((size_t)lrow * cw + i0 + k) * SW, v[k]);
    }
    if (t == kBigT - 1) {
#pragma unroll
        for (int w = 0; w < W; w++) dst_tot[((size_t)lrow * nchunks + chunk) * W + w] = v[kBigE - 1][w];
    }
}

// PASS 3: out[i] = sign_extend(s2_local[i] + prefix of its chunk), out32 words per entry
template <int W, int SW>
__global__ void __launch_bounds__(kBigT)
    raa_big_out_kernel(const uint32_t *__restrict__ src_local, const uint32_t *__restrict__ src_tot, uint32_t *__restrict__ rows_out,
                       uint32_t row0, uint32_t cw, uint32_t nchunks, uint32_t out32) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *pref = smem;
    const uint32_t t = threadIdx.x, chunk = blockIdx.x, lrow = blockIdx.y;
    const size_t row = (size_t)row0 + lrow;
    chunk_prefixes<W>(src_tot + (size_t)lrow * nchunks * W, nchunks, pref, t);
    uint32_t p[W];
#pragma unroll
    for (int w = 0; w < W; w++) p[w] = pref[chunk * W + w];
    for (uint32_t q = t; q < (uint32_t)kBigChunk; q += kBigT) {  // consecutive lanes write consecutive entries
        const uint32_t i = chunk * kBigChunk + q;
        if (i >= cw) break;
        uint32_t x[W];
        load_entry<W, SW>(src_local + ((size_t)lrow * cw + i) * SW, x);
        add_limbs<W>(x, p);
        const uint32_t sign = (uint32_t)((int32_t)x[W - 1] >> 31);
        uint32_t *d = rows_out + (row * cw + i) * out32;
        if ((out32 & 7u) == 0) {
            for (uint32_t o = 0; o < out32; o += 8) {
                uint32_t rec[8];
#pragma unroll
                for (int j = 0; j < 8; j++) rec[j] = (o + j < (uint32_t)W) ? x[(o + j < (uint32_t)W) ? o + j : 0] : sign;
                st_global_v8(d + o, rec);
            }
        } else {
            for (uint32_t o = 0; o < out32; o++) d[o] = o < (uint32_t)W ? x[o < (uint32_t)W ? o : 0] : sign;
        }
    }
}

// ---- the row-per-CTA form (Int<1> inputs, W <= 4, cw a multiple of 16384) ------------------------------------------
// One persistent 1024-thread CTA per SM takes whole rows, a row in segments of 16384 positions (16 consecutive positions per
// thread and segment, the running total carried from segment to segment):
//   pass 1: y1[i] = widen(row[perm1[i] mod row_len]) straight from global memory (the row is L2 / L1 resident), prefix
//           sum, s1 as 16-byte records into THIS CTA's slice of the scratch -- cw * 16 bytes, written and re-read by the
//           same SM, so it never leaves L2;
//   pass 2: y2[i] = s1[perm2[i]] (one ld.global.cg per entry: the slice is rewritten for every row, so L1 must be
//           bypassed), prefix sum, the finished entry stored sign-extended with one streaming 256-bit store.
// No chunk totals, no third pass, no intermediate vector in DRAM: the compulsory traffic only (8 B in, out32 * 4 B out per
// position) in principle -- measured (ncu, 2048 rows of cw = 32768): the 78 MB of scratch slices do not all stay in L2 under
// the 2 GiB of streaming output (DRAM 2.1 GB read + 4.9 GB written for 2.4 GB compulsory), and the random 8- and 16-byte
// gathers keep the load/store unit's queue full.  nv = 27 (8192 rows): encode 7.78 ms (three chunked launches) -> 6.28 ms,
// commit 22.5 -> 21.0 ms; the BLAKE3 passes that follow take 14.7 ms.  ZIPGPU_BIG_CHUNKED=1 selects the chunked form.
constexpr int kRowT = 1024, kRowE = 16, kRowSeg = kRowT * kRowE;

// FUSE (W == 3, Int<4> entries): the commit form -- after the second prefix sum of a segment every thread also hashes the
// 16 entries it owns (BLAKE3 leaves + tree levels 1..4, read back from its warp's tile) and stores entry, leaf digest and
// nodes itself, like the serial fused kernel of raa_encode.cu: the leaf pass that would re-read the whole codeword from
// HBM (8 GiB at nv = 27) disappears.
template <int W, bool FUSE>
__global__ void __launch_bounds__(kRowT, 1)
    raa_big_row_kernel(const uint32_t *__restrict__ evals, const uint32_t *__restrict__ perm1, const uint32_t *__restrict__ perm2,
                       uint4 *scratch, uint32_t *__restrict__ rows_out, uint32_t num_rows, uint32_t row_len, uint32_t cw,
                       uint32_t out32, uint8_t *__restrict__ layers, uint32_t one) {
    static_assert(!FUSE || W == 3, "the fused form hashes from the 3-limb tiles");
    constexpr int IN32 = 2;
    // W == 3: a warp's 512 entries go through a [3][16][32] tile of its own (the XOR swizzle of slot_of<16>: lane-major
    // writes and position-major reads are both conflict free) so that every global store instruction of a warp covers
    // 32 CONSECUTIVE positions -- 512 contiguous bytes of s1 records, 1 KiB of output.  (Stored straight from the
    // registers that hold 16 consecutive positions per lane, each instruction touched 32 different lines: ncu showed the
    // load/store unit's queue as the limiter, lg_throttle 27 of 77 stalled warps per issue.)
    constexpr bool TILE = W == 3;
    extern __shared__ __align__(16) uint32_t tiles[];  // TILE: [32 warps][3][512] words
    __shared__ uint32_t aux[64 * W];
    __shared__ uint32_t s_tot[W];
    const uint32_t t = threadIdx.x;
    const uint32_t lane = t & 31u;
    uint32_t *tile = tiles + (TILE ? (t >> 5) * (3 * 512) : 0);
    uint4 *s1 = scratch + (size_t)blockIdx.x * cw;
    const uint32_t nseg = cw / kRowSeg;
    const bool pow2 = (row_len & (row_len - 1)) == 0;
    for (uint32_t row = blockIdx.x; row < num_rows; row += gridDim.x) {
        const uint2 *erow = reinterpret_cast<const uint2 *>(evals + (size_t)row * row_len * IN32);
        uint32_t carry[W];
#pragma unroll
        for (int w = 0; w < W; w++) carry[w] = 0u;
        // ---- pass 1 ----
#pragma unroll 1
        for (uint32_t seg = 0; seg < nseg; seg++) {
            const uint32_t i0 = seg * kRowSeg + t * kRowE;
            const uint4 *p4 = reinterpret_cast<const uint4 *>(perm1 + i0);
            uint32_t v[kRowE][W];
#pragma unroll
            for (int q = 0; q < kRowE / 4; q++) {
                const uint4 pi = __ldg(p4 + q);
                const uint32_t src[4] = {pi.x, pi.y, pi.z, pi.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t e = pow2 ? (src[j] & (row_len - 1)) : (src[j] % row_len);
                    const uint2 x = __ldg(erow + e);
                    v[4 * q + j][0] = x.x;
                    v[4 * q + j][1] = x.y;
                    const uint32_t sign = (uint32_t)((int32_t)x.y >> 31);
#pragma unroll
                    for (int w = IN32; w < W; w++) v[4 * q + j][w] = sign;
                }
            }
            uint32_t pre[W];
            block_scan<W, kRowE>(v, pre, aux, t, kRowT / 32);
            add_limbs<W>(pre, carry);
            if constexpr (TILE) {
#pragma unroll
                for (int k = 0; k < kRowE; k++) {
                    add_limbs<W>(v[k], pre);
                    const uint32_t sl = slot_of<kRowE>(lane, k, 32);
#pragma unroll
                    for (int w = 0; w < W; w++) tile[w * 512 + sl] = v[k][w];
                }
                __syncwarp();
                uint4 *dst = s1 + seg * kRowSeg + (t >> 5) * 512;  // this warp's 512 consecutive positions
#pragma unroll
                for (int j = 0; j < kRowE; j++) {
                    const uint32_t p = j * 32 + lane, sl = slot_of<kRowE>(p / kRowE, p % kRowE, 32);
                    dst[p] = make_uint4(tile[sl], tile[512 + sl], tile[1024 + sl], 0u);
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int k = 0; k < kRowE; k++) {
                    add_limbs<W>(v[k], pre);
                    s1[i0 + k] = make_uint4(v[k][0], v[k][1], v[k][2], W > 3 ? v[k][W > 3 ? 3 : 0] : 0u);
                }
            }
            if (t == kRowT - 1) {
#pragma unroll
                for (int w = 0; w < W; w++) s_tot[w] = v[kRowE - 1][w];
            }
            __syncthreads();
#pragma unroll
            for (int w = 0; w < W; w++) carry[w] = s_tot[w];
            __syncthreads();  // (s_tot and aux are reused by the next segment)
        }
        // every s1 record of the row is written (the barriers above order the CTA's own global writes for its own reads)
#pragma unroll
        for (int w = 0; w < W; w++) carry[w] = 0u;
        uint32_t *orow = rows_out + (size_t)row * cw * out32;
        // ---- pass 2 ----
#pragma unroll 1
        for (uint32_t seg = 0; seg < nseg; seg++) {
            const uint32_t i0 = seg * kRowSeg + t * kRowE;
            const uint4 *p4 = reinterpret_cast<const uint4 *>(perm2 + i0);
            uint32_t v[kRowE][W];
#pragma unroll
            for (int q = 0; q < kRowE / 4; q++) {
                const uint4 pi = __ldg(p4 + q);
                const uint32_t src[4] = {pi.x, pi.y, pi.z, pi.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint4 r;
                    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                                 : "l"(s1 + src[j]));
                    v[4 * q + j][0] = r.x;
                    v[4 * q + j][1] = r.y;
                    v[4 * q + j][2] = r.z;
                    if constexpr (W > 3) v[4 * q + j][3] = r.w;
                }
            }
            uint32_t pre[W];
            block_scan<W, kRowE>(v, pre, aux, t, kRowT / 32);
            add_limbs<W>(pre, carry);
            if constexpr (TILE) {
#pragma unroll
                for (int k = 0; k < kRowE; k++) {
                    add_limbs<W>(v[k], pre);
                    const uint32_t sl = slot_of<kRowE>(lane, k, 32);
#pragma unroll
                    for (int w = 0; w < W; w++) tile[w * 512 + sl] = v[k][w];
                }
                __syncwarp();
                if constexpr (FUSE) {
                    uint8_t *lay_row = layers + (size_t)row * (2 * (size_t)cw - 2) * 32;
                    b3::Digest stack[4];
#pragma unroll 1
                    for (uint32_t k = 0; k < (uint32_t)kRowE; k++) {
                        const uint32_t sl = slot_of<kRowE>(lane, k, 32);
                        uint32_t x[8];
                        x[0] = tile[sl];
                        x[1] = tile[512 + sl];
                        x[2] = tile[1024 + sl];
                        const uint32_t sign = (uint32_t)((int32_t)x[2] >> 31);
#pragma unroll
                        for (int w = 3; w < 8; w++) x[w] = sign;
                        const uint32_t idx = i0 + k;  // leaf index within the row
                        st_stream_v8(orow + (size_t)idx * 8, x);
                        b3::Digest d;
                        b3::hash_leaf<8>(x, d.w, one);
                        st_global_v8(lay_row + (size_t)idx * 32, d.w);
#pragma unroll 1
                        for (int l = 0; l < 4; l++) {
                            if ((k >> l) & 1u) {
                                d = b3::hash_node_call(stack[l], d, one);
                                const size_t off = 2 * (size_t)cw - ((2 * (size_t)cw) >> (l + 1));
                                st_global_v8(lay_row + (off + (idx >> (l + 1))) * 32, d.w);
                            } else {
                                stack[l] = d;
                                break;
                            }
                        }
                    }
                } else {
                uint32_t *dw = orow + (size_t)(seg * kRowSeg + (t >> 5) * 512) * out32;
#pragma unroll
                for (int j = 0; j < kRowE; j++) {
                    const uint32_t p = j * 32 + lane, sl = slot_of<kRowE>(p / kRowE, p % kRowE, 32);
                    const uint32_t a0 = tile[sl], a1 = tile[512 + sl], a2 = tile[1024 + sl];
                    const uint32_t sign = (uint32_t)((int32_t)a2 >> 31);
                    uint32_t *d = dw + (size_t)p * out32;
                    const uint32_t rec0[8] = {a0, a1, a2, sign, sign, sign, sign, sign};
                    st_stream_v8(d, rec0);
                    const uint32_t recs[8] = {sign, sign, sign, sign, sign, sign, sign, sign};
                    for (uint32_t o = 8; o < out32; o += 8) st_stream_v8(d + o, recs);
                }
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int k = 0; k < kRowE; k++) {
                    add_limbs<W>(v[k], pre);
                    const uint32_t sign = (uint32_t)((int32_t)v[k][W - 1] >> 31);
                    uint32_t *d = orow + (size_t)(i0 + k) * out32;
                    for (uint32_t o = 0; o < out32; o += 8) {
                        uint32_t rec[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) rec[j] = (o + j < (uint32_t)W) ? v[k][(o + j < (uint32_t)W) ? o + j : 0] : sign;
                        st_stream_v8(d + o, rec);
                    }
                }
            }
            if (t == kRowT - 1) {
#pragma unroll
                for (int w = 0; w < W; w++) s_tot[w] = v[kRowE - 1][w];
            }
            __syncthreads();
#pragma unroll
            for (int w = 0; w < W; w++) carry[w] = s_tot[w];
            __syncthreads();
        }
        // (the next row's pass 1 overwrites s1: every gather of this row is done -- the barriers of the last segment)
    }
}

template <int W, bool FUSE>
cudaError_t launch_big_rows(const BigEncodeArgs &a, uint32_t grid) {
    const size_t smem = W == 3 ? (size_t)(kRowT / 32) * 3 * 512 * sizeof(uint32_t) : 0;  // the warps' transposition tiles
    if (smem) {
        cudaError_t e = cudaFuncSetAttribute(raa_big_row_kernel<W, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    raa_big_row_kernel<W, FUSE><<<grid, kRowT, smem, a.stream>>>(a.evals, a.perm1, a.perm2, reinterpret_cast<uint4 *>(a.scratch),
                                                              a.rows_out, a.num_rows, a.row_len, a.cw, a.out32, a.fuse_layers, 1u);
    return cudaGetLastError();
}

template <int IN32, int W>
cudaError_t launch_big_w(const BigEncodeArgs &a) {
    constexpr int SW = W <= 4 ? 4 : 8;
    const uint32_t nchunks = (a.cw + kBigChunk - 1) / kBigChunk;
    const size_t smem_scan = (64 * W + (size_t)nchunks * W) * sizeof(uint32_t), smem_out = (size_t)nchunks * W * sizeof(uint32_t);
    if (smem_scan > 200 * 1024) return cudaErrorInvalidConfiguration;
    auto k1 = raa_big_scan_kernel<IN32, W, SW, false>;
    auto k2 = raa_big_scan_kernel<IN32, W, SW, true>;
    auto k3 = raa_big_out_kernel<W, SW>;
    cudaError_t err;
    if ((err = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scan)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_out, 16))) != cudaSuccess)
        return err;
    uint32_t *s1 = a.scratch, *s2 = s1 + (size_t)a.batch_rows * a.cw * SW;
    uint32_t *t1 = s2 + (size_t)a.batch_rows * a.cw * SW, *t2 = t1 + (size_t)a.batch_rows * nchunks * W;
    for (uint32_t r0 = 0; r0 < a.num_rows; r0 += a.batch_rows) {
        const uint32_t nr = std::min(a.batch_rows, a.num_rows - r0);
        const dim3 grid(nchunks, nr);
        k1<<<grid, kBigT, 64 * W * sizeof(uint32_t), a.stream>>>(a.evals, a.perm1, nullptr, nullptr, s1, t1, r0, a.row_len, a.cw, nchunks);
        k2<<<grid, kBigT, smem_scan, a.stream>>>(nullptr, a.perm2, s1, t1, s2, t2, r0, a.row_len, a.cw, nchunks);
        k3<<<grid, kBigT, std::max<size_t>(smem_out, 16), a.stream>>>(s2, t2, a.rows_out, r0, a.cw, nchunks, a.out32);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    return cudaSuccess;
}

}  // namespace

// rows per batch and scratch bytes for a big-codeword encode (s1 + s2 local scans, chunk totals): ~1 GiB of scratch
void raa_big_plan(int in_limbs, uint32_t cw, uint32_t num_rows, uint32_t *batch_rows, size_t *scratch_bytes) {
    const int W = encode_compute_limbs(in_limbs, cw), SW = W <= 4 ? 4 : 8;
    const size_t per_row = (size_t)cw * SW * 4 * 2 + (size_t)((cw + kBigChunk - 1) / kBigChunk) * W * 4 * 2;
    size_t rows = std::max<size_t>(1, ((size_t)1 << 30) / per_row);
    rows = std::min<size_t>(rows, std::max<uint32_t>(num_rows, 1));
    *batch_rows = (uint32_t)rows;
    *scratch_bytes = rows * per_row + 256;
}

// tree levels the fused commit form of the row-per-CTA encoder builds (0: no fused form for this shape)
int raa_big_fused_levels(int in_limbs, int out_limbs, uint32_t cw) {
    if (getenv("ZIPGPU_BIG_CHUNKED") || getenv("ZIPGPU_BIG_NO_FUSE")) return 0;
    if (in_limbs != 1 || out_limbs != 4 || cw % kRowSeg != 0 || (cw & (cw - 1)) != 0) return 0;
    return encode_compute_limbs(in_limbs, cw) <= 3 ? 4 : 0;
}

bool raa_big_supported(int in_limbs, uint32_t cw) {
    const int W = encode_compute_limbs(in_limbs, cw);
    return cw <= (1u << 24) && ((in_limbs == 1 && (W == 3 || W == 4)) || (in_limbs == 2 && (W == 5 || W == 6)));
}

cudaError_t launch_raa_encode_big(const BigEncodeArgs &a, int *launches) {
    const int W = encode_compute_limbs(a.in_limbs, a.cw);
    // the row-per-CTA form: Int<1> inputs, whole segments, 32-byte-vector output records, scratch for one row per CTA
    if (a.in_limbs == 1 && W <= 4 && a.cw % kRowSeg == 0 && (a.out32 & 7u) == 0 && a.num_sms > 0 && !getenv("ZIPGPU_BIG_CHUNKED")) {
        uint32_t grid = std::min<uint32_t>((uint32_t)a.num_sms, a.num_rows);
        const size_t per_cta = (size_t)a.cw * sizeof(uint4);
        if (a.scratch_bytes >= per_cta) {
            grid = (uint32_t)std::min<size_t>(grid, a.scratch_bytes / per_cta);
            if (launches) *launches = 1;
            if (a.fuse_layers) {
                if (W > 3 || a.out32 != 8) return cudaErrorInvalidConfiguration;
                return launch_big_rows<3, true>(a, grid);
            }
            return W <= 3 ? launch_big_rows<3, false>(a, grid) : launch_big_rows<4, false>(a, grid);
        }
    }
    if (a.fuse_layers) return cudaErrorInvalidConfiguration;  // (raa_big_fused_levels() says where the fused form exists)
    if (launches) *launches = 3 * (int)((a.num_rows + a.batch_rows - 1) / a.batch_rows);
    if (a.in_limbs == 1 && W <= 3) return launch_big_w<2, 3>(a);
    if (a.in_limbs == 1 && W == 4) return launch_big_w<2, 4>(a);
    if (a.in_limbs == 2 && W <= 5) return launch_big_w<4, 5>(a);
    if (a.in_limbs == 2 && W == 6) return launch_big_w<4, 6>(a);
    return cudaErrorInvalidConfiguration;
}

}  // namespace zipgpu
