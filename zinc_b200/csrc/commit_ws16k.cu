// zinc_b200/csrc/commit_ws16k.cu -- warp-specialised commit kernel for cw = 16384 (nv = 25 / 26), Int<1> -> Int<4>.
//
// Same job as commit_ws.cu -- RAA encode (code_raa.rs:89-105) + BLAKE3 leaves + the lowest Merkle levels
// (pcs/utils.rs:87-118) of every row in one launch -- for the codeword length whose planes (3 x 16384 words = 192 KiB)
// leave no room for a second plane set, so that commit_ws.cu's "hash the parked row while the next one is encoded into
// the other set" cannot work.  Here the hash warps do not read the planes at all:
//   warps  0..15 (ENC)  : a row in TWO half-passes of 512 threads x 16 entries (thread t acts as "virtual threads" t and
//                         t + 512 of the 1024-thread layout the per-pp tables are built for).  Pass 1 gathers
//                         row[perm1[i] mod row_len] straight from global memory (the 64 KiB row is L2/L1-resident; no
//                         staging copy), scans and parks s1 in the planes at the edge-coloured addresses; pass 2 gathers
//                         s1 through tab2, scans, and writes the finished entries FROM REGISTERS to `rows_out` (one
//                         256-bit store per entry) -- s2 is never parked.  The second half-pass of each scan carries in
//                         the total of the first.
//   warps 16..31 (HASH) : BLAKE3 leaves of the row just written, re-read from `rows_out` (the group's own CTA stored
//                         them a moment ago: L2 hits), 32 consecutive leaves per thread -> one level-5 node, levels 0..5
//                         to `layers`.
// A two-slot ring of row ids and mbarriers hands rows from ENC to HASH; the planes are private to the ENC group, which
// therefore runs up to two rows ahead while the hash warps never leave the alu pipe.  Before (round 1) this shape ran the
// two-CTA fused kernel with ONE 1024-thread CTA per SM, i.e. encode and hash phases in series.  Measured at nv = 26
// (8192 rows): 8.21 -> 8.10 ms per commit (fused part 7.92 + upper passes 0.28 -> 7.94 + 0.15); the gain is small because
// the serial form hashed with all 32 warps of the SM -- the hash group here reaches 0.83 of the alu-pipe peak, against
// 0.87 for the shared-memory-fed groups of commit_ws.cu.
#include <cstdlib>

#include "raa_common.cuh"

namespace zipgpu {

namespace {

constexpr int kEnc = 512, kVT = 1024, kE = 16, kW = 3, kIn32 = 2, kOut32 = 8;
constexpr uint32_t kCw = kVT * kE, kRowLen = kCw / 2, kP = kCw;  // 16384 positions, 8192 evaluations, plane size
constexpr int kBarEnc16k = 1;

__global__ void __launch_bounds__(2 * kEnc, 1)
    commit_ws16k_kernel(const uint32_t *__restrict__ evals, uint32_t *rows_out, const uint32_t *__restrict__ perm1,
                        const uint16_t *__restrict__ tab2, const uint8_t *__restrict__ colw, uint32_t num_rows,
                        uint8_t *__restrict__ layers, uint32_t one, uint32_t *__restrict__ row_counter) {
    using T16 = Tab16<kE>;
    using T8 = Tab8<kE>;
    using EncBar = NamedBarrier<kBarEnc16k, kEnc>;
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *planes = smem;           // [3][16384]: s1 only
    uint32_t *aux = smem + kW * kP;    // scan scratch of the ENC group
    __shared__ volatile uint32_t s_row[2];   // ring of rows handed to the hash warps (0xffffffff: no more rows)
    __shared__ volatile uint32_t s_next;
    __shared__ uint32_t s_tot[kW];           // total of the first half-pass of a scan
    __shared__ unsigned long long s_full[2], s_empty[2];
    const uint32_t tid = threadIdx.x, t = tid & (kEnc - 1);
    if (tid == 0) {
        mbar_init(&s_full[0], kEnc);
        mbar_init(&s_full[1], kEnc);
        mbar_init(&s_empty[0], kEnc);
        mbar_init(&s_empty[1], kEnc);
    }
    __syncthreads();

    if (tid < kEnc) {
        // ============================== ENC ==============================
        uint32_t row = blockIdx.x, it = 0;
        for (; row < num_rows; it++) {
            const uint32_t buf = it & 1u;
            if (it >= 2) {  // the ring slot is free again (one warp polls, the others sleep on the barrier)
                if (t < 32) mbar_wait(&s_empty[buf], ((it >> 1) & 1u) ^ 1u);
                EncBar::sync();
            }
            // the claim for the row after this one is made now (its latency is off the critical path) but published
            // only after the barriers of pass 1: until then slower threads may still be reading s_next for THIS row
            uint32_t early = row + gridDim.x;
            if (t == 0 && row_counter) early = gridDim.x + atomicAdd(row_counter, 1u) + 1u;
            const uint32_t *erow = evals + (size_t)row * (kRowLen * kIn32);
            uint32_t carry[kW] = {0u, 0u, 0u};
            // ---- pass 1: y1 = widen(row[perm1[i] mod row_len]), s1 = prefix sum, parked at the coloured addresses ----
#pragma unroll 1
            for (uint32_t h = 0; h < 2; h++) {
                const uint32_t vt = t + kEnc * h;  // virtual thread of the 1024-thread layout
                uint32_t v[kE][kW];
                {
                    const uint4 *p4 = reinterpret_cast<const uint4 *>(perm1 + (size_t)vt * kE);
#pragma unroll
                    for (int q = 0; q < kE / 4; q++) {
                        const uint4 pi = __ldg(p4 + q);
                        const uint32_t src[4] = {pi.x, pi.y, pi.z, pi.w};
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint2 x = __ldg(reinterpret_cast<const uint2 *>(erow) + (src[j] & (kRowLen - 1)));
                            v[4 * q + j][0] = x.x;
                            v[4 * q + j][1] = x.y;
                            v[4 * q + j][2] = (uint32_t)((int32_t)x.y >> 31);
                        }
                    }
                }
                uint32_t cc[T8::NR];
                T8::load(colw, vt, kVT, cc);
                uint32_t pre[kW];
                block_scan<kW, kE, EncBar>(v, pre, aux, t, kEnc >> 5);
                add_limbs<kW>(pre, carry);
                const uint32_t wbase = (vt >> 5) * (kE * 32);
#pragma unroll
                for (int k = 0; k < kE; k++) {
                    add_limbs<kW>(v[k], pre);
                    const uint32_t s1 = wbase + k * 32 + T8::get(cc, k);
#pragma unroll
                    for (int w = 0; w < kW; w++) planes[w * kP + s1] = v[k][w];
                }
                if (t == kEnc - 1) {
#pragma unroll
                    for (int w = 0; w < kW; w++) s_tot[w] = v[kE - 1][w];
                }
                EncBar::sync();  // the half's s1 entries are parked; s_tot is published
                if (h == 0) {
#pragma unroll
                    for (int w = 0; w < kW; w++) carry[w] = s_tot[w];
                }
            }
            if (t == 0) {
                if (early < num_rows) prefetch_l2_bulk(evals + (size_t)early * (kRowLen * kIn32), kRowLen * kIn32 * 4u);
                s_next = early;
            }
            // ---- pass 2: y2 = s1[perm2[i]] (conflict free by the colouring), s2 = prefix sum, stored from registers ----
#pragma unroll
            for (int w = 0; w < kW; w++) carry[w] = 0u;
#pragma unroll 1
            for (uint32_t h = 0; h < 2; h++) {
                const uint32_t vt = t + kEnc * h;
                uint32_t c2[T16::NR];
                T16::load(tab2, vt, kVT, c2);
                uint32_t v[kE][kW];
#pragma unroll
                for (int k = 0; k < kE; k++) {
                    const uint32_t sl = T16::get(c2, k);
#pragma unroll
                    for (int w = 0; w < kW; w++) v[k][w] = planes[w * kP + sl];
                }
                uint32_t pre[kW];
                block_scan<kW, kE, EncBar>(v, pre, aux, t, kEnc >> 5);
                add_limbs<kW>(pre, carry);
                uint32_t *dst = rows_out + ((size_t)row * kCw + (size_t)vt * kE) * kOut32;
#pragma unroll
                for (int k = 0; k < kE; k++) {
                    add_limbs<kW>(v[k], pre);
                    const uint32_t sign = (uint32_t)((int32_t)v[k][kW - 1] >> 31);
                    const uint32_t rec[8] = {v[k][0], v[k][1], v[k][2], sign, sign, sign, sign, sign};
                    st_global_v8(dst + k * kOut32, rec);
                }
                if (t == kEnc - 1) {
#pragma unroll
                    for (int w = 0; w < kW; w++) s_tot[w] = v[kE - 1][w];
                }
                EncBar::sync();  // after h == 1: every gather of this row is done, the planes may be refilled
                if (h == 0) {
#pragma unroll
                    for (int w = 0; w < kW; w++) carry[w] = s_tot[w];
                }
            }
            if (t == 0) s_row[buf] = row;
            mbar_arrive(&s_full[buf]);  // every thread's arrive releases its own stores to rows_out
            row = s_next;
        }
        {   // no more rows
            const uint32_t buf = it & 1u;
            if (it >= 2) mbar_wait(&s_empty[buf], ((it >> 1) & 1u) ^ 1u);
            if (t == 0) s_row[buf] = 0xffffffffu;
            mbar_arrive(&s_full[buf]);
        }
    } else {
        // ============================== HASH ==============================
        constexpr int EH = kCw / kEnc, H = 5;  // 32 consecutive leaves per thread -> one level-5 node
        static_assert((1 << H) == EH, "leaves per hash thread");
        for (uint32_t it = 0;; it++) {
            const uint32_t buf = it & 1u;
            mbar_wait(&s_full[buf], (it >> 1) & 1u);
            const uint32_t row = s_row[buf];
            if (row == 0xffffffffu) break;
            const uint32_t *cw_row = rows_out + (size_t)row * kCw * kOut32;
            uint8_t *lay_row = layers + (size_t)row * (2 * (size_t)kCw - 2) * 32;
            b3::Digest stack[H];
#ifndef ZIPGPU_WS16K_NOHASH  // (measurement builds: ENC alone = 32 us per row and SM, against 143 us of hashing)
            uint32_t xn[kOut32];  // the next leaf is fetched from L2 while the current one is hashed (7.94 vs 8.06 ms at nv = 26)
            ld_global_v8(cw_row + (size_t)(t * EH) * kOut32, xn);
#pragma unroll 1
            for (uint32_t k = 0; k < (uint32_t)EH; k++) {
                const uint32_t idx = t * EH + k;  // leaf index within the row
                uint32_t x[kOut32];
#pragma unroll
                for (int w = 0; w < kOut32; w++) x[w] = xn[w];
                ld_global_v8(cw_row + (size_t)(k + 1 < (uint32_t)EH ? idx + 1 : idx) * kOut32, xn);
                b3::Digest d;
                b3::hash_leaf<kOut32>(x, d.w, one);
                st_global_v8(lay_row + (size_t)idx * 32, d.w);
#pragma unroll 1
                for (int l = 0; l < H; l++) {
                    if ((k >> l) & 1u) {
                        b3::Digest o;
                        b3::hash_node(stack[l].w, d.w, o.w, one);
                        d = o;
                        const size_t off = 2 * (size_t)kCw - ((2 * (size_t)kCw) >> (l + 1));
                        st_global_v8(lay_row + (off + (idx >> (l + 1))) * 32, d.w);
                    } else {
                        stack[l] = d;
                        break;
                    }
                }
            }
#endif
            mbar_arrive(&s_empty[buf]);
        }
    }
}

}  // namespace

bool commit_ws16k_supported(uint32_t row_len, uint32_t cw) { return cw == kCw && row_len == kRowLen; }
int commit_ws16k_levels() { return 5; }

cudaError_t launch_commit_ws16k(const EncodeArgs &a) {
    const size_t smem = ((size_t)kW * kP + 64 * kW) * sizeof(uint32_t);
    cudaError_t err = cudaFuncSetAttribute(commit_ws16k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    uint32_t grid = (uint32_t)a.num_sms;
    if (grid > a.num_rows) grid = a.num_rows;
    uint32_t *row_counter = a.num_rows >= 2 * grid ? a.row_counter : nullptr;
    if (row_counter) {
        err = cudaMemsetAsync(row_counter, 0xff, 2 * sizeof(uint32_t), a.stream);
        if (err != cudaSuccess) return err;
    }
    commit_ws16k_kernel<<<grid, 2 * kEnc, smem, a.stream>>>(a.evals, a.rows_out, a.perm1_raw, a.tab2, a.colw, a.num_rows,
                                                           a.fuse_layers, 1u, row_counter);
    return cudaGetLastError();
}

}  // namespace zipgpu
