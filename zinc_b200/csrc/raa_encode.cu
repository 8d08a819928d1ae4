// zinc_b200/csrc/raa_encode.cu -- K1: row-batched RAA encoder for sm_100a.
//
// Replaces, for every row of the evaluation matrix (commit.rs:158-183):
//     repeat -> shuffle_seeded(perm_1_seed) -> accumulate -> shuffle_seeded(perm_2_seed) -> accumulate
// (code_raa.rs:89-105,142-171; zip/utils.rs:139-142) over Int<N> -> Int<K> (field/int.rs).
//
// Design (B200-first, not a translation of the Rust loops):
//   * One CTA owns one row at a time and keeps the whole codeword on chip: the only HBM traffic is the
//     compulsory 8 B read + 64 B written per evaluation (INT_LIMBS=1).  The CTA is persistent over rows; the
//     next row's evaluations are fetched into registers while the current row is being written out, so the
//     DRAM latency of the input hides behind the output stores.
//   * repeat o perm1 folds into a gather from the staged input row: y1[i] = widen(row[perm1[i] mod row_len]).
//   * Values are held in W 32-bit limbs (W=3, 96 bit, is exact for Int<1> inputs up to cw = 2^16 because
//     |s2| < 2^63 * cw^2); the stored Int<K> is the sign extension, produced on the way out.
//   * Each thread owns E consecutive codeword positions: a serial carry-chain scan in registers, a
//     warp-shuffle scan of the per-thread totals and one cross-warp step give the row prefix sum.
//   * s1 lives in shared memory as W planes of u32.  Its layout is chosen per pp on the host so that BOTH the
//     owner-thread writes and the perm2 gather are bank-conflict free: elements are the edges of a 32-regular
//     bipartite multigraph (write group = warp x step on one side, read group = warp x step of the gathering
//     position on the other); a proper 32-edge-colouring (Koenig; Euler splits) gives every element a bank
//     such that no group sees a bank twice.  address = write_group * 32 + colour.
//   * s2 uses a thread-striped XOR-swizzled layout slot(t,k) = k*T + (t ^ (k << log2(32/E))) which is
//     conflict free for the owner-thread writes and for the coalesced read-out (consecutive i = t*E + k), so
//     that consecutive lanes store consecutive 32-byte Int<4> values with one 256-bit store each.
#include <cstdlib>
#include <vector>

#include "raa_common.cuh"

namespace zipgpu {

// IN32 : u32 words per input value          W    : u32 limbs carried through the scans
// E    : codeword positions per thread      OUT32: u32 words per output value (0 = run-time `out32`)
// CACHE: the three tables stay in registers for every row of this persistent CTA (E <= 8)
// EXACT: cw == T*E, rep == 2, E*IN32 >= 8: no padding predicates; warp-local staging and write-out (see above)
// FUSE : (EXACT only) the commit kernel: after a row's codeword is written out, every thread BLAKE3-hashes the E
//        entries it owns straight from shared memory and reduces them to one node of level log2(E) of the row's
//        Merkle tree (pcs/utils.rs:87-118), writing levels 0..log2(E) to `layers`.  With two CTAs per SM one CTA's
//        memory-bound encode phases run under the other's ALU-bound hash phase, and the codeword is never re-read
//        from HBM.  The levels above log2(E) are left to the batched passes of merkle.cu.
// BULK : (EXACT, Int<4> outputs) the write-out goes through 1 KiB record tiles in shared memory and bulk asynchronous
//        copies (TMA) instead of 256-bit STG: the LSU data stage, which the stores would occupy for 8 k cycles per row,
//        stays free for the shared-memory traffic of the next row's gathers
template <int IN32, int W, int E, int OUT32, bool CACHE, bool EXACT, bool FUSE, bool BULK, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
    raa_encode_kernel(const uint32_t *__restrict__ evals, uint32_t *__restrict__ rows_out,
                      const uint16_t *__restrict__ tab1, const uint16_t *__restrict__ tab2,
                      const uint8_t *__restrict__ colw, uint32_t num_rows, uint32_t row_len, uint32_t cw,
                      uint32_t out32_rt, uint8_t *__restrict__ layers, uint32_t one, uint32_t *__restrict__ evals_copy,
                      uint32_t *__restrict__ row_counter) {
    static_assert(!FUSE || (EXACT && OUT32 != 0), "the fused commit kernel exists for exact shapes only");
    static_assert(!BULK || (EXACT && OUT32 == 8 && W <= 4), "bulk write-out: exact shapes, 32-byte records");
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t T = blockDim.x, t = threadIdx.x;
    const uint32_t P = T * E;   // plane size in words (>= cw)
    uint32_t *planes = smem;    // [W][P]
    uint32_t *stage = smem;     // input row, aliases the planes (dead before s1 is written)
    uint32_t *aux = smem + (size_t)W * P;
    uint32_t *tile = aux + 64 * W + (t >> 5) * 256;  // BULK: this warp's 1 KiB record tile
    const uint32_t nwarps = T >> 5;
    const uint32_t in_words = row_len * IN32;
    const uint32_t out32 = OUT32 ? (uint32_t)OUT32 : out32_rt;
    using T16 = Tab16<E>;
    using T8 = Tab8<E>;
    // CACHE: the three tables are loaded once.  !CACHE (E = 16 runs at a 64-register budget): tab1 is (re)loaded
    // while the previous row is written out, colw and tab2 while the scans run, so their L2 latency is never
    // exposed and their registers are only live when v[] is not
    uint32_t c1[T16::NR], c2[T16::NR], cc[T8::NR];
    T16::load(tab1, t, T, c1);
    if constexpr (CACHE) {
        T16::load(tab2, t, T, c2);
        T8::load(colw, t, T, cc);
    }
    const uint32_t wbase = (t >> 5) * (E * 32);  // s1 address of this thread's write group for k = 0

    // ---- prologue: stage the first row ----
    uint32_t row = blockIdx.x;
    if (row < num_rows) {
        if constexpr (EXACT) {
            WarpStage<IN32, E> ws;
            ws.load(evals + (size_t)row * in_words, t);
            ws.store(stage, P, T, t);
            if (evals_copy) ws.copy_out(evals_copy + (size_t)row * in_words, t);
        } else {
            stage_row(evals + (size_t)row * in_words, stage, in_words, t, T);
        }
    }
    __syncthreads();

    // Rows are claimed dynamically (the first gridDim.x rows are static): the warp schedulers do not share an SM fairly
    // between its two CTAs -- measured with static rows, one CTA of every pair finished its 14 rows in 1.05 ms and
    // left the other alone for the remaining 0.9 ms, with nobody to hide its encode phases.  Claiming keeps every
    // pair together to the end.  The claim for the row after this one is made early and published through a barrier.
    // FUSE: the claim is made only after the hash phase (a starved CTA spends ~0.4 ms in it and would otherwise sit on
    // a claimed row for that long); the other kernels claim a row ahead, which gives the L2 prefetch of the input a
    // whole row-time of lead.  Every claim is a plain atomicAdd: a row is processed if and only if somebody claimed it,
    // and nobody ever declines one (round 1 had a timing heuristic here that let slow CTAs stop claiming near the end;
    // it made whether a row gets processed depend on %globaltimer and is gone).
    __shared__ volatile uint32_t s_next;
    constexpr uint32_t kPending = 0xffffffffu;
    auto publish_next = [&](uint32_t nx) {
        // pull that row's input into L2 now: one bulk prefetch, no registers
        if (nx < num_rows && !evals_copy)  // (a no-op on system memory)
            prefetch_l2_bulk(evals + (size_t)nx * in_words, in_words * 4u);
        s_next = nx;
    };
    auto claim_late = [&]() {  // thread 0 only, FUSE
        publish_next(gridDim.x + atomicAdd(row_counter, 1u) + 1u);
    };
    constexpr bool LATE_CLAIM = FUSE && MINB >= 2;  // with one CTA per SM nobody hides the unprefetched input load
    while (row < num_rows) {
        // early claim: the atomic is issued here and its result consumed after the first gather, so its latency is
        // off the critical path of the row
        uint32_t early = row + gridDim.x;
        if (t == 0) {
            if (LATE_CLAIM && row_counter) s_next = kPending;
            else if (row_counter) early = gridDim.x + atomicAdd(row_counter, 1u) + 1u;
        }
        // ---- 1. y1 = widen(row[perm1[i] mod row_len]) ----
        uint32_t v[E][W];
        {
#pragma unroll
            for (int k = 0; k < E; k++) {
                const uint32_t so = T16::get(c1, k) * IN32;
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
                if (!EXACT && (t * E + k) >= cw) {  // padding positions contribute nothing
#pragma unroll
                    for (int w = 0; w < W; w++) v[k][w] = 0u;
                }
            }
        }

        if (t == 0 && !(LATE_CLAIM && row_counter)) publish_next(early);  // before the scan's barriers
        // ---- 2. s1 = prefix sum(y1), parked at the edge-coloured addresses ----
        // (the barriers inside block_scan also order every thread's stage reads before the plane writes)
        uint32_t pre[W];
        if constexpr (!CACHE) T8::load(colw, t, T, cc);  // in flight during the scan
        block_scan<W, E>(v, pre, aux, t, nwarps);
        {
#pragma unroll
            for (int k = 0; k < E; k++) {
                add_limbs<W>(v[k], pre);
                const uint32_t s = wbase + k * 32 + T8::get(cc, k);
#pragma unroll
                for (int w = 0; w < W; w++) planes[w * P + s] = v[k][w];
            }
        }
        if constexpr (!CACHE) T16::load(tab2, t, T, c2);  // in flight across the barrier
        __syncthreads();

        // ---- 3. y2 = s1[perm2[i]] (conflict free by construction of the colouring) ----
#pragma unroll
        for (int k = 0; k < E; k++) {
            const uint32_t sl = T16::get(c2, k);
            const bool valid = EXACT || (t * E + k) < cw;
#pragma unroll
            for (int w = 0; w < W; w++) v[k][w] = valid ? planes[w * P + sl] : 0u;
        }

        // ---- 4. s2 = prefix sum(y2), parked in the swizzled planes ----
        block_scan<W, E>(v, pre, aux, t, nwarps);
#pragma unroll
        for (int k = 0; k < E; k++) {
            add_limbs<W>(v[k], pre);
            const uint32_t s = slot_of<E>(t, k, T);
#pragma unroll
            for (int w = 0; w < W; w++) planes[w * P + s] = v[k][w];
        }
        // EXACT: a warp reads back only what it wrote itself
        if constexpr (EXACT) __syncwarp();
        else __syncthreads();

        if constexpr (!CACHE) T16::load(tab1, t, T, c1);  // for the next row; in flight during the write-out

        // ---- 5. coalesced write-out with sign extension to out32 words: consecutive lanes read consecutive
        //         codeword entries back (conflict-free by the XOR swizzle) and each 32-byte Int<4> leaves as ONE
        //         256-bit store, so a warp request covers 1 KiB of contiguous output (8 full 128-byte lines). ----
        uint32_t *dst_row = rows_out + (size_t)row * cw * out32;
        // (FUSE: no write-out phase -- the hash phase below stores each entry as it loads it for the leaf hash)
        if constexpr (!FUSE) {
        if constexpr (BULK) {
            // warp w writes out positions [w*32E, (w+1)*32E), 32 consecutive ones (1 KiB of output) per step: the three
            // limbs come back from the planes, the sign-extended 32-byte records go into the warp's tile (lane L puts
            // the half (L>>2)&1 of its record first: the quarter-warps then hit 8 distinct 16-byte bank groups), and
            // one lane hands the tile to the TMA engine.
            const uint32_t lane = t & 31u;
            const uint32_t ha = (lane >> 2) & 1u;
            uint8_t *dst_w = reinterpret_cast<uint8_t *>(dst_row) + (size_t)(t >> 5) * (32 * E) * 32;
#pragma unroll
            for (int it = 0; it < E; it++) {
                const uint32_t i = (t >> 5) * (32 * E) + it * 32 + lane;
                const uint32_t s = slot_of<E>(i / E, i % E, T);
                uint32_t val[W];
#pragma unroll
                for (int w = 0; w < W; w++) val[w] = planes[w * P + s];
                const uint32_t sign = (uint32_t)((int32_t)val[W - 1] >> 31);
                const uint4 lo = make_uint4(val[0], W > 1 ? val[W > 1 ? 1 : 0] : sign, W > 2 ? val[W > 2 ? 2 : 0] : sign,
                                            W > 3 ? val[W > 3 ? 3 : 0] : sign);
                const uint4 hi = make_uint4(sign, sign, sign, sign);
                if (lane == 0) bulk_wait_read_all();  // the previous step's copy has read the tile
                __syncwarp();
                uint4 *rec = reinterpret_cast<uint4 *>(tile) + lane * 2;
                rec[ha] = ha ? hi : lo;
                rec[ha ^ 1u] = ha ? lo : hi;
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) bulk_store_s2g(dst_w + (size_t)it * 1024, tile, 1024);
            }
        } else {
#pragma unroll
        for (int it = 0; it < E; it++) {
            // EXACT: warp w writes out positions [w*32E, (w+1)*32E), 32 consecutive ones per step
            const uint32_t i = EXACT ? ((t >> 5) * (32 * E) + it * 32 + (t & 31u)) : (it * T + t);
            if (EXACT || i < cw) {
                const uint32_t s = slot_of<E>(i / E, i % E, T);
                uint32_t val[W];
#pragma unroll
                for (int w = 0; w < W; w++) val[w] = planes[w * P + s];
                const uint32_t sign = (uint32_t)((int32_t)val[W - 1] >> 31);
                uint32_t *d = dst_row + (size_t)i * out32;
                if (OUT32 ? (OUT32 % 8 == 0) : ((out32 & 7u) == 0)) {
                    constexpr int QV = (W + 7) / 8;  // 32-byte vectors that still carry value words
#pragma unroll
                    for (int qv = 0; qv < QV; qv++) {
                        uint32_t o[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) o[j] = (8 * qv + j < W) ? val[(8 * qv + j < W) ? 8 * qv + j : 0] : sign;
                        st_global_v8(d + 8 * qv, o);
                    }
                    const uint32_t sg[8] = {sign, sign, sign, sign, sign, sign, sign, sign};
                    if constexpr (OUT32 != 0) {
#pragma unroll
                        for (int q = 8 * QV; q < OUT32; q += 8) st_global_v8(d + q, sg);
                    } else {
                        for (uint32_t q = 8 * QV; q < out32; q += 8) st_global_v8(d + q, sg);
                    }
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
        }
        }
        // ---- 5b. FUSE: leaf hashes and the lowest log2(E) tree levels of the entries this thread owns ----
        if constexpr (FUSE) {
            constexpr int H = E >= 16 ? 4 : E >= 8 ? 3 : E >= 4 ? 2 : E >= 2 ? 1 : 0;
            static_assert((1 << H) == E, "E is a power of two");
            uint8_t *lay_row = layers + (size_t)row * (2 * (size_t)cw - 2) * 32;
            b3::Digest stack[H > 0 ? H : 1];  // binary-counter stack; indexed at run time, so it lives in local memory
#pragma unroll 1
            for (uint32_t k = 0; k < (uint32_t)E; k++) {
                const uint32_t s = slot_of<E>(t, k, T);
                uint32_t x[OUT32];
#pragma unroll
                for (int w = 0; w < W; w++) x[w] = planes[w * P + s];
                const uint32_t sign = (uint32_t)((int32_t)x[W - 1] >> 31);
#pragma unroll
                for (int w = W; w < OUT32; w++) x[w] = sign;
                const uint32_t idx = t * E + k;  // leaf index within the row
                {   // the codeword entry itself: x IS the sign-extended record
                    static_assert(OUT32 % 8 == 0, "fused commit: whole 32-byte vectors per entry");
                    uint8_t *rec = reinterpret_cast<uint8_t *>(dst_row) + (size_t)idx * (OUT32 * 4);
#pragma unroll
                    for (int q = 0; q < OUT32 / 8; q++) {
                        const uint32_t o[8] = {x[8 * q], x[8 * q + 1], x[8 * q + 2], x[8 * q + 3],
                                               x[8 * q + 4], x[8 * q + 5], x[8 * q + 6], x[8 * q + 7]};
                        st_stream_v8(rec + 32 * q, o);
                    }
                }
                b3::Digest d;
                b3::hash_leaf<OUT32>(x, d.w, one);
                st_global_v8(lay_row + (size_t)idx * 32, d.w);
#pragma unroll 1
                for (int l = 0; l < H; l++) {
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
        }
        if constexpr (LATE_CLAIM) {
            if (t == 0 && row_counter) {
                claim_late();
                __threadfence_block();
            }
        }
        // ---- 6. stage the next row (its lines were pulled into L2 at claim time); the planes are reused ----
        uint32_t next = s_next;  // written before the scans' barriers, or (late claim) by thread 0 just now
        if constexpr (LATE_CLAIM) {
            while (next == kPending) {
                __nanosleep(64);
                next = s_next;
            }
        }
        if constexpr (EXACT) {
            __syncwarp();  // this warp's slots are free once IT has written its share out
            // (issuing these loads before the write-out was measured slower: the 16 extra live registers spill)
            if (next < num_rows) {
                WarpStage<IN32, E> ws;
                ws.load(evals + (size_t)next * in_words, t);
                ws.store(stage, P, T, t);
                if (evals_copy) ws.copy_out(evals_copy + (size_t)next * in_words, t);
            }
        } else {
            __syncthreads();
            if (next < num_rows) stage_row(evals + (size_t)next * in_words, stage, in_words, t, T);
        }
        __syncthreads();
        row = next;
    }
    if constexpr (BULK) {
        if ((t & 31u) == 0) bulk_wait_all();  // the tile must stay alive until the TMA engine is done with it
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
namespace {

struct EncodeCfg {
    int E, T;
    bool cache;
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

// the fast variant (and its table layout): exact power-of-two shape, rep = 2, ZipTypes' K = 4N output width
bool cfg_exact(const EncodeCfg &c, uint32_t row_len, uint32_t cw, int in_limbs, uint32_t out32) {
    return (uint32_t)c.T * (uint32_t)c.E == cw && 2 * row_len == cw && (c.E * in_limbs * 2) % 8 == 0 &&
           out32 == 8u * (uint32_t)in_limbs;
}

template <int IN32, int W, int E, int OUT32, bool CACHE, bool EXACT, bool FUSE, int MAXT, int MINB>
cudaError_t launch_one(const EncodeArgs &a, int T, size_t smem) {
    constexpr bool BULK = EXACT && OUT32 == 8 && W <= 4;
    if (BULK) smem += (size_t)(T / 32) * 1024;  // one record tile per warp
    auto kern = raa_encode_kernel<IN32, W, E, OUT32, CACHE, EXACT, FUSE, BULK, MAXT, MINB>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int occ = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem);
    if (err != cudaSuccess) return err;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    static const int env_occ = getenv("ZIPGPU_ENC_CTAS_PER_SM") ? atoi(getenv("ZIPGPU_ENC_CTAS_PER_SM")) : 0;  // tuning knob
    if (env_occ > 0 && env_occ < occ) occ = env_occ;
    uint32_t grid = (uint32_t)a.num_sms * (uint32_t)occ;
    if (grid > a.num_rows) grid = a.num_rows;
    // few rows per CTA: static rows (arming and claiming cost more than the imbalance they remove; measured at nv = 20)
    uint32_t *row_counter = a.num_rows >= 6 * grid ? a.row_counter : nullptr;
    if (row_counter) {  // word 0: claims so far - 1 (word 1 unused)
        err = cudaMemsetAsync(row_counter, 0xff, 2 * sizeof(uint32_t), a.stream);
        if (err != cudaSuccess) return err;
    }
    kern<<<grid, T, smem, a.stream>>>(a.evals, a.rows_out, a.tab1, a.tab2, a.colw, a.num_rows, a.row_len, a.cw,
                                      a.out32, a.fuse_layers, 1u, a.evals_copy, row_counter);
    return cudaGetLastError();
}

// EXACT variants only exist where a lane's share of the input row is whole 16-byte vectors (cfg_exact checks it)
template <int IN32, int W, int E, int OUT32, bool CACHE, bool EXACT, bool FUSE, int MAXT, int MINB>
cudaError_t launch_if_valid(const EncodeArgs &a, int T, size_t smem) {
    if constexpr (EXACT && (E * IN32) % 8 != 0) return cudaErrorInvalidConfiguration;
    else return launch_one<IN32, W, E, OUT32, CACHE, EXACT, FUSE, MAXT, MINB>(a, T, smem);
}

template <int IN32, int W, int OUT32, bool EXACT, bool FUSE>
cudaError_t launch_e(const EncodeArgs &a, const EncodeCfg &c, size_t smem) {
    switch (c.E) {
        case 16:
            // W >= 5 (Int<2> inputs): 16 x W live limbs do not fit 64 registers and the planes only allow one CTA per
            // SM anyway, so those variants are compiled for one resident CTA (up to 128 registers, no spills)
            if (c.T <= 512) return launch_if_valid<IN32, W, 16, OUT32, false, EXACT, FUSE, 512, (W >= 5 ? 1 : 2)>(a, c.T, smem);
            return launch_if_valid<IN32, W, 16, OUT32, false, EXACT, FUSE, 1024, 1>(a, c.T, smem);
        case 8: return launch_if_valid<IN32, W, 8, OUT32, true, EXACT, FUSE, 512, 2>(a, c.T, smem);
        case 4: return launch_if_valid<IN32, W, 4, OUT32, true, EXACT, FUSE, 512, 2>(a, c.T, smem);
        case 2: return launch_if_valid<IN32, W, 2, OUT32, true, EXACT, FUSE, 512, 2>(a, c.T, smem);
        default: return launch_if_valid<IN32, W, 1, OUT32, true, EXACT, FUSE, 512, 2>(a, c.T, smem);
    }
}

// cw = 16384: the single-SM warp-specialised kernel (commit_ws16k.cu) is OPT-IN (ZIPGPU_WS16K_MIN_ROWS=<rows>).  While the
// serial fused kernel had its own write-out phase, ws16k won from 2048 rows (8192 rows: 8.081 vs 8.178 ms, the gain
// being the fifth tree level its 32-leaf hash threads fold in).  Since the serial kernel stores each entry in the hash
// phase (no write-out phase at all) it is faster at every row count (scripts/shard_sweep.py --row-len 8192, ms per
// commit, serial / ws16k): 1024 rows 1.000 / 1.064, 2048: 1.966 / 2.069, 4096: 3.899 / 4.071, 8192: 7.760 / 8.081.
// cw = 16384: the 2-CTA-cluster kernel (commit_wsc.cu) is OPT-IN too (ZIPGPU_WSC=1, from ZIPGPU_WSC_MIN_ROWS rows): three
// variants measured, the best 8.16 ms per nv = 26 commit (commit_wsc.cu has the history).
static bool wsc_enabled() {
    const char *e = getenv("ZIPGPU_WSC");
    return e && e[0] == '1';
}
static uint32_t wsc_min_rows() {
    const char *e = getenv("ZIPGPU_WSC_MIN_ROWS");
    return e ? (uint32_t)atol(e) : 1024u;
}
static uint32_t ws16k_min_rows() {
    const char *e = getenv("ZIPGPU_WS16K_MIN_ROWS");  // read per launch: the tests switch it
    return e ? (uint32_t)atol(e) : 0xffffffffu;
}

template <int IN32, int W>
cudaError_t launch_w(const EncodeArgs &a) {
    const EncodeCfg c = pick_cfg(a.cw);
    const size_t P = (size_t)c.T * c.E;
    const size_t smem = (W * P + 64 * W) * sizeof(uint32_t);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    const bool exact = cfg_exact(c, a.row_len, a.cw, a.in_limbs, a.out32);
#ifdef ZIPGPU_DEV_HOT_ONLY  // development builds: only the nv=24 instantiation (fast ptxas -v iterations)
    if (a.fuse_layers) return launch_one<2, 3, 16, 8, false, true, true, 512, 2>(a, c.T, smem);
    return launch_one<2, 3, 16, 8, false, true, false, 512, 2>(a, c.T, smem);
#else
    // the hot instantiations (ZipTypes K = 4N limbs, exact power-of-two shapes) get a compile-time output width,
    // no padding predicates and the register prefetch of the next row
    // (zero-copy input, evals_copy != NULL, stays on the two-CTA fused kernel, which also writes the HBM copy)
    if (a.fuse_layers && !a.evals_copy && IN32 == 2 && W == 3 && exact && a.out32 == 8 && !getenv("ZIPGPU_NO_WS") &&
        commit_ws_supported(c.E, c.T)) {
        // the warp-specialised commit kernel (commit_ws.cu): an encode group and a hash group per CTA, two plane sets
        return launch_commit_ws(a, c.E, c.T, a.fused_levels_out);
    }
    if (a.fuse_layers && !a.evals_copy && IN32 == 2 && W == 3 && exact && a.out32 == 8 &&
        commit_wsc_supported(a.row_len, a.cw) && wsc_enabled() && !getenv("ZIPGPU_NO_WS") && a.num_rows >= wsc_min_rows()) {
        if (a.fused_levels_out) *a.fused_levels_out = commit_wsc_levels();
        return launch_commit_wsc(a);
    }
    if (a.fuse_layers && !a.evals_copy && IN32 == 2 && W == 3 && exact && a.out32 == 8 && a.perm1_raw &&
        commit_ws16k_supported(a.row_len, a.cw) && !getenv("ZIPGPU_NO_WS") && a.num_rows >= ws16k_min_rows()) {
        if (a.fused_levels_out) *a.fused_levels_out = commit_ws16k_levels();
        return launch_commit_ws16k(a);
    }
    if (a.fuse_layers && a.fused_levels_out) {
        int h = 0;
        while ((1 << h) < c.E) h++;
        *a.fused_levels_out = h;
    }
    if (a.fuse_layers) {
        if (!exact) return cudaErrorInvalidConfiguration;
        return launch_e<IN32, W, 4 * IN32, true, true>(a, c, smem);
    }
    if (exact) return launch_e<IN32, W, 4 * IN32, true, false>(a, c, smem);
    return launch_e<IN32, W, 0, false, false>(a, c, smem);
#endif
}

// Proper edge colouring of a d-regular bipartite multigraph (d a power of two) by recursive Euler splits.
// Edge e joins left node eu[e] and right node ev[e]; colour[e] in [0, d).
void euler_colour(const std::vector<uint32_t> &edges, const std::vector<uint32_t> &eu, const std::vector<uint32_t> &ev,
                  uint32_t n_left, uint32_t n_right, uint32_t d, uint32_t base, std::vector<uint8_t> &colour) {
    if (d == 1) {
        for (uint32_t e : edges) colour[e] = (uint8_t)base;
        return;
    }
    // adjacency over the combined node space [0, n_left) + [n_left, n_left + n_right)
    const uint32_t n = n_left + n_right, m = (uint32_t)edges.size();
    std::vector<uint32_t> head(n + 1, 0);
    for (uint32_t e : edges) {
        head[eu[e] + 1]++;
        head[n_left + ev[e] + 1]++;
    }
    for (uint32_t i = 0; i < n; i++) head[i + 1] += head[i];
    std::vector<uint32_t> adj(2 * (size_t)m), fill(head.begin(), head.end() - 1);
    for (uint32_t j = 0; j < m; j++) {
        const uint32_t e = edges[j];
        adj[fill[eu[e]]++] = j;
        adj[fill[n_left + ev[e]]++] = j;
    }
    std::vector<uint32_t> cur(head.begin(), head.end() - 1);
    std::vector<uint8_t> used(m, 0);
    std::vector<uint32_t> half[2];
    half[0].reserve(m / 2);
    half[1].reserve(m / 2);
    for (uint32_t s = 0; s < n_left; s++) {
        uint32_t u = s;
        for (;;) {  // closed trail from s: edges walked left->right go to half 0, right->left to half 1
            while (cur[u] < head[u + 1] && used[adj[cur[u]]]) cur[u]++;
            if (cur[u] == head[u + 1]) break;  // all degrees are even: a trail can only get stuck at its start
            const uint32_t j = adj[cur[u]++];
            used[j] = 1;
            const uint32_t e = edges[j];
            if (u < n_left) {
                half[0].push_back(e);
                u = n_left + ev[e];
            } else {
                half[1].push_back(e);
                u = eu[e];
            }
        }
    }
    euler_colour(half[0], eu, ev, n_left, n_right, d / 2, base, colour);
    euler_colour(half[1], eu, ev, n_left, n_right, d / 2, base + d / 2, colour);
}

}  // namespace

size_t encode_perm_padded_len(uint32_t cw) {
    const EncodeCfg c = pick_cfg(cw);
    return (size_t)c.T * c.E;
}

// host side, once per pp: translate the two gather permutations into the kernel's tables
void build_encode_tables(const uint32_t *perm1, const uint32_t *perm2, uint32_t row_len, uint32_t cw, int in_limbs,
                         int out_limbs, uint16_t *tab1, uint16_t *tab2, uint8_t *colw) {
    const EncodeCfg c = pick_cfg(cw);
    const uint32_t E = (uint32_t)c.E, T = (uint32_t)c.T, P = T * E;
    const uint32_t in32 = (uint32_t)in_limbs * 2;
    const bool exact = cfg_exact(c, row_len, cw, in_limbs, (uint32_t)out_limbs * 2);
    // where input element e sits in the stage, in units of one element (kernel: word offset = tab1 * IN32)
    auto stage_unit = [&](uint32_t e) -> uint32_t {
        if (!exact) return e;  // linear
        const uint32_t per_warp = 16 * E, epr = 32 / in32;  // elements per warp / per 32-word slot row
        const uint32_t w = e / per_warp, m = e % per_warp, r = m / epr, q = m % epr;
        return ((r / E) * P + (r % E) * T + 32 * w + in32 * q) / in32;
    };
    auto at16 = [&](uint32_t t, uint32_t k) -> size_t {
        const uint32_t G = E < 8 ? E : 8;
        return ((size_t)(k / G) * T + t) * G + (k % G);
    };
    auto at8 = [&](uint32_t t, uint32_t k) -> size_t {
        const uint32_t G = E < 16 ? E : 16;
        return ((size_t)(k / G) * T + t) * G + (k % G);
    };
    auto group_of = [&](uint32_t pos) { return (pos / E / 32) * E + pos % E; };  // (warp, step) of a position

    // bank of every s1 element: identity (the owner's lane) unless the shape is exact, where the colouring
    // makes the gather conflict free as well
    std::vector<uint8_t> colour(P);
    for (uint32_t j = 0; j < P; j++) colour[j] = (uint8_t)((j / E) & 31u);
    if (P == cw && T % 32 == 0) {
        std::vector<uint32_t> eu(P), ev(P), edges(P);
        for (uint32_t i = 0; i < cw; i++) {
            const uint32_t j = perm2[i];  // element j is written by its owner and gathered by position i
            eu[j] = group_of(j);
            ev[j] = group_of(i);
        }
        for (uint32_t j = 0; j < P; j++) edges[j] = j;
        std::vector<uint8_t> col(P, 0xff);
        const uint32_t groups = (T / 32) * E;
        euler_colour(edges, eu, ev, groups, groups, 32, 0, col);
        // verify (cheap) before trusting it: every write group and every read group sees 32 distinct banks
        std::vector<uint32_t> seen_w(groups, 0), seen_r(groups, 0);
        bool ok = true;
        for (uint32_t j = 0; j < P && ok; j++) {
            if (col[j] > 31) { ok = false; break; }
            const uint32_t bit = 1u << col[j];
            if ((seen_w[eu[j]] & bit) || (seen_r[ev[j]] & bit)) ok = false;
            seen_w[eu[j]] |= bit;
            seen_r[ev[j]] |= bit;
        }
        if (ok) colour.swap(col);
    }
    auto addr1 = [&](uint32_t j) { return group_of(j) * 32 + colour[j]; };
    for (uint32_t t = 0; t < T; t++) {
        for (uint32_t k = 0; k < E; k++) {
            const uint32_t i = t * E + k;
            if (i < cw) {
                tab1[at16(t, k)] = (uint16_t)stage_unit(perm1[i] % row_len);
                tab2[at16(t, k)] = (uint16_t)addr1(perm2[i]);
            } else {
                tab1[at16(t, k)] = 0;
                tab2[at16(t, k)] = 0;
            }
            colw[at8(t, k)] = colour[i];
        }
    }
}

// Merkle levels the fused commit kernel produces (0 = no fused variant for this shape): log2 of the entries per thread
int encode_fused_levels(int in_limbs, int out_limbs, uint32_t row_len, uint32_t cw) {
    const EncodeCfg c = pick_cfg(cw);
    if (!cfg_exact(c, row_len, cw, in_limbs, (uint32_t)out_limbs * 2)) return 0;
    if (cw & (cw - 1)) return 0;
    int h = 0;
    while ((1 << h) < c.E) h++;
    return h;
}

int encode_compute_limbs(int in_limbs, uint32_t cw) {
    int lg = 0;
    while ((1ull << lg) < cw) lg++;
    const int bits = 64 * in_limbs + 2 * lg;
    return (bits + 31) / 32;
}

bool encode_supported(int in_limbs, uint32_t cw, uint32_t row_len) {
    const int W = encode_compute_limbs(in_limbs, cw);
    if (!((in_limbs == 1 && (W == 3 || W == 4)) || (in_limbs == 2 && (W == 5 || W == 6)))) return false;
    const EncodeCfg c = pick_cfg(cw);
    if (c.T > 1024) return false;
    if (row_len > 65535 || (size_t)c.T * c.E > 65536) return false;  // u16 tables
    const size_t P = (size_t)c.T * c.E;
    // the staged input row aliases the planes
    if ((size_t)row_len * in_limbs * 2 > (size_t)W * P) return false;
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
