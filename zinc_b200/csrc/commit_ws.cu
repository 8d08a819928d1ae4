// zinc_b200/csrc/commit_ws.cu -- the warp-specialised commit kernel: RAA encode + BLAKE3 leaves + the lowest Merkle
// levels of every row in ONE launch (Int<1> -> Int<4>, cw = 512 ... 8192); what zipgpu_commit* launches for exact shapes.
//
// Replaces commit.rs:69-74 (encode_rows, then MerkleTree::new per row) up to tree level log2(E / U); the batched passes of
// merkle.cu finish the trees.  Two thread groups per CTA, two plane sets in shared memory:
//   warps 0..n-1  (ENC)  : stage -> gather -> scan -> gather -> scan -> park s2 in plane set `buf`, then straight on to
//                          the next row in the other plane set
//   warps n..2n-1 (HASH) : the row parked in `buf`, straight from the planes: every entry is stored to `rows_out` (the
//                          codeword) and hashed (BLAKE3 leaf), the lowest tree levels follow in registers
// Two mbarriers per plane set hand it back and forth (full: ENC arrives / HASH waits; empty: HASH arrives / ENC waits).
// (Round 1 used named barriers, bar.arrive / bar.sync: a bar.sync only completes when ALL waiters have arrived, which put
// the 16 hash warps in lockstep at every row boundary -- the fast ones idled until the slowest was done.  With mbarriers
// a hash warp starts the next row the moment it is parked, and the warps may drift apart by up to a row.)
// This is what the hardware made of the two-CTA fused kernel anyway -- its warp schedulers let one CTA of
// every pair run as if alone (75 us per row, 62 of them hashing) and starved the other -- minus the 13 us per row the
// favoured CTA spent not hashing: here the hash warps never leave the alu pipe.
//
// (Tried and measured slower, round 2: letting the ENC group also reduce the level-log2(E/U) nodes of a unit to its root
// through a shared-memory digest buffer, which would make a commit ONE launch.  A compression is a ~200-step dependent
// chain; in a few ENC warps competing with 16 busy hash warps for the alu pipe it takes ~6 us per tree level, ~55 us
// for the 9 levels of a row -- the ENC group became the critical path: nv = 24 went from 1.94 to 2.19 ms.  A second attempt
// gave the tops to the HASH group, batched over up to 32 units and level by level with children re-read from `layers`:
// correct and one launch per commit, but it only beat the separate passes by 1-3 % below ~8 rows per CTA on two shapes,
// lost 3-6 % on large jobs (the batched passes of merkle.cu run the latency-bound levels at full occupancy), and its mere
// presence in the kernel slowed the main hash loop by 3.6 %.  The narrow tops stay with merkle.cu.
// Numbers: profiles/r2_hash_ab.md.)
//
// Work units.  U = 1: a CTA claims whole rows dynamically and hash thread t continues with the E entries ENC thread t
// produced (one level-log2(E) node per thread and row).  That is right for thousands of rows, but a row is ~65 us of
// hashing, and with 3-4 rows per SM (one 2^24 commit sharded over 8 GPUs; a 2^20 commit on one GPU) the last wave leaves
// a quarter of the machine idle.  U = 2 / 4 hashes a row as U units: in a unit every hash thread takes E / U consecutive
// entries, so all hash warps work on 1/U of the row (half / a quarter of the time), and the units -- not the rows -- are
// split evenly and statically over the CTAs.  A row whose units fall into two CTAs is encoded by both (encoding is a
// tenth of the hashing and runs in the otherwise idle ENC warps); each writes out only the part of the codeword its
// units cover.  The fused part then stops one or two tree levels lower (level log2(E / U)).
#include <algorithm>
#include <cstdlib>

#include "peer_sync.cuh"
#include "raa_common.cuh"

namespace zipgpu {

constexpr int kBarEnc = 1;  // named barrier of the ENC group


// ---- the "tops" epilogue: whole trees in the one launch ----------------------------------------------------------------
// When the main loops are done both thread groups of the CTA (2T threads, the ENC warps no longer idle) finish the trees
// of the units this CTA hashed.  A unit left T nodes at level L0 = log2(E / U) in `layers`; eight units at a time:
//   step A  every thread re-reads 4 consecutive nodes (128 bytes, L2 hits: this CTA wrote them) and reduces them by two
//           levels in registers -- three quarters of all compressions above L0, at full occupancy and without a barrier;
//   step B  the remaining log2(T) - 2 levels through a transposed digest buffer in the (now free) plane memory, one
//           compression of latency and two barriers per level, all eight units side by side.
// U = 2: a row is two units.  If both are this CTA's, one thread joins them; if the CTA boundary runs through the row,
// each CTA publishes its half's root (level depth - 1, in `layers`), fences and bumps the boundary's counter -- the
// one that arrives second finds the other half there and makes the root.
// Replaces the 2-3 latency-bound passes of merkle.cu that otherwise follow (36 us of a 0.27 ms commit of 512 rows).
struct WsTops {
    uint8_t *roots;
    uint32_t *pair_flags;
    const RootsFanout *fan;
    unsigned long long fan_step;
    uint32_t fan_row_begin;
};

constexpr uint32_t kTopsMaxRowsPerCta = 64;  // U == 1: rows a CTA may take when it also finishes their trees

template <int T, int U>
__device__ __forceinline__ void ws_tree_tops(uint32_t *sbuf, uint8_t *__restrict__ layers, const WsTops tp, uint32_t cw,
                                          uint32_t L0, uint32_t depth, const volatile uint32_t *row_list, uint32_t row_count,
                                          uint32_t u0, uint32_t u1, uint32_t one) {
    constexpr uint32_t Q = T / 4;        // nodes per unit after step A
    constexpr uint32_t SB = 8 * Q + 8;   // words per digest word in sbuf
    const uint32_t j = threadIdx.x;
    const size_t row_stride = (2 * (size_t)cw - 2) * 32;
    // units of this CTA, numbered q = 0 .. m-1
    //   U == 1: the rows this CTA took, in the order it took them (row_list, written by the ENC group);
    //   U == 2: units ua + q, ua = u0 rounded down to even (q = 0 may be foreign)
    const uint32_t ua = U == 1 ? 0u : (u0 & ~1u);
    const uint32_t m = U == 1 ? row_count : u1 - ua;
#define UNIT_ROW(q) (U == 1 ? row_list[(q)] : (ua + (q)) >> 1)
#define UNIT_HALF(q) (U == 1 ? 0u : (ua + (q)) & 1u)
#define UNIT_OURS(q) ((q) < m && (U == 1 || ua + (q) >= u0))
#define LEVEL_OFF(l) (2 * (size_t)cw - ((2 * (size_t)cw) >> (l)))  /* first digest of level l (l >= 1) */

    for (uint32_t q0 = 0; q0 < m; q0 += 8) {
        {   // ---- step A
            const uint32_t b = j / Q, i4 = j % Q, q = q0 + b;
            if (UNIT_OURS(q)) {
                const uint32_t row = UNIT_ROW(q), half = UNIT_HALF(q);
                uint8_t *lay_row = layers + (size_t)row * row_stride;
                const uint32_t idx0 = half * T + 4 * i4;  // within level L0
                const uint8_t *src = lay_row + (LEVEL_OFF(L0) + idx0) * 32;
                if (L0 == 0) src = lay_row + (size_t)idx0 * 32;
                b3::Digest n0, n1, c01, c23;
                ld_global_v8_sync(src, n0.w);
                ld_global_v8_sync(src + 32, n1.w);
                c01 = b3::hash_node_call(n0, n1, one);
                ld_global_v8_sync(src + 64, n0.w);
                ld_global_v8_sync(src + 96, n1.w);
                c23 = b3::hash_node_call(n0, n1, one);
                uint8_t *d1 = lay_row + (LEVEL_OFF(L0 + 1) + (idx0 >> 1)) * 32;
                st_global_v8(d1, c01.w);
                st_global_v8(d1 + 32, c23.w);
                n0 = b3::hash_node_call(c01, c23, one);
                st_global_v8(lay_row + (LEVEL_OFF(L0 + 2) + (idx0 >> 2)) * 32, n0.w);
#pragma unroll
                for (int w = 0; w < 8; w++) sbuf[w * SB + b * Q + i4] = n0.w[w];
            }
        }
        __syncthreads();
        // ---- step B: n nodes per unit are produced at level l
        uint32_t l = L0 + 3;
        for (uint32_t n = Q / 2; n >= 1; n >>= 1, l++) {
            const uint32_t b = j / n, i = j % n, q = q0 + b;
            const bool act = b < 8 && UNIT_OURS(q);
            b3::Digest o;
            if (act) {
                b3::Digest lft, rgt;
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    const uint2 v = *reinterpret_cast<const uint2 *>(&sbuf[w * SB + b * Q + 2 * i]);
                    lft.w[w] = v.x;
                    rgt.w[w] = v.y;
                }
                o = b3::hash_node_call(lft, rgt, one);
            }
            __syncthreads();
            if (act) {
                const uint32_t row = UNIT_ROW(q), half = UNIT_HALF(q);
#pragma unroll
                for (int w = 0; w < 8; w++) sbuf[w * SB + b * Q + i] = o.w[w];
                if (l == depth) {  // (U == 1, n == 1) the root
                    st_global_v8(tp.roots + (size_t)row * 32, o.w);
                    if (tp.fan) fan_store_root(tp.fan, tp.fan_step, tp.fan_row_begin + row, o.w);
                } else {
                    st_global_v8(layers + (size_t)row * row_stride + (LEVEL_OFF(l) + half * n + i) * 32, o.w);
                }
            }
            __syncthreads();
        }
        if constexpr (U == 2) {  // ---- the two unit roots of a row (level depth - 1; thread b stored the one of unit b)
            if (j < 8 && UNIT_OURS(q0 + j)) {
                const uint32_t q = q0 + j, row = UNIT_ROW(q), half = UNIT_HALF(q);
                const bool sibling_here = half ? UNIT_OURS(q - 1) : UNIT_OURS(q + 1);  // q0 and ua are even: same batch
                b3::Digest mine, other, root;
                bool make = false;
#pragma unroll
                for (int w = 0; w < 8; w++) mine.w[w] = sbuf[w * SB + j * Q];
                if (sibling_here) {
                    if (half == 0) {
#pragma unroll
                        for (int w = 0; w < 8; w++) other.w[w] = sbuf[w * SB + (j + 1) * Q];
                        make = true;
                    }
                } else {
                    uint32_t *flag = tp.pair_flags + blockIdx.x + (half ? 0u : 1u);  // the boundary this row straddles
                    __threadfence();  // this thread stored the half's root above
                    if (atomicAdd(flag, 1u) == 1u) {
                        __threadfence();
                        ld_global_cg_v8(layers + (size_t)row * row_stride + (LEVEL_OFF(depth - 1) + (half ^ 1u)) * 32, other.w);
                        *flag = 0u;  // re-armed for the next launch
                        make = true;
                    }
                }
                if (make) {
                    root = half ? b3::hash_node_call(other, mine, one) : b3::hash_node_call(mine, other, one);
                    st_global_v8(tp.roots + (size_t)row * 32, root.w);
                    if (tp.fan) fan_store_root(tp.fan, tp.fan_step, tp.fan_row_begin + row, root.w);
                }
            }
            __syncthreads();
        }
    }
    if (tp.fan) fan_finish(tp.fan, tp.fan_step);
#undef UNIT_ROW
#undef UNIT_HALF
#undef UNIT_OURS
#undef LEVEL_OFF
}

// E entries per ENC thread, kWsEnc threads per group: (16, 512) = cw 8192, (8, 512) = cw 4096, (8, 256) = cw 2048,
// (4, 256) = cw 1024, (4, 128) = cw 512 -- the (E, T) of the plain encoder for those shapes, so the same pre-translated
// tables serve both kernels.  The 512-thread CTAs leave room for two CTAs per SM.
template <int E, int kWsEnc, int U, bool TOPS>
__global__ void __launch_bounds__(2 * kWsEnc, kWsEnc == 512 ? 1 : 2)
    commit_ws_kernel(const uint32_t *__restrict__ evals, uint32_t *__restrict__ rows_out,
                     const uint16_t *__restrict__ tab1, const uint16_t *__restrict__ tab2,
                     const uint8_t *__restrict__ colw, uint32_t num_rows, uint8_t *__restrict__ layers, uint32_t one,
                     uint32_t *__restrict__ row_counter, const WsTops tops) {
    constexpr int IN32 = 2, W = 3, OUT32 = 8;
    constexpr uint32_t T = kWsEnc, P = T * E, cw = P, in_words = (P / 2) * IN32;
    constexpr int EH = E / U;                   // entries per hash thread and unit
    static_assert(EH >= 2 && EH * U == E, "units must divide the entries per thread");
    using T16 = Tab16<E>;
    using T8 = Tab8<E>;
    using EncBar = NamedBarrier<kBarEnc, kWsEnc>;
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *planes = smem;                    // [2][W][P]
    uint32_t *aux = smem + 2 * W * P;           // scan scratch of the ENC group
    __shared__ volatile uint32_t s_row[2];      // row parked in each plane set (0xffffffff: no more rows)
    __shared__ volatile uint32_t s_next;
    __shared__ unsigned long long s_full[2], s_empty[2];  // mbarriers: plane set parked / plane set free again
    __shared__ volatile uint32_t s_list[TOPS && U == 1 ? kTopsMaxRowsPerCta : 1];  // TOPS: the rows this CTA took
    __shared__ volatile uint32_t s_count;
    const uint32_t tid = threadIdx.x;
    const uint32_t t = tid & (kWsEnc - 1);      // index within the group
    // U > 1: this CTA's static share of the num_rows * U units
    const uint32_t total_units = num_rows * U;
    const uint32_t u0 = U == 1 ? 0u : (uint32_t)(((uint64_t)blockIdx.x * total_units) / gridDim.x);
    const uint32_t u1 = U == 1 ? 0u : (uint32_t)(((uint64_t)(blockIdx.x + 1) * total_units) / gridDim.x);

    if (tid == 0) {
        mbar_init(&s_full[0], kWsEnc);
        mbar_init(&s_full[1], kWsEnc);
        mbar_init(&s_empty[0], kWsEnc);
        mbar_init(&s_empty[1], kWsEnc);
    }
    __syncthreads();
    // use number n (0, 1, ..) of plane set b is iteration it = 2n + b: its "full" phase has parity n & 1, and the ENC
    // group may refill the set once the "empty" phase of use n - 1 has completed

    if (tid < kWsEnc) {
        // ============================== ENC ==============================
        uint32_t c1[T16::NR], c2[T16::NR], cc[T8::NR];
        T16::load(tab1, t, T, c1);
        const uint32_t wbase = (t >> 5) * (E * 32);
        uint32_t row = U == 1 ? blockIdx.x : u0 / U, it = 0;
        const uint32_t row_end = U == 1 ? num_rows : (u1 + U - 1) / U;  // U > 1: rows [u0 / U, ceil(u1 / U))
        for (; row < row_end; it++) {
            const uint32_t buf = it & 1u;
            uint32_t *pl = planes + buf * (W * P);
            if (it >= 2) {  // the hash warps are done with this plane set: one warp polls, the others sleep on the barrier
                if (t < 32) mbar_wait(&s_empty[buf], ((it >> 1) & 1u) ^ 1u);
                EncBar::sync();
            }
            uint32_t early = U == 1 ? row + gridDim.x : row + 1;
            if constexpr (TOPS && U == 1) {
                // the epilogue needs the rows afterwards; a CTA whose list is full takes no more (the others do: the
                // launcher makes sure the lists of all CTAs together hold every row several times over)
                if (t == 0) {
                    s_list[it] = row;
                    if (it + 1 >= kTopsMaxRowsPerCta) early = 0xffffffffu;
                    else if (row_counter) early = gridDim.x + atomicAdd(row_counter, 1u) + 1u;
                }
            } else {
                if (U == 1 && t == 0 && row_counter) early = gridDim.x + atomicAdd(row_counter, 1u) + 1u;
            }
            {
                WarpStage<IN32, E> ws;
                ws.load(evals + (size_t)row * in_words, t);
                ws.store(pl, P, T, t);
            }
            EncBar::sync();
            uint32_t v[E][W];
#pragma unroll
            for (int k = 0; k < E; k++) {
                const uint32_t so = T16::get(c1, k) * IN32;
                const uint2 x = *reinterpret_cast<const uint2 *>(pl + so);
                v[k][0] = x.x;
                v[k][1] = x.y;
                v[k][2] = (uint32_t)((int32_t)x.y >> 31);
            }
            if (t == 0) {
                if (early < row_end) prefetch_l2_bulk(evals + (size_t)early * in_words, in_words * 4u);
                s_next = early;
            }
            uint32_t pre[W];
            T8::load(colw, t, T, cc);
            block_scan<W, E, EncBar>(v, pre, aux, t, T >> 5);
#pragma unroll
            for (int k = 0; k < E; k++) {
                add_limbs<W>(v[k], pre);
                const uint32_t s1 = wbase + k * 32 + T8::get(cc, k);
#pragma unroll
                for (int w = 0; w < W; w++) pl[w * P + s1] = v[k][w];
            }
            T16::load(tab2, t, T, c2);
            EncBar::sync();
#pragma unroll
            for (int k = 0; k < E; k++) {
                const uint32_t sl = T16::get(c2, k);
#pragma unroll
                for (int w = 0; w < W; w++) v[k][w] = pl[w * P + sl];
            }
            block_scan<W, E, EncBar>(v, pre, aux, t, T >> 5);
#pragma unroll
            for (int k = 0; k < E; k++) {
                add_limbs<W>(v[k], pre);
                const uint32_t s2 = slot_of<E>(t, k, T);
#pragma unroll
                for (int w = 0; w < W; w++) pl[w * P + s2] = v[k][w];
            }
            if (t == 0) s_row[buf] = row;
            mbar_arrive(&s_full[buf]);  // hand the row to the hash warps (every thread's arrive releases its own writes)
            __syncwarp();
            T16::load(tab1, t, T, c1);   // for the next row
            row = s_next;  // published before this iteration's ENC barriers
        }
        {   // no more rows: tell the hash warps through the next plane set
            const uint32_t buf = it & 1u;
            if (it >= 2) mbar_wait(&s_empty[buf], ((it >> 1) & 1u) ^ 1u);
            if (t == 0) {
                s_row[buf] = 0xffffffffu;
                if constexpr (TOPS) s_count = it;
            }
            mbar_arrive(&s_full[buf]);
        }
    } else {
        // ============================== HASH ==============================
        constexpr int H = EH >= 16 ? 4 : EH >= 8 ? 3 : EH >= 4 ? 2 : 1;
        static_assert((1 << H) == EH, "entries per hash thread: a power of two");
        for (uint32_t it = 0;; it++) {
            const uint32_t buf = it & 1u;
            mbar_wait(&s_full[buf], (it >> 1) & 1u);
            const uint32_t row = s_row[buf];
            if (row == 0xffffffffu) break;
            const uint32_t *pl = planes + buf * (W * P);
            uint8_t *lay_row = layers + (size_t)row * (2 * (size_t)cw - 2) * 32;
#ifdef WS_NO_HASH  // A/B builds: the encoder alone
            if (one == 1u) {
                mbar_arrive(&s_empty[buf]);
                continue;
            }
#endif
            // the units of this row that are ours (U == 1: the whole row)
            uint32_t uu = 0, uu_end = 1;
            if constexpr (U > 1) {
                uu = u0 > row * U ? u0 - row * U : 0u;
                uu_end = u1 < (row + 1) * U ? u1 - row * U : (uint32_t)U;
            }
#pragma unroll 1
            for (; uu < uu_end; uu++) {
                const uint32_t pbase = (uu * T + t) * EH;  // first of this thread's EH consecutive positions
                b3::Digest stack[H];
#pragma unroll 1  // (unrolled by 2: 1.847 -> 2.045 ms, the loop no longer fits the instruction cache)
                for (uint32_t k = 0; k < (uint32_t)EH; k++) {
                    const uint32_t idx = pbase + k;  // leaf index within the row
                    const uint32_t s = slot_of<E>(idx / E, idx % E, T);
                    // (building the message from the 3 value words + the sign word directly -- a byte-swapped sign word is
                    // itself -- saves 5 PRMT per leaf and measured 3 % SLOWER: ptxas schedules the compression differently)
                    uint32_t x[OUT32];
#pragma unroll
                    for (int w = 0; w < W; w++) x[w] = pl[w * P + s];
                    const uint32_t sign = (uint32_t)((int32_t)x[W - 1] >> 31);
#pragma unroll
                    for (int w = W; w < OUT32; w++) x[w] = sign;
                    // the codeword entry itself: x IS the 32-byte Int<4> record (round 2: the ENC group used to write the
                    // codeword out through record tiles and bulk stores -- 3 % of the kernel, on the same issue slots;
                    // one streaming store here: fused kernel 1.836 -> 1.803 ms)
                    st_stream_v8(reinterpret_cast<uint8_t *>(rows_out) + ((size_t)row * cw + idx) * 32, x);
                    b3::Digest d;
                    b3::hash_leaf<OUT32>(x, d.w, one);
                    st_global_v8(lay_row + (size_t)idx * 32, d.w);
#pragma unroll 1
                    for (int l = 0; l < H; l++) {
                        if ((k >> l) & 1u) {
#ifdef ZIPGPU_WS_CALL_NODE  // A/B builds
                            d = b3::hash_node_call(stack[l], d, one);
#else
                            {   // inlined: this is the only call site in the loop, so nothing is duplicated, and the call,
                                // its register shuffling and the late stack loads go away (fused kernel 1.847 -> 1.833 ms)
                                b3::Digest o;
                                b3::hash_node(stack[l].w, d.w, o.w, one);
                                d = o;
                            }
#endif
                            const size_t off = 2 * (size_t)cw - ((2 * (size_t)cw) >> (l + 1));
                            st_global_v8(lay_row + (off + (idx >> (l + 1))) * 32, d.w);
                        } else {
                            stack[l] = d;
                            break;
                        }
                    }
                }
            }
            mbar_arrive(&s_empty[buf]);  // the plane set may be overwritten once all hash threads have said so
        }
    }
    if constexpr (TOPS) {
        __syncthreads();  // every unit of this CTA is hashed and its level-L0 nodes are in `layers`; the planes are free
        constexpr int L0 = EH >= 16 ? 4 : EH >= 8 ? 3 : EH >= 4 ? 2 : 1;
        uint32_t depth = 0;
        while ((1u << depth) < cw) depth++;
        ws_tree_tops<kWsEnc, U>(planes, layers, tops, cw, (uint32_t)L0, depth, s_list, s_count, u0, u1, one);
    }
}

namespace {

// Rows up to which the fused launch also finishes the trees (ZIPGPU_WS_TOPS_MAX_ROWS; ZIPGPU_WS_TOPS=0 disables, =1 forces
// it for every shape).  Measured on B200 (scripts/shard_sweep.py, ms per commit, tops / separate passes):
//   cw = 8192: 256 rows 0.161 / 0.164, 512: 0.263 / 0.273, 1024: 0.4995 / 0.508, 2048: 0.9755 / 0.978, 4096: 1.936 / 1.927
//   cw = 4096: 256 rows 0.091 / 0.094, 512: 0.144 / 0.154, 1024: 0.260 / 0.269, 2048: 0.504 / 0.505
// A batch of 8 units costs the epilogue ~22 us (15 us of alu work + the latency of the narrow levels), the separate passes
// ~36 us at 512 rows but, being spread over the whole GPU at full occupancy, no more than the epilogue from ~3000 rows.
// The two-CTA-per-SM variants (cw <= 2048) gain 1-2 % up to ~3.5 rows per CTA and lose 5 % at 7 (cw = 2048: 1024 rows
// 0.149 / 0.151, 2048 rows 0.286 / 0.271): they keep the separate passes.
static uint32_t ws_tops_max_rows() {
    const char *e = getenv("ZIPGPU_WS_TOPS_MAX_ROWS");
    return e ? (uint32_t)atol(e) : 2048u;
}

template <int E, int TENC, int U, bool TOPS>
cudaError_t launch_ws_u(const EncodeArgs &a, uint32_t grid) {
    const size_t ws_smem = (2 * 3 * (size_t)(E * TENC) + 64 * 3) * sizeof(uint32_t);
    static_assert(!TOPS || 8 * (8 * (TENC / 4) + 8) <= 2 * 3 * E * TENC, "the digest buffer of the tops fits in the planes");
    auto kern = commit_ws_kernel<E, TENC, U, TOPS>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws_smem);
    if (err != cudaSuccess) return err;
    uint32_t *row_counter = nullptr;
    if (U == 1) {
        if (grid > a.num_rows) grid = a.num_rows;
        row_counter = a.num_rows >= 2 * grid ? a.row_counter : nullptr;
        if (row_counter) {
            err = cudaMemsetAsync(row_counter, 0xff, 2 * sizeof(uint32_t), a.stream);
            if (err != cudaSuccess) return err;
        }
    } else {
        const uint32_t units = a.num_rows * U;
        if (grid > units) grid = units;
    }
    WsTops tp{};
    if (TOPS) {
        tp.roots = a.tops_roots;
        tp.pair_flags = a.pair_flags;
        tp.fan = a.fan;
        tp.fan_step = a.fan_step;
        tp.fan_row_begin = a.fan_row_begin;
    }
    kern<<<grid, 2 * TENC, ws_smem, a.stream>>>(a.evals, a.rows_out, a.tab1, a.tab2, a.colw, a.num_rows, a.fuse_layers, 1u,
                                                row_counter, tp);
    return cudaGetLastError();
}

// Units per row (measured on B200, scripts/shard_sweep.py, profiles/r2_shard_sweep.md): whole rows while every CTA gets
// ~6 or more of them; halves when a one-CTA-per-SM kernel gets 3..6 rows per CTA (cw = 8192: 512 rows 0.304 -> 0.277 ms,
// cw = 4096: 0.168 -> 0.159 ms).  Below 3 rows per CTA, and for the two-CTA-per-SM variants (whose co-resident CTAs
// already even each other out), finer units only move work into the passes that follow and measure equal or slower.
// ZIPGPU_WS_UNITS forces 1 / 2 for experiments.
template <int E, int TENC>
cudaError_t launch_ws(const EncodeArgs &a, int *fused_levels) {
    const uint32_t grid = (uint32_t)a.num_sms * (TENC == 512 ? 1u : 2u);
    int U = 1;
    if (TENC == 512 && a.num_rows >= 3 * grid && a.num_rows < 6 * grid) U = 2;
    if (const char *env = getenv("ZIPGPU_WS_UNITS")) U = atoi(env) >= 2 ? 2 : 1;
    // whole trees in this launch: when the caller wants the roots and has a boundary-flag array for split rows
    // (U == 1: a CTA lists at most kTopsMaxRowsPerCta rows; keep the total capacity at 4x the rows or more)
    const bool tops_ok = a.tops_roots != nullptr && a.pair_flags != nullptr && grid + 1 <= 512 &&
                         (U == 2 || (uint64_t)a.num_rows * 4 <= (uint64_t)std::min(grid, a.num_rows) * kTopsMaxRowsPerCta);
    bool tops = tops_ok && TENC == 512 && a.num_rows <= ws_tops_max_rows();
    if (const char *env = getenv("ZIPGPU_WS_TOPS")) tops = tops_ok && env[0] != '0';
    int h = 0;
    while ((1 << h) < E / U) h++;
    if (tops) {
        h = 0;
        while ((1u << h) < a.cw) h++;
    }
    if (fused_levels) *fused_levels = h;
    if (a.fan_fused) *a.fan_fused = tops && a.fan != nullptr;
    if (tops) return U == 2 ? launch_ws_u<E, TENC, 2, true>(a, grid) : launch_ws_u<E, TENC, 1, true>(a, grid);
    return U == 2 ? launch_ws_u<E, TENC, 2, false>(a, grid) : launch_ws_u<E, TENC, 1, false>(a, grid);
}

}  // namespace

// the dispatch rule of launch_ws for callers that plan around it (the chunked host pipeline): does a commit of `rows`
// rows of this shape finish its trees in the fused launch?
bool commit_ws_whole_trees(uint32_t cw, uint32_t rows) {
    if (const char *env = getenv("ZIPGPU_WS_TOPS")) return env[0] != '0';
    return (cw == 4096 || cw == 8192) && rows <= ws_tops_max_rows();
}

bool commit_ws_supported(int E, int T) {
    return (E == 16 && T == 512) || (E == 8 && T == 512) || (E == 8 && T == 256) || (E == 4 && T == 256) ||
           (E == 4 && T == 128);
}

cudaError_t launch_commit_ws(const EncodeArgs &a, int E, int T, int *fused_levels) {
    if (E == 16 && T == 512) return launch_ws<16, 512>(a, fused_levels);
    if (E == 8 && T == 512) return launch_ws<8, 512>(a, fused_levels);
    if (E == 8 && T == 256) return launch_ws<8, 256>(a, fused_levels);
    if (E == 4 && T == 256) return launch_ws<4, 256>(a, fused_levels);
    if (E == 4 && T == 128) return launch_ws<4, 128>(a, fused_levels);
    return cudaErrorInvalidConfiguration;
}

}  // namespace zipgpu
