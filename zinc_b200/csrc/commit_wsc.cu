// zinc_b200/csrc/commit_wsc.cu -- the warp-specialised commit kernel for cw = 16384 (nv = 25 / 26) as a 2-CTA CLUSTER:
// commit_ws.cu's arrangement (ENC warps encode row r + 1 into one plane set while HASH warps hash row r straight from
// the other) for the codeword length whose two plane sets (2 x 192 KiB) do not fit one SM -- they are split over the
// shared memories of the two SMs of a cluster (distributed shared memory).
//
// Same job: RAA encode (code_raa.rs:89-105) + BLAKE3 leaves + tree levels 1..4 (pcs/utils.rs:87-118) of every row in
// one launch, Int<1> -> Int<4>.  The per-pp tables are those of the (E = 16, T = 1024) encoder layout; CTA c of a
// cluster runs virtual threads 512 c .. 512 c + 511 of it:
//   * a plane in the [k][t] layout (staged row, finished entries s2: slot k * 1024 + t') is split by COLUMN -- CTA c
//     holds the columns of its own threads as a [16][512] half-plane -- and a plane in the [warp][k * 32 + colour] layout
//     (parked s1) by WARP.  Either way everything a thread WRITES is in its own CTA's shared memory, the hash warps read
//     only local memory (the leaves of columns 512 c .. are exactly the 8192 leaves 8192 c ..), and the code for staging,
//     parking, write-out and hashing is the (E = 16, T = 512) code of commit_ws.cu on the half-plane;
//   * only the two GATHERS of the encoder cross over: through tab1 from the staged row, through tab2 from s1 -- half of
//     them land in the peer's shared memory (ld.shared::cluster; ~215 cycles instead of ~30, 16 independent loads per
//     thread, and the ENC warps have three quarters of a row time to spare);
//   * the two prefix sums are CTA-local scans plus the total of CTA 0 handed to CTA 1 (one remote store);
//   * the ENC groups of the two CTAs meet four times per row (staged / gather 1 done + totals / s1 parked / gather 2
//     done + totals) through a 2-arrival mbarrier in each CTA: local named barrier, one thread arrives on its own and
//     (mbarrier.arrive.release.cluster on the mapa'd address) on the peer's barrier, one warp waits with acquire.cluster,
//     named barrier again.  The HASH groups never take part: barrier.cluster would stall them.
// Rows are claimed per cluster (CTA 0 claims and stores the row into the peer's s_next).
//
// Why: the single-SM form for this shape (commit_ws16k.cu) keeps s1 only and lets the hash warps re-read the codeword
// from L2 -- they reach 0.83 of the alu-pipe peak, the shared-memory-fed hash loop 0.87-0.88.
//
// RESULT (B200, nv = 26, 8192 rows): bit-exact, but the fused launch takes 8.14 ms against 7.94 ms for commit_ws16k.cu and
// 7.90 ms for the serial fused kernel -- 73.5 us per row and CTA where the hash loop alone needs 65.5.  The encoder is the
// critical path: alone (hashing compiled out) it needs 38 us per row, against ~15 us for the same 8192 positions inside
// one SM.  Its two gathers are 8192 eight-byte and 24576 four-byte scattered shared::cluster loads per row and CTA, half
// of them remote -- the SM-to-SM network moves ~20 B/clk in wide accesses but far less in 4-byte ones -- and the four
// meetings cost a named barrier, a remote mbarrier round trip and another named barrier each.  With 16 hash warps
// taking four of five issue slots the 38 us stretch past the 65 us budget.  (A random permutation over the whole row
// means half of every intermediate vector must cross between the two CTAs; in bulk that would be 2.6 us per row, but
// there is no room for a receive buffer next to two plane sets.)  The kernel therefore stays OPT-IN (ZIPGPU_WSC=1).
#include <cstdlib>

#include "raa_common.cuh"

namespace zipgpu {

namespace {

constexpr int kT = 512, kE = 16, kW = 3, kIn32 = 2, kOut32 = 8;   // per CTA: 512 ENC + 512 HASH threads
constexpr uint32_t kVT = 1024, kCw = kVT * kE, kRowLen = kCw / 2;   // cluster: 1024 virtual threads, 16384 positions
constexpr uint32_t kPL = kT * kE;                                   // half-plane: 8192 words
constexpr int kBarEncC = 1;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t cta_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t ld_cluster_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 ld_cluster_v2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared::cluster.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long *bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void cluster_barrier_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(2 * kT, 1)
    commit_wsc_kernel(const uint32_t *__restrict__ evals, uint32_t *__restrict__ rows_out,
                      const uint16_t *__restrict__ tab1, const uint16_t *__restrict__ tab2,
                      const uint8_t *__restrict__ colw, uint32_t num_rows, uint8_t *__restrict__ layers, uint32_t one,
                      uint32_t *__restrict__ row_counter) {
    using T16 = Tab16<kE>;
    using T8 = Tab8<kE>;
    using EncBar = NamedBarrier<kBarEncC, kT>;
    constexpr uint32_t in_words_half = (kRowLen / 2) * kIn32;  // words of the input row a CTA stages
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *planes = smem;                   // [2][W][kPL]: this CTA's halves of the two plane sets
    uint32_t *aux = smem + 2 * kW * kPL;       // scan scratch of the ENC group
    uint32_t *tiles = aux + 64 * kW;           // one 1 KiB record tile per ENC warp
    __shared__ volatile uint32_t s_row[2];     // row parked in each plane set (0xffffffff: no more rows)
    __shared__ volatile uint32_t s_next;       // written by CTA 0's claimer (locally and into the peer)
    __shared__ volatile uint32_t s_xtot[kW];   // CTA 1: the scan total of CTA 0
    __shared__ unsigned long long s_full[2], s_empty[2], s_cs;
    const uint32_t tid = threadIdx.x;
    const uint32_t t = tid & (kT - 1);         // index within the group
    const uint32_t crank = cluster_ctarank(), peer = crank ^ 1u;
    const uint32_t cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (tid == 0) {
        mbar_init(&s_full[0], kT);
        mbar_init(&s_full[1], kT);
        mbar_init(&s_empty[0], kT);
        mbar_init(&s_empty[1], kT);
        mbar_init(&s_cs, 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_barrier_all();  // both CTAs are resident and their barriers initialised before anyone touches the peer

    if (tid < kT) {
        // ============================== ENC ==============================
        const uint32_t vt = crank * kT + t;    // virtual thread of the 1024-thread table layout
        const uint32_t planes_sa = (uint32_t)__cvta_generic_to_shared(planes);
        const uint32_t base_c[2] = {mapa_shared(planes_sa, 0), mapa_shared(planes_sa, 1)};
        const uint32_t cs_own = mapa_shared((uint32_t)__cvta_generic_to_shared(&s_cs), crank);
        const uint32_t cs_peer = mapa_shared((uint32_t)__cvta_generic_to_shared(&s_cs), peer);
        const uint32_t next_peer = mapa_shared((uint32_t)__cvta_generic_to_shared(const_cast<uint32_t *>(&s_next)), peer);
        const uint32_t xtot_peer = mapa_shared((uint32_t)__cvta_generic_to_shared(const_cast<uint32_t *>(&s_xtot[0])), peer);
        uint32_t cs_phase = 0;
        // the ENC groups of both CTAs meet: everything either wrote to its own shared memory before is visible to the other
        auto enc_cluster_sync = [&]() {
            EncBar::sync();
            if (t == 0) {
                mbar_arrive_cluster(cs_own);
                mbar_arrive_cluster(cs_peer);
            }
            if (t < 32) mbar_wait_cluster(&s_cs, cs_phase);
            EncBar::sync();
            cs_phase ^= 1u;
        };
        // CTA 1 adds the scan total of CTA 0 (handed over before the meeting that precedes this call)
        auto add_peer_total = [&](uint32_t (&pre)[kW]) {
            if (crank == 1) {
                uint32_t x[kW];
#pragma unroll
                for (int w = 0; w < kW; w++) x[w] = s_xtot[w];
                add_limbs<kW>(pre, x);
            }
        };

        uint32_t c1[T16::NR], c2[T16::NR], cc[T8::NR];
        T16::load(tab1, vt, kVT, c1);
        const uint32_t wbase = (t >> 5) * (kE * 32);
        uint32_t *tile = tiles + (t >> 5) * 256;
        const uint32_t lane = t & 31u, ha = (lane >> 2) & 1u;
        uint32_t row = cluster_id, it = 0;
        for (; row < num_rows; it++) {
            const uint32_t buf = it & 1u;
            uint32_t *pl = planes + buf * (kW * kPL);
            const uint32_t set_off = buf * (kW * kPL) * 4u;  // byte offset of the plane set inside a CTA's planes
            if (it >= 2) {  // this CTA's hash warps are done with its half of the plane set
                if (t < 32) mbar_wait(&s_empty[buf], ((it >> 1) & 1u) ^ 1u);
                EncBar::sync();
            }
            uint32_t early = row + num_clusters;
            if (crank == 0 && t == 0 && row_counter) early = num_clusters + atomicAdd(row_counter, 1u) + 1u;
            {   // this CTA's half of the input row into its own slots ([k][512] layout of plane 0)
                WarpStage<kIn32, kE> ws;
                ws.load(evals + (size_t)row * (kRowLen * kIn32) + (size_t)crank * in_words_half, t);
                ws.store(pl, kPL, kT, t);
            }
            enc_cluster_sync();  // (1) the whole row is staged
            uint32_t v[kE][kW];
#pragma unroll
            for (int k = 0; k < kE; k++) {
                const uint32_t so = T16::get(c1, k) * kIn32;  // word k' * 1024 + column of the [16][1024] staged layout
                const uint32_t col = so & 1023u, loc = ((so >> 10) << 9) | (col & 511u);
                const uint2 x = ld_cluster_v2(base_c[col >> 9] + set_off + loc * 4u);
                v[k][0] = x.x;
                v[k][1] = x.y;
                v[k][2] = (uint32_t)((int32_t)x.y >> 31);
            }
            if (crank == 0 && t == 0) {
                if (early < num_rows) prefetch_l2_bulk(evals + (size_t)early * (kRowLen * kIn32), kRowLen * kIn32 * 4u);
                s_next = early;
                st_cluster_u32(next_peer, early);
            }
            uint32_t pre[kW];
            T8::load(colw, vt, kVT, cc);
            block_scan<kW, kE, EncBar>(v, pre, aux, t, kT >> 5);
            if (crank == 0 && t == kT - 1) {  // the total of CTA 0 -> CTA 1
                uint32_t tot[kW];
#pragma unroll
                for (int w = 0; w < kW; w++) tot[w] = v[kE - 1][w];
                add_limbs<kW>(tot, pre);
#pragma unroll
                for (int w = 0; w < kW; w++) st_cluster_u32(xtot_peer + 4u * w, tot[w]);
            }
            enc_cluster_sync();  // (2) every gather from the staged row is done; the total has arrived
            add_peer_total(pre);
#pragma unroll
            for (int k = 0; k < kE; k++) {
                add_limbs<kW>(v[k], pre);
                const uint32_t s1 = wbase + k * 32 + T8::get(cc, k);
#pragma unroll
                for (int w = 0; w < kW; w++) pl[w * kPL + s1] = v[k][w];
            }
            T16::load(tab2, vt, kVT, c2);
            enc_cluster_sync();  // (3) s1 is parked
#pragma unroll
            for (int k = 0; k < kE; k++) {
                const uint32_t sl = T16::get(c2, k);  // s1 address of the [32 warps][512] layout
                const uint32_t a = base_c[sl >> 13] + set_off + (sl & 8191u) * 4u;
#pragma unroll
                for (int w = 0; w < kW; w++) v[k][w] = ld_cluster_u32(a + (uint32_t)w * (kPL * 4u));
            }
            block_scan<kW, kE, EncBar>(v, pre, aux, t, kT >> 5);
            if (crank == 0 && t == kT - 1) {
                uint32_t tot[kW];
#pragma unroll
                for (int w = 0; w < kW; w++) tot[w] = v[kE - 1][w];
                add_limbs<kW>(tot, pre);
#pragma unroll
                for (int w = 0; w < kW; w++) st_cluster_u32(xtot_peer + 4u * w, tot[w]);
            }
            enc_cluster_sync();  // (4) every gather from s1 is done; the total has arrived
            add_peer_total(pre);
#pragma unroll
            for (int k = 0; k < kE; k++) {
                add_limbs<kW>(v[k], pre);
                const uint32_t s2 = slot_of<kE>(t, k, kT);
#pragma unroll
                for (int w = 0; w < kW; w++) pl[w * kPL + s2] = v[k][w];
            }
            if (t == 0) s_row[buf] = row;
            mbar_arrive(&s_full[buf]);  // this CTA's half of the row to its hash warps
            __syncwarp();
            T16::load(tab1, vt, kVT, c1);  // for the next row; in flight during the write-out
            {   // write-out of this warp's 512 positions (32-byte records into the warp's tile, one bulk store per KiB)
                uint8_t *dst_w = reinterpret_cast<uint8_t *>(rows_out) +
                                 ((size_t)row * kCw + (size_t)crank * (kCw / 2) + (size_t)(t >> 5) * (32 * kE)) * 32;
#pragma unroll
                for (int j = 0; j < kE; j++) {
                    const uint32_t i = (t >> 5) * (32 * kE) + j * 32 + lane;
                    const uint32_t s = slot_of<kE>(i / kE, i % kE, kT);
                    const uint32_t a0 = pl[s], a1 = pl[kPL + s], a2 = pl[2 * kPL + s];
                    const uint32_t sign = (uint32_t)((int32_t)a2 >> 31);
                    const uint4 lo = make_uint4(a0, a1, a2, sign), hi = make_uint4(sign, sign, sign, sign);
                    if (lane == 0) bulk_wait_read_all();
                    __syncwarp();
                    uint4 *rec = reinterpret_cast<uint4 *>(tile) + lane * 2;
                    rec[ha] = ha ? hi : lo;
                    rec[ha ^ 1u] = ha ? lo : hi;
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) bulk_store_s2g(dst_w + (size_t)j * 1024, tile, 1024);
                }
            }
            row = s_next;  // published before meeting (2) of this iteration
        }
        {   // no more rows: tell the hash warps through the next plane set
            const uint32_t buf = it & 1u;
            if (it >= 2) mbar_wait(&s_empty[buf], ((it >> 1) & 1u) ^ 1u);
            if (t == 0) s_row[buf] = 0xffffffffu;
            mbar_arrive(&s_full[buf]);
        }
        if (lane == 0) bulk_wait_all();
    } else {
        // ============================== HASH ==============================
        // the (E = 16, T = 512) hash loop of commit_ws.cu on this CTA's half-planes: leaves 8192 * crank ...
        constexpr int H = 4;
        const uint32_t gbase = crank * (kCw / 2);
        for (uint32_t it = 0;; it++) {
            const uint32_t buf = it & 1u;
            mbar_wait(&s_full[buf], (it >> 1) & 1u);
            const uint32_t row = s_row[buf];
            if (row == 0xffffffffu) break;
            const uint32_t *pl = planes + buf * (kW * kPL);
            uint8_t *lay_row = layers + (size_t)row * (2 * (size_t)kCw - 2) * 32;
            b3::Digest stack[H];
#ifdef WSC_NO_HASH  // A/B builds: the encoder alone
            if (one == 1u) {
                mbar_arrive(&s_empty[buf]);
                continue;
            }
#endif
#pragma unroll 1
            for (uint32_t k = 0; k < (uint32_t)kE; k++) {
                const uint32_t s = slot_of<kE>(t, k, kT);
                const uint32_t idx = gbase + t * kE + k;  // leaf index within the row
                uint32_t x[kOut32];
#pragma unroll
                for (int w = 0; w < kW; w++) x[w] = pl[w * kPL + s];
                const uint32_t sign = (uint32_t)((int32_t)x[kW - 1] >> 31);
#pragma unroll
                for (int w = kW; w < kOut32; w++) x[w] = sign;
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
            mbar_arrive(&s_empty[buf]);
        }
    }
    cluster_barrier_all();  // neither CTA leaves while the other may still read its shared memory
}

}  // namespace

bool commit_wsc_supported(uint32_t row_len, uint32_t cw) { return cw == kCw && row_len == kRowLen; }
int commit_wsc_levels() { return 4; }

cudaError_t launch_commit_wsc(const EncodeArgs &a) {
    const size_t smem = ((size_t)2 * kW * kPL + 64 * kW + (kT / 32) * 256) * sizeof(uint32_t);
    cudaError_t err = cudaFuncSetAttribute(commit_wsc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    uint32_t clusters = (uint32_t)a.num_sms / 2;
    if (clusters > a.num_rows) clusters = a.num_rows;
    if (clusters == 0) return cudaSuccess;
    uint32_t *row_counter = a.num_rows >= 2 * clusters ? a.row_counter : nullptr;
    if (row_counter) {
        err = cudaMemsetAsync(row_counter, 0xff, 2 * sizeof(uint32_t), a.stream);
        if (err != cudaSuccess) return err;
    }
    commit_wsc_kernel<<<2 * clusters, 2 * kT, smem, a.stream>>>(a.evals, a.rows_out, a.tab1, a.tab2, a.colw, a.num_rows,
                                                                a.fuse_layers, 1u, row_counter);
    return cudaGetLastError();
}

}  // namespace zipgpu
