// zinc_b200/csrc/commit_wsc.cu -- the warp-specialised commit kernel for cw = 16384 (nv = 25 / 26) as a 2-CTA CLUSTER
// (opt-in: ZIPGPU_WSC=1).
//
// commit_ws.cu's arrangement -- ENC warps encode row r + 1 while HASH warps hash row r straight from shared memory --
// for the codeword length whose two plane sets (2 x 192 KiB) do not fit one SM: the two CTAs of a cluster take one row
// together, CTA c owning positions 8192 c .. 8192 c + 8191 (virtual threads 512 c .. of a 1024-thread layout, 16
// consecutive positions each).  Same job: RAA encode (code_raa.rs:89-105) + BLAKE3 leaves + tree levels 1..4
// (pcs/utils.rs:87-118) of every row in one launch, Int<1> -> Int<4>.
//   * shared memory holds ONLY the finished entries s2 of a CTA's half of two rows ([16][512] half-planes; the
//     (E = 16, T = 512) parking and hash code of commit_ws.cu) -- the hash warps read local memory only, store each
//     entry to `rows_out` and hash it;
//   * pass 1 gathers row[perm1[i] mod row_len] straight from global memory (the 64 KiB row is L2 / L1 resident), scans
//     inside the CTA's half and stores s1 as 16-byte records to an L2-resident scratch (two rows per cluster);
//   * the ENC groups of the two CTAs meet ONCE per row (s1 of the whole row is in the scratch, CTA 0's total of the first
//     prefix sum in both CTAs): named barrier, one thread arrives (mbarrier.arrive.release.cluster on mapa'd addresses)
//     on its own and the peer's 2-arrival mbarrier, one warp waits with acquire.cluster, named barrier;
//   * pass 2 gathers s1[perm2[i]] from the scratch (one 16-byte ld.global.cg per entry; entries of CTA 1's half get CTA
//     0's total added), scans, and CTA 1 alone waits for CTA 0's total of the second prefix sum (remote store + remote
//     arrive; CTA 0 does not wait).  The HASH groups never take part in any of this (barrier.cluster would stall them).
// Rows are claimed per cluster (CTA 0 claims and stores the row into the peer's s_next).
//
// History and RESULT (B200, nv = 26, 8192 rows; commit_ws16k.cu: fused launch 7.93 ms + 0.15 ms upper passes):
//   v1  both plane sets split over the two SMs' shared memories, the encoder's two gathers through distributed shared
//       memory (ld.shared::cluster), four meetings per row: bit-exact on the first run, fused launch 8.14 ms -- the
//       encoder alone needed 38 us per row (8192 eight-byte + 24576 four-byte scattered DSMEM loads per row and CTA, half
//       remote) and, competing with 16 hash warps, became the critical path;
//   v2  s1 through the L2 scratch instead (this file): encoder alone 41 us -- the record-tile / bulk-store write-out
//       of the codeword, not the gathers, was the expensive part;
//   v3  the hash threads store the codeword (as in commit_ws.cu now): encoder alone 22 us, fused launch 7.88 ms.
// That is 0.7 % better than commit_ws16k.cu per launch, but this kernel stops at tree level 4 (16 leaves per hash thread)
// where ws16k folds in a fifth, so the commit is 8.16 vs 8.08 ms.  All forms of this shape end within 3 % of each other
// (serial: hash at 0.95 of the alu peak plus a serial encode; ws16k: hidden encode, L2-fed hash at 0.83; cluster:
// shared-memory-fed hash, but twice the encoder work per hash warp of the cw = 8192 kernel).
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

__device__ __forceinline__ uint4 ld_global_cg_v4(const void *p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(2 * kT, 1)
    commit_wsc_kernel(const uint32_t *__restrict__ evals, uint32_t *__restrict__ rows_out,
                      const uint32_t *__restrict__ perm1, const uint32_t *__restrict__ perm2, uint4 *scratch,
                      uint32_t num_rows, uint8_t *__restrict__ layers, uint32_t one, uint32_t *__restrict__ row_counter) {
    using EncBar = NamedBarrier<kBarEncC, kT>;
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *planes = smem;                   // [2][W][kPL]: this CTA's halves of two rows' finished entries (s2)
    uint32_t *aux = smem + 2 * kW * kPL;       // scan scratch of the ENC group
    __shared__ volatile uint32_t s_row[2];     // row parked in each plane set (0xffffffff: no more rows)
    __shared__ volatile uint32_t s_next;       // written by CTA 0's claimer (locally and into the peer)
    __shared__ volatile uint32_t s_tot1[2][kW];  // total of CTA 0's half of the first prefix sum, by row parity (both CTAs)
    __shared__ volatile uint32_t s_tot2[kW];     // CTA 1: total of CTA 0's half of the second prefix sum
    __shared__ unsigned long long s_full[2], s_empty[2], s_cs, s_cs2;
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
        mbar_init(&s_cs2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_barrier_all();  // both CTAs are resident and their barriers initialised before anyone touches the peer

    if (tid < kT) {
        // ============================== ENC ==============================
        const uint32_t vt = crank * kT + t;    // virtual thread: positions 16 vt .. 16 vt + 15 of the row
        const uint32_t cs_own = mapa_shared((uint32_t)__cvta_generic_to_shared(&s_cs), crank);
        const uint32_t cs_peer = mapa_shared((uint32_t)__cvta_generic_to_shared(&s_cs), peer);
        const uint32_t cs2_peer = mapa_shared((uint32_t)__cvta_generic_to_shared(&s_cs2), peer);
        const uint32_t next_peer = mapa_shared((uint32_t)__cvta_generic_to_shared(const_cast<uint32_t *>(&s_next)), peer);
        const uint32_t tot1_peer = mapa_shared((uint32_t)__cvta_generic_to_shared(const_cast<uint32_t *>(&s_tot1[0][0])), peer);
        const uint32_t tot2_peer = mapa_shared((uint32_t)__cvta_generic_to_shared(const_cast<uint32_t *>(&s_tot2[0])), peer);
        uint32_t cs_phase = 0, cs2_phase = 0;
        const uint4 *p1v = reinterpret_cast<const uint4 *>(perm1 + (size_t)vt * kE);
        const uint4 *p2v = reinterpret_cast<const uint4 *>(perm2 + (size_t)vt * kE);
        uint32_t row = cluster_id, it = 0;
        for (; row < num_rows; it++) {
            const uint32_t buf = it & 1u;
            uint32_t *pl = planes + buf * (kW * kPL);
            uint4 *s1g = scratch + ((size_t)cluster_id * 2 + buf) * kCw;  // this row's s1 (16-byte records), L2-resident
            uint32_t early = row + num_clusters;
            if (crank == 0 && t == 0 && row_counter) early = num_clusters + atomicAdd(row_counter, 1u) + 1u;
            // ---- pass 1: y1 = widen(row[perm1[i] mod row_len]) straight from global memory (the 64 KiB row is L2 / L1
            // resident), prefix sum inside this CTA's half
            const uint32_t *erow = evals + (size_t)row * (kRowLen * kIn32);
            uint32_t v[kE][kW];
#pragma unroll
            for (int q = 0; q < kE / 4; q++) {
                const uint4 pi = __ldg(p1v + q);
                const uint32_t src[4] = {pi.x, pi.y, pi.z, pi.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint2 x = __ldg(reinterpret_cast<const uint2 *>(erow) + (src[j] & (kRowLen - 1)));
                    v[4 * q + j][0] = x.x;
                    v[4 * q + j][1] = x.y;
                    v[4 * q + j][2] = (uint32_t)((int32_t)x.y >> 31);
                }
            }
            uint32_t pre[kW];
            block_scan<kW, kE, EncBar>(v, pre, aux, t, kT >> 5);
            if (crank == 0 && t == kT - 1) {  // the total of CTA 0's half: to both CTAs (entries of CTA 1's half lack it)
                uint32_t tot[kW];
#pragma unroll
                for (int w = 0; w < kW; w++) tot[w] = v[kE - 1][w];
                add_limbs<kW>(tot, pre);
#pragma unroll
                for (int w = 0; w < kW; w++) {
                    s_tot1[buf][w] = tot[w];
                    st_cluster_u32(tot1_peer + 4u * (buf * kW + w), tot[w]);
                }
            }
            {   // s1 of this thread's 16 consecutive positions (prefix within the CTA's half) -> the row's scratch
                uint4 *dst = s1g + (size_t)vt * kE;
#pragma unroll
                for (int k = 0; k < kE; k++) {
                    add_limbs<kW>(v[k], pre);
                    dst[k] = make_uint4(v[k][0], v[k][1], v[k][2], 0u);
                }
            }
            __threadfence();     // the records are in L2 before this thread meets the others
            // the ENC groups of both CTAs meet: s1 of the whole row is in the scratch, CTA 0's total in both CTAs
            EncBar::sync();
            if (t == 0) {
                mbar_arrive_cluster(cs_own);
                mbar_arrive_cluster(cs_peer);
            }
            if (t < 32) mbar_wait_cluster(&s_cs, cs_phase);
            EncBar::sync();
            cs_phase ^= 1u;
            if (crank == 0 && t == 0) {  // (the peer read the previous claim before it came to this meeting)
                if (early < num_rows) prefetch_l2_bulk(evals + (size_t)early * (kRowLen * kIn32), kRowLen * kIn32 * 4u);
                s_next = early;
                st_cluster_u32(next_peer, early);
            }
            // ---- pass 2: y2 = s1[perm2[i]] from the scratch (one 16-byte L2 read per entry), second prefix sum
            {
                uint32_t t1[kW];
#pragma unroll
                for (int w = 0; w < kW; w++) t1[w] = s_tot1[buf][w];
#pragma unroll
                for (int q = 0; q < kE / 4; q++) {
                    const uint4 pi = __ldg(p2v + q);
                    const uint32_t src[4] = {pi.x, pi.y, pi.z, pi.w};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint4 r = ld_global_cg_v4(s1g + src[j]);
                        const uint32_t m = src[j] >= kCw / 2 ? 0xffffffffu : 0u;  // an entry of CTA 1's half: + CTA 0's total
                        uint32_t add[kW];
#pragma unroll
                        for (int w = 0; w < kW; w++) add[w] = t1[w] & m;
                        v[4 * q + j][0] = r.x;
                        v[4 * q + j][1] = r.y;
                        v[4 * q + j][2] = r.z;
                        add_limbs<kW>(v[4 * q + j], add);
                    }
                }
            }
            block_scan<kW, kE, EncBar>(v, pre, aux, t, kT >> 5);
            if (crank == 0) {
                if (t == kT - 1) {  // CTA 0's total -> CTA 1, which alone waits for it
                    uint32_t tot[kW];
#pragma unroll
                    for (int w = 0; w < kW; w++) tot[w] = v[kE - 1][w];
                    add_limbs<kW>(tot, pre);
#pragma unroll
                    for (int w = 0; w < kW; w++) st_cluster_u32(tot2_peer + 4u * w, tot[w]);
                    mbar_arrive_cluster(cs2_peer);
                }
            } else {
                if (t < 32) mbar_wait_cluster(&s_cs2, cs2_phase);
                EncBar::sync();
                cs2_phase ^= 1u;
                uint32_t x[kW];
#pragma unroll
                for (int w = 0; w < kW; w++) x[w] = s_tot2[w];
                add_limbs<kW>(pre, x);
            }
            if (it >= 2) {  // this CTA's hash warps are done with the plane set (the encoder ran ahead of them until here)
                if (t < 32) mbar_wait(&s_empty[buf], ((it >> 1) & 1u) ^ 1u);
                EncBar::sync();
            }
#pragma unroll
            for (int k = 0; k < kE; k++) {
                add_limbs<kW>(v[k], pre);
                const uint32_t s2 = slot_of<kE>(t, k, kT);
#pragma unroll
                for (int w = 0; w < kW; w++) pl[w * kPL + s2] = v[k][w];
            }
            if (t == 0) s_row[buf] = row;
            mbar_arrive(&s_full[buf]);  // this CTA's half of the row to its hash warps
            row = s_next;  // published after this iteration's meeting
        }
        {   // no more rows: tell the hash warps through the next plane set
            const uint32_t buf = it & 1u;
            if (it >= 2) mbar_wait(&s_empty[buf], ((it >> 1) & 1u) ^ 1u);
            if (t == 0) s_row[buf] = 0xffffffffu;
            mbar_arrive(&s_full[buf]);
        }
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
                st_stream_v8(reinterpret_cast<uint8_t *>(rows_out) + ((size_t)row * kCw + idx) * 32, x);  // the codeword entry
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

size_t commit_wsc_scratch_bytes(int num_sms) { return (size_t)(num_sms / 2) * 2 * kCw * sizeof(uint4); }

cudaError_t launch_commit_wsc(const EncodeArgs &a) {
    if (!a.perm1_raw || !a.perm2_raw || !a.wsc_scratch) return cudaErrorInvalidValue;
    const size_t smem = ((size_t)2 * kW * kPL + 64 * kW) * sizeof(uint32_t);
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
    commit_wsc_kernel<<<2 * clusters, 2 * kT, smem, a.stream>>>(a.evals, a.rows_out, a.perm1_raw, a.perm2_raw,
                                                                reinterpret_cast<uint4 *>(a.wsc_scratch), a.num_rows,
                                                                a.fuse_layers, 1u, row_counter);
    return cudaGetLastError();
}

}  // namespace zipgpu
