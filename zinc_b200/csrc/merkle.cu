// zinc_b200/csrc/merkle.cu -- K2/K3: BLAKE3 leaf hashing and per-row Merkle levels for sm_100a.
//
// Replaces MerkleTree::new batched over all rows of a commit (zip/pcs/utils.rs:74-118, called once per row,
// serially, from commit.rs:71-74).  The layers buffer reproduces `MerkleTree::layers` exactly:
//   per row [leaf hashes (cw) | level depth-1 (cw/2) | ... | level 1 (2)], the root goes to roots[row].
//
// Design: the work is pure INT32 ALU (one compression per hash), so the kernels are organised to keep every
// lane busy rather than around memory.  Each thread reduces 2^H *consecutive* children to one node with a
// binary-counter stack held in registers (no shuffles, no shared memory, no idle lanes inside the subtree),
// writing every intermediate level to `layers` on the way.  Level 0..3 (15/16 of all compressions) run in
// the first pass; each further pass consumes 3 levels (the last one up to 4), batched over ALL rows so the
// narrow top of the trees still fills the GPU (4096 rows x 1 node at nv=24).
#include "blake3_dev.cuh"
#include "common.cuh"
#include "kernels.h"
#include "peer_sync.cuh"

#include <algorithm>
#include <cstdlib>

namespace zipgpu {

namespace {

struct TreeGeom {
    uint32_t cw;        // leaves per row = 1 << depth
    uint32_t depth;
    size_t row_stride;  // digests per row in layers = 2*cw - 2
};

// address of node `idx` of level `level` (0 = leaf hashes, depth = root) of row `row`
__device__ __forceinline__ uint4 *node_ptr(uint8_t *layers, uint8_t *roots, const TreeGeom &g, uint32_t row,
                                           uint32_t level, uint32_t idx) {
    if (level == g.depth) return reinterpret_cast<uint4 *>(roots + (size_t)row * 32);
    const size_t off = 2 * (size_t)g.cw - ((2 * (size_t)g.cw) >> level);
    return reinterpret_cast<uint4 *>(layers + ((size_t)row * g.row_stride + off + idx) * 32);
}

__device__ __forceinline__ void store_digest(uint4 *p, const uint32_t (&d)[8]) { st_global_v8(p, d); }
__device__ __forceinline__ void load_digest(const uint4 *p, uint32_t (&d)[8]) { ld_global_v8(p, d); }

template <int LEAF32>
__device__ __forceinline__ void load_leaf(const uint32_t *p, uint32_t (&x)[LEAF32]) {
    if constexpr (LEAF32 % 8 == 0) {
#pragma unroll
        for (int i = 0; i < LEAF32 / 8; i++) {
            uint32_t a[8];
            ld_stream_v8(p + 8 * i, a);
#pragma unroll
            for (int j = 0; j < 8; j++) x[8 * i + j] = a[j];
        }
    } else if constexpr (LEAF32 % 4 == 0) {
        const uint4 *p4 = reinterpret_cast<const uint4 *>(p);
#pragma unroll
        for (int i = 0; i < LEAF32 / 4; i++) {
            const uint4 a = ld_stream_v4(p4 + i);
            x[4 * i] = a.x; x[4 * i + 1] = a.y; x[4 * i + 2] = a.z; x[4 * i + 3] = a.w;
        }
    } else {
        const uint2 *p2 = reinterpret_cast<const uint2 *>(p);
#pragma unroll
        for (int i = 0; i < LEAF32 / 2; i++) {
            const uint2 a = __ldg(p2 + i);
            x[2 * i] = a.x; x[2 * i + 1] = a.y;
        }
    }
}

// LEAF32 > 0: inputs are raw leaves (LEAF32 u32 words each), level_in must be 0 and level-0 digests are produced.
// LEAF32 == 0: inputs are the digests of level `level_in`, already in `layers`.
template <int LEAF32, int H, int MINB>
__global__ void __launch_bounds__(128, MINB)
    merkle_subtree_kernel(const uint32_t *__restrict__ leaves, uint8_t *layers, uint8_t *roots, uint32_t num_rows,
                          TreeGeom g, uint32_t level_in, uint32_t one, const RootsFanout *__restrict__ fan,
                          unsigned long long fan_step, uint32_t fan_row_begin) {
    pdl_wait();  // (launched with programmatic stream serialisation: the producer of `leaves` / `layers` may still be draining)
    const uint32_t chunks_per_row = (g.cw >> level_in) >> H;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = gid < (size_t)num_rows * chunks_per_row;
    if (!active && !fan) return;
    const uint32_t row = active ? (uint32_t)(gid / chunks_per_row) : 0u;
    const uint32_t c = (uint32_t)(gid % chunks_per_row);

    uint32_t stack[(H > 0) ? H : 1][8];
#pragma unroll 1
    for (uint32_t i = 0; i < (active ? (1u << H) : 0u); i++) {
        const uint32_t idx = (c << H) | i;  // index within level_in
        uint32_t d[8];
        if constexpr (LEAF32 > 0) {
            uint32_t x[LEAF32];
            load_leaf<LEAF32>(leaves + ((size_t)row * g.cw + idx) * LEAF32, x);
            b3::hash_leaf<LEAF32>(x, d, one);
            store_digest(node_ptr(layers, roots, g, row, 0, idx), d);
        } else {
            load_digest(node_ptr(layers, roots, g, row, level_in, idx), d);
        }
        if (H == 0 && fan && level_in == g.depth) fan_store_root(fan, fan_step, fan_row_begin + row, d);  // depth 0
        bool parked = false;
#pragma unroll
        for (int l = 0; l < H; l++) {
            if (!parked) {
                if ((i >> l) & 1u) {
                    b3::Digest lft, rgt;
#pragma unroll
                    for (int w = 0; w < 8; w++) {
                        lft.w[w] = stack[l][w];
                        rgt.w[w] = d[w];
                    }
                    const b3::Digest o = b3::hash_node_call(lft, rgt, one);
#pragma unroll
                    for (int w = 0; w < 8; w++) d[w] = o.w[w];
                    store_digest(node_ptr(layers, roots, g, row, level_in + l + 1, idx >> (l + 1)), d);
                    if (fan && level_in + l + 1 == g.depth) fan_store_root(fan, fan_step, fan_row_begin + row, d);
                } else {
#pragma unroll
                    for (int w = 0; w < 8; w++) stack[l][w] = d[w];
                    parked = true;
                }
            }
        }
    }
    if (fan) fan_finish(fan, fan_step);  // every thread of every CTA (inactive ones included) gets here
}

// Latency path for small jobs (a prover's 2^12..2^16 commits, the reference's own criterion shapes): ONE CTA walks a
// whole subtree of up to 2^10 inputs level by level through shared memory, one compression of latency per level,
// instead of a chain of subtree passes whose launches and per-thread sequential subtrees dominate when the GPU is
// not full.  Digests live transposed in shared memory ([word][node]) so that the (2t, 2t+1) reads are conflict-free.
constexpr int CTA_TREE_MAX_LEVELS = 10;
constexpr int CTA_TREE_MAX_ROWS = 32;  // narrow subtrees: up to 32 rows share a CTA so that the warps stay full
constexpr int CTA_TREE_BUF = (1 << CTA_TREE_MAX_LEVELS) + 2 * CTA_TREE_MAX_ROWS;
// RL > 0: the CTA takes 2^RL whole rows (the subtree is everything that is left of a row); every row owns a region of
// 2^S + 2 words per digest word (the +2 keeps pair reads 8-byte aligned and spreads the rows over the banks).
template <int LEAF32>
__global__ void __launch_bounds__(512)
    merkle_cta_tree_kernel(const uint32_t *__restrict__ leaves, uint8_t *layers, uint8_t *roots, uint32_t num_rows, TreeGeom g,
                           uint32_t level_in, uint32_t S, uint32_t RL, uint32_t one, const RootsFanout *__restrict__ fan,
                           unsigned long long fan_step, uint32_t fan_row_begin) {
    __shared__ __align__(8) uint32_t buf[8][CTA_TREE_BUF];
    pdl_wait();
    const uint32_t subtrees_per_row = (g.cw >> level_in) >> S;
    const uint32_t row0 = RL ? blockIdx.x << RL : blockIdx.x / subtrees_per_row;
    const uint32_t b = RL ? 0u : blockIdx.x % subtrees_per_row;
    const uint32_t t = threadIdx.x, width = 1u << S, stride = width + 2;
    for (uint32_t e = t; e < (width << RL); e += blockDim.x) {
        const uint32_t rl = e >> S, i = e & (width - 1), row = row0 + rl;
        if (row >= num_rows) continue;
        const uint32_t idx = (b << S) | i;
        uint32_t d[8];
        if constexpr (LEAF32 > 0) {
            uint32_t x[LEAF32];
            load_leaf<LEAF32>(leaves + ((size_t)row * g.cw + idx) * LEAF32, x);
            b3::hash_leaf<LEAF32>(x, d, one);
            store_digest(node_ptr(layers, roots, g, row, 0, idx), d);
        } else {
            load_digest(node_ptr(layers, roots, g, row, level_in, idx), d);
        }
#pragma unroll
        for (int w = 0; w < 8; w++) buf[w][rl * stride + i] = d[w];
    }
    __syncthreads();
    for (uint32_t l = 1; l <= S; l++) {
        const uint32_t nlog = S - l, n = 1u << nlog;  // nodes per row at this level
        const uint32_t rl = t >> nlog, i = t & (n - 1), row = row0 + rl;
        const bool act = t < (n << RL) && row < num_rows;
        b3::Digest o;
        if (act) {
            b3::Digest lft, rgt;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint2 v = *reinterpret_cast<const uint2 *>(&buf[w][rl * stride + 2 * i]);
                lft.w[w] = v.x;
                rgt.w[w] = v.y;
            }
            o = b3::hash_node_call(lft, rgt, one);
        }
        __syncthreads();
        if (act) {
#pragma unroll
            for (int w = 0; w < 8; w++) buf[w][rl * stride + i] = o.w[w];
            store_digest(node_ptr(layers, roots, g, row, level_in + l, (b << nlog) | i), o.w);
            if (fan && level_in + l == g.depth) fan_store_root(fan, fan_step, fan_row_begin + row, o.w);
        }
        __syncthreads();
    }
    if (fan) fan_finish(fan, fan_step);
}

template <int LEAF32>
cudaError_t launch_cta_tree(const MerkleArgs &a, const TreeGeom &g, uint32_t level_in, uint32_t S, bool fan = false) {
    const uint32_t subtrees_per_row = (g.cw >> level_in) >> S;
    uint32_t RL = 0;
    // whole rows: pack rows while the CTA stays within 2^10 inputs and 32 rows -- but only as long as two CTAs per SM
    // remain (a tiny job is faster spread over many SMs than packed into full warps: commit of 2^12, 14.5 vs 19.6 us)
    if (subtrees_per_row == 1)
        while (RL < 5 && (S + RL + 1) <= (uint32_t)CTA_TREE_MAX_LEVELS && (a.num_rows >> (RL + 1)) >= 2u * (uint32_t)a.num_sms) RL++;
    const size_t grid = RL ? ((size_t)a.num_rows + (1u << RL) - 1) >> RL : (size_t)a.num_rows * subtrees_per_row;
    if (grid == 0) return cudaSuccess;
    if (grid > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    const uint32_t half = (1u << (S + RL)) / 2, block = half < 32 ? 32 : half;
    static const bool pdl = getenv("ZIPGPU_NO_PDL") == nullptr;
    if (pdl)
        return launch_pdl(merkle_cta_tree_kernel<LEAF32>, dim3((uint32_t)grid), dim3(block), 0, a.stream, a.leaves, a.layers, a.roots,
                          a.num_rows, g, level_in, S, RL, 1u, fan ? a.fan : nullptr, a.fan_step, a.fan_row_begin);
    merkle_cta_tree_kernel<LEAF32><<<(uint32_t)grid, block, 0, a.stream>>>(a.leaves, a.layers, a.roots, a.num_rows, g, level_in,
                                                                     S, RL, 1u, fan ? a.fan : nullptr, a.fan_step,
                                                                     a.fan_row_begin);
    return cudaGetLastError();
}

// MINB = 5: ptxas then keeps the whole working set in 92 registers without spills; measured on B200 the leaf pass
// runs 2 % faster at 5 resident blocks/SM than squeezed into 80 registers for 6 (1.837 vs 1.880 ms at nv = 24)
template <int LEAF32, int H, int MINB = 5>
cudaError_t launch_pass(const MerkleArgs &a, const TreeGeom &g, uint32_t level_in, bool fan = false) {
    const size_t threads = (size_t)a.num_rows * ((g.cw >> level_in) >> H);
    const uint32_t block = 128;
    const size_t grid = (threads + block - 1) / block;
    if (grid == 0) return cudaSuccess;
    if (grid > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    static const bool pdl = getenv("ZIPGPU_NO_PDL") == nullptr;
    if (pdl)
        return launch_pdl(merkle_subtree_kernel<LEAF32, H, MINB>, dim3((uint32_t)grid), dim3(block), 0, a.stream, a.leaves, a.layers,
                          a.roots, a.num_rows, g, level_in, 1u, fan ? a.fan : nullptr, a.fan_step, a.fan_row_begin);
    merkle_subtree_kernel<LEAF32, H, MINB><<<(uint32_t)grid, block, 0, a.stream>>>(a.leaves, a.layers, a.roots, a.num_rows,
                                                                           g, level_in, 1u, fan ? a.fan : nullptr,
                                                                           a.fan_step, a.fan_row_begin);
    return cudaGetLastError();
}

template <int LEAF32>
cudaError_t launch_leaf_pass(const MerkleArgs &a, const TreeGeom &g, int h, bool fan) {
    switch (h) {
        case 0: return launch_pass<LEAF32, 0>(a, g, 0, fan);
        case 1: return launch_pass<LEAF32, 1>(a, g, 0, fan);
        case 2: return launch_pass<LEAF32, 2>(a, g, 0, fan);
        default: return launch_pass<LEAF32, 3>(a, g, 0, fan);
    }
}

cudaError_t launch_node_pass(const MerkleArgs &a, const TreeGeom &g, uint32_t level_in, int h, bool fan) {
    switch (h) {
        case 1: return launch_pass<0, 1>(a, g, level_in, fan);
        case 2: return launch_pass<0, 2>(a, g, level_in, fan);
        case 3: return launch_pass<0, 3>(a, g, level_in, fan);
        default: return launch_pass<0, 4>(a, g, level_in, fan);
    }
}

}  // namespace

bool merkle_supported(int leaf32) {
    return leaf32 == 2 || leaf32 == 4 || leaf32 == 6 || leaf32 == 8 || leaf32 == 16;
}

cudaError_t launch_merkle_levels(const MerkleArgs &a, int from_level, int until_level, int *reached, int *launches) {
    TreeGeom g;
    g.depth = (uint32_t)a.depth;
    g.cw = 1u << a.depth;
    g.row_stride = 2 * (size_t)g.cw - 2;
    if (until_level < 0 || until_level > a.depth) until_level = a.depth;
    int n = 0, level = from_level;
    cudaError_t err = cudaSuccess;
    // the launch that reaches the roots carries the multi-GPU roots exchange (RootsFanout), if one was asked for
    // nodes (over all rows) from which the remaining levels go to the CTA-per-subtree kernel (ZIPGPU_CTA_TREE_LOG2: knob)
    static const int cta_tree_log2 = getenv("ZIPGPU_CTA_TREE_LOG2") ? atoi(getenv("ZIPGPU_CTA_TREE_LOG2")) : 18;
    const size_t cta_tree_max = (size_t)1 << cta_tree_log2;
    const bool want_fan = a.fan != nullptr && until_level == a.depth && a.num_rows > 0;
    bool fused = false;
    if (a.fan_fused) *a.fan_fused = false;
    // small whole trees: the CTA-per-subtree latency path (at most 2^18 leaves in total, i.e. <= 256 CTAs of work)
    if (from_level == 0 && until_level == a.depth && a.depth >= 1 && ((size_t)a.num_rows << a.depth) <= cta_tree_max &&
        !getenv("ZIPGPU_NO_CTA_TREE")) {
        while (level < a.depth) {
            const uint32_t S = (uint32_t)std::min(CTA_TREE_MAX_LEVELS, a.depth - level);
            const bool fan = want_fan && level + (int)S == a.depth;
            fused |= fan;
            if (level == 0) {
                switch (a.leaf32) {
                    case 2: err = launch_cta_tree<2>(a, g, 0, S, fan); break;
                    case 4: err = launch_cta_tree<4>(a, g, 0, S, fan); break;
                    case 6: err = launch_cta_tree<6>(a, g, 0, S, fan); break;
                    case 8: err = launch_cta_tree<8>(a, g, 0, S, fan); break;
                    case 16: err = launch_cta_tree<16>(a, g, 0, S, fan); break;
                    default: return cudaErrorInvalidValue;
                }
            } else {
                err = launch_cta_tree<0>(a, g, (uint32_t)level, S, fan);
            }
            if (err != cudaSuccess) return err;
            n++;
            level += (int)S;
        }
        if (reached) *reached = level;
        if (launches) *launches = n;
        if (a.fan_fused) *a.fan_fused = fused;
        return cudaSuccess;
    }
    if (level == 0 && (until_level > 0 || a.depth == 0)) {
        // the leaf pass covers levels 0..min(3, depth) (15/16 of all compressions)
        const int h = a.depth < 3 ? a.depth : 3;
        const bool fan = want_fan && h == a.depth;
        fused |= fan;
        switch (a.leaf32) {
            case 2: err = launch_leaf_pass<2>(a, g, h, fan); break;
            case 4: err = launch_leaf_pass<4>(a, g, h, fan); break;
            case 6: err = launch_leaf_pass<6>(a, g, h, fan); break;
            case 8: err = launch_leaf_pass<8>(a, g, h, fan); break;
            case 16: err = launch_leaf_pass<16>(a, g, h, fan); break;
            default: return cudaErrorInvalidValue;
        }
        if (err != cudaSuccess) return err;
        n++;
        level = h;
    }
    const bool cta_top_ok = until_level == a.depth && !getenv("ZIPGPU_NO_CTA_TREE");
    while (level < until_level) {  // every further pass 3 levels, the last one up to 4
        // once the trees have narrowed to <= 2^18 nodes in total the remaining passes are latency-bound: finish with
        // the CTA-per-subtree kernel (one launch per 10 levels, one compression of latency per level)
        if (cta_top_ok && level > 0 && ((size_t)a.num_rows << (a.depth - level)) <= cta_tree_max) {
            const uint32_t S = (uint32_t)std::min(CTA_TREE_MAX_LEVELS, a.depth - level);
            const bool fan = want_fan && level + (int)S == a.depth;
            fused |= fan;
            err = launch_cta_tree<0>(a, g, (uint32_t)level, S, fan);
            if (err != cudaSuccess) return err;
            n++;
            level += (int)S;
            continue;
        }
        const int remaining = a.depth - level;
        const int h = remaining <= 4 ? remaining : 3;
        const bool fan = want_fan && level + h == a.depth;
        fused |= fan;
        err = launch_node_pass(a, g, (uint32_t)level, h, fan);
        if (err != cudaSuccess) return err;
        n++;
        level += h;
    }
    if (reached) *reached = level;
    if (launches) *launches = n;
    if (a.fan_fused) *a.fan_fused = fused;
    return cudaSuccess;
}

cudaError_t launch_merkle_rows(const MerkleArgs &a, int *launches) {
    return launch_merkle_levels(a, 0, -1, nullptr, launches);
}

}  // namespace zipgpu
