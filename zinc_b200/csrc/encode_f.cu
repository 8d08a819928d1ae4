// zinc_b200/csrc/encode_f.cu -- RaaCode::encode_f: the RAA code over field elements (next item f-3).
//
// code_raa.rs:133-138 runs the same encode_inner as the integer code -- repeat -> shuffle_seeded(perm_1_seed) ->
// accumulate -> shuffle_seeded(perm_2_seed) -> accumulate (code_raa.rs:89-105) -- with Out = F, so every `+=` is the
// field's addition: add the stored residues, subtract the modulus once if the sum overflowed or is >= modulus
// (RandomField AddAssign -> FieldConfig::add_assign / reduce_modulus, field/arithmetic.rs:66-77, field/config.rs:53-76).
// The verifier calls it on ONE combined row per opening (verify_z.rs:141-142), so this is a latency kernel: one CTA per
// row, every thread owns a contiguous segment, serial modular running sums inside the segment, a block scan of the
// segment totals (modular addition is associative for residues < modulus), intermediate vector in global scratch.
//
// The same kernel with MOD = false is RaaCode::encode_wide for any In = Int<in>, Out = Int<out> (code_raa.rs:125-131):
// entries sign-extended on load (`Out::from(&In)`, int.rs:194-199), wrap-around adds at the output width (the
// reference's checked adds never fire for values the protocol produces).  The verifier re-encodes the combined row of
// every proximity test with In = Out = M = Int<8> (verify_z.rs:74-78) -- a width the tiled encoder of raa_encode.cu,
// built for the prover's Int<1>/Int<2> evaluations, does not carry.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace zipgpu {

namespace {

constexpr int kEfT = 512;

template <int NW>
struct Fe {
    uint32_t w[NW];
};

// field/config.rs:53-76: s = a + b (wrapping, carry c); if (c || s >= p) s -= p (wrapping)
template <int NW>
__device__ __forceinline__ Fe<NW> add_mod(const Fe<NW> &a, const Fe<NW> &b, const Fe<NW> &p) {
    Fe<NW> s, d;
    uint32_t carry = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        const unsigned long long t = (unsigned long long)a.w[i] + b.w[i] + carry;
        s.w[i] = (uint32_t)t;
        carry = (uint32_t)(t >> 32);
    }
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        const unsigned long long t = (unsigned long long)s.w[i] - p.w[i] - borrow;
        d.w[i] = (uint32_t)t;
        borrow = (uint32_t)(t >> 63);
    }
    const bool reduce = carry || !borrow;  // !borrow: s >= p
#pragma unroll
    for (int i = 0; i < NW; i++) s.w[i] = reduce ? d.w[i] : s.w[i];
    return s;
}

// MOD: the field's addition; otherwise two's-complement wrap-around at NW words
template <int NW, bool MOD>
__device__ __forceinline__ Fe<NW> add_fe(const Fe<NW> &a, const Fe<NW> &b, const Fe<NW> &p) {
    if constexpr (MOD) {
        return add_mod<NW>(a, b, p);
    } else {
        Fe<NW> s;
        uint32_t carry = 0;
#pragma unroll
        for (int i = 0; i < NW; i++) {
            const unsigned long long t = (unsigned long long)a.w[i] + b.w[i] + carry;
            s.w[i] = (uint32_t)t;
            carry = (uint32_t)(t >> 32);
        }
        return s;
    }
}

// an input entry of `inw` <= NW words, sign-extended (MOD: inw == NW)
template <int NW>
__device__ __forceinline__ Fe<NW> load_in(const uint32_t *p, uint32_t inw) {
    Fe<NW> x;
    const uint32_t fill = (uint32_t)((int32_t)p[inw - 1] >> 31);
#pragma unroll
    for (int i = 0; i < NW; i++) x.w[i] = (uint32_t)i < inw ? p[i] : fill;
    return x;
}

template <int NW>
__device__ __forceinline__ Fe<NW> load_fe(const uint32_t *p) {
    Fe<NW> x;
#pragma unroll
    for (int i = 0; i < NW; i += 2) {
        const uint2 v = *reinterpret_cast<const uint2 *>(p + i);
        x.w[i] = v.x;
        x.w[i + 1] = v.y;
    }
    return x;
}
template <int NW>
__device__ __forceinline__ void store_fe(uint32_t *p, const Fe<NW> &x) {
#pragma unroll
    for (int i = 0; i < NW; i += 2) *reinterpret_cast<uint2 *>(p + i) = make_uint2(x.w[i], x.w[i + 1]);
}

// exclusive modular scan of one value per thread over the CTA (Hillis-Steele in shared memory); returns the prefix
template <int NW, bool MOD>
__device__ __forceinline__ Fe<NW> block_exclusive_scan(const Fe<NW> &mine, const Fe<NW> &p, uint32_t *sh, uint32_t t) {
    uint32_t *cur = sh, *nxt = sh + kEfT * NW;
    store_fe<NW>(cur + t * NW, mine);
    __syncthreads();
    for (uint32_t off = 1; off < (uint32_t)kEfT; off <<= 1) {
        Fe<NW> v = load_fe<NW>(cur + t * NW);
        if (t >= off) v = add_fe<NW, MOD>(load_fe<NW>(cur + (t - off) * NW), v, p);
        store_fe<NW>(nxt + t * NW, v);
        __syncthreads();
        uint32_t *tmp = cur;
        cur = nxt;
        nxt = tmp;
    }
    Fe<NW> ex;
#pragma unroll
    for (int i = 0; i < NW; i++) ex.w[i] = 0u;
    if (t > 0) ex = load_fe<NW>(cur + (t - 1) * NW);
    __syncthreads();
    return ex;
}

template <int NW, bool MOD>
__global__ void __launch_bounds__(kEfT)
    encode_f_kernel(const uint32_t *__restrict__ rows_in, uint32_t *__restrict__ out, const uint32_t *__restrict__ perm1,
                    const uint32_t *__restrict__ perm2, const uint32_t *__restrict__ modulus, uint32_t *__restrict__ scratch,
                    uint32_t row_len, uint32_t cw, uint32_t inw) {
    extern __shared__ __align__(8) uint32_t sh[];
    const uint32_t t = threadIdx.x;
    const size_t row = blockIdx.x;
    const uint32_t seg = (cw + kEfT - 1) / kEfT, i0 = min(t * seg, cw), i1 = min(i0 + seg, cw);
    Fe<NW> p;
    if constexpr (MOD) {
        p = load_fe<NW>(modulus);
    } else {
#pragma unroll
        for (int i = 0; i < NW; i++) p.w[i] = 0u;
    }
    const uint32_t *in = rows_in + row * row_len * inw;
    uint32_t *s1 = scratch + row * (size_t)cw * NW, *o = out + row * (size_t)cw * NW;
    Fe<NW> acc;
    // ---- repeat o perm1, accumulate ----
#pragma unroll
    for (int i = 0; i < NW; i++) acc.w[i] = 0u;
    for (uint32_t i = i0; i < i1; i++) {
        acc = add_fe<NW, MOD>(acc, load_in<NW>(in + (size_t)(__ldg(perm1 + i) % row_len) * inw, inw), p);
        store_fe<NW>(s1 + (size_t)i * NW, acc);
    }
    Fe<NW> pre = block_exclusive_scan<NW, MOD>(acc, p, sh, t);
    for (uint32_t i = i0; i < i1; i++)
        store_fe<NW>(s1 + (size_t)i * NW, add_fe<NW, MOD>(pre, load_fe<NW>(s1 + (size_t)i * NW), p));
    __threadfence_block();
    __syncthreads();
    // ---- perm2, accumulate ----
#pragma unroll
    for (int i = 0; i < NW; i++) acc.w[i] = 0u;
    for (uint32_t i = i0; i < i1; i++) {
        acc = add_fe<NW, MOD>(acc, load_fe<NW>(s1 + (size_t)__ldg(perm2 + i) * NW), p);
        store_fe<NW>(o + (size_t)i * NW, acc);
    }
    pre = block_exclusive_scan<NW, MOD>(acc, p, sh, t);
    for (uint32_t i = i0; i < i1; i++)
        store_fe<NW>(o + (size_t)i * NW, add_fe<NW, MOD>(pre, load_fe<NW>(o + (size_t)i * NW), p));
}

template <int NW, bool MOD>
cudaError_t launch_ef(const EncodeFArgs &a, uint32_t inw) {
    const size_t smem = 2 * (size_t)kEfT * NW * sizeof(uint32_t);
    auto kern = encode_f_kernel<NW, MOD>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<a.num_rows, kEfT, smem, a.stream>>>(a.rows_in, a.out, a.perm1, a.perm2, a.modulus, a.scratch, a.row_len, a.cw, inw);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_encode_f(const EncodeFArgs &a) {
    if (a.num_rows == 0) return cudaSuccess;
    const int NW = 2 * a.limbs;
    if (a.in_limbs > 0) {  // encode_wide: Int<in_limbs> -> Int<limbs>, wrap-around adds
        const uint32_t inw = 2u * (uint32_t)a.in_limbs;
        switch (NW) {
            case 2: return launch_ef<2, false>(a, inw);
            case 4: return launch_ef<4, false>(a, inw);
            case 6: return launch_ef<6, false>(a, inw);
            case 8: return launch_ef<8, false>(a, inw);
            case 10: return launch_ef<10, false>(a, inw);
            case 12: return launch_ef<12, false>(a, inw);
            case 14: return launch_ef<14, false>(a, inw);
            case 16: return launch_ef<16, false>(a, inw);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (NW) {
        case 2: return launch_ef<2, true>(a, 2);
        case 4: return launch_ef<4, true>(a, 4);
        case 6: return launch_ef<6, true>(a, 6);
        case 8: return launch_ef<8, true>(a, 8);
        case 10: return launch_ef<10, true>(a, 10);
        case 12: return launch_ef<12, true>(a, 12);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace zipgpu
