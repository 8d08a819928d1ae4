// zinc_b200/csrc/combine_rows.cu -- K5: the proximity-test row combination u' = sum_i coeff_i * row_i.
//
// Replaces `combine_rows(coeffs.map(expand::<N, M>), evals.map(expand::<N, M>), row_len)` of the testing phase of
// `open` (zip/pcs/open_z.rs:100-113; zip/utils.rs:94-127,129-137): an integer mat-vec over the UNENCODED
// evaluation matrix with Fiat-Shamir integer challenges as coefficients, evaluated in M = Int<8>.  For N = Int<1>
// every product fits 128 bits and the sum over up to 2^32 rows fits 160, so the arithmetic is done exactly in a
// 192-bit two's-complement accumulator (three u64 with one carry chain) and sign-extended to out_limbs on the way out
// -- bit-identical to the reference's checked Int<8> arithmetic, which cannot overflow here.
//
// HBM-bound: 8 bytes read per evaluation, once.  Pass 1: thread = one column of a slice of rows (consecutive lanes
// read consecutive columns: 256 B per warp and row), coefficients broadcast from shared memory.  Pass 2: sums the
// per-slice partials and sign-extends.
#include "common.cuh"
#include "kernels.h"

namespace zipgpu {

namespace {

struct Acc192 {
    unsigned long long w0, w1, w2;
};

__device__ __forceinline__ void acc_mul_add(Acc192 &a, long long x, long long y) {
    const unsigned long long lo = (unsigned long long)x * (unsigned long long)y;
    const long long hi = __mul64hi(x, y);
    const unsigned long long ext = (unsigned long long)(hi >> 63);
    asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u64 %2, %2, %5;"
        : "+l"(a.w0), "+l"(a.w1), "+l"(a.w2)
        : "l"(lo), "l"((unsigned long long)hi), "l"(ext));
}
__device__ __forceinline__ void acc_add(Acc192 &a, const Acc192 &b) {
    asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u64 %2, %2, %5;"
        : "+l"(a.w0), "+l"(a.w1), "+l"(a.w2)
        : "l"(b.w0), "l"(b.w1), "l"(b.w2));
}

constexpr int kCols = 128;       // columns (threads) per block
constexpr int kSliceRows = 128;  // rows per block

__global__ void __launch_bounds__(kCols)
    combine_rows_partial_kernel(const long long *__restrict__ evals, const long long *__restrict__ coeffs,
                                unsigned long long *__restrict__ partial, uint32_t num_rows, uint32_t row_len) {
    __shared__ long long c[kSliceRows];
    const uint32_t r0 = blockIdx.y * kSliceRows;
    const uint32_t nr = min((uint32_t)kSliceRows, num_rows - r0);
    for (uint32_t i = threadIdx.x; i < nr; i += blockDim.x) c[i] = coeffs[r0 + i];
    __syncthreads();
    const uint32_t col = blockIdx.x * kCols + threadIdx.x;
    if (col >= row_len) return;
    Acc192 a{0, 0, 0};
    const long long *p = evals + (size_t)r0 * row_len + col;
#pragma unroll 16  // 16 independent 8-byte loads in flight per thread
    for (uint32_t i = 0; i < nr; i++) acc_mul_add(a, c[i], __ldg(p + (size_t)i * row_len));
    unsigned long long *o = partial + ((size_t)blockIdx.y * row_len + col) * 3;
    o[0] = a.w0;
    o[1] = a.w1;
    o[2] = a.w2;
}

__global__ void __launch_bounds__(32)
    combine_rows_final_kernel(const unsigned long long *__restrict__ partial, unsigned long long *__restrict__ out,
                              uint32_t slices, uint32_t row_len, uint32_t out_limbs) {
    pdl_wait();  // launched with programmatic stream serialisation behind the partial sums
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= row_len) return;
    Acc192 a{0, 0, 0};
#pragma unroll 8  // (not unrolled, the 3 x slices dependent-looking loads of a thread went out one slice at a time: ~20 us)
    for (uint32_t s = 0; s < slices; s++) {
        const unsigned long long *p = partial + ((size_t)s * row_len + col) * 3;
        Acc192 b{p[0], p[1], p[2]};
        acc_add(a, b);
    }
    const unsigned long long sign = (unsigned long long)((long long)a.w2 >> 63);
    unsigned long long *o = out + (size_t)col * out_limbs;
    const unsigned long long w[3] = {a.w0, a.w1, a.w2};
    for (uint32_t l = 0; l < out_limbs; l++) o[l] = l < 3 ? w[l] : sign;
}

}  // namespace

size_t combine_rows_scratch_bytes(uint32_t num_rows, uint32_t row_len) {
    const size_t slices = (num_rows + kSliceRows - 1) / kSliceRows;
    return slices * row_len * 3 * sizeof(unsigned long long);
}

cudaError_t launch_combine_rows(const CombineArgs &a, int *launches) {
    if (a.num_rows == 0 || a.row_len == 0) return cudaSuccess;
    const uint32_t slices = (a.num_rows + kSliceRows - 1) / kSliceRows;
    dim3 grid((a.row_len + kCols - 1) / kCols, slices);
    combine_rows_partial_kernel<<<grid, kCols, 0, a.stream>>>(reinterpret_cast<const long long *>(a.evals),
                                                               reinterpret_cast<const long long *>(a.coeffs),
                                                               reinterpret_cast<unsigned long long *>(a.scratch),
                                                               a.num_rows, a.row_len);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (launches) *launches = 2;
    return launch_pdl(combine_rows_final_kernel, dim3((a.row_len + 31) / 32), dim3(32), 0, a.stream,
                      reinterpret_cast<const unsigned long long *>(a.scratch), reinterpret_cast<unsigned long long *>(a.out),
                      slices, a.row_len, a.out_limbs);
}

}  // namespace zipgpu
