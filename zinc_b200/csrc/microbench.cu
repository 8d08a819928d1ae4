// zinc_b200/csrc/microbench.cu -- dependency-light INT32 issue-rate micro-benchmark.
//
// the hasher's roofline is INT32 ALU issue, which has no datasheet figure that can be trusted
// under load; this measures it on the box.  kind 0: xor / funnel shift / add (ptxas issues the 2-input add as
// IMAD.IADD: 3 alu + 1 fma); kind 1: the same plus IMAD (the "fma" pipe) in the ratio a BLAKE3 G uses when its adds
// are issued as IMAD (8 alu : 6 fma); kind 2: LOP3 and SHF only -- nothing but the "alu" pipe, where every xor and
// rotate of BLAKE3 must go: the measured peak the hasher's roofline is quoted against.
#include "common.cuh"
#include "kernels.h"

namespace zipgpu {

template <int KIND>
__global__ void __launch_bounds__(256) int32_bench_kernel(uint32_t *sink, int iters, uint32_t seed) {
    uint32_t a[8], one = seed | 1u;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 8 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                // 4 alu ops per lane-iteration: xor, funnel shift, add3, xor
                uint32_t x = a[i] ^ a[(i + 1) & 7];
                x = __funnelshift_r(x, x, 7 + r);
                if (KIND == 2) {
                    // 4 alu-pipe ops, no other pipe: lop3, shf, lop3, shf
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a[(i + 3) & 7]), "r"(one));
                    x = __funnelshift_r(x, x, 12);
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a[(i + 5) & 7]), "r"(one));
                    a[i] = __funnelshift_r(x, x, 8);
                } else if (KIND == 0) {
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a[(i + 3) & 7]));
                    a[i] = x ^ one;
                } else {
                    // 3 IMAD per 4 alu (== 6 : 8)
                    uint32_t y;
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(y) : "r"(x), "r"(one), "r"(a[(i + 3) & 7]));
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(y) : "r"(y), "r"(one), "r"(a[(i + 5) & 7]));
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(y) : "r"(y), "r"(one), "r"(a[(i + 6) & 7]));
                    x = y ^ a[(i + 2) & 7];
                    x = __funnelshift_r(x, x, 12);
                    a[i] = x ^ one;
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i];
    if (s == 0x12345u) sink[0] = s;  // practically never true: keeps the chain alive without traffic
}

cudaError_t launch_microbench_int32(int kind, int iters, int num_sms, cudaStream_t stream, uint32_t *d_sink,
                                    double *lane_ops) {
    const int grid = num_sms * 8;
    const double per_thread = (double)iters * 64.0 * (kind == 0 ? 4.0 : kind == 2 ? 6.0 : 7.0);
    *lane_ops = per_thread * 256.0 * grid;
    if (kind == 0) int32_bench_kernel<0><<<grid, 256, 0, stream>>>(d_sink, iters, 0x9e3779b9u);
    else if (kind == 2) int32_bench_kernel<2><<<grid, 256, 0, stream>>>(d_sink, iters, 0x9e3779b9u);
    else int32_bench_kernel<1><<<grid, 256, 0, stream>>>(d_sink, iters, 0x9e3779b9u);
    return cudaGetLastError();
}

}  // namespace zipgpu
