// scratch/pipes.cu -- which INT32 ops go to which pipe and at what rate (B200)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define N 8
template <int KIND>
__global__ void __launch_bounds__(256) k(uint32_t *sink, int iters, uint32_t seed) {
    uint32_t a[N], b[N];
    for (int i = 0; i < N; i++) { a[i] = seed + threadIdx.x * N + i; b[i] = a[i] * 3 + 1; }
    uint32_t one = (seed >> 31) + 1;  // == 1 at run time, unknown at compile time
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int i = 0; i < N; i++) {
                if (KIND == 0) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); }
                if (KIND == 1) { asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); }
                if (KIND == 2) { asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i])); }
                if (KIND == 3) { asm volatile("prmt.b32 %0, %0, %0, 0x0321;" : "+r"(a[i])); }
                if (KIND == 4) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(one), "r"(b[i])); }
                if (KIND == 5) { asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(a[i]) : "r"(b[i])); }
                if (KIND == 6) { asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(one)); }
                if (KIND == 7) {
                    uint32_t &A = a[i], &B = b[i], &C = a[(i + 1) % N], &D = b[(i + 1) % N];
                    A = A + B + one; D = __funnelshift_r(D ^ A, D ^ A, 16); C = C + D; B = __funnelshift_r(B ^ C, B ^ C, 12);
                }
                if (KIND == 8) {
                    uint32_t &A = a[i], &B = b[i], &C = a[(i + 1) % N], &D = b[(i + 1) % N];
                    A = A + B + one; D = __funnelshift_r(D ^ A, D ^ A, 16);
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(C) : "r"(D), "r"(one));
                    B = __funnelshift_r(B ^ C, B ^ C, 12);
                }
                if (KIND == 15) {  // half G, all three adds as IMAD: 4 alu + 3 fma
                    uint32_t &A = a[i], &B = b[i], &C = a[(i + 1) % N], &D = b[(i + 1) % N];
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(A) : "r"(one), "r"(B));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(A) : "r"(one), "r"(seed));
                    D = __funnelshift_r(D ^ A, D ^ A, 16);
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(C) : "r"(D), "r"(one));
                    B = __funnelshift_r(B ^ C, B ^ C, 12);
                }
                if (KIND == 16) {  // half G: a+b on IMAD, +m folded as IADD (alu): 5 alu + 2 fma
                    uint32_t &A = a[i], &B = b[i], &C = a[(i + 1) % N], &D = b[(i + 1) % N];
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(A) : "r"(one), "r"(B));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(A) : "r"(seed));
                    D = __funnelshift_r(D ^ A, D ^ A, 16);
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(C) : "r"(D), "r"(one));
                    B = __funnelshift_r(B ^ C, B ^ C, 12);
                }
                if (KIND == 17) {  // 4 alu + 3 fma, all INDEPENDENT chains (no cross dependencies)
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(seed));
                    asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i]));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(one));
                    asm volatile("shf.r.wrap.b32 %0, %0, %0, 12;" : "+r"(a[i]));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(one), "r"(seed));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(one), "r"(seed));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(one), "r"(seed));
                }
                if (KIND == 9) {
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                    asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i]));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(one));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(one), "r"(a[i]));
                }
                if (KIND == 10) {
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                    asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i]));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(one), "r"(a[i]));
                }
                if (KIND == 11) {
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(one), "r"(a[i]));
                }
                if (KIND == 18) {  // IMAD.WIDE.U32 imm
                    asm volatile("{ .reg .b64 t; mul.wide.u32 t, %0, 0x2000000; mov.b64 {%0,%1}, t; }" : "+r"(a[i]), "=r"(b[i]));
                }
                if (KIND == 19) {  // IMAD.WIDE.U32 reg
                    asm volatile("{ .reg .b64 t; mul.wide.u32 t, %0, %2; mov.b64 {%0,%1}, t; }" : "+r"(a[i]), "=r"(b[i]) : "r"(one));
                }
                if (KIND == 20) {  // IMAD.HI.U32 reg
                    asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(one), "r"(b[i]));
                }
                if (KIND == 21) {  // LEA.HI (shift+add on alu)
                    asm volatile("{ .reg .b32 t; shf.l.wrap.b32 t, %0, %0, 25; add.u32 %0, t, %1; }" : "+r"(a[i]) : "r"(b[i]));
                }
                if (KIND == 22) {  // 1 alu : 2 imad-imm
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                    asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(b[i]) : "r"(a[i]));
                    asm volatile("mad.lo.u32 %0, %0, 5, %1;" : "+r"(b[i]) : "r"(one));
                }
                if (KIND == 23) {  // imad-imm only
                    asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(a[i]) : "r"(b[i]));
                }
                if (KIND == 12) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(one)); }
                if (KIND == 13) { asm volatile("vadd.u32.u32.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); }
                if (KIND == 14) {
                    unsigned long long x = ((unsigned long long)b[i] << 32) | a[i];
                    asm volatile("add.u64 %0, %0, %1;" : "+l"(x) : "l"((unsigned long long)one));
                    a[i] = (uint32_t)x; b[i] = (uint32_t)(x >> 32);
                }
            }
        }
    }
    uint32_t s = 0;
    for (int i = 0; i < N; i++) s ^= a[i] ^ b[i];
    if (s == 0x12345u) sink[0] = s;
}

static const double OPS[] = {1, 1, 1, 1, 1, 1, 2, 6, 6, 4, 3, 2, 1, 1, 2, 7, 7, 7, 1, 1, 1, 1, 3, 1};
static const char *NAME[] = {"add.u32", "xor (LOP3)", "shf", "prmt", "mad.lo reg-mult", "mad.lo x1", "add3 (2 adds)",
                             "halfG plain (6 ops)", "halfG c+d via mad (6 ops)", "3alu:1imad", "2alu:1imad", "1alu:1imad",
                             "lop3 3-in", "vadd", "add.u64 (2 ops)", "halfG all-IMAD (4alu+3fma)", "halfG a+b IMAD,+m IADD (5+2)", "4alu+3fma independent", "IMAD.WIDE imm", "IMAD.WIDE reg", "IMAD.HI reg", "shf+add (LEA.HI?)", "1alu:2imad-imm", "imad-imm x3"};

template <int KIND> void run(uint32_t *sink, int sms) {
    const int iters = 400, grid = sms * 8;
    k<KIND><<<grid, 256>>>(sink, iters, 0x12345678u);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<KIND><<<grid, 256>>>(sink, iters, 0x12345678u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)iters * 16 * N * OPS[KIND] * 256.0 * grid;
    double per_clk_sm = ops / (ms * 1e-3) / (1.965e9 * sms);
    printf("%-28s %8.3f ms  %7.2f Tlane-op/s  %6.1f lane-ops/clk/SM (at 1965 MHz)\n", NAME[KIND], ms, ops / ms / 1e9, per_clk_sm);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint32_t *sink; cudaMalloc(&sink, 64);
    int sms = p.multiProcessorCount;
    run<0>(sink, sms); run<1>(sink, sms); run<2>(sink, sms); run<3>(sink, sms); run<4>(sink, sms); run<5>(sink, sms);
    run<6>(sink, sms); run<7>(sink, sms); run<8>(sink, sms); run<9>(sink, sms); run<10>(sink, sms); run<11>(sink, sms);
    run<12>(sink, sms); run<15>(sink, sms); run<16>(sink, sms); run<17>(sink, sms);
    run<18>(sink, sms); run<19>(sink, sms); run<20>(sink, sms); run<21>(sink, sms); run<22>(sink, sms); run<23>(sink, sms);
    return 0;
}
