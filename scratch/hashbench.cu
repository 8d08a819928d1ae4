// scratch/hashbench.cu -- BLAKE3 single-compression throughput for different instruction-selection variants of G
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

struct Schedule { unsigned char s[7][16]; };
constexpr Schedule make_schedule() {
    constexpr unsigned char perm[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
    Schedule r{};
    for (int i = 0; i < 16; i++) r.s[0][i] = (unsigned char)i;
    for (int k = 1; k < 7; k++) for (int i = 0; i < 16; i++) r.s[k][i] = r.s[k - 1][perm[i]];
    return r;
}
__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ uint32_t rot16p(uint32_t x) { return __byte_perm(x, 0, 0x1032); }
__device__ __forceinline__ uint32_t rot8p(uint32_t x) { return __byte_perm(x, 0, 0x0321); }
__device__ __forceinline__ uint32_t madd(uint32_t a, uint32_t b, uint32_t one) {
    uint32_t r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b)); return r;
}

template <int N>
__device__ __forceinline__ uint32_t rotw(uint32_t x) {  // rotate right by N on the fma pipe: IMAD.WIDE + add
    uint32_t lo, hi;
    asm("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0,%1}, t; }" : "=r"(lo), "=r"(hi) : "r"(x), "n"(1u << (32 - N)));
    return lo + hi;
}
template <int N>
__device__ __forceinline__ uint32_t rotwm(uint32_t x, uint32_t one) {  // same, add forced to IMAD
    uint32_t lo, hi;
    asm("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0,%1}, t; }" : "=r"(lo), "=r"(hi) : "r"(x), "n"(1u << (32 - N)));
    return madd(lo, hi, one);
}
__device__ int g_ctr;
template <int V>
__device__ __forceinline__ void G(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d, uint32_t x, uint32_t y, uint32_t one) {
    if (V == 0) {
        a = a + b + x; d = rotr(d ^ a, 16); c = c + d; b = rotr(b ^ c, 12);
        a = a + b + y; d = rotr(d ^ a, 8); c = c + d; b = rotr(b ^ c, 7);
    } else if (V == 1) {
        a = madd(madd(a, b, one), x, one); d = rotr(d ^ a, 16); c = madd(c, d, one); b = rotr(b ^ c, 12);
        a = madd(madd(a, b, one), y, one); d = rotr(d ^ a, 8); c = madd(c, d, one); b = rotr(b ^ c, 7);
    } else if (V == 2) {
        a = madd(a, b, one) + x; d = rotr(d ^ a, 16); c = c + d; b = rotr(b ^ c, 12);
        a = madd(a, b, one) + y; d = rotr(d ^ a, 8); c = c + d; b = rotr(b ^ c, 7);
    } else if (V == 3) {
        a = madd(a, b + x, one); d = rotr(d ^ a, 16); c = c + d; b = rotr(b ^ c, 12);
        a = madd(a, b + y, one); d = rotr(d ^ a, 8); c = c + d; b = rotr(b ^ c, 7);
    } else if (V == 4) {
        a = a + b + x; d = rotr(d ^ a, 16); c = madd(c, d, one); b = rotr(b ^ c, 12);
        a = a + b + y; d = rotr(d ^ a, 8); c = madd(c, d, one); b = rotr(b ^ c, 7);
    } else if (V == 5) {
        a = a + b + x; d = rot16p(d ^ a); c = c + d; b = rotr(b ^ c, 12);
        a = a + b + y; d = rot8p(d ^ a); c = c + d; b = rotr(b ^ c, 7);
    } else if (V == 6) {  // first 3-input add split on fma, second stays IADD3
        a = madd(madd(a, b, one), x, one); d = rotr(d ^ a, 16); c = c + d; b = rotr(b ^ c, 12);
        a = a + b + y; d = rotr(d ^ a, 8); c = c + d; b = rotr(b ^ c, 7);
    } else if (V == 8) {  // V2 + rot7 on the fma pipe
        a = madd(a, b, one) + x; d = rotr(d ^ a, 16); c = c + d; b = rotr(b ^ c, 12);
        a = madd(a, b, one) + y; d = rotr(d ^ a, 8); c = c + d; b = rotw<7>(b ^ c);
    } else if (V == 9) {  // V2 + rot7 and rot12 on the fma pipe
        a = madd(a, b, one) + x; d = rotr(d ^ a, 16); c = c + d; b = rotw<12>(b ^ c);
        a = madd(a, b, one) + y; d = rotr(d ^ a, 8); c = c + d; b = rotw<7>(b ^ c);
    } else if (V == 10) {  // plain adds + rot7 on fma
        a = a + b + x; d = rotr(d ^ a, 16); c = c + d; b = rotr(b ^ c, 12);
        a = a + b + y; d = rotr(d ^ a, 8); c = c + d; b = rotw<7>(b ^ c);
    } else if (V == 11) {  // plain adds + rot7, rot12 on fma
        a = a + b + x; d = rotr(d ^ a, 16); c = c + d; b = rotw<12>(b ^ c);
        a = a + b + y; d = rotr(d ^ a, 8); c = c + d; b = rotw<7>(b ^ c);
    } else if (V == 12) {  // V2 + rot7 on fma with forced IMAD add
        a = madd(a, b, one) + x; d = rotr(d ^ a, 16); c = c + d; b = rotr(b ^ c, 12);
        a = madd(a, b, one) + y; d = rotr(d ^ a, 8); c = c + d; b = rotwm<7>(b ^ c, one);
    } else if (V == 13) {  // plain adds + all four rotations on fma
        a = a + b + x; d = rotw<16>(d ^ a); c = c + d; b = rotw<12>(b ^ c);
        a = a + b + y; d = rotw<8>(d ^ a); c = c + d; b = rotw<7>(b ^ c);
    } else if (V == 7) {  // a+b on fma, then +x via IADD (alu) -- both halves; c+d compiler's choice
        a = madd(a, b, one); a = a + x; d = rotr(d ^ a, 16); c = c + d; b = rotr(b ^ c, 12);
        a = a + b + y; d = rotr(d ^ a, 8); c = c + d; b = rotr(b ^ c, 7);
    }
}

template <int V, int POS> struct Pick { static constexpr int v = V; };
// V14: rot7 on fma in column Gs only (r = 0.5); V15: in 5 of 8 Gs; V16: 6 of 8
template <int POS> struct Pick<14, POS> { static constexpr int v = (POS < 4) ? 8 : 2; };
template <int POS> struct Pick<15, POS> { static constexpr int v = (POS < 5) ? 8 : 2; };
template <int POS> struct Pick<16, POS> { static constexpr int v = (POS < 6) ? 8 : 2; };
template <int POS> struct Pick<17, POS> { static constexpr int v = (POS % 2 == 0) ? 9 : 2; };
template <int V>
__device__ __forceinline__ void compress(const uint32_t (&m)[16], uint32_t (&out)[8], uint32_t one) {
    constexpr Schedule S = make_schedule();
    uint32_t s0 = 0x6A09E667u, s1 = 0xBB67AE85u, s2 = 0x3C6EF372u, s3 = 0xA54FF53Au, s4 = 0x510E527Fu, s5 = 0x9B05688Cu,
             s6 = 0x1F83D9ABu, s7 = 0x5BE0CD19u, s8 = 0x6A09E667u, s9 = 0xBB67AE85u, s10 = 0x3C6EF372u, s11 = 0xA54FF53Au,
             s12 = 0, s13 = 0, s14 = 64, s15 = 11;
#pragma unroll
    for (int r = 0; r < 7; r++) {
        G<Pick<V,0>::v>(s0, s4, s8, s12, m[S.s[r][0]], m[S.s[r][1]], one);  G<Pick<V,1>::v>(s1, s5, s9, s13, m[S.s[r][2]], m[S.s[r][3]], one);
        G<Pick<V,2>::v>(s2, s6, s10, s14, m[S.s[r][4]], m[S.s[r][5]], one); G<Pick<V,3>::v>(s3, s7, s11, s15, m[S.s[r][6]], m[S.s[r][7]], one);
        G<Pick<V,4>::v>(s0, s5, s10, s15, m[S.s[r][8]], m[S.s[r][9]], one); G<Pick<V,5>::v>(s1, s6, s11, s12, m[S.s[r][10]], m[S.s[r][11]], one);
        G<Pick<V,6>::v>(s2, s7, s8, s13, m[S.s[r][12]], m[S.s[r][13]], one); G<Pick<V,7>::v>(s3, s4, s9, s14, m[S.s[r][14]], m[S.s[r][15]], one);
    }
    out[0] = s0 ^ s8; out[1] = s1 ^ s9; out[2] = s2 ^ s10; out[3] = s3 ^ s11;
    out[4] = s4 ^ s12; out[5] = s5 ^ s13; out[6] = s6 ^ s14; out[7] = s7 ^ s15;
}

template <int V>
__global__ void __launch_bounds__(128) bench(uint32_t *out, int iters, uint32_t one) {
    uint32_t m[16];
    for (int i = 0; i < 16; i++) m[i] = threadIdx.x * 16 + i + blockIdx.x;
    for (int it = 0; it < iters; it++) {
        uint32_t o[8];
        compress<V>(m, o, one);
#pragma unroll
        for (int i = 0; i < 8; i++) { m[i] = o[i]; m[8 + i] ^= o[i]; }
    }
    uint32_t s = 0;
    for (int i = 0; i < 16; i++) s ^= m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V> void run(uint32_t *out, int sms, uint32_t *ref) {
    const int iters = 300, grid = sms * 16;
    bench<V><<<grid, 128>>>(out, iters, 1u);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<V><<<grid, 128>>>(out, iters, 1u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    uint32_t h[4]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    if (V == 0) { ref[0] = h[0]; ref[1] = h[1]; }
    double comps = (double)iters * 128.0 * grid;
    double cyc_per_G = ms * 1e-3 * 1.965e9 * sms * 4 * 32 / comps / 56.0;
    printf("V%d: %7.3f ms  %6.2f Gcomp/s  %5.2f clk/G/warp/SMSP  %s\n", V, ms, comps / ms / 1e6, cyc_per_G,
           (h[0] == ref[0] && h[1] == ref[1]) ? "ok" : "MISMATCH");
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint32_t *out; cudaMalloc(&out, p.multiProcessorCount * 16 * 128 * 4);
    uint32_t ref[2];
    int sms = p.multiProcessorCount;
    run<0>(out, sms, ref); run<1>(out, sms, ref); run<2>(out, sms, ref); run<3>(out, sms, ref);
    run<4>(out, sms, ref); run<5>(out, sms, ref); run<6>(out, sms, ref); run<7>(out, sms, ref);
    run<8>(out, sms, ref); run<9>(out, sms, ref); run<10>(out, sms, ref); run<11>(out, sms, ref); run<12>(out, sms, ref); run<13>(out, sms, ref); run<14>(out, sms, ref); run<15>(out, sms, ref); run<16>(out, sms, ref); run<17>(out, sms, ref);
    return 0;
}
