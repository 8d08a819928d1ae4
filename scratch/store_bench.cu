// scratch/store_bench.cu -- what store pattern reaches the B200 write bandwidth (1 GiB per launch)?
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ void st8(void *p, uint32_t a) {
    asm volatile("st.global.v8.u32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(a) : "memory");
}
__device__ __forceinline__ void st4(void *p, uint32_t a) {
    asm volatile("st.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(a) : "memory");
}
// MODE 0: encoder pattern: CTA = one 256 KiB row at a time; warp w owns [w*16K, +16K), 16 steps of 1 KiB (STG.256)
// MODE 1: same rows, but the CTA sweeps the row front to back: step it covers [it*16K, +16K) with all 16 warps (STG.256)
// MODE 2: grid-wide linear sweep with STG.256     MODE 3: grid-wide linear sweep with STG.128
template <int MODE>
__global__ void __launch_bounds__(512, 2) k(uint8_t *out, uint32_t num_rows, uint32_t v) {
    const uint32_t t = threadIdx.x, w = t >> 5, L = t & 31;
    if (MODE <= 1) {
        for (uint32_t row = blockIdx.x; row < num_rows; row += gridDim.x) {
            uint8_t *r = out + (size_t)row * 262144;
#pragma unroll
            for (int it = 0; it < 16; it++) {
                const uint32_t i = MODE == 0 ? (w * 512 + it * 32 + L) : (it * 512 + t);
                st8(r + (size_t)i * 32, v + i);
            }
        }
    } else {
        const size_t total = (size_t)num_rows * 262144;
        const size_t step = MODE == 2 ? 32 : 16;
        for (size_t off = ((size_t)blockIdx.x * blockDim.x + t) * step; off < total; off += (size_t)gridDim.x * blockDim.x * step) {
            if (MODE == 2) st8(out + off, v); else st4(out + off, v);
        }
    }
}
// MODE 4: MODE 0 with the data coming from 3 LDS per store (96 KiB planes);  MODE 5: MODE 4 + one CTA barrier per row
// MODE 6: MODE 5 + the planes refilled each row (48 STS per thread), like the encoder's park of s2
template <int MODE>
__global__ void __launch_bounds__(512, 2) k2(uint8_t *out, uint32_t num_rows, uint32_t v) {
    extern __shared__ uint32_t planes[];
    const uint32_t t = threadIdx.x, w = t >> 5, L = t & 31;
    for (uint32_t i = t; i < 3 * 8192; i += 512) planes[i] = i * v;
    __syncthreads();
    for (uint32_t row = blockIdx.x; row < num_rows; row += gridDim.x) {
        uint8_t *r = out + (size_t)row * 262144;
        if (MODE == 6) {
#pragma unroll
            for (int kk = 0; kk < 16; kk++) {
                const uint32_t s = kk * 512 + (t ^ ((kk << 1) & 31));
                planes[s] = row + kk; planes[8192 + s] = row ^ kk; planes[16384 + s] = row * kk;
            }
            __syncwarp();
        }
#pragma unroll
        for (int it = 0; it < 16; it++) {
            const uint32_t i = w * 512 + it * 32 + L;
            const uint32_t kk = i % 16, tt = i / 16;
            const uint32_t s = kk * 512 + (tt ^ ((kk << 1) & 31));
            const uint32_t a = planes[s], b = planes[8192 + s], c = planes[16384 + s];
            const uint32_t sg = (uint32_t)((int32_t)c >> 31);
            asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%4,%4,%4,%4};" ::"l"(r + (size_t)i * 32), "r"(a), "r"(b), "r"(c), "r"(sg) : "memory");
        }
        if (MODE >= 5) __syncthreads();
    }
}
template <int MODE> void run2(uint8_t *out, int grid, const char *name) {
    cudaFuncSetAttribute(k2<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 99072);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; i++) k2<MODE><<<grid, 512, 99072>>>(out, 4096, 1);
    cudaEventRecord(a);
    for (int i = 0; i < 20; i++) k2<MODE><<<grid, 512, 99072>>>(out, 4096, 1);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 20;
    printf("%-40s grid %4d: %.4f ms = %.0f GB/s  (%s)\n", name, grid, ms, 1073741824.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
template <int MODE> void run(uint8_t *out, int grid, const char *name) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; i++) k<MODE><<<grid, 512>>>(out, 4096, 1);
    cudaEventRecord(a);
    for (int i = 0; i < 20; i++) k<MODE><<<grid, 512>>>(out, 4096, 1);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 20;
    printf("%-40s grid %4d: %.4f ms = %.0f GB/s\n", name, grid, ms, 1073741824.0 / ms / 1e6);
}
int main() {
    uint8_t *out; cudaMalloc(&out, 1u << 30);
    for (int grid : {148, 296}) {
        run<0>(out, grid, "row per CTA, warp-local 16K spans");
        run<1>(out, grid, "row per CTA, CTA-wide front");
        run<2>(out, grid, "linear sweep STG.256");
        run<3>(out, grid, "linear sweep STG.128");
    }
    for (int grid : {148, 296}) {
        run2<4>(out, grid, "MODE0 + 3 LDS per store");
        run2<5>(out, grid, "  + CTA barrier per row");
        run2<6>(out, grid, "  + 48 STS per thread per row");
    }
    run<2>(out, 148 * 8, "linear sweep STG.256"); run<3>(out, 148 * 8, "linear sweep STG.128");
    return 0;
}
