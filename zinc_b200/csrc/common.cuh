// zinc_b200/csrc/common.cuh -- shared helpers for the sm_100a kernels of libzipgpu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace zipgpu {

constexpr int kWarp = 32;

// streaming (read-once) 16-byte global load / store: keep L1 for the tables that are reused
__device__ __forceinline__ uint4 ld_stream_v4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_v4(uint4 *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

// asynchronous bulk prefetch of [p, p+bytes) into L2 (sm_90+: one instruction, no registers, no completion to wait for).
// p and bytes must be multiples of 16.
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes) {
    if ((bytes & 15u) == 0 && bytes != 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---- bulk asynchronous shared -> global copies (the TMA engine; sm_90+).  The stores of the encoder leave the SM
// through this path instead of the LSU, whose data stage is the kernel's bottleneck.  `bytes` and both addresses must
// be multiples of 16.  Completion is tracked per thread in bulk async-groups. ----
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_s2g(void *gmem, const void *smem, uint32_t bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem), "r"(s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// all bulk groups of this thread have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all bulk groups of this thread are complete (the global writes are done)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- programmatic dependent launch (sm_90+).  A kernel launched through launch_pdl may be scheduled while its
// predecessor on the stream is still draining; it must execute pdl_wait() before touching anything the predecessor wrote
// (a no-op when the kernel was launched the ordinary way).  Hides the ~2-3 us launch gap between the short, dependent
// kernels of a small commit. ----
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- mbarrier (shared-memory arrive/wait objects, sm_80+): unlike bar.sync, waiters do not synchronise with each
// other -- every thread (warp) proceeds as soon as the phase it waits for has completed ----
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {  // release.cta
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {  // acquire.cta
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!done);
}

// 32-byte global accesses (sm_100: LDG/STG.E.ENL2.256): one full sector per lane, so a 32-byte digest or Int<4>
// never reaches L2 as two partial-sector writes
struct __align__(32) u32x8 {
    uint32_t w[8];
};
__device__ __forceinline__ void st_global_v8(void *p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
// streaming (evict-first) 256-bit store: the 1 GiB of codewords passes through L2 once and must not push out the
// next rows' prefetched inputs or the permutation tables
__device__ __forceinline__ void st_stream_v8(void *p, const uint32_t (&v)[8]) {
    asm volatile("st.global.cs.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld_global_v8(const void *p, uint32_t (&v)[8]) {
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
// ordered forms: not moved across barriers / fences by the compiler.  _sync: data written earlier by this CTA;
// _cg: data written by another CTA (L2, bypassing the L1 of this SM)
__device__ __forceinline__ void ld_global_v8_sync(const void *p, uint32_t (&v)[8]) {
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void ld_global_cg_v8(const void *p, uint32_t (&v)[8]) {
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void ld_stream_v8(const void *p, uint32_t (&v)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}

// ---- multi-limb (W x u32, little-endian) wrap-around add: one carry chain per call ---------------------
template <int W>
__device__ __forceinline__ void add_limbs(uint32_t (&a)[W], const uint32_t (&b)[W]);

template <>
__device__ __forceinline__ void add_limbs<2>(uint32_t (&a)[2], const uint32_t (&b)[2]) {
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(a[0]), "+r"(a[1]) : "r"(b[0]), "r"(b[1]));
}
template <>
__device__ __forceinline__ void add_limbs<3>(uint32_t (&a)[3], const uint32_t (&b)[3]) {
    asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, %5;"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2])
        : "r"(b[0]), "r"(b[1]), "r"(b[2]));
}
template <>
__device__ __forceinline__ void add_limbs<4>(uint32_t (&a)[4], const uint32_t (&b)[4]) {
    asm("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %6;\n\taddc.u32 %3, %3, %7;"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3])
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
}
template <>
__device__ __forceinline__ void add_limbs<5>(uint32_t (&a)[5], const uint32_t (&b)[5]) {
    asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, %7;\n\taddc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %4, %9;"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4])
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]));
}
template <>
__device__ __forceinline__ void add_limbs<6>(uint32_t (&a)[6], const uint32_t (&b)[6]) {
    asm("add.cc.u32 %0, %0, %6;\n\taddc.cc.u32 %1, %1, %7;\n\taddc.cc.u32 %2, %2, %8;\n\taddc.cc.u32 %3, %3, %9;\n\t"
        "addc.cc.u32 %4, %4, %10;\n\taddc.u32 %5, %5, %11;"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5])
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]));
}

}  // namespace zipgpu
