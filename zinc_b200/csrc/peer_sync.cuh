// zinc_b200/csrc/peer_sync.cuh -- device side of the roots exchange of a row-sharded commit (RootsFanout, kernels.h).
//
// The reference's commitment is the list of per-row roots (commit.rs:71-81), so a commit sharded by row range has ONE
// exchange step: every GPU needs every other GPU's 32-byte roots.  It is fused into the kernel that produces them:
//   * fan_store_root: the thread that holds a finished root stores it into every rank's result buffer (P2P stores
//     through NVLink / NVSwitch; the own buffer is one of them) and fences at system scope;
//   * fan_finish: every CTA counts itself done; the LAST CTA of the launch publishes this rank's step counter into
//     every peer's flag words (st.release.sys) and waits until every peer's counter in the own flag words has reached
//     the step (ld.acquire.sys).  When the kernel completes, the own result
//     buffer holds all roots.  The wait is bounded: a peer that does not show up within timeout_ns is reported through
//     the mapped status word and the kernel ends -- the context stays usable and the host call returns an error.
#pragma once
#include <stdint.h>

#include "common.cuh"
#include "kernels.h"

namespace zipgpu {

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// publish + wait, by the threads tid < world of ONE CTA (called by the last CTA of the launch, or by the stand-alone
// exchange kernel).  The caller's writes to the peers must happen-before this call (a system-scope fence by the calling
// thread, or by a thread it has synchronised with).  st.release.sys / ld.acquire.sys carry the ordering themselves: no
// extra fences around them (each system fence costs ~3-7 us on a B200; measured, scripts/fan_probe.py).
__device__ __forceinline__ void fan_handshake(const RootsFanout *fan, unsigned long long step, uint32_t tid) {
    if (tid < (uint32_t)fan->world) {
        st_release_sys_u64(fan->flags[tid] + fan->rank, step);  // flag word [my rank] of peer `tid`
        const unsigned long long *mine = fan->flags[fan->rank] + tid;
        const unsigned long long t0 = global_timer_ns();
        unsigned int spins = 0;
        while (ld_acquire_sys_u64(mine) < step) {
            if ((++spins & 63u) == 0) {
                if (global_timer_ns() - t0 > fan->timeout_ns) {
                    *reinterpret_cast<volatile unsigned int *>(fan->status) = 1u + tid;  // mapped host memory
                    __threadfence_system();
                    break;
                }
                __nanosleep(200);
            }
        }
    }
}

// The thread that holds a finished root stores it into every rank's result buffer (P2P stores through NVLink; the own
// buffer is one of them) while the rest of the grid is still hashing.  No fence here: the CTA's thread 0 fences once for
// all of them in fan_finish.
__device__ __forceinline__ void fan_store_root(const RootsFanout *fan, unsigned long long step, uint32_t global_row,
                                               const uint32_t (&d)[8]) {
    const int par = (int)(step & 1ull);
    for (int p = 0; p < fan->world; p++) st_global_v8(fan->bufs[par][p] + (size_t)global_row * 32, d);
}

// End of a kernel whose threads called fan_store_root.  EVERY thread of EVERY CTA must reach it (it contains barriers).
// Per CTA: barrier (the storing threads' writes happen-before thread 0), ONE system-scope fence by thread 0 (cumulative
// over them), then the CTA counts itself done with a device-scope atomic.  The CTA that counts last has therefore
// synchronised with every storing CTA after its system fence, and publishes the flags with release semantics.
// Cost on one B200 (world = 1, scripts/fan_probe.py): +13 us per commit; the first version (a system fence per stored
// root, fences around the handshake) cost +24 us.  (Also measured: only local stores during the launch and the last CTA
// copying all roots to the peers behind one fence -- 0.318 vs 0.308 ms per sharded commit at N = 8.)
__device__ __forceinline__ void fan_finish(const RootsFanout *fan, unsigned long long step) {
    __shared__ unsigned int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        const unsigned int prev = atomicAdd(fan->done, 1u);
        s_last = (prev == total - 1u) ? 1u : 0u;
        if (s_last) {
            *fan->done = 0u;  // re-armed for the next launch on this stream (ordered by the kernel boundary)
            __threadfence();  // acquire side of the counter: the other CTAs' fenced stores happen-before the flags below
        }
    }
    __syncthreads();
    if (s_last) fan_handshake(fan, step, threadIdx.x);
}

}  // namespace zipgpu
