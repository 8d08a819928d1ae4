// zinc_b200/csrc/peer_roots.cu -- multi-GPU: the stand-alone form of the roots exchange over NVLink peer memory.
//
// A row-range-sharded commit (reference: the commitment is the list of per-row roots, commit.rs:71-81) has a
// single exchange step: every GPU needs every other GPU's 32-byte roots.  Normally that step is part of the kernel that
// produces the roots (merkle.cu + peer_sync.cuh).  This kernel is the same exchange for the cases where no single launch
// produces them (a chunked host pipeline that also returns the layers, a rank that owns no rows): it copies the local
// roots into every peer's result buffer (P2P stores through NVSwitch), publishes this rank's step counter in every
// peer's flag words with release semantics and waits, bounded, until the counters of all peers have reached this step.
// Buffers are double-buffered by step parity, so a rank that runs one step ahead never overwrites roots a slower rank
// is still reading.
#include "common.cuh"
#include "kernels.h"
#include "peer_sync.cuh"

namespace zipgpu {

__global__ void __launch_bounds__(1024) peer_roots_allgather_kernel(const RootsFanout *__restrict__ fan, unsigned long long step,
                                                                    const uint8_t *__restrict__ src_bytes, size_t offset,
                                                                    size_t nbytes) {
    const uint32_t tid = threadIdx.x;
    const size_t chunks = nbytes / 16;
    const uint4 *src = reinterpret_cast<const uint4 *>(src_bytes);
    const int par = (int)(step & 1ull);
    for (int p = 0; p < fan->world; p++) {
        uint4 *dst = reinterpret_cast<uint4 *>(fan->bufs[par][p] + offset);
        if (reinterpret_cast<const uint4 *>(dst) == src) continue;  // already in place in the own buffer
        for (size_t i = tid; i < chunks; i += blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();  // every copying thread fences its own stores; the barrier orders them before the publication
    __syncthreads();
    fan_handshake(fan, step, tid);
}

cudaError_t launch_peer_roots_allgather(const RootsFanout *fan, unsigned long long step, const uint8_t *src, size_t offset,
                                        size_t nbytes, cudaStream_t stream) {
    peer_roots_allgather_kernel<<<1, 1024, 0, stream>>>(fan, step, src, offset, nbytes);
    return cudaGetLastError();
}

}  // namespace zipgpu
