// zinc_b200/csrc/peer_roots.cu -- multi-GPU: the all-gather of the row roots as ONE kernel over NVLink peer memory.
//
// A row-range-sharded commit (reference: the commitment is the list of per-row roots, commit.rs:71-81) has a
// single exchange step: every GPU needs every other GPU's 32-byte roots.  Instead of an NCCL all-gather (a separate
// launch plus its proxy/handshake latency, which is a tenth of a 0.33 ms sharded commit at N = 8), one small kernel
// stores this GPU's roots straight into every peer's result buffer (P2P stores through NVSwitch), publishes a per-rank
// step counter in every peer's flag words with release semantics, and waits until the counters of all peers have
// reached this step.  Buffers are double-buffered by step parity, so a rank that runs one step ahead never overwrites
// roots a slower rank is still reading.
#include "common.cuh"
#include "kernels.h"

namespace zipgpu {

__global__ void __launch_bounds__(1024) peer_roots_allgather_kernel(PeerRootsArgs a) {
    const uint32_t tid = threadIdx.x;
    const size_t chunks = a.nbytes / 16;
    const uint4 *src = reinterpret_cast<const uint4 *>(a.src);
    for (int p = 0; p < a.world; p++) {
        uint4 *dst = reinterpret_cast<uint4 *>(a.peer_roots[p] + a.offset);
        if (reinterpret_cast<const uint4 *>(dst) == src) continue;  // already in place in the own buffer
        for (size_t i = tid; i < chunks; i += blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < (uint32_t)a.world) {
        // publish: flag word [my rank] of peer `tid`
        volatile unsigned long long *f = a.peer_flags[tid] + a.rank;
        *f = a.step;
        // wait for peer `tid`'s publication in the own flag words; bounded so that a missing rank traps instead of hanging
        volatile unsigned long long *mine = a.peer_flags[a.rank] + tid;
        unsigned long long spins = 0;
        while (*mine < a.step) {
            if (++spins > (1ull << 27)) __trap();  // tens of seconds
        }
    }
    __threadfence_system();
    __syncthreads();
}

cudaError_t launch_peer_roots_allgather(const PeerRootsArgs &a, cudaStream_t stream) {
    peer_roots_allgather_kernel<<<1, 1024, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace zipgpu
