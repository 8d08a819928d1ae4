// zinc_b200/csrc/zipgpu.cu -- C ABI of libzipgpu (include/zipgpu.h): contexts, per-pp code state, the
// host<->device pipeline around the kernels, device-resident prover data.
//
// Host-pointer entry points split the rows into chunks and run  H2D(chunk k+1) || kernels(chunk k) || D2H(chunk k-1)
// on three streams, so that the PCIe copies of a commit hide behind the hashing.  Device buffers are recycled by a
// per-context cache (dev_alloc/dev_free), so steady-state calls never reach cudaMalloc.
#include <map>
#include <sys/mman.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/zipgpu.h"
#include "kernels.h"

namespace zipgpu {
void perm_from_seed(uint64_t seed, uint32_t n, uint32_t *perm);
void chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]);
}

using namespace zipgpu;

// ------------------------------------------------------------------------------------------------------
// state
// ------------------------------------------------------------------------------------------------------
struct ProfRec {
    cudaEvent_t e0, e1, e2;  // before encode, between, after hash
    bool has_enc, has_hash;
    bool is_part;            // a continuation of an earlier record (deferred top passes): not a new call
};

// Device buffers are recycled through a small per-context cache (a freed buffer carries the event after which
// it may be reused on another stream), so steady-state calls never reach cudaMalloc/cudaFree.
struct CachedBuf {
    void *p;
    size_t bytes;
    cudaEvent_t ready;
    cudaStream_t last = nullptr;  // the stream `ready` was recorded on: a reuse on the same stream needs no wait
};

constexpr size_t kPairFlagRing = 64, kPairFlagWords = 512;  // one array per fused launch in flight, one word per CTA boundary

struct zipgpu_ctx {
    std::vector<CachedBuf> cache_free;
    std::vector<CachedBuf> cache_live;
    std::mutex cache_mu;
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;  // kernels
    cudaStream_t stream2 = nullptr; // kernels of every other chunk of a host job (tails overlap the next chunk)
    cudaStream_t stream_hi = nullptr;  // high priority: the sparse code's tensor-core GEMM, so that its one CTA per SM is
                                       // placed as soon as hash CTAs of the previous row chunk retire
    cudaStream_t h2d = nullptr;
    cudaStream_t d2h = nullptr;
    std::vector<cudaEvent_t> ring;  // timing-disabled events for cross-stream ordering
    size_t ring_pos = 0;
    std::atomic<uint64_t> launches{0};
    bool profile = false;
    std::vector<ProfRec> prof_pending;
    std::vector<ProfRec> prof_free;
    double enc_ms = 0, hash_ms = 0;
    uint64_t prof_calls = 0;
    uint32_t *d_sink = nullptr;
    uint32_t *d_row_counters = nullptr;  // ring of row-claim counters, one per encoder launch in flight
    size_t row_counter_pos = 0;
    uint32_t *d_pair_flags = nullptr;    // ring of zeroed boundary-flag arrays for the tops epilogue of commit_ws_kernel
    size_t pair_flags_pos = 0;
    std::mutex mu;
    // Public entry points serialise on this (recursive: some call each other), so several host threads may share one
    // context; their jobs are then enqueued one after the other on the context's streams.
    std::recursive_mutex api_mu;
};
#define API_LOCK(c) std::lock_guard<std::recursive_mutex> api_lock__((c)->api_mu)

struct zipgpu_code {
    zipgpu_ctx *ctx;
    size_t row_len, rep, cw;
    int in_limbs, out_limbs;
    int depth;  // -1 when cw is not a power of two (encode only)
    int fused_levels;  // Merkle levels the fused commit kernel produces (0: no fused variant for this shape)
    void *d_tables = nullptr;   // one cached buffer holding the three tables below
    uint16_t *d_tab1, *d_tab2;  // pre-translated gather tables (raa_encode.cu)
    uint8_t *d_colw;
    // codewords longer than one SM's shared memory holds (cw > 16384): chunked encoder of raa_big.cu, raw permutations
    bool big = false;
    uint32_t *d_perm1 = nullptr, *d_perm2 = nullptr;  // (also uploaded on first use by encode_f)
    std::vector<uint32_t> h_perm1, h_perm2;           // host copy of the permutations
    // ZipLinearCode (zipgpu_sparse_code_create): either the ELL tables of the generic kernel or the dense 0/1 matrix
    // of the tensor-core kernel (sparse_encode.cu)
    bool sparse = false;
    uint32_t sp_d = 0;
    uint32_t *d_sp_cols = nullptr;
    int64_t *d_sp_coef = nullptr;
    uint8_t *d_sp_dense = nullptr;
    uint32_t *d_sp_bias = nullptr;
};

struct zipgpu_data {
    zipgpu_ctx *ctx;
    size_t num_rows, cw, row_len;
    int in_limbs, out_limbs, depth;
    uint64_t *d_evals;  // the unencoded evaluations (the proximity test combines them, open_z.rs:100-113)
    uint64_t *d_rows;
    uint8_t *d_layers;
    uint8_t *d_roots;
    uint8_t *d_roots_all = nullptr;  // row-sharded commit: ALL roots of the commitment (after the exchange)
    size_t total_rows = 0;
};

static thread_local std::string g_err;

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
static int cuda_fail(cudaError_t e, const char *what) {
    g_err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return e == cudaErrorMemoryAllocation ? ZIPGPU_ERR_NOMEM : ZIPGPU_ERR_CUDA;
}
#define CU(call)                                           \
    do {                                                   \
        cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

static bool is_pow2(size_t x) { return x && !(x & (x - 1)); }
static int ilog2(size_t x) {
    int l = 0;
    while (x >>= 1) l++;
    return l;
}
static size_t layers_per_row(int depth) { return ((size_t)2 << depth) - 2; }

static cudaEvent_t next_event(zipgpu_ctx *c) {
    cudaEvent_t e = c->ring[c->ring_pos];
    c->ring_pos = (c->ring_pos + 1) % c->ring.size();
    return e;
}
// make `waiter` wait for everything enqueued so far on `src`
static cudaError_t chain(zipgpu_ctx *c, cudaStream_t src, cudaStream_t waiter) {
    if (src == waiter) return cudaSuccess;
    cudaEvent_t e = next_event(c);
    cudaError_t err = cudaEventRecord(e, src);
    if (err != cudaSuccess) return err;
    return cudaStreamWaitEvent(waiter, e, 0);
}

static cudaError_t dev_alloc(zipgpu_ctx *c, void **out, size_t bytes, cudaStream_t s) {
    bytes = std::max<size_t>((bytes + 511) & ~(size_t)511, 512);
    std::lock_guard<std::mutex> lk(c->cache_mu);
    size_t best = (size_t)-1;
    for (size_t i = 0; i < c->cache_free.size(); i++) {
        const size_t b = c->cache_free[i].bytes;
        if (b >= bytes && b <= bytes + bytes / 4 && (best == (size_t)-1 || b < c->cache_free[best].bytes)) best = i;
    }
    CachedBuf buf;
    if (best != (size_t)-1) {
        buf = c->cache_free[best];
        c->cache_free.erase(c->cache_free.begin() + best);
        if (buf.last != s) {  // (stream order already covers a reuse on the stream it was released on)
            cudaError_t e = cudaStreamWaitEvent(s, buf.ready, 0);
            if (e != cudaSuccess) return e;
        }
    } else {
        cudaError_t e = cudaMalloc(&buf.p, bytes);
        if (e == cudaErrorMemoryAllocation) {  // drop the cache and retry once
            cudaGetLastError();
            cudaDeviceSynchronize();
            for (auto &f : c->cache_free) { cudaFree(f.p); cudaEventDestroy(f.ready); }
            c->cache_free.clear();
            e = cudaMalloc(&buf.p, bytes);
        }
        if (e != cudaSuccess) return e;
        buf.bytes = bytes;
        e = cudaEventCreateWithFlags(&buf.ready, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    c->cache_live.push_back(buf);
    *out = buf.p;
    return cudaSuccess;
}
// the buffer may be handed out again once everything enqueued on `s` so far has completed
static cudaError_t dev_free(zipgpu_ctx *c, void *p, cudaStream_t s) {
    if (!p) return cudaSuccess;
    std::lock_guard<std::mutex> lk(c->cache_mu);
    for (size_t i = 0; i < c->cache_live.size(); i++) {
        if (c->cache_live[i].p == p) {
            CachedBuf buf = c->cache_live[i];
            c->cache_live.erase(c->cache_live.begin() + i);
            cudaError_t e = cudaEventRecord(buf.ready, s);
            buf.last = s;
            c->cache_free.push_back(buf);
            return e;
        }
    }
    return cudaErrorInvalidValue;
}
// Error-path hygiene: every function that takes buffers from the cache declares a DevGuard; buffers it allocated and
// neither freed nor handed over (release) by the time it returns -- i.e. on an early error return -- go back to the
// cache instead of staying "live" for the life of the context.
struct DevGuard {
    zipgpu_ctx *ctx;
    cudaStream_t s;
    std::vector<void *> owned;
    DevGuard *prev;
    static thread_local DevGuard *current;
    DevGuard(zipgpu_ctx *c, cudaStream_t st) : ctx(c), s(st), prev(current) { current = this; }
    ~DevGuard() {
        current = prev;
        if (!owned.empty()) {
            // an early error return: the copy / second kernel streams may still be using these buffers, and the cache
            // only orders a reuse after `s` -- wait for all of the context's streams before handing them back
            cudaStreamSynchronize(ctx->h2d);
            cudaStreamSynchronize(ctx->d2h);
            cudaStreamSynchronize(ctx->stream2);
            cudaStreamSynchronize(ctx->stream_hi);
            cudaStreamSynchronize(ctx->stream);
            if (s != ctx->stream) cudaStreamSynchronize(s);
            cudaGetLastError();
        }
        for (void *p : owned) dev_free(ctx, p, s);
    }
    void add(void *p) { owned.push_back(p); }
    void drop(void *p) {
        for (size_t i = 0; i < owned.size(); i++)
            if (owned[i] == p) {
                owned.erase(owned.begin() + i);
                return;
            }
    }
    void release(void *p) { drop(p); }  // ownership moves elsewhere (a zipgpu_data handle)
};
thread_local DevGuard *DevGuard::current = nullptr;

static cudaError_t dev_alloc_tracked(zipgpu_ctx *c, void **out, size_t bytes, cudaStream_t s) {
    cudaError_t e = dev_alloc(c, out, bytes, s);
    if (e == cudaSuccess && DevGuard::current) DevGuard::current->add(*out);
    return e;
}
static cudaError_t dev_free_tracked(zipgpu_ctx *c, void *p, cudaStream_t s) {
    for (DevGuard *g = DevGuard::current; g; g = g->prev) g->drop(p);
    return dev_free(c, p, s);
}
#define DEV_ALLOC(ctx, ptr, bytes, s) CU(dev_alloc_tracked((ctx), (void **)(ptr), (bytes), (s)))
#define DEV_FREE(ctx, ptr, s) CU(dev_free_tracked((ctx), (ptr), (s)))

// ------------------------------------------------------------------------------------------------------
// library / context
// ------------------------------------------------------------------------------------------------------
extern "C" const char *zipgpu_version(void) { return "zipgpu 0.1 (sm_100a)"; }
extern "C" const char *zipgpu_last_error(void) { return g_err.c_str(); }
namespace zipgpu {
// mgpu.cu runs the per-device calls on worker threads and hands their (thread-local) message to the calling thread
void set_last_error(const std::string &msg) { g_err = msg; }
}

extern "C" int zipgpu_device_count(int *count) {
    if (!count) return fail(ZIPGPU_ERR_INVALID, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
        return fail(ZIPGPU_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *count = n;
    return ZIPGPU_OK;
}

extern "C" void zipgpu_ctx_destroy(zipgpu_ctx *c);

extern "C" int zipgpu_ctx_create(int device, zipgpu_ctx **out) {
    if (!out) return fail(ZIPGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(ZIPGPU_ERR_NO_DEVICE, "no CUDA device visible: libzipgpu has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(ZIPGPU_ERR_INVALID, "device index out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        return fail(ZIPGPU_ERR_NO_DEVICE, std::string("libzipgpu is built for sm_100a only; device is ") + prop.name +
                                              " (sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")");
    }
    zipgpu_ctx *c = new (std::nothrow) zipgpu_ctx();
    if (!c) return fail(ZIPGPU_ERR_NOMEM, "host allocation failed");
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithPriority(&c->stream_hi, cudaStreamNonBlocking, -5)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking)) != cudaSuccess) {
        zipgpu_ctx_destroy(c);  // releases whatever was created so far
        return cuda_fail(e, "cudaStreamCreate");
    }
    c->ring.assign(256, nullptr);
    for (auto &ev : c->ring) {
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) {
            zipgpu_ctx_destroy(c);
            return cuda_fail(e, "cudaEventCreate");
        }
    }
    if ((e = cudaMalloc(&c->d_pair_flags, kPairFlagRing * kPairFlagWords * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMemset(c->d_pair_flags, 0, kPairFlagRing * kPairFlagWords * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&c->d_row_counters, 2 * 256 * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&c->d_sink, 256)) != cudaSuccess) {
        zipgpu_ctx_destroy(c);
        return cuda_fail(e, "cudaMalloc");
    }
    *out = c;
    return ZIPGPU_OK;
}

extern "C" void zipgpu_ctx_destroy(zipgpu_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto &r : c->prof_pending) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); cudaEventDestroy(r.e2); }
    for (auto &r : c->prof_free) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); cudaEventDestroy(r.e2); }
    for (auto &ev : c->ring)
        if (ev) cudaEventDestroy(ev);
    for (auto &b : c->cache_free) { cudaFree(b.p); cudaEventDestroy(b.ready); }
    for (auto &b : c->cache_live) { cudaFree(b.p); cudaEventDestroy(b.ready); }
    if (c->d_sink) cudaFree(c->d_sink);
    if (c->d_row_counters) cudaFree(c->d_row_counters);
    if (c->d_pair_flags) cudaFree(c->d_pair_flags);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream_hi) cudaStreamDestroy(c->stream_hi);
    if (c->h2d) cudaStreamDestroy(c->h2d);
    if (c->d2h) cudaStreamDestroy(c->d2h);
    cudaGetLastError();
    delete c;
}

extern "C" int zipgpu_ctx_device(const zipgpu_ctx *c) { return c ? c->device : -1; }
extern "C" uint64_t zipgpu_ctx_launch_count(const zipgpu_ctx *c) { return c ? c->launches.load() : 0; }

extern "C" int zipgpu_ctx_sync(zipgpu_ctx *c) {
    if (!c) return fail(ZIPGPU_ERR_INVALID, "ctx is NULL");
    API_LOCK(c);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->h2d));
    CU(cudaStreamSynchronize(c->stream2));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->d2h));
    return ZIPGPU_OK;
}

// after a host job: its copy / second kernel streams have been joined back into the kernel stream, so waiting for that
// one stream is waiting for the whole job (three fewer driver calls than zipgpu_ctx_sync on the latency path)
static int sync_job(zipgpu_ctx *c) {
    CU(cudaStreamSynchronize(c->stream));
    return ZIPGPU_OK;
}

// Pinned host memory for evaluations / roots / proof streams.  The DMA rate of a pinned buffer depends on what backs it:
// measured on the B200 box (a VM behind an IOMMU), an 8 MiB buffer that cudaHostAlloc carved out of a fragmented
// process (4 KiB pages) fed the GPU at 20-27 GB/s, one backed by 2 MiB pages at 46-55 GB/s -- the nv = 20 host-to-host
// commit took 0.55 vs 0.26 ms.  So: a 2 MiB-aligned anonymous mapping with transparent huge pages requested,
// pre-faulted, then page-locked (cudaHostRegister); cudaHostAlloc only as the fallback (ZIPGPU_HOST_ALLOC_PLAIN=1
// forces it).
namespace {
struct HostBlock { void *map; size_t map_len, len; bool registered; };
std::mutex g_host_mu;
std::map<void *, HostBlock> g_host_blocks;
}  // namespace

extern "C" int zipgpu_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(ZIPGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    constexpr size_t kHuge = (size_t)2 << 20;
    static const bool plain = getenv("ZIPGPU_HOST_ALLOC_PLAIN") != nullptr;
    if (!plain && bytes >= kHuge) {
        const size_t len = (bytes + kHuge - 1) & ~(kHuge - 1), map_len = len + kHuge;
        void *map = mmap(nullptr, map_len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (map != MAP_FAILED) {
            void *p = (void *)(((uintptr_t)map + kHuge - 1) & ~(uintptr_t)(kHuge - 1));
            madvise(p, len, MADV_HUGEPAGE);  // advisory: without THP this is an ordinary pre-faulted pinned mapping
            for (size_t off = 0; off < len; off += 4096) ((volatile char *)p)[off] = 0;
            if (cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess) {
                std::lock_guard<std::mutex> lk(g_host_mu);
                g_host_blocks[p] = HostBlock{map, map_len, len, true};
                *out = p;
                return ZIPGPU_OK;
            }
            cudaGetLastError();
            munmap(map, map_len);
        }
    }
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc");
    return ZIPGPU_OK;
}
extern "C" int zipgpu_host_free(void *p) {
    if (!p) return ZIPGPU_OK;
    HostBlock b{};
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        auto it = g_host_blocks.find(p);
        if (it != g_host_blocks.end()) {
            b = it->second;
            g_host_blocks.erase(it);
        }
    }
    if (b.map) {
        cudaError_t e = cudaHostUnregister(p);
        munmap(b.map, b.map_len);
        if (e != cudaSuccess) return cuda_fail(e, "cudaHostUnregister");
        return ZIPGPU_OK;
    }
    CU(cudaFreeHost(p));
    return ZIPGPU_OK;
}
extern "C" int zipgpu_host_register(void *p, size_t bytes) {
    CU(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
    return ZIPGPU_OK;
}
extern "C" int zipgpu_host_unregister(void *p) {
    CU(cudaHostUnregister(p));
    return ZIPGPU_OK;
}

// ------------------------------------------------------------------------------------------------------
// code geometry (host only)
// ------------------------------------------------------------------------------------------------------
extern "C" int zipgpu_perm_from_seed(uint64_t seed, uint32_t n, uint32_t *perm_out) {
    if (!perm_out && n) return fail(ZIPGPU_ERR_INVALID, "perm_out is NULL");
    perm_from_seed(seed, n, perm_out);
    return ZIPGPU_OK;
}

extern "C" int zipgpu_chacha_block(const uint32_t *key, uint64_t counter, int rounds, uint32_t *out) {
    if (!key || !out) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (rounds < 2 || rounds > 20 || (rounds & 1)) return fail(ZIPGPU_ERR_INVALID, "rounds must be even, 2..20");
    chacha_block(key, counter, rounds, out);
    return ZIPGPU_OK;
}

static uint64_t isqrt64(uint64_t x) {
    uint64_t r = 0;
    for (uint64_t bit = 1ull << 31; bit; bit >>= 1) {
        const uint64_t t = r | bit;
        if (t * t <= x) r = t;
    }
    return r;
}
static size_t next_pow2(size_t x) {
    size_t p = 1;
    while (p < x) p <<= 1;
    return p;
}
// code_raa.rs:42-43
extern "C" size_t zipgpu_raa_row_len(size_t poly_size) {
    if (poly_size == 0) return 0;
    const int num_vars = ilog2(poly_size);
    return next_pow2((size_t)isqrt64(1ull << num_vars));
}
// pcs/structs.rs:79-90
extern "C" size_t zipgpu_num_rows(size_t poly_size, size_t row_len) {
    if (poly_size == 0 || row_len == 0) return 0;
    const int num_vars = ilog2(poly_size);
    return next_pow2(((size_t)1 << num_vars) / row_len);
}
// code_raa.rs:53-67
extern "C" int zipgpu_raa_codeword_width_bits(int in_limbs, size_t poly_size, size_t rep) {
    const int num_vars = poly_size ? ilog2(poly_size) : 0;
    const int nv_even = (num_vars % 2 == 0) ? num_vars : num_vars + 1;
    return 64 * in_limbs + nv_even + 2 * ilog2(next_pow2(rep ? rep : 1));
}

// ------------------------------------------------------------------------------------------------------
// code (per-pp state)
// ------------------------------------------------------------------------------------------------------
static bool is_permutation(const uint32_t *p, size_t n) {
    std::vector<uint8_t> seen(n, 0);
    for (size_t i = 0; i < n; i++) {
        if (p[i] >= n || seen[p[i]]) return false;
        seen[p[i]] = 1;
    }
    return true;
}

extern "C" int zipgpu_code_create(zipgpu_ctx *ctx, size_t row_len, size_t rep, int in_limbs, int out_limbs,
                                  const uint32_t *perm1, const uint32_t *perm2, zipgpu_code **out) {
    if (!out) return fail(ZIPGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!ctx || !perm1 || !perm2) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (row_len == 0 || rep == 0) return fail(ZIPGPU_ERR_INVALID, "row_len and repetition_factor must be positive");
    if (in_limbs < 1 || out_limbs < in_limbs) return fail(ZIPGPU_ERR_INVALID, "need 1 <= in_limbs <= out_limbs");
    const size_t cw = row_len * rep;
    if (cw > (1u << 24)) return fail(ZIPGPU_ERR_UNSUPPORTED, "codeword longer than 2^24");
    // the narrowest width the reference's assert (code_raa.rs:53-72) can accept for this codeword length
    const int need_bits = 64 * in_limbs + 2 * ilog2(next_pow2(cw));
    if (64 * out_limbs < need_bits)
        return fail(ZIPGPU_ERR_WIDTH, "Cannot fit " + std::to_string(need_bits) + "-bit wide codeword entries in " +
                                          std::to_string(64 * out_limbs) + " bits integers");
    if (!is_permutation(perm1, cw) || !is_permutation(perm2, cw))
        return fail(ZIPGPU_ERR_INVALID, "perm1/perm2 must be permutations of [0, codeword_len)");
    const bool big = !encode_supported(in_limbs, (uint32_t)cw, (uint32_t)row_len);
    if (big && !raa_big_supported(in_limbs, (uint32_t)cw))
        return fail(ZIPGPU_ERR_UNSUPPORTED, "no encoder kernel for in_limbs=" + std::to_string(in_limbs) +
                                                " codeword_len=" + std::to_string(cw));
    API_LOCK(ctx);
    CU(cudaSetDevice(ctx->device));
    zipgpu_code *c = new (std::nothrow) zipgpu_code();
    if (!c) return fail(ZIPGPU_ERR_NOMEM, "host allocation failed");
    c->ctx = ctx;
    c->row_len = row_len;
    c->rep = rep;
    c->cw = cw;
    c->in_limbs = in_limbs;
    c->out_limbs = out_limbs;
    c->depth = is_pow2(cw) ? ilog2(cw) : -1;
    c->fused_levels = c->depth > 0 ? encode_fused_levels(in_limbs, out_limbs, (uint32_t)row_len, (uint32_t)cw) : 0;
    if (!merkle_supported(out_limbs * 2)) c->fused_levels = 0;
    c->h_perm1.assign(perm1, perm1 + cw);
    c->h_perm2.assign(perm2, perm2 + cw);
    c->d_tab1 = c->d_tab2 = nullptr;
    c->d_colw = nullptr;
    cudaError_t e;
    if (big) {  // the chunked encoder reads the permutations as they are
        c->big = true;
        c->fused_levels = c->depth > 0 && merkle_supported(out_limbs * 2) ? raa_big_fused_levels(in_limbs, out_limbs, (uint32_t)cw) : 0;
        if ((e = cudaMalloc(&c->d_perm1, cw * 4)) != cudaSuccess || (e = cudaMalloc(&c->d_perm2, cw * 4)) != cudaSuccess ||
            (e = cudaMemcpy(c->d_perm1, perm1, cw * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaMemcpy(c->d_perm2, perm2, cw * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
            cudaFree(c->d_perm1);
            cudaFree(c->d_perm2);
            delete c;
            return cuda_fail(e, "cudaMalloc/cudaMemcpy(permutations)");
        }
        *out = c;
        return ZIPGPU_OK;
    }
    const size_t padded = encode_perm_padded_len((uint32_t)cw);
    // the kernel consumes pre-translated, lane-major tables (raa_encode.cu), not the raw permutations.  One buffer from
    // the context's cache (a prover builds a new code per proof, zinc/prover.rs:313: no cudaMalloc in steady state), one copy.
    std::vector<uint8_t> host(padded * 5, 0);
    uint16_t *t1 = reinterpret_cast<uint16_t *>(host.data()), *t2 = reinterpret_cast<uint16_t *>(host.data() + padded * 2);
    uint8_t *cl = host.data() + padded * 4;
    build_encode_tables(perm1, perm2, (uint32_t)row_len, (uint32_t)cw, in_limbs, out_limbs, t1, t2, cl);
    void *d_tables = nullptr;
    if ((e = dev_alloc(ctx, &d_tables, padded * 5, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_tables, host.data(), padded * 5, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
        if (d_tables) dev_free(ctx, d_tables, ctx->stream);
        delete c;
        return cuda_fail(e, "code tables");
    }
    if (in_limbs == 1 && out_limbs == 4 && commit_ws16k_supported((uint32_t)row_len, (uint32_t)cw)) {
        // the cw = 16384 commit kernel gathers pass 1 through the permutation as uploaded
        if ((e = cudaMalloc(&c->d_perm1, cw * 4)) != cudaSuccess || (e = cudaMalloc(&c->d_perm2, cw * 4)) != cudaSuccess ||
            (e = cudaMemcpy(c->d_perm1, perm1, cw * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaMemcpy(c->d_perm2, perm2, cw * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
            cudaFree(c->d_perm1);
            cudaFree(c->d_perm2);
            dev_free(ctx, d_tables, ctx->stream);
            delete c;
            return cuda_fail(e, "cudaMalloc/cudaMemcpy(permutations)");
        }
    }
    c->d_tables = d_tables;
    c->d_tab1 = reinterpret_cast<uint16_t *>(d_tables);
    c->d_tab2 = reinterpret_cast<uint16_t *>(static_cast<uint8_t *>(d_tables) + padded * 2);
    c->d_colw = static_cast<uint8_t *>(d_tables) + padded * 4;
    *out = c;
    return ZIPGPU_OK;
}

// ZipLinearCode::new's product (zip/code.rs:100-147): the two sampled matrices are inputs, see include/zipgpu.h
extern "C" int zipgpu_sparse_code_create(zipgpu_ctx *ctx, size_t row_len, size_t codeword_len, size_t cells_per_row,
                                         int in_limbs, int out_limbs, const uint32_t *cols_a, const int64_t *coef_a,
                                         const uint32_t *cols_b, const int64_t *coef_b, zipgpu_code **out) {
    if (!out) return fail(ZIPGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!ctx || !cols_a || !coef_a || !cols_b || !coef_b) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (row_len == 0 || codeword_len == 0 || (codeword_len & 1))
        return fail(ZIPGPU_ERR_INVALID, "row_len must be positive and codeword_len a positive even number");
    if (cells_per_row == 0 || cells_per_row > row_len)
        return fail(ZIPGPU_ERR_INVALID, "cells_per_row must be in [1, row_len]");
    if (in_limbs < 1 || out_limbs < in_limbs || out_limbs > 8)
        return fail(ZIPGPU_ERR_INVALID, "need 1 <= in_limbs <= out_limbs <= 8");
    if (codeword_len > (1u << 24) || row_len > (1u << 24)) return fail(ZIPGPU_ERR_UNSUPPORTED, "codeword longer than 2^24");
    const size_t n = codeword_len / 2, d = cells_per_row, cw = codeword_len;
    // the tensor-core kernel takes coefficients 0..cmax with 255*cmax*row_len < 2^31 (exact s32 accumulation)
    int64_t cmin = 0, cmax = 0;
    for (size_t i = 0; i < n * d; i++) {
        if (cols_a[i] >= row_len || cols_b[i] >= row_len)
            return fail(ZIPGPU_ERR_INVALID, "sparse matrix column index out of range");
        cmin = std::min(cmin, std::min(coef_a[i], coef_b[i]));
        cmax = std::max(cmax, std::max(coef_a[i], coef_b[i]));
    }
    bool dense = cmin >= 0 && cmax <= 255 && (uint64_t)255 * (uint64_t)std::max<int64_t>(cmax, 1) * row_len < (1ull << 31) &&
                 sparse_gemm_supported(in_limbs, out_limbs, (uint32_t)row_len, (uint32_t)cw) &&
                 (uint64_t)cw * row_len <= (1ull << 31) &&  // the dense byte matrix must stay a sane size
                 !getenv("ZIPGPU_SPARSE_GENERIC");
    std::vector<uint8_t> hd;
    std::vector<uint32_t> hbias;
    if (dense) {
        hd.assign(cw * row_len, 0);
        hbias.assign(cw, 0);
        for (size_t j = 0; j < cw && dense; j++) {
            const uint32_t *cols = j < n ? cols_a + j * d : cols_b + (j - n) * d;
            const int64_t *coef = j < n ? coef_a + j * d : coef_b + (j - n) * d;
            for (size_t k = 0; k < d; k++) {
                const uint32_t v = hd[j * row_len + cols[k]] + (uint32_t)coef[k];  // repeated columns add up
                if (v > (uint32_t)std::max<int64_t>(cmax, 1)) { dense = false; break; }
                hd[j * row_len + cols[k]] = (uint8_t)v;
                hbias[j] += (uint32_t)coef[k];
            }
        }
    }
    API_LOCK(ctx);
    CU(cudaSetDevice(ctx->device));
    zipgpu_code *c = new (std::nothrow) zipgpu_code();
    if (!c) return fail(ZIPGPU_ERR_NOMEM, "host allocation failed");
    c->ctx = ctx;
    c->row_len = row_len;
    c->rep = cw / row_len;
    c->cw = cw;
    c->in_limbs = in_limbs;
    c->out_limbs = out_limbs;
    c->depth = is_pow2(cw) ? ilog2(cw) : -1;
    c->fused_levels = 0;
    c->d_tab1 = c->d_tab2 = nullptr;
    c->d_colw = nullptr;
    c->sparse = true;
    c->sp_d = (uint32_t)d;
    cudaError_t e = cudaSuccess;
    if (dense) {
        if ((e = cudaMalloc(&c->d_sp_dense, cw * row_len)) == cudaSuccess && (e = cudaMalloc(&c->d_sp_bias, cw * 4)) == cudaSuccess &&
            (e = cudaMemcpy(c->d_sp_dense, hd.data(), cw * row_len, cudaMemcpyHostToDevice)) == cudaSuccess)
            e = cudaMemcpy(c->d_sp_bias, hbias.data(), cw * 4, cudaMemcpyHostToDevice);
    } else {
        // ELL, transposed to [d][cw] so that a warp's 32 codeword entries read consecutive words
        std::vector<uint32_t> tc(d * cw);
        std::vector<int64_t> tf(d * cw);
        for (size_t j = 0; j < cw; j++) {
            const uint32_t *cols = j < n ? cols_a + j * d : cols_b + (j - n) * d;
            const int64_t *coef = j < n ? coef_a + j * d : coef_b + (j - n) * d;
            for (size_t k = 0; k < d; k++) {
                tc[k * cw + j] = cols[k];
                tf[k * cw + j] = coef[k];
            }
        }
        if ((e = cudaMalloc(&c->d_sp_cols, d * cw * 4)) == cudaSuccess && (e = cudaMalloc(&c->d_sp_coef, d * cw * 8)) == cudaSuccess &&
            (e = cudaMemcpy(c->d_sp_cols, tc.data(), d * cw * 4, cudaMemcpyHostToDevice)) == cudaSuccess)
            e = cudaMemcpy(c->d_sp_coef, tf.data(), d * cw * 8, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        cudaFree(c->d_sp_dense);
        cudaFree(c->d_sp_bias);
        cudaFree(c->d_sp_cols);
        cudaFree(c->d_sp_coef);
        delete c;
        return cuda_fail(e, "sparse code tables");
    }
    *out = c;
    return ZIPGPU_OK;
}
/* 1 when the code runs on the tensor-core kernel (coefficients 0..255 and a tileable shape), 0 generic, -1 not sparse */
extern "C" int zipgpu_code_sparse_kind(const zipgpu_code *c) { return !c || !c->sparse ? -1 : c->d_sp_dense ? 1 : 0; }

extern "C" void zipgpu_code_destroy(zipgpu_code *c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaDeviceSynchronize();
    if (c->d_tables) dev_free(c->ctx, c->d_tables, c->ctx->stream);  // back to the context's cache
    cudaFree(c->d_perm1);
    cudaFree(c->d_perm2);
    cudaFree(c->d_sp_dense);
    cudaFree(c->d_sp_bias);
    cudaFree(c->d_sp_cols);
    cudaFree(c->d_sp_coef);
    delete c;
}
extern "C" size_t zipgpu_code_row_len(const zipgpu_code *c) { return c ? c->row_len : 0; }
extern "C" size_t zipgpu_code_codeword_len(const zipgpu_code *c) { return c ? c->cw : 0; }
extern "C" int zipgpu_code_merkle_depth(const zipgpu_code *c) { return c ? ilog2(next_pow2(c->cw)) : -1; }

// ------------------------------------------------------------------------------------------------------
// multi-GPU: the roots exchange of a row-sharded commit (peer_sync.cuh, peer_roots.cu, merkle.cu)
// ------------------------------------------------------------------------------------------------------
struct zipgpu_peer_roots {
    zipgpu_ctx *ctx;
    size_t total_rows, buf_bytes;
    int rank, world;
    uint8_t *base = nullptr;                 // own allocation: [buffer 0 | buffer 1 | flags (PEER_MAX u64) | done counter]
    uint8_t *peer_base[PEER_MAX] = {};       // every rank's allocation as addressable from this device (own = base)
    bool ipc_opened[PEER_MAX] = {};          // mapped with cudaIpcOpenMemHandle (to be closed)
    RootsFanout *d_fan = nullptr;            // the device-side descriptor, written once at connect time
    unsigned int *h_status = nullptr;        // mapped pinned host word the kernels report a missing peer through
    bool connected = false;
    unsigned long long step = 0;             // steps launched so far (incremented only after a successful launch)
};

static unsigned long long peer_timeout_ns() {
    if (const char *env = getenv("ZIPGPU_PEER_TIMEOUT_MS")) return (unsigned long long)atoll(env) * 1000000ull;
    return 20ull * 1000000000ull;
}

extern "C" void zipgpu_peer_roots_destroy(zipgpu_peer_roots *p);

extern "C" int zipgpu_peer_roots_create(zipgpu_ctx *ctx, size_t total_rows, int rank, int world, zipgpu_peer_roots **out,
                                        uint8_t *ipc_out) {
    if (!out) return fail(ZIPGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!ctx) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world) return fail(ZIPGPU_ERR_INVALID, "bad rank / world");
    static_assert(sizeof(cudaIpcMemHandle_t) <= ZIPGPU_IPC_BYTES, "IPC handle size");
    API_LOCK(ctx);
    CU(cudaSetDevice(ctx->device));
    zipgpu_peer_roots *p = new (std::nothrow) zipgpu_peer_roots();
    if (!p) return fail(ZIPGPU_ERR_NOMEM, "host allocation failed");
    p->ctx = ctx;
    p->total_rows = total_rows;
    p->buf_bytes = (total_rows * 32 + 255) & ~(size_t)255;
    p->rank = rank;
    p->world = world;
    const size_t bytes = 2 * p->buf_bytes + PEER_MAX * sizeof(unsigned long long) + 256;
    cudaError_t e = cudaMalloc(&p->base, bytes);  // IPC needs a cudaMalloc allocation of its own
    if (e == cudaSuccess) e = cudaMemset(p->base, 0, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_fan, sizeof(RootsFanout));
    if (e == cudaSuccess) e = cudaHostAlloc(&p->h_status, sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess) *p->h_status = 0;
    if (e == cudaSuccess && ipc_out) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, p->base);
        if (e == cudaSuccess) {
            memset(ipc_out, 0, ZIPGPU_IPC_BYTES);
            memcpy(ipc_out, &h, sizeof(h));
        }
    }
    if (e != cudaSuccess) {
        zipgpu_peer_roots_destroy(p);
        return cuda_fail(e, "peer roots buffers");
    }
    p->peer_base[rank] = p->base;
    *out = p;
    return ZIPGPU_OK;
}

// every rank's base pointer is known: write the device-side descriptor
static int peer_roots_finish_connect(zipgpu_peer_roots *p) {
    RootsFanout f;
    memset(&f, 0, sizeof(f));
    for (int r = 0; r < p->world; r++) {
        f.bufs[0][r] = p->peer_base[r];
        f.bufs[1][r] = p->peer_base[r] + p->buf_bytes;
        f.flags[r] = reinterpret_cast<unsigned long long *>(p->peer_base[r] + 2 * p->buf_bytes);
    }
    f.done = reinterpret_cast<unsigned int *>(p->base + 2 * p->buf_bytes + PEER_MAX * sizeof(unsigned long long));
    void *d_status = nullptr;
    CU(cudaHostGetDevicePointer(&d_status, p->h_status, 0));
    f.status = static_cast<unsigned int *>(d_status);
    f.timeout_ns = peer_timeout_ns();
    f.rank = p->rank;
    f.world = p->world;
    CU(cudaMemcpy(p->d_fan, &f, sizeof(f), cudaMemcpyHostToDevice));
    p->connected = true;
    return ZIPGPU_OK;
}

// one process per GPU: the peers' buffers arrive as CUDA IPC handles
extern "C" int zipgpu_peer_roots_connect(zipgpu_peer_roots *p, const uint8_t *ipc_all) {
    if (!p || (!ipc_all && p->world > 1)) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    API_LOCK(p->ctx);
    CU(cudaSetDevice(p->ctx->device));
    for (int r = 0; r < p->world; r++) {
        if (r == p->rank || p->peer_base[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, ipc_all + (size_t)r * ZIPGPU_IPC_BYTES, sizeof(h));
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle (peer roots)");
        p->peer_base[r] = static_cast<uint8_t *>(ptr);
        p->ipc_opened[r] = true;
    }
    return peer_roots_finish_connect(p);
}

// all ranks in ONE process (one context per device): direct peer access, no IPC
extern "C" int zipgpu_peer_roots_connect_local(zipgpu_peer_roots *const *all, int n) {
    if (!all || n < 1 || n > PEER_MAX) return fail(ZIPGPU_ERR_INVALID, "bad argument");
    for (int r = 0; r < n; r++)
        if (!all[r] || all[r]->rank != r || all[r]->world != n || all[r]->total_rows != all[0]->total_rows)
            return fail(ZIPGPU_ERR_INVALID, "connect_local: the objects must be ranks 0..n-1 of one exchange");
    for (int r = 0; r < n; r++) {
        zipgpu_peer_roots *p = all[r];
        API_LOCK(p->ctx);
        CU(cudaSetDevice(p->ctx->device));
        for (int q = 0; q < n; q++) {
            if (q == r) continue;
            const int peer_dev = all[q]->ctx->device;
            if (peer_dev != p->ctx->device) {
                int can = 0;
                CU(cudaDeviceCanAccessPeer(&can, p->ctx->device, peer_dev));
                if (!can)
                    return fail(ZIPGPU_ERR_UNSUPPORTED, "device " + std::to_string(p->ctx->device) + " cannot access device " +
                                                            std::to_string(peer_dev) + " (no peer access)");
                cudaError_t e = cudaDeviceEnablePeerAccess(peer_dev, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
            }
            p->peer_base[q] = all[q]->base;
        }
        int rc = peer_roots_finish_connect(p);
        if (rc) return rc;
    }
    return ZIPGPU_OK;
}

// 0, or ZIPGPU_ERR_PEER_TIMEOUT if a kernel of this exchange gave up waiting for a peer (call after synchronising)
extern "C" int zipgpu_peer_roots_status(zipgpu_peer_roots *p) {
    if (!p) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    const unsigned int st = *reinterpret_cast<volatile unsigned int *>(p->h_status);
    if (st == 0) return ZIPGPU_OK;
    return fail(ZIPGPU_ERR_PEER_TIMEOUT, "roots exchange: rank " + std::to_string(st - 1) + " did not publish its roots within " +
                                             std::to_string(peer_timeout_ns() / 1000000ull) +
                                             " ms; the exchange object must be recreated on all ranks");
}

extern "C" int zipgpu_peer_roots_allgather(zipgpu_peer_roots *p, size_t row_begin, size_t count, const uint8_t *d_local_roots,
                                           void *stream, uint8_t **d_all_out) {
    if (!p || !d_all_out || (count && !d_local_roots)) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (!p->connected) return fail(ZIPGPU_ERR_INVALID, "zipgpu_peer_roots_connect has not been called");
    if (row_begin + count > p->total_rows) return fail(ZIPGPU_ERR_INVALID, "row range outside the commitment");
    if (((row_begin * 32) | (uintptr_t)d_local_roots) & 15) return fail(ZIPGPU_ERR_INVALID, "roots must be 16-byte aligned");
    API_LOCK(p->ctx);
    CU(cudaSetDevice(p->ctx->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : p->ctx->stream;
    const unsigned long long step = p->step + 1;
    cudaError_t e = launch_peer_roots_allgather(p->d_fan, step, d_local_roots, row_begin * 32, count * 32, s);
    if (e != cudaSuccess) return cuda_fail(e, "launch_peer_roots_allgather");
    p->step = step;  // only now: a failed launch must not leave this rank a step ahead of its peers
    p->ctx->launches++;
    *d_all_out = p->base + (step & 1) * p->buf_bytes;
    return ZIPGPU_OK;
}

extern "C" void zipgpu_peer_roots_destroy(zipgpu_peer_roots *p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < p->world; r++)
        if (r != p->rank && p->peer_base[r] && p->ipc_opened[r]) cudaIpcCloseMemHandle(p->peer_base[r]);
    cudaFree(p->base);
    cudaFree(p->d_fan);
    if (p->h_status) cudaFreeHost(p->h_status);
    cudaGetLastError();
    delete p;
}

// ------------------------------------------------------------------------------------------------------
// profiling records
// ------------------------------------------------------------------------------------------------------
static bool prof_begin(zipgpu_ctx *c, ProfRec *r) {
    if (!c->profile) return false;
    if (!c->prof_free.empty()) {
        *r = c->prof_free.back();
        c->prof_free.pop_back();
    } else {
        if (cudaEventCreate(&r->e0) != cudaSuccess || cudaEventCreate(&r->e1) != cudaSuccess ||
            cudaEventCreate(&r->e2) != cudaSuccess)
            return false;
    }
    r->has_enc = r->has_hash = r->is_part = false;
    return true;
}

// ------------------------------------------------------------------------------------------------------
// device-level building blocks
// ------------------------------------------------------------------------------------------------------
static int check_align16(const void *p, const char *name) {  // 32 bytes: rows, digests move as 256-bit accesses
    if (((uintptr_t)p & 31) != 0) return fail(ZIPGPU_ERR_INVALID, std::string(name) + " must be 32-byte aligned");
    return ZIPGPU_OK;
}

// fuse_layers != NULL: the fused commit kernel (encode + Merkle levels 0..code->fused_levels into fuse_layers)
// A roots exchange requested for the rows of one commit_dev / merkle_dev call (row-sharded multi-GPU commit): the
// launch that produces the roots carries it when there is exactly one such launch (`fused` reports it); otherwise
// finish_exchange() runs the stand-alone kernel.
struct FanReq {
    zipgpu_peer_roots *pr;
    size_t row_begin;    // index, in the whole commitment, of the first local row
    bool fused = false;
};

static int encode_dev(zipgpu_code *code, size_t num_rows, const uint64_t *d_evals, uint64_t *d_rows, cudaStream_t s,
                      uint8_t *fuse_layers = nullptr, uint64_t *evals_copy = nullptr, int *fused_levels = nullptr,
                      uint8_t *tops_roots = nullptr, FanReq *fan = nullptr) {
    if (num_rows == 0) return ZIPGPU_OK;
    if (num_rows > 0xffffffffull) return fail(ZIPGPU_ERR_UNSUPPORTED, "too many rows");
    int rc;
    if ((rc = check_align16(d_evals, "evals")) || (rc = check_align16(d_rows, "rows_out"))) return rc;
    if (code->sparse) {
        if (fuse_layers || evals_copy) return fail(ZIPGPU_ERR_INVALID, "the sparse code has no fused commit kernel");
        zipgpu_ctx *ctx = code->ctx;
        SparseEncodeArgs sa;
        sa.evals = d_evals;
        sa.rows_out = d_rows;
        sa.num_rows = (uint32_t)num_rows;
        sa.row_len = (uint32_t)code->row_len;
        sa.cw = (uint32_t)code->cw;
        sa.in_limbs = code->in_limbs;
        sa.out_limbs = code->out_limbs;
        sa.num_sms = ctx->num_sms;
        sa.cols_t = code->d_sp_cols;
        sa.coef_t = code->d_sp_coef;
        sa.d = code->sp_d;
        sa.dense = code->d_sp_dense;
        sa.nnz = code->d_sp_bias;
        sa.use_umma = !getenv("ZIPGPU_SPARSE_MMA_SYNC");
        sa.stream = s;
        DevGuard guard(ctx, s);
        if (sa.dense) DEV_ALLOC(ctx, &sa.planes, sparse_planes_bytes(sa.num_rows, sa.row_len, sa.in_limbs), s);
        int n = 0;
        cudaError_t e = launch_sparse_encode(sa, &n);
        if (e != cudaSuccess) return cuda_fail(e, "launch_sparse_encode");
        ctx->launches += n;
        if (sa.planes) DEV_FREE(ctx, sa.planes, s);
        return ZIPGPU_OK;
    }
    if (code->big) {
        if (evals_copy || (fuse_layers && code->fused_levels <= 0))
            return fail(ZIPGPU_ERR_INVALID, "no fused commit kernel for this codeword length");
        zipgpu_ctx *ctx = code->ctx;
        BigEncodeArgs b;
        b.fuse_layers = fuse_layers;
        if (fuse_layers && fused_levels) *fused_levels = code->fused_levels;
        b.evals = reinterpret_cast<const uint32_t *>(d_evals);
        b.rows_out = reinterpret_cast<uint32_t *>(d_rows);
        b.perm1 = code->d_perm1;
        b.perm2 = code->d_perm2;
        b.num_rows = (uint32_t)num_rows;
        b.row_len = (uint32_t)code->row_len;
        b.cw = (uint32_t)code->cw;
        b.out32 = (uint32_t)code->out_limbs * 2;
        b.in_limbs = code->in_limbs;
        b.stream = s;
        size_t scratch_bytes = 0;
        raa_big_plan(code->in_limbs, b.cw, b.num_rows, &b.batch_rows, &scratch_bytes);
        b.num_sms = ctx->num_sms;
        b.scratch_bytes = scratch_bytes;
        DevGuard guard(ctx, s);
        DEV_ALLOC(ctx, &b.scratch, scratch_bytes, s);
        int n = 0;
        cudaError_t e = launch_raa_encode_big(b, &n);
        if (e != cudaSuccess) return cuda_fail(e, "launch_raa_encode_big");
        ctx->launches += n;
        DEV_FREE(ctx, b.scratch, s);
        return ZIPGPU_OK;
    }
    EncodeArgs a;
    a.evals = reinterpret_cast<const uint32_t *>(d_evals);
    a.rows_out = reinterpret_cast<uint32_t *>(d_rows);
    a.tab1 = code->d_tab1;
    a.tab2 = code->d_tab2;
    a.colw = code->d_colw;
    a.perm1_raw = code->d_perm1;
    a.perm2_raw = code->d_perm2;
    void *wsc_scratch = nullptr;
    const char *wsc_env = getenv("ZIPGPU_WSC");  // the cluster kernel is opt-in (read per call: the tests switch it)
    if (wsc_env && wsc_env[0] == '1' && fuse_layers && code->d_perm2 && code->in_limbs == 1 && code->out_limbs == 4 &&
        commit_wsc_supported((uint32_t)code->row_len, (uint32_t)code->cw)) {
        // the cluster commit kernel of this shape keeps s1 of the rows in flight in an L2-resident scratch
        cudaError_t ea = dev_alloc(code->ctx, &wsc_scratch, commit_wsc_scratch_bytes(code->ctx->num_sms), s);
        if (ea != cudaSuccess) return cuda_fail(ea, "dev_alloc(wsc scratch)");
        a.wsc_scratch = wsc_scratch;
    }
    a.num_rows = (uint32_t)num_rows;
    a.row_len = (uint32_t)code->row_len;
    a.cw = (uint32_t)code->cw;
    a.out32 = (uint32_t)code->out_limbs * 2;
    a.in_limbs = code->in_limbs;
    a.num_sms = code->ctx->num_sms;
    a.fuse_layers = fuse_layers;
    a.fused_levels_out = fused_levels;
    a.evals_copy = reinterpret_cast<uint32_t *>(evals_copy);
    if (!getenv("ZIPGPU_STATIC_ROWS")) {
        a.row_counter = code->ctx->d_row_counters + 2 * (code->ctx->row_counter_pos++ % 256);
    }
    bool fan_fused = false;
    if (tops_roots && fuse_layers) {  // the warp-specialised kernel may finish the trees (and carry the roots exchange)
        a.tops_roots = tops_roots;
        a.pair_flags = code->ctx->d_pair_flags + kPairFlagWords * (code->ctx->pair_flags_pos++ % kPairFlagRing);
        if (fan && !fan->fused) {
            a.fan = fan->pr->d_fan;
            a.fan_step = fan->pr->step + 1;
            a.fan_row_begin = (uint32_t)fan->row_begin;
            a.fan_fused = &fan_fused;
        }
    }
    a.stream = s;
    cudaError_t e = launch_raa_encode(a);
    if (wsc_scratch) dev_free(code->ctx, wsc_scratch, s);
    if (e != cudaSuccess) return cuda_fail(e, "launch_raa_encode");
    code->ctx->launches++;
    if (fan_fused) {  // the exchange of this step is in flight
        fan->fused = true;
        fan->pr->step++;
    }
    return ZIPGPU_OK;
}

// tree passes from `from_level` (0 = from the raw leaves) until a level >= until_level (-1: the roots)
static int merkle_dev(zipgpu_ctx *ctx, size_t num_rows, int depth, int leaf_limbs, const uint64_t *d_leaves,
                      uint8_t *d_layers, uint8_t *d_roots, cudaStream_t s, int from_level = 0, int until_level = -1,
                      int *reached = nullptr, FanReq *fan = nullptr) {
    if (reached) *reached = from_level;
    if (num_rows == 0) return ZIPGPU_OK;
    if (num_rows > 0xffffffffull) return fail(ZIPGPU_ERR_UNSUPPORTED, "too many rows");
    if (depth < 0 || depth > 30) return fail(ZIPGPU_ERR_INVALID, "depth out of range");
    if (!merkle_supported(leaf_limbs * 2))
        return fail(ZIPGPU_ERR_UNSUPPORTED, "leaf_limbs must be one of 1,2,3,4,8");
    MerkleArgs a;
    a.leaves = reinterpret_cast<const uint32_t *>(d_leaves);
    a.layers = d_layers;
    a.roots = d_roots;
    a.num_rows = (uint32_t)num_rows;
    a.depth = depth;
    a.leaf32 = leaf_limbs * 2;
    a.stream = s;
    a.num_sms = ctx->num_sms;
    if (fan && !fan->fused) {
        a.fan = fan->pr->d_fan;
        a.fan_step = fan->pr->step + 1;
        a.fan_row_begin = (uint32_t)fan->row_begin;
        a.fan_fused = &fan->fused;
    }
    int n = 0;
    cudaError_t e = launch_merkle_levels(a, from_level, until_level, reached, &n);
    if (e != cudaSuccess) return cuda_fail(e, "launch_merkle_levels");
    ctx->launches += (uint64_t)n;
    if (fan && fan->fused && a.fan) fan->pr->step++;  // the exchange of this step is in flight
    return ZIPGPU_OK;
}

// After the kernels of a sharded commit: if no launch carried the exchange, run the stand-alone kernel on the local
// roots.  *d_all = this rank's result buffer of the step (all total_rows roots once the stream has passed this point).
static int finish_exchange(FanReq &fan, const uint8_t *d_local_roots, size_t count, cudaStream_t s, uint8_t **d_all) {
    zipgpu_peer_roots *p = fan.pr;
    if (!fan.fused) {
        const unsigned long long step = p->step + 1;
        cudaError_t e = launch_peer_roots_allgather(p->d_fan, step, d_local_roots, fan.row_begin * 32, count * 32, s);
        if (e != cudaSuccess) return cuda_fail(e, "launch_peer_roots_allgather");
        p->step = step;
        p->ctx->launches++;
        fan.fused = true;
    }
    if (d_all) *d_all = p->base + (p->step & 1) * p->buf_bytes;
    return ZIPGPU_OK;
}

static bool fusion_enabled() {
    return getenv("ZIPGPU_NO_FUSE") == nullptr;  // A/B knob: the two-kernel path (encode, then hash)
}
// rows from which commit_dev takes the fused kernel (ZIPGPU_FUSE_MIN_ROWS overrides: tests force it at small shapes)
static size_t fuse_min_rows(const zipgpu_ctx *ctx, const zipgpu_code *code) {
    if (const char *env = getenv("ZIPGPU_FUSE_MIN_ROWS")) return (size_t)atol(env);
    // measured break-even (scripts/shard_sweep.py, round 2): the warp-specialised kernel (Int<1> -> Int<4>) beats encode +
    // leaf pass from 128 rows for cw = 2048 / 4096 / 8192 (cw = 8192: 256 rows 0.162 vs 0.166 ms, 512 rows 0.277 vs 0.302;
    // cw = 2048: 128 rows 0.043 vs 0.052) and from 256 rows for cw = 1024 (nv = 17: 0.031 vs 0.035 ms); the two-CTA fused
    // kernel of the other shapes from ~10 rows per CTA slot
    const bool ws = code->in_limbs == 1 && code->out_limbs == 4 &&
                    (code->cw == 1024 || code->cw == 2048 || code->cw == 4096 || code->cw == 8192);
    if (ws) return code->cw == 1024 ? 256 : 128;
    return (size_t)(code->cw >= 8192 ? 6 : 10) * ctx->num_sms;
}

// Encode + Merkle of a row range on stream s, with optional profiling events.  until_level >= 0 stops the trees at the
// first pass boundary >= until_level (chunked pipelines finish them with merkle_top_dev); *reached reports it.
// Exact shapes run the fused commit kernel: encode + the lowest log2(E) tree levels in one launch.
static int commit_dev(zipgpu_code *code, size_t num_rows, const uint64_t *d_evals, uint64_t *d_rows, uint8_t *d_layers,
                      uint8_t *d_roots, cudaStream_t s, int until_level = -1, int *reached = nullptr,
                      uint64_t *evals_copy = nullptr, FanReq *fan = nullptr) {
    zipgpu_ctx *ctx = code->ctx;
    ProfRec r;
    const bool prof = prof_begin(ctx, &r);
    if (prof) cudaEventRecord(r.e0, s);
    // The fused kernel is one persistent CTA pair per SM: it needs a few rows per CTA for one CTA's encode phases to
    // run under the other's hashing; below that the two-kernel path, whose hash passes balance at subtree granularity,
    // is faster.
    bool rows_ok = num_rows >= fuse_min_rows(ctx, code);
    if (!rows_ok && !getenv("ZIPGPU_FUSE_MIN_ROWS") && !code->big && !code->sparse && code->in_limbs == 1 && code->out_limbs == 4 &&
        commit_ws16k_supported((uint32_t)code->row_len, (uint32_t)code->cw)) {
        // cw = 16384: one 1024-thread CTA per SM takes a row at a time, so below the threshold the serial fused kernel
        // still wins whenever the rows fill whole waves (296 rows: 0.309 vs 0.340 ms, 740: 0.723 vs 0.798; but 512 rows =
        // 3.46 waves: 0.584 vs 0.576)
        const size_t sms = (size_t)ctx->num_sms, waves = (num_rows + sms - 1) / sms;
        rows_ok = waves * sms * 100 <= num_rows * 112;
    }
    const bool fuse = d_roots && d_layers && code->fused_levels > 0 && code->fused_levels <= code->depth && rows_ok &&
                      fusion_enabled();
    if (evals_copy && !fuse) return fail(ZIPGPU_ERR_INVALID, "zero-copy input needs the fused commit kernel");
    // Sparse code on the tensor cores: the GEMM (TMA + tcgen05, 6 warps per SM, hardly any INT32 work) and the BLAKE3
    // passes (INT32-alu-bound, no shared memory) use different parts of an SM.  Row chunks are encoded back to back on
    // the high-priority stream while the trees of the chunks already encoded are hashed on `s`; the narrow top passes
    // run once over all rows at the end.  Everything is joined back into `s`.  Measured: the hash warps take issue
    // slots from the MMA-issuing and epilogue warps, so the gain is small -- +7 % at nv = 26, +2 % at nv = 24, a loss
    // at nv = 22 (8 x 4 short launches) -- hence only for >= 2^26 codeword entries (ZIPGPU_SPARSE_OVERLAP=1/0 forces).
    const char *ov = getenv("ZIPGPU_SPARSE_OVERLAP");
    const bool overlap = ov ? ov[0] == '1' : num_rows * code->cw >= ((size_t)1 << 26);
    if (code->sparse && code->d_sp_dense && d_roots && d_layers && code->depth > 7 && num_rows >= 1024 && overlap) {
        const size_t chunks = 8;
        const size_t per = ((num_rows + chunks - 1) / chunks + 31) & ~(size_t)31;
        const size_t per_row_layers = (((size_t)2 << code->depth) - 2) * 32;
        cudaStream_t hi = ctx->stream_hi;
        CU(chain(ctx, s, hi));
        int split_level = 0, chunk_level = 0;
        bool first_chunk = true;
        for (size_t r0 = 0; r0 < num_rows; r0 += per) {
            const size_t nr = std::min(per, num_rows - r0);
            int rc = encode_dev(code, nr, d_evals + r0 * code->row_len * code->in_limbs,
                                d_rows + r0 * code->cw * code->out_limbs, hi);
            if (rc) return rc;
            CU(chain(ctx, hi, s));
            rc = merkle_dev(ctx, nr, code->depth, code->out_limbs, d_rows + r0 * code->cw * code->out_limbs,
                            d_layers + r0 * per_row_layers, d_roots + r0 * 32, s, 0, until_level >= 0 ? until_level : 6,
                            &chunk_level);
            if (rc) return rc;
            split_level = first_chunk ? chunk_level : std::min(split_level, chunk_level);  // lowest level any chunk stopped at
            first_chunk = false;
        }
        if (until_level < 0 && split_level < code->depth) {
            int rc = merkle_dev(ctx, num_rows, code->depth, code->out_limbs, d_rows, d_layers, d_roots, s, split_level, -1,
                                nullptr, fan);
            if (rc) return rc;
            split_level = code->depth;
        }
        if (reached) *reached = split_level;
        if (prof) {
            cudaEventRecord(r.e1, s);  // the two phases overlap: the whole commit is reported as one interval
            cudaEventRecord(r.e2, s);
            r.has_enc = true;
            r.has_hash = true;
            std::lock_guard<std::mutex> lk(ctx->mu);
            ctx->prof_pending.push_back(r);
        }
        return ZIPGPU_OK;
    }
    int fused_levels = code->fused_levels;  // the launch reports how far it really built the trees (sub-row units stop lower)
    // (whole trees in the fused launch only when this call goes all the way to the roots)
    uint8_t *tops_roots = fuse && d_roots && d_layers && until_level < 0 ? d_roots : nullptr;
    int rc = encode_dev(code, num_rows, d_evals, d_rows, s, fuse ? d_layers : nullptr, evals_copy, &fused_levels, tops_roots,
                        fan);
    if (rc) return rc;
    if (prof) {
        cudaEventRecord(r.e1, s);
        r.has_enc = true;
    }
    if (reached) *reached = 0;
    if (d_roots) {
        rc = merkle_dev(ctx, num_rows, code->depth, code->out_limbs, d_rows, d_layers, d_roots, s,
                        fuse ? fused_levels : 0, until_level, reached, fan);
        if (rc) return rc;
        if (prof) {
            cudaEventRecord(r.e2, s);
            r.has_hash = true;
        }
    }
    if (prof) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->prof_pending.push_back(r);
    }
    return ZIPGPU_OK;
}

// the passes above `from_level` of the trees of `num_rows` rows whose lower levels are already in d_layers
static int merkle_top_dev(zipgpu_code *code, size_t num_rows, const uint64_t *d_rows, uint8_t *d_layers, uint8_t *d_roots,
                          cudaStream_t s, int from_level, FanReq *fan = nullptr) {
    zipgpu_ctx *ctx = code->ctx;
    ProfRec r;
    const bool prof = prof_begin(ctx, &r);
    if (prof) cudaEventRecord(r.e1, s);
    int rc = merkle_dev(ctx, num_rows, code->depth, code->out_limbs, d_rows, d_layers, d_roots, s, from_level, -1, nullptr,
                        fan);
    if (rc) return rc;
    if (prof) {
        cudaEventRecord(r.e2, s);
        r.has_hash = true;
        r.is_part = true;
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->prof_pending.push_back(r);
    }
    return ZIPGPU_OK;
}

// ------------------------------------------------------------------------------------------------------
// encode_rows
// ------------------------------------------------------------------------------------------------------
extern "C" int zipgpu_encode_rows_device(zipgpu_code *code, size_t num_rows, const uint64_t *d_evals,
                                         uint64_t *d_rows_out, void *stream) {
    if (!code || (num_rows && (!d_evals || !d_rows_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    API_LOCK(code->ctx);
    CU(cudaSetDevice(code->ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : code->ctx->stream;
    return commit_dev(code, num_rows, d_evals, d_rows_out, nullptr, nullptr, s);
}

static size_t pick_chunk_rows(size_t num_rows, size_t bytes_per_row_in, int num_sms) {
    if (const char *env = getenv("ZIPGPU_CHUNK_ROWS")) {  // tuning knob for experiments
        const long v = atol(env);
        if (v > 0) return std::min<size_t>((size_t)v, std::max<size_t>(num_rows, 1));
    }
    // Measured on B200 (scratch/e2e_mid.py, e2e_probe.py): below 4 MiB of input one chunk is best (the launches of a
    // chunk cost more than the overlap gains); above, chunks of 1/16 of the input clamped to 2..8 MiB and never fewer
    // than 256 rows keep the copies back to back and the exposed tail short (nv = 20: 0.34 -> 0.29 ms).
    const size_t total = num_rows * std::max<size_t>(bytes_per_row_in, 1);
    if (total < (4u << 20)) return std::max<size_t>(num_rows, 1);
    const size_t target = std::min<size_t>(std::max<size_t>(total / 16, 2u << 20), 8u << 20);
    size_t rows = std::max<size_t>(target / std::max<size_t>(bytes_per_row_in, 1), 256);
    (void)num_sms;
    return std::min(rows, std::max<size_t>(num_rows, 1));
}

// the pipelined host path shared by encode_rows / commit / commit_resident / batch_commit.  Does not synchronise.
struct HostJob {
    const uint64_t *evals;
    uint64_t *rows_out;   // host, nullable
    uint8_t *layers_out;  // host, nullable
    uint8_t *roots_out;   // host, nullable (encode only)
    bool want_roots;
    zipgpu_data **keep;   // nullable
    // row-sharded multi-GPU commit: exchange the roots with the peers (the launch that produces them carries it) and
    // copy ALL total_rows roots of the commitment to roots_all_out (host, nullable)
    FanReq *fan = nullptr;
    uint8_t *roots_all_out = nullptr;
};

// OPT-IN (ZIPGPU_ZEROCOPY=1).  Pinned host evaluations + a full-size commit that keeps its prover data on the device:
// the fused commit kernel reads the evaluation rows IN PLACE over PCIe (zero-copy; pinned memory is device-addressable
// under UVA), so the transfer streams underneath the ALU-bound hashing of the same persistent kernel -- no pipeline
// fill, no drain -- and drops a copy of every row into HBM for the opening phase.  Measured on B200 (nv = 24): 2.89 ms
// against 2.72 ms for the chunked DMA pipeline below (SM-initiated PCIe reads reach ~48 GB/s, the copy engine 55 GB/s,
// and an L2 prefetch of system memory is a no-op), so the DMA pipeline stays the default.
static bool zero_copy_eligible(zipgpu_code *code, size_t num_rows, const HostJob &job, const uint64_t **dev_alias) {
    if (!job.want_roots || job.rows_out || job.layers_out || job.fan) return false;
    if (code->fused_levels <= 0 || code->depth < code->fused_levels || num_rows < fuse_min_rows(code->ctx, code)) return false;
    if (!fusion_enabled() || !getenv("ZIPGPU_ZEROCOPY")) return false;
    if (((uintptr_t)job.evals & 31) != 0) return false;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, job.evals) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (attr.type != cudaMemoryTypeHost || !attr.devicePointer) return false;
    *dev_alias = reinterpret_cast<const uint64_t *>(attr.devicePointer);
    return true;
}

static int run_zero_copy_job(zipgpu_code *code, size_t num_rows, const HostJob &job, const uint64_t *evals_alias) {
    zipgpu_ctx *ctx = code->ctx;
    const size_t in_row_bytes = code->row_len * code->in_limbs * 8;
    const size_t out_row_bytes = code->cw * code->out_limbs * 8;
    const size_t lay_row_bytes = layers_per_row(code->depth) * 32;
    if (job.keep) *job.keep = nullptr;
    uint64_t *d_evals = nullptr, *d_rows = nullptr;
    uint8_t *d_layers = nullptr, *d_roots = nullptr;
    cudaStream_t s = ctx->stream;
    DevGuard guard(ctx, s);
    DEV_ALLOC(ctx, &d_evals, num_rows * in_row_bytes, s);
    DEV_ALLOC(ctx, &d_rows, num_rows * out_row_bytes, s);
    DEV_ALLOC(ctx, &d_layers, num_rows * lay_row_bytes, s);
    DEV_ALLOC(ctx, &d_roots, num_rows * 32, s);
    int rc = commit_dev(code, num_rows, evals_alias, d_rows, d_layers, d_roots, s, -1, nullptr, d_evals);
    if (rc) return rc;
    if (job.roots_out) CU(cudaMemcpyAsync(job.roots_out, d_roots, num_rows * 32, cudaMemcpyDeviceToHost, s));
    if (job.keep) {
        zipgpu_data *d = new (std::nothrow) zipgpu_data();
        if (!d) return fail(ZIPGPU_ERR_NOMEM, "host allocation failed");
        d->ctx = ctx;
        d->num_rows = num_rows;
        d->cw = code->cw;
        d->row_len = code->row_len;
        d->in_limbs = code->in_limbs;
        d->d_evals = d_evals;
        d->out_limbs = code->out_limbs;
        d->depth = code->depth;
        d->d_rows = d_rows;
        d->d_layers = d_layers;
        d->d_roots = d_roots;
        guard.release(d_evals);
        guard.release(d_rows);
        guard.release(d_layers);
        guard.release(d_roots);
        *job.keep = d;
    } else {
        DEV_FREE(ctx, d_evals, s);
        DEV_FREE(ctx, d_rows, s);
        DEV_FREE(ctx, d_layers, s);
        DEV_FREE(ctx, d_roots, s);
    }
    return ZIPGPU_OK;
}

static int run_host_job(zipgpu_code *code, size_t num_rows, const HostJob &job) {
    zipgpu_ctx *ctx = code->ctx;
    {
        const uint64_t *alias = nullptr;
        if (num_rows && zero_copy_eligible(code, num_rows, job, &alias)) return run_zero_copy_job(code, num_rows, job, alias);
    }
    const size_t in_row_bytes = code->row_len * code->in_limbs * 8;
    const size_t out_row_bytes = code->cw * code->out_limbs * 8;
    const bool merkle = job.want_roots;
    if (merkle && code->depth < 0)
        return fail(ZIPGPU_ERR_INVALID, "leaves.len().is_power_of_two(): codeword_len is not a power of two");
    const size_t lay_row_bytes = merkle ? layers_per_row(code->depth) * 32 : 0;
    if (job.keep) *job.keep = nullptr;
    if (num_rows == 0) {
        if (job.fan) {  // a rank without rows still takes part in the exchange
            uint8_t *d_all = nullptr;
            int rc = finish_exchange(*job.fan, nullptr, 0, ctx->stream, &d_all);
            if (rc) return rc;
            if (job.roots_all_out)
                CU(cudaMemcpyAsync(job.roots_all_out, d_all, job.fan->pr->total_rows * 32, cudaMemcpyDeviceToHost, ctx->stream));
        }
        return ZIPGPU_OK;
    }

    uint64_t *d_evals = nullptr, *d_rows = nullptr;
    uint8_t *d_layers = nullptr, *d_roots = nullptr;
    cudaStream_t s = ctx->stream;
    DevGuard guard(ctx, s);
    static const bool trace = getenv("ZIPGPU_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    DEV_ALLOC(ctx, &d_evals, num_rows * in_row_bytes, s);
    DEV_ALLOC(ctx, &d_rows, num_rows * out_row_bytes, s);
    if (merkle) {
        DEV_ALLOC(ctx, &d_layers, std::max<size_t>(num_rows * lay_row_bytes, 32), s);
        DEV_ALLOC(ctx, &d_roots, num_rows * 32, s);
    }
    // Chunk schedule: uniform chunks with a short last one.  The H2D copies run back to back and the kernels of a chunk
    // finish about one chunk-time after its copy, so what stays exposed after the last copy is the kernel work of the
    // LAST chunk -- which cannot take less than the ~50 us latency of its three launches, hence exactly one short chunk.
    const size_t chunk = pick_chunk_rows(num_rows, in_row_bytes, ctx->num_sms);
    std::vector<std::pair<size_t, size_t>> sched;
    for (size_t r0 = 0; r0 < num_rows; r0 += chunk) sched.emplace_back(r0, std::min(chunk, num_rows - r0));
    {
        size_t tail = 64;
        if (const char *env = getenv("ZIPGPU_TAIL_ROWS")) tail = (size_t)atol(env);
        if (tail > 0 && sched.size() >= 2 && sched.back().second >= 2 * tail) {
            const std::pair<size_t, size_t> last = sched.back();
            sched.pop_back();
            sched.emplace_back(last.first, last.second - tail);
            sched.emplace_back(last.first + last.second - tail, tail);
        }
    }
    // one chunk: everything on the kernel stream (no cross-stream events; a small commit is all launch latency)
    const bool single = sched.size() == 1;
    cudaStream_t h2d = single ? s : ctx->h2d, d2h = single ? s : ctx->d2h, st2 = single ? s : ctx->stream2;
    cudaError_t e;
    const double t_alloc = now();
    if ((e = chain(ctx, s, h2d)) != cudaSuccess) return cuda_fail(e, "chain");
    if ((e = chain(ctx, s, d2h)) != cudaSuccess) return cuda_fail(e, "chain");
    if ((e = chain(ctx, s, st2)) != cudaSuccess) return cuda_fail(e, "chain");
    // consecutive chunks alternate between two kernel streams: the tail of one chunk's kernels (partial waves, the
    // latency-bound narrow passes) overlaps the head of the next chunk's
    static const bool one_stream = getenv("ZIPGPU_ONE_STREAM") != nullptr;
    cudaStream_t ks[2] = {s, one_stream ? s : st2};
    size_t chunk_no = 0;

    // Per chunk the trees are taken to the first pass boundary >= level 6 (the wide passes); the rest is deferred.
    // ... unless the fused launch of a chunk finishes its trees anyway (commit_ws_kernel's tops epilogue, cw 4096 / 8192):
    // then nothing is left after the last chunk but its own launch (nv = 24 e2e 2.70 -> 2.67 ms, nv = 22 0.777 -> 0.753)
    const bool chunk_whole_trees = code->in_limbs == 1 && code->out_limbs == 4 && !code->sparse && !code->big &&
                                   fusion_enabled() && chunk >= fuse_min_rows(ctx, code) &&
                                   commit_ws_whole_trees((uint32_t)code->cw, (uint32_t)chunk);
    static const bool force_defer = getenv("ZIPGPU_DEFER_TOP") != nullptr;  // A/B knob
    const bool defer_top = merkle && !job.layers_out && sched.size() > 1 && code->depth > 7 &&
                           (!chunk_whole_trees || force_defer);
    int split_level = -1, min_split = 1 << 30;  // the deferred top passes start at the LOWEST level any chunk stopped at
    // ZIPGPU_TIMELINE=1: timing events after every copy / chunk, printed relative to the first (diagnostics only)
    static const bool timeline = getenv("ZIPGPU_TIMELINE") != nullptr;
    std::vector<cudaEvent_t> tl_copy, tl_kern;
    cudaEvent_t tl0 = nullptr;
    if (timeline) {
        cudaEventCreate(&tl0);
        cudaEventRecord(tl0, h2d);
    }
    for (const auto &ch : sched) {
        const size_t r0 = ch.first, n = ch.second;
        CU(cudaMemcpyAsync((uint8_t *)d_evals + r0 * in_row_bytes, (const uint8_t *)job.evals + r0 * in_row_bytes,
                           n * in_row_bytes, cudaMemcpyHostToDevice, h2d));
        if (timeline) {
            cudaEvent_t ev;
            cudaEventCreate(&ev);
            cudaEventRecord(ev, h2d);
            tl_copy.push_back(ev);
        }
        cudaStream_t k = ks[chunk_no++ & 1];
        if ((e = chain(ctx, h2d, k)) != cudaSuccess) return cuda_fail(e, "chain");
        int rc = commit_dev(code, n, (const uint64_t *)((uint8_t *)d_evals + r0 * in_row_bytes),
                            (uint64_t *)((uint8_t *)d_rows + r0 * out_row_bytes),
                            merkle ? d_layers + r0 * lay_row_bytes : nullptr, merkle ? d_roots + r0 * 32 : nullptr, k,
                            defer_top ? 6 : -1, &split_level, nullptr, single ? job.fan : nullptr);
        if (rc) return rc;
        min_split = std::min(min_split, split_level);
        if (timeline) {
            cudaEvent_t ev;
            cudaEventCreate(&ev);
            cudaEventRecord(ev, k);
            tl_kern.push_back(ev);
        }
        if (job.rows_out || job.layers_out) {
            if ((e = chain(ctx, k, d2h)) != cudaSuccess) return cuda_fail(e, "chain");
            if (job.rows_out)
                CU(cudaMemcpyAsync((uint8_t *)job.rows_out + r0 * out_row_bytes, (uint8_t *)d_rows + r0 * out_row_bytes,
                                   n * out_row_bytes, cudaMemcpyDeviceToHost, d2h));
            if (job.layers_out && lay_row_bytes)
                CU(cudaMemcpyAsync(job.layers_out + r0 * lay_row_bytes, d_layers + r0 * lay_row_bytes,
                                   n * lay_row_bytes, cudaMemcpyDeviceToHost, d2h));
        }
    }
    if ((e = chain(ctx, st2, s)) != cudaSuccess) return cuda_fail(e, "chain");
    if (defer_top && min_split < code->depth) {
        int rc = merkle_top_dev(code, num_rows, d_rows, d_layers, d_roots, s, min_split, job.fan);
        if (rc) return rc;
    }
    uint8_t *d_all_roots = nullptr, *d_roots_all_keep = nullptr;
    if (merkle && job.fan) {  // (a no-op when one of the launches above carried the exchange)
        int rc = finish_exchange(*job.fan, d_roots, num_rows, s, &d_all_roots);
        if (rc) return rc;
        if (job.keep) {  // the exchange buffer is recycled two steps later: the handle keeps its own copy
            DEV_ALLOC(ctx, &d_roots_all_keep, job.fan->pr->total_rows * 32, s);
            CU(cudaMemcpyAsync(d_roots_all_keep, d_all_roots, job.fan->pr->total_rows * 32, cudaMemcpyDeviceToDevice, s));
        }
    }
    if (timeline) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, s);
        cudaEventSynchronize(ev);
        float ms = 0;
        fprintf(stderr, "[zipgpu timeline] chunk: rows, copy done (ms), kernels done (ms)\n");
        for (size_t i = 0; i < sched.size(); i++) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, tl0, tl_copy[i]);
            cudaEventElapsedTime(&b, tl0, tl_kern[i]);
            fprintf(stderr, "  %2zu: %5zu  %7.3f  %7.3f\n", i, sched[i].second, a, b);
            cudaEventDestroy(tl_copy[i]);
            cudaEventDestroy(tl_kern[i]);
        }
        cudaEventElapsedTime(&ms, tl0, ev);
        fprintf(stderr, "  all kernels done %7.3f ms\n", ms);
        cudaEventDestroy(ev);
        cudaEventDestroy(tl0);
    }
    if (merkle && job.roots_out) {
        if ((e = chain(ctx, s, d2h)) != cudaSuccess) return cuda_fail(e, "chain");
        CU(cudaMemcpyAsync(job.roots_out, d_roots, num_rows * 32, cudaMemcpyDeviceToHost, d2h));
    }
    if (merkle && job.fan && job.roots_all_out) {
        if ((e = chain(ctx, s, d2h)) != cudaSuccess) return cuda_fail(e, "chain");
        CU(cudaMemcpyAsync(job.roots_all_out, d_all_roots, job.fan->pr->total_rows * 32, cudaMemcpyDeviceToHost, d2h));
    }
    // join the copy streams back into the kernel stream so the frees are ordered after every use
    if ((e = chain(ctx, d2h, s)) != cudaSuccess) return cuda_fail(e, "chain");
    if ((e = chain(ctx, h2d, s)) != cudaSuccess) return cuda_fail(e, "chain");
    const double t_enq = now();
    if (!job.keep) DEV_FREE(ctx, d_evals, s);
    if (trace)
        fprintf(stderr, "[zipgpu] host job: %zu rows x %zu, %zu chunks: alloc %.3f ms, enqueue %.3f ms\n", num_rows,
                (size_t)code->row_len, sched.size(), t_alloc - t_begin, t_enq - t_alloc);
    if (job.keep) {
        zipgpu_data *d = new (std::nothrow) zipgpu_data();
        if (!d) return fail(ZIPGPU_ERR_NOMEM, "host allocation failed");
        d->ctx = ctx;
        d->num_rows = num_rows;
        d->cw = code->cw;
        d->row_len = code->row_len;
        d->in_limbs = code->in_limbs;
        d->d_evals = d_evals;
        d->out_limbs = code->out_limbs;
        d->depth = code->depth;
        d->d_rows = d_rows;
        d->d_layers = d_layers;
        d->d_roots = d_roots;
        d->d_roots_all = d_roots_all_keep;
        d->total_rows = job.fan ? job.fan->pr->total_rows : num_rows;
        guard.release(d_evals);
        guard.release(d_rows);
        guard.release(d_layers);
        guard.release(d_roots);
        if (d_roots_all_keep) guard.release(d_roots_all_keep);
        *job.keep = d;
    } else {
        DEV_FREE(ctx, d_rows, s);
        if (d_layers) DEV_FREE(ctx, d_layers, s);
        if (d_roots) DEV_FREE(ctx, d_roots, s);
    }
    return ZIPGPU_OK;
}

extern "C" int zipgpu_encode_rows(zipgpu_code *code, size_t num_rows, const uint64_t *evals, uint64_t *rows_out) {
    if (!code || (num_rows && (!evals || !rows_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    API_LOCK(code->ctx);
    CU(cudaSetDevice(code->ctx->device));
    HostJob job{evals, rows_out, nullptr, nullptr, false, nullptr};
    int rc = run_host_job(code, num_rows, job);
    int rc2 = sync_job(code->ctx);
    return rc ? rc : rc2;
}

// the latency encoder of encode_f.cu on host buffers: in_limbs == 0: field elements of `limbs` limbs modulo `modulus`;
// in_limbs > 0: Int<in_limbs> entries widened to Int<limbs>, wrap-around adds
static int encode_rows_generic(zipgpu_code *code, size_t num_rows, int limbs, int in_limbs, const uint64_t *modulus,
                               const uint64_t *rows, uint64_t *out) {
    if (num_rows == 0) return ZIPGPU_OK;
    zipgpu_ctx *ctx = code->ctx;
    API_LOCK(ctx);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    if (!code->d_perm1) {  // first use: the gather form of the permutations, as uploaded
        cudaError_t e;
        if ((e = cudaMalloc(&code->d_perm1, code->cw * 4)) != cudaSuccess || (e = cudaMalloc(&code->d_perm2, code->cw * 4)) != cudaSuccess ||
            (e = cudaMemcpy(code->d_perm1, code->h_perm1.data(), code->cw * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaMemcpy(code->d_perm2, code->h_perm2.data(), code->cw * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
            cudaFree(code->d_perm1);
            cudaFree(code->d_perm2);
            code->d_perm1 = code->d_perm2 = nullptr;
            return cuda_fail(e, "cudaMalloc/cudaMemcpy(permutations)");
        }
    }
    DevGuard guard(ctx, s);
    const size_t in_bytes = num_rows * code->row_len * (in_limbs ? in_limbs : limbs) * 8;
    const size_t out_bytes = num_rows * code->cw * limbs * 8;
    uint32_t *d_in = nullptr, *d_out = nullptr, *d_scr = nullptr, *d_mod = nullptr;
    DEV_ALLOC(ctx, &d_in, in_bytes, s);
    DEV_ALLOC(ctx, &d_out, out_bytes, s);
    DEV_ALLOC(ctx, &d_scr, out_bytes, s);
    DEV_ALLOC(ctx, &d_mod, 64, s);
    CU(cudaMemcpyAsync(d_in, rows, in_bytes, cudaMemcpyHostToDevice, s));
    if (modulus) CU(cudaMemcpyAsync(d_mod, modulus, (size_t)limbs * 8, cudaMemcpyHostToDevice, s));
    EncodeFArgs a;
    a.rows_in = d_in;
    a.out = d_out;
    a.perm1 = code->d_perm1;
    a.perm2 = code->d_perm2;
    a.modulus = d_mod;
    a.scratch = d_scr;
    a.num_rows = (uint32_t)num_rows;
    a.row_len = (uint32_t)code->row_len;
    a.cw = (uint32_t)code->cw;
    a.limbs = limbs;
    a.in_limbs = in_limbs;
    a.stream = s;
    cudaError_t e = launch_encode_f(a);
    if (e != cudaSuccess) return cuda_fail(e, "launch_encode_f");
    ctx->launches++;
    CU(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, s));
    DEV_FREE(ctx, d_in, s);
    DEV_FREE(ctx, d_out, s);
    DEV_FREE(ctx, d_scr, s);
    DEV_FREE(ctx, d_mod, s);
    CU(cudaStreamSynchronize(s));
    return ZIPGPU_OK;
}

// ------------------------------------------------------------------------------------------------------
// encode_f (code_raa.rs:133-138): the RAA code over field elements
// ------------------------------------------------------------------------------------------------------
extern "C" int zipgpu_encode_f(zipgpu_code *code, size_t num_rows, int limbs, const uint64_t *modulus, const uint64_t *rows,
                               uint64_t *out) {
    if (!code || !modulus || (num_rows && (!rows || !out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (code->sparse) return fail(ZIPGPU_ERR_UNSUPPORTED, "encode_f is implemented for the RAA code");
    if (limbs < 1 || limbs > 6) return fail(ZIPGPU_ERR_UNSUPPORTED, "field elements of 1..6 u64 limbs");
    if (num_rows > 0x7fffffffull) return fail(ZIPGPU_ERR_UNSUPPORTED, "too many rows");
    bool nonzero = false;
    for (int l = 0; l < limbs; l++) nonzero |= modulus[l] != 0;
    if (!nonzero) return fail(ZIPGPU_ERR_INVALID, "modulus is zero");
    // the parallel scan equals the reference's sequential accumulate because addition of residues < modulus is
    // associative: insist on reduced inputs (the reference's RandomField values always are)
    for (size_t i = 0; i < num_rows * code->row_len; i++) {
        int cmp = 0;
        for (int l = limbs - 1; l >= 0 && cmp == 0; l--)
            cmp = rows[i * limbs + l] < modulus[l] ? -1 : rows[i * limbs + l] > modulus[l] ? 1 : 0;
        if (cmp >= 0) return fail(ZIPGPU_ERR_INVALID, "encode_f: element " + std::to_string(i) + " is not reduced modulo the field modulus");
    }
    return encode_rows_generic(code, num_rows, limbs, 0, modulus, rows, out);
}

// encode_wide (code_raa.rs:125-131) for any In = Int<in_limbs>, Out = Int<out_limbs>: what the verifier runs on the
// combined row of a proximity test with In = Out = M (verify_z.rs:74-78)
extern "C" int zipgpu_encode_wide(zipgpu_code *code, size_t num_rows, int in_limbs, int out_limbs, const uint64_t *rows,
                                  uint64_t *out) {
    if (!code || (num_rows && (!rows || !out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (code->sparse) return fail(ZIPGPU_ERR_UNSUPPORTED, "encode_wide is implemented for the RAA code");
    if (in_limbs < 1 || out_limbs < in_limbs || out_limbs > 8)
        return fail(ZIPGPU_ERR_INVALID, "need 1 <= in_limbs <= out_limbs <= 8");
    if (num_rows > 0x7fffffffull) return fail(ZIPGPU_ERR_UNSUPPORTED, "too many rows");
    return encode_rows_generic(code, num_rows, out_limbs, in_limbs, nullptr, rows, out);
}

// ------------------------------------------------------------------------------------------------------
// merkle_rows
// ------------------------------------------------------------------------------------------------------
extern "C" int zipgpu_merkle_rows_device(zipgpu_ctx *ctx, size_t num_rows, int depth, int leaf_limbs,
                                         const uint64_t *d_leaves, uint8_t *d_layers_out, uint8_t *d_roots_out,
                                         void *stream) {
    if (!ctx || (num_rows && (!d_leaves || !d_roots_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    API_LOCK(ctx);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    DevGuard guard(ctx, s);
    int rc;
    if ((rc = check_align16(d_leaves, "leaves")) || (rc = check_align16(d_roots_out, "roots_out")) ||
        (rc = check_align16(d_layers_out, "layers_out")))
        return rc;
    uint8_t *scratch = nullptr;
    if (!d_layers_out && depth > 0) {
        DEV_ALLOC(ctx, &scratch, num_rows * layers_per_row(depth) * 32, s);
        d_layers_out = scratch;
    }
    ProfRec r;
    const bool prof = prof_begin(ctx, &r);
    if (prof) cudaEventRecord(r.e1, s);
    rc = merkle_dev(ctx, num_rows, depth, leaf_limbs, d_leaves, d_layers_out, d_roots_out, s);
    if (prof && rc == 0) {
        cudaEventRecord(r.e2, s);
        r.has_hash = true;
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->prof_pending.push_back(r);
    }
    if (scratch) DEV_FREE(ctx, scratch, s);
    return rc;
}

extern "C" int zipgpu_merkle_rows(zipgpu_ctx *ctx, size_t num_rows, int depth, int leaf_limbs, const uint64_t *leaves,
                                  uint8_t *layers_out, uint8_t *roots_out) {
    if (!ctx || (num_rows && (!leaves || !roots_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (depth < 0 || depth > 30) return fail(ZIPGPU_ERR_INVALID, "depth out of range");
    if (num_rows == 0) return ZIPGPU_OK;
    API_LOCK(ctx);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    DevGuard guard(ctx, s);
    const size_t nleaves = num_rows << depth;
    const size_t leaf_bytes = nleaves * leaf_limbs * 8;
    const size_t lay_bytes = num_rows * layers_per_row(depth) * 32;
    uint64_t *d_leaves = nullptr;
    uint8_t *d_layers = nullptr, *d_roots = nullptr;
    DEV_ALLOC(ctx, &d_leaves, leaf_bytes, s);
    DEV_ALLOC(ctx, &d_layers, std::max<size_t>(lay_bytes, 32), s);
    DEV_ALLOC(ctx, &d_roots, num_rows * 32, s);
    CU(cudaMemcpyAsync(d_leaves, leaves, leaf_bytes, cudaMemcpyHostToDevice, s));
    int rc = zipgpu_merkle_rows_device(ctx, num_rows, depth, leaf_limbs, d_leaves, d_layers, d_roots, s);
    if (rc == 0) {
        if (layers_out && lay_bytes) CU(cudaMemcpyAsync(layers_out, d_layers, lay_bytes, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(roots_out, d_roots, num_rows * 32, cudaMemcpyDeviceToHost, s));
    }
    DEV_FREE(ctx, d_leaves, s);
    DEV_FREE(ctx, d_layers, s);
    DEV_FREE(ctx, d_roots, s);
    CU(cudaStreamSynchronize(s));
    return rc;
}

// ------------------------------------------------------------------------------------------------------
// commit
// ------------------------------------------------------------------------------------------------------
extern "C" int zipgpu_commit_device(zipgpu_code *code, size_t num_rows, const uint64_t *d_evals, uint64_t *d_rows_out,
                                    uint8_t *d_layers_out, uint8_t *d_roots_out, void *stream) {
    if (!code || (num_rows && (!d_evals || !d_roots_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (code->depth < 0)
        return fail(ZIPGPU_ERR_INVALID, "leaves.len().is_power_of_two(): codeword_len is not a power of two");
    if (num_rows == 0) return ZIPGPU_OK;
    zipgpu_ctx *ctx = code->ctx;
    API_LOCK(ctx);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    DevGuard guard(ctx, s);
    int rc;
    if ((rc = check_align16(d_roots_out, "roots_out")) || (rc = check_align16(d_layers_out, "layers_out"))) return rc;
    uint64_t *rows_scratch = nullptr;
    uint8_t *layers_scratch = nullptr;
    if (!d_rows_out) {
        DEV_ALLOC(ctx, &rows_scratch, num_rows * code->cw * code->out_limbs * 8, s);
        d_rows_out = rows_scratch;
    }
    if (!d_layers_out && code->depth > 0) {
        DEV_ALLOC(ctx, &layers_scratch, num_rows * layers_per_row(code->depth) * 32, s);
        d_layers_out = layers_scratch;
    }
    rc = commit_dev(code, num_rows, d_evals, d_rows_out, d_layers_out, d_roots_out, s);
    if (rows_scratch) DEV_FREE(ctx, rows_scratch, s);
    if (layers_scratch) DEV_FREE(ctx, layers_scratch, s);
    return rc;
}

extern "C" int zipgpu_commit(zipgpu_code *code, size_t num_rows, const uint64_t *evals, uint64_t *rows_out,
                             uint8_t *layers_out, uint8_t *roots_out) {
    if (!code || (num_rows && (!evals || !roots_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    API_LOCK(code->ctx);
    CU(cudaSetDevice(code->ctx->device));
    HostJob job{evals, rows_out, layers_out, roots_out, true, nullptr};
    int rc = run_host_job(code, num_rows, job);
    int rc2 = sync_job(code->ctx);
    return rc ? rc : rc2;
}

// batch_commit as ONE matrix: rows are independent and the polynomials share the code, so the batch is the commit of a
// (num_polys * num_rows)-row matrix whose row blocks arrive from different host pointers.  Groups of whole polynomials
// (~8 MiB of evaluations) form the chunks of the same pipeline as run_host_job: H2D per polynomial, kernels per group
// on alternating streams, the narrow tree passes once over the whole batch, outputs copied back per polynomial.
static int run_batch_job(zipgpu_code *code, size_t num_polys, size_t num_rows, const uint64_t *const *evals,
                         uint64_t *const *rows_out, uint8_t *const *layers_out, uint8_t *const *roots_out) {
    zipgpu_ctx *ctx = code->ctx;
    const size_t in_row_bytes = code->row_len * code->in_limbs * 8;
    const size_t out_row_bytes = code->cw * code->out_limbs * 8;
    if (code->depth < 0)
        return fail(ZIPGPU_ERR_INVALID, "leaves.len().is_power_of_two(): codeword_len is not a power of two");
    const size_t lay_row_bytes = layers_per_row(code->depth) * 32;
    const size_t total_rows = num_polys * num_rows;
    const size_t poly_in = num_rows * in_row_bytes, poly_out = num_rows * out_row_bytes, poly_lay = num_rows * lay_row_bytes;
    bool want_rows = false, want_layers = false;
    for (size_t p = 0; p < num_polys; p++) {
        want_rows |= rows_out && rows_out[p];
        want_layers |= layers_out && layers_out[p];
    }
    uint64_t *d_evals = nullptr, *d_rows = nullptr;
    uint8_t *d_layers = nullptr, *d_roots = nullptr;
    cudaStream_t s = ctx->stream;
    DevGuard guard(ctx, s);
    DEV_ALLOC(ctx, &d_evals, total_rows * in_row_bytes, s);
    DEV_ALLOC(ctx, &d_rows, total_rows * out_row_bytes, s);
    DEV_ALLOC(ctx, &d_layers, std::max<size_t>(total_rows * lay_row_bytes, 32), s);
    DEV_ALLOC(ctx, &d_roots, total_rows * 32, s);
    cudaError_t e;
    if ((e = chain(ctx, s, ctx->h2d)) != cudaSuccess) return cuda_fail(e, "chain");
    if ((e = chain(ctx, s, ctx->d2h)) != cudaSuccess) return cuda_fail(e, "chain");
    if ((e = chain(ctx, s, ctx->stream2)) != cudaSuccess) return cuda_fail(e, "chain");
    cudaStream_t ks[2] = {s, ctx->stream2};
    const size_t group = std::max<size_t>(1, (8u << 20) / std::max<size_t>(poly_in, 1));
    const bool defer_top = !want_layers && num_polys > group && code->depth > 7;
    int split_level = -1, min_split = 1 << 30;
    size_t chunk_no = 0;
    for (size_t p0 = 0; p0 < num_polys; p0 += group) {
        const size_t np = std::min(group, num_polys - p0);
        for (size_t p = p0; p < p0 + np; p++)
            CU(cudaMemcpyAsync((uint8_t *)d_evals + p * poly_in, evals[p], poly_in, cudaMemcpyHostToDevice, ctx->h2d));
        cudaStream_t k = ks[chunk_no++ & 1];
        if ((e = chain(ctx, ctx->h2d, k)) != cudaSuccess) return cuda_fail(e, "chain");
        const size_t r0 = p0 * num_rows;
        int rc = commit_dev(code, np * num_rows, (const uint64_t *)((uint8_t *)d_evals + r0 * in_row_bytes),
                            (uint64_t *)((uint8_t *)d_rows + r0 * out_row_bytes), d_layers + r0 * lay_row_bytes,
                            d_roots + r0 * 32, k, defer_top ? 6 : -1, &split_level);
        if (rc) return rc;
        min_split = std::min(min_split, split_level);
        if (want_rows || want_layers) {
            if ((e = chain(ctx, k, ctx->d2h)) != cudaSuccess) return cuda_fail(e, "chain");
            for (size_t p = p0; p < p0 + np; p++) {
                if (rows_out && rows_out[p])
                    CU(cudaMemcpyAsync(rows_out[p], (uint8_t *)d_rows + p * poly_out, poly_out, cudaMemcpyDeviceToHost, ctx->d2h));
                if (layers_out && layers_out[p] && poly_lay)
                    CU(cudaMemcpyAsync(layers_out[p], d_layers + p * poly_lay, poly_lay, cudaMemcpyDeviceToHost, ctx->d2h));
            }
        }
    }
    if ((e = chain(ctx, ctx->stream2, s)) != cudaSuccess) return cuda_fail(e, "chain");
    if (defer_top && min_split < code->depth) {
        int rc = merkle_top_dev(code, total_rows, d_rows, d_layers, d_roots, s, min_split);
        if (rc) return rc;
    }
    if ((e = chain(ctx, s, ctx->d2h)) != cudaSuccess) return cuda_fail(e, "chain");
    for (size_t p = 0; p < num_polys; p++)
        CU(cudaMemcpyAsync(roots_out[p], d_roots + p * num_rows * 32, num_rows * 32, cudaMemcpyDeviceToHost, ctx->d2h));
    if ((e = chain(ctx, ctx->d2h, s)) != cudaSuccess) return cuda_fail(e, "chain");
    if ((e = chain(ctx, ctx->h2d, s)) != cudaSuccess) return cuda_fail(e, "chain");
    DEV_FREE(ctx, d_evals, s);
    DEV_FREE(ctx, d_rows, s);
    DEV_FREE(ctx, d_layers, s);
    DEV_FREE(ctx, d_roots, s);
    return ZIPGPU_OK;
}

extern "C" int zipgpu_batch_commit(zipgpu_code *code, size_t num_polys, size_t num_rows, const uint64_t *const *evals,
                                   uint64_t *const *rows_out, uint8_t *const *layers_out, uint8_t *const *roots_out) {
    if (!code || (num_polys && (!evals || !roots_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    API_LOCK(code->ctx);
    CU(cudaSetDevice(code->ctx->device));
    for (size_t p = 0; p < num_polys; p++)
        if (num_rows && (!evals[p] || !roots_out[p]))
            return fail(ZIPGPU_ERR_INVALID, "NULL polynomial or roots pointer in batch");
    if (num_polys == 0 || num_rows == 0) return ZIPGPU_OK;
    int rc = ZIPGPU_OK;
    // the whole batch as one matrix while its prover data fits comfortably (rows + layers: 192 B per evaluation);
    // beyond that, polynomial by polynomial (still pipelined: poly p+1's H2D overlaps poly p's kernels)
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const size_t need = num_polys * num_rows * (code->row_len * code->in_limbs * 8 + code->cw * code->out_limbs * 8 +
                                                (code->depth >= 0 ? layers_per_row(code->depth) * 32 : 0) + 32);
    if (code->depth >= 0 && need < total_b / 2) {
        rc = run_batch_job(code, num_polys, num_rows, evals, rows_out, layers_out, roots_out);
    } else {
        for (size_t p = 0; p < num_polys && rc == 0; p++) {
            HostJob job{evals[p], rows_out ? rows_out[p] : nullptr, layers_out ? layers_out[p] : nullptr, roots_out[p], true,
                        nullptr};
            rc = run_host_job(code, num_rows, job);
        }
    }
    int rc2 = sync_job(code->ctx);
    return rc ? rc : rc2;
}

// ------------------------------------------------------------------------------------------------------
// device-resident prover data
// ------------------------------------------------------------------------------------------------------
extern "C" int zipgpu_commit_resident(zipgpu_code *code, size_t num_rows, const uint64_t *evals, uint8_t *roots_out,
                                      zipgpu_data **handle) {
    if (!code || !handle || (num_rows && (!evals || !roots_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    API_LOCK(code->ctx);
    CU(cudaSetDevice(code->ctx->device));
    HostJob job{evals, nullptr, nullptr, roots_out, true, handle};
    int rc = run_host_job(code, num_rows, job);
    int rc2 = sync_job(code->ctx);
    return rc ? rc : rc2;
}

// ------------------------------------------------------------------------------------------------------
// row-sharded commit: this GPU's row range of ONE commitment, roots exchanged with the peers in the same launch
// ------------------------------------------------------------------------------------------------------
static int check_shard(const zipgpu_code *code, const zipgpu_peer_roots *pr, size_t row_begin, size_t count) {
    if (!pr->connected) return fail(ZIPGPU_ERR_INVALID, "zipgpu_peer_roots_connect has not been called");
    if (pr->ctx != code->ctx) return fail(ZIPGPU_ERR_INVALID, "code and peer_roots belong to different contexts");
    if (row_begin + count > pr->total_rows) return fail(ZIPGPU_ERR_INVALID, "row range outside the commitment");
    if (code->depth < 0)
        return fail(ZIPGPU_ERR_INVALID, "leaves.len().is_power_of_two(): codeword_len is not a power of two");
    return ZIPGPU_OK;
}

extern "C" int zipgpu_commit_device_sharded(zipgpu_code *code, zipgpu_peer_roots *pr, size_t row_begin, size_t count,
                                            const uint64_t *d_evals, uint64_t *d_rows_out, uint8_t *d_layers_out,
                                            void *stream, uint8_t **d_all_roots_out) {
    if (!code || !pr || (count && !d_evals)) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    int rc = check_shard(code, pr, row_begin, count);
    if (rc) return rc;
    zipgpu_ctx *ctx = code->ctx;
    API_LOCK(ctx);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    DevGuard guard(ctx, s);
    if ((rc = check_align16(d_layers_out, "layers_out"))) return rc;
    FanReq fan{pr, row_begin};
    uint64_t *rows_scratch = nullptr;
    uint8_t *layers_scratch = nullptr, *d_roots = nullptr;
    if (count) {
        if (!d_rows_out) {
            DEV_ALLOC(ctx, &rows_scratch, count * code->cw * code->out_limbs * 8, s);
            d_rows_out = rows_scratch;
        }
        if (!d_layers_out && code->depth > 0) {
            DEV_ALLOC(ctx, &layers_scratch, count * layers_per_row(code->depth) * 32, s);
            d_layers_out = layers_scratch;
        }
        DEV_ALLOC(ctx, &d_roots, count * 32, s);
        rc = commit_dev(code, count, d_evals, d_rows_out, d_layers_out, d_roots, s, -1, nullptr, nullptr, &fan);
        if (rc) return rc;
    }
    rc = finish_exchange(fan, d_roots, count, s, d_all_roots_out);
    if (rows_scratch) DEV_FREE(ctx, rows_scratch, s);
    if (layers_scratch) DEV_FREE(ctx, layers_scratch, s);
    if (d_roots) DEV_FREE(ctx, d_roots, s);
    return rc;
}

extern "C" int zipgpu_commit_resident_sharded(zipgpu_code *code, zipgpu_peer_roots *pr, size_t row_begin, size_t count,
                                              const uint64_t *evals, uint8_t *roots_all_out, zipgpu_data **handle) {
    if (!code || !pr || (count && !evals)) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    int rc = check_shard(code, pr, row_begin, count);
    if (rc) return rc;
    API_LOCK(code->ctx);
    CU(cudaSetDevice(code->ctx->device));
    FanReq fan{pr, row_begin};
    HostJob job{evals, nullptr, nullptr, nullptr, true, handle, &fan, roots_all_out};
    rc = run_host_job(code, count, job);
    int rc2 = sync_job(code->ctx);
    if (rc == 0 && rc2 == 0) rc = zipgpu_peer_roots_status(pr);
    return rc ? rc : rc2;
}

extern "C" void zipgpu_data_free(zipgpu_data *d) {
    if (!d) return;
    API_LOCK(d->ctx);
    cudaSetDevice(d->ctx->device);
    cudaStream_t s = d->ctx->stream;
    dev_free(d->ctx, d->d_evals, s);
    dev_free(d->ctx, d->d_rows, s);
    dev_free(d->ctx, d->d_layers, s);
    dev_free(d->ctx, d->d_roots, s);
    dev_free(d->ctx, d->d_roots_all, s);
    delete d;
}
extern "C" size_t zipgpu_data_num_rows(const zipgpu_data *d) { return d ? d->num_rows : 0; }
extern "C" const uint64_t *zipgpu_data_rows_device(const zipgpu_data *d) { return d ? d->d_rows : nullptr; }
extern "C" const uint8_t *zipgpu_data_layers_device(const zipgpu_data *d) { return d ? d->d_layers : nullptr; }
extern "C" const uint8_t *zipgpu_data_roots_device(const zipgpu_data *d) { return d ? d->d_roots : nullptr; }
extern "C" const uint8_t *zipgpu_data_all_roots_device(const zipgpu_data *d) { return d ? d->d_roots_all : nullptr; }

extern "C" int zipgpu_data_read_rows(const zipgpu_data *d, size_t row_begin, size_t row_count, uint64_t *rows_out) {
    if (!d || (row_count && !rows_out)) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (row_begin + row_count > d->num_rows) return fail(ZIPGPU_ERR_INVALID, "row range out of bounds");
    API_LOCK(d->ctx);
    CU(cudaSetDevice(d->ctx->device));
    const size_t rb = d->cw * d->out_limbs * 8;
    CU(cudaMemcpyAsync(rows_out, (const uint8_t *)d->d_rows + row_begin * rb, row_count * rb, cudaMemcpyDeviceToHost,
                       d->ctx->stream));
    CU(cudaStreamSynchronize(d->ctx->stream));
    return ZIPGPU_OK;
}
extern "C" int zipgpu_data_read_layers(const zipgpu_data *d, size_t row_begin, size_t row_count, uint8_t *layers_out) {
    if (!d || (row_count && !layers_out)) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (row_begin + row_count > d->num_rows) return fail(ZIPGPU_ERR_INVALID, "row range out of bounds");
    API_LOCK(d->ctx);
    CU(cudaSetDevice(d->ctx->device));
    const size_t lb = layers_per_row(d->depth) * 32;
    if (lb != 0 && row_count != 0) {
        CU(cudaMemcpyAsync(layers_out, d->d_layers + row_begin * lb, row_count * lb, cudaMemcpyDeviceToHost,
                           d->ctx->stream));
        CU(cudaStreamSynchronize(d->ctx->stream));
    }
    return ZIPGPU_OK;
}

extern "C" int zipgpu_data_open_columns_strided(const zipgpu_data *d, size_t num_cols, const uint32_t *columns,
                                                uint64_t *col_values_out, uint8_t *paths_out, size_t total_rows,
                                                size_t row_offset) {
    if (!d || (num_cols && (!columns || !col_values_out || (!paths_out && d->depth > 0))))
        return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (row_offset + d->num_rows > total_rows) return fail(ZIPGPU_ERR_INVALID, "shard rows outside the commitment");
    if (num_cols == 0) return ZIPGPU_OK;
    for (size_t i = 0; i < num_cols; i++)
        if (columns[i] >= d->cw) return fail(ZIPGPU_ERR_INVALID, "column index out of range");
    zipgpu_ctx *ctx = d->ctx;
    API_LOCK(ctx);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    DevGuard guard(ctx, s);
    const size_t val_bytes = num_cols * d->num_rows * d->out_limbs * 8;
    const size_t path_bytes = num_cols * d->num_rows * (size_t)d->depth * 32;
    uint32_t *d_cols = nullptr, *d_vals = nullptr;
    uint8_t *d_paths = nullptr;
    DEV_ALLOC(ctx, &d_cols, num_cols * 4, s);
    DEV_ALLOC(ctx, &d_vals, val_bytes, s);
    DEV_ALLOC(ctx, &d_paths, std::max<size_t>(path_bytes, 32), s);
    CU(cudaMemcpyAsync(d_cols, columns, num_cols * 4, cudaMemcpyHostToDevice, s));
    OpenArgs a;
    a.rows = reinterpret_cast<const uint32_t *>(d->d_rows);
    a.layers = d->d_layers;
    a.columns = d_cols;
    a.col_values = d_vals;
    a.paths = d_paths;
    a.num_rows = (uint32_t)d->num_rows;
    a.cw = (uint32_t)d->cw;
    a.out32 = (uint32_t)d->out_limbs * 2;
    a.num_cols = (uint32_t)num_cols;
    a.depth = d->depth;
    a.stream = s;
    cudaError_t e = launch_open_columns(a);
    if (e != cudaSuccess) return cuda_fail(e, "launch_open_columns");
    ctx->launches++;
    // per column the shard's rows land at rows [row_offset, row_offset + num_rows) of a total_rows-row block
    const size_t vrow = (size_t)d->out_limbs * 8, prow = (size_t)d->depth * 32;
    CU(cudaMemcpy2DAsync((uint8_t *)col_values_out + row_offset * vrow, total_rows * vrow, d_vals, d->num_rows * vrow,
                         d->num_rows * vrow, num_cols, cudaMemcpyDeviceToHost, s));
    if (path_bytes)
        CU(cudaMemcpy2DAsync(paths_out + row_offset * prow, total_rows * prow, d_paths, d->num_rows * prow,
                             d->num_rows * prow, num_cols, cudaMemcpyDeviceToHost, s));
    DEV_FREE(ctx, d_cols, s);
    DEV_FREE(ctx, d_vals, s);
    DEV_FREE(ctx, d_paths, s);
    CU(cudaStreamSynchronize(s));
    return ZIPGPU_OK;
}

extern "C" int zipgpu_data_open_columns(const zipgpu_data *d, size_t num_cols, const uint32_t *columns,
                                        uint64_t *col_values_out, uint8_t *paths_out) {
    return zipgpu_data_open_columns_strided(d, num_cols, columns, col_values_out, paths_out, d ? d->num_rows : 0, 0);
}

extern "C" size_t zipgpu_data_open_columns_wire_bytes(const zipgpu_data *d) {
    return d ? open_columns_wire_bytes((uint32_t)d->num_rows, (uint32_t)d->out_limbs, d->depth) : 0;
}

namespace zipgpu {
int data_open_columns_wire_strided(const zipgpu_data *d, size_t num_cols, const uint32_t *columns, uint8_t *stream_out,
                                   size_t total_rows, size_t row_offset);
}
extern "C" int zipgpu_data_open_columns_wire(const zipgpu_data *d, size_t num_cols, const uint32_t *columns,
                                             uint8_t *stream_out) {
    return zipgpu::data_open_columns_wire_strided(d, num_cols, columns, stream_out, d ? d->num_rows : 0, 0);
}

// a shard's part of the proof stream of a total_rows-row commitment (mgpu.cu): per column the values block and the
// proofs block of the shard's rows go to their places inside the column's [values | proofs] record
int zipgpu::data_open_columns_wire_strided(const zipgpu_data *d, size_t num_cols, const uint32_t *columns,
                                           uint8_t *stream_out, size_t total_rows, size_t row_offset) {
    if (!d || (num_cols && (!columns || !stream_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (row_offset + d->num_rows > total_rows) return fail(ZIPGPU_ERR_INVALID, "shard rows outside the commitment");
    if (num_cols == 0) return ZIPGPU_OK;
    for (size_t i = 0; i < num_cols; i++)
        if (columns[i] >= d->cw) return fail(ZIPGPU_ERR_INVALID, "column index out of range");
    zipgpu_ctx *ctx = d->ctx;
    API_LOCK(ctx);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    DevGuard guard(ctx, s);
    const size_t bytes = num_cols * zipgpu_data_open_columns_wire_bytes(d);
    uint32_t *d_cols = nullptr;
    uint8_t *d_out = nullptr;
    DEV_ALLOC(ctx, &d_cols, num_cols * 4, s);
    DEV_ALLOC(ctx, &d_out, bytes, s);
    CU(cudaMemcpyAsync(d_cols, columns, num_cols * 4, cudaMemcpyHostToDevice, s));
    OpenArgs a;
    a.rows = reinterpret_cast<const uint32_t *>(d->d_rows);
    a.layers = d->d_layers;
    a.columns = d_cols;
    a.col_values = nullptr;
    a.paths = nullptr;
    a.num_rows = (uint32_t)d->num_rows;
    a.cw = (uint32_t)d->cw;
    a.out32 = (uint32_t)d->out_limbs * 2;
    a.num_cols = (uint32_t)num_cols;
    a.depth = d->depth;
    a.stream = s;
    cudaError_t e = launch_open_columns_wire(a, d_out);
    if (e != cudaSuccess) return cuda_fail(e, "launch_open_columns_wire");
    ctx->launches++;
    {
        const size_t vrow = (size_t)d->out_limbs * 8, prow = 8 + (size_t)d->depth * 32;
        const size_t col_local = d->num_rows * (vrow + prow), col_total = total_rows * (vrow + prow);
        if (col_local == col_total) {  // the whole commitment is here: the device buffer IS the stream, one linear copy
            CU(cudaMemcpyAsync(stream_out, d_out, bytes, cudaMemcpyDeviceToHost, s));
        } else {
            CU(cudaMemcpy2DAsync(stream_out + row_offset * vrow, col_total, d_out, col_local, d->num_rows * vrow, num_cols,
                                 cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpy2DAsync(stream_out + total_rows * vrow + row_offset * prow, col_total, d_out + d->num_rows * vrow,
                                 col_local, d->num_rows * prow, num_cols, cudaMemcpyDeviceToHost, s));
        }
    }
    DEV_FREE(ctx, d_cols, s);
    DEV_FREE(ctx, d_out, s);
    CU(cudaStreamSynchronize(s));
    return ZIPGPU_OK;
}

// ------------------------------------------------------------------------------------------------------
// proximity-test row combination
// ------------------------------------------------------------------------------------------------------
extern "C" int zipgpu_combine_rows_device(zipgpu_ctx *ctx, size_t num_rows, size_t row_len, const uint64_t *d_evals,
                                          const uint64_t *d_coeffs, int out_limbs, uint64_t *d_out, void *stream) {
    if (!ctx || (num_rows && row_len && (!d_evals || !d_coeffs || !d_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (out_limbs < 3) return fail(ZIPGPU_ERR_WIDTH, "combine_rows needs out_limbs >= 3 (products of Int<1> are 128 bits wide)");
    if (num_rows > 0xffffffffull || row_len > 0xffffffffull) return fail(ZIPGPU_ERR_UNSUPPORTED, "shape too large");
    if (num_rows == 0 || row_len == 0) return ZIPGPU_OK;
    API_LOCK(ctx);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    DevGuard guard(ctx, s);
    uint64_t *scratch = nullptr;
    DEV_ALLOC(ctx, &scratch, combine_rows_scratch_bytes((uint32_t)num_rows, (uint32_t)row_len), s);
    CombineArgs a;
    a.evals = d_evals;
    a.coeffs = d_coeffs;
    a.scratch = scratch;
    a.out = d_out;
    a.num_rows = (uint32_t)num_rows;
    a.row_len = (uint32_t)row_len;
    a.out_limbs = (uint32_t)out_limbs;
    a.stream = s;
    int n = 0;
    cudaError_t e = launch_combine_rows(a, &n);
    if (e != cudaSuccess) return cuda_fail(e, "launch_combine_rows");
    ctx->launches += (uint64_t)n;
    DEV_FREE(ctx, scratch, s);
    return ZIPGPU_OK;
}

extern "C" int zipgpu_data_combine_rows(const zipgpu_data *d, const uint64_t *coeffs, int out_limbs, uint64_t *combined_out) {
    if (!d || !coeffs || !combined_out) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (d->in_limbs != 1) return fail(ZIPGPU_ERR_UNSUPPORTED, "combine_rows is implemented for Int<1> evaluations");
    zipgpu_ctx *ctx = d->ctx;
    API_LOCK(ctx);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    DevGuard guard(ctx, s);
    uint64_t *d_coeffs = nullptr, *d_out = nullptr;
    DEV_ALLOC(ctx, &d_coeffs, d->num_rows * 8, s);
    DEV_ALLOC(ctx, &d_out, d->row_len * (size_t)out_limbs * 8, s);
    CU(cudaMemcpyAsync(d_coeffs, coeffs, d->num_rows * 8, cudaMemcpyHostToDevice, s));
    int rc = zipgpu_combine_rows_device(ctx, d->num_rows, d->row_len, d->d_evals, d_coeffs, out_limbs, d_out, s);
    if (rc == 0) CU(cudaMemcpyAsync(combined_out, d_out, d->row_len * (size_t)out_limbs * 8, cudaMemcpyDeviceToHost, s));
    DEV_FREE(ctx, d_coeffs, s);
    DEV_FREE(ctx, d_out, s);
    CU(cudaStreamSynchronize(s));
    return rc;
}

// ------------------------------------------------------------------------------------------------------
// measurement helpers
// ------------------------------------------------------------------------------------------------------
extern "C" int zipgpu_profile_enable(zipgpu_ctx *c, int on) {
    if (!c) return fail(ZIPGPU_ERR_INVALID, "ctx is NULL");
    c->profile = on != 0;
    return ZIPGPU_OK;
}

extern "C" int zipgpu_profile_read(zipgpu_ctx *c, double *encode_ms, double *hash_ms, uint64_t *calls, int reset) {
    if (!c) return fail(ZIPGPU_ERR_INVALID, "ctx is NULL");
    API_LOCK(c);
    CU(cudaSetDevice(c->device));
    std::lock_guard<std::mutex> lk(c->mu);
    for (auto &r : c->prof_pending) {
        float ms = 0;
        if (r.has_enc) {
            CU(cudaEventSynchronize(r.e1));
            CU(cudaEventElapsedTime(&ms, r.e0, r.e1));
            c->enc_ms += ms;
        }
        if (r.has_hash) {
            CU(cudaEventSynchronize(r.e2));
            CU(cudaEventElapsedTime(&ms, r.e1, r.e2));
            c->hash_ms += ms;
        }
        if (!r.is_part) c->prof_calls++;
        c->prof_free.push_back(r);
    }
    c->prof_pending.clear();
    if (encode_ms) *encode_ms = c->enc_ms;
    if (hash_ms) *hash_ms = c->hash_ms;
    if (calls) *calls = c->prof_calls;
    if (reset) {
        c->enc_ms = c->hash_ms = 0;
        c->prof_calls = 0;
    }
    return ZIPGPU_OK;
}

extern "C" int zipgpu_microbench_int32(zipgpu_ctx *c, int kind, int iters, double *lane_ops_per_s) {
    if (!c || !lane_ops_per_s) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (kind < 0 || kind > 2 || iters < 1) return fail(ZIPGPU_ERR_INVALID, "bad kind/iters");
    API_LOCK(c);
    CU(cudaSetDevice(c->device));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    double ops = 0;
    cudaError_t e = launch_microbench_int32(kind, iters, c->num_sms, c->stream, c->d_sink, &ops);  // warm-up
    if (e != cudaSuccess) return cuda_fail(e, "microbench");
    CU(cudaEventRecord(e0, c->stream));
    e = launch_microbench_int32(kind, iters, c->num_sms, c->stream, c->d_sink, &ops);
    if (e != cudaSuccess) return cuda_fail(e, "microbench");
    CU(cudaEventRecord(e1, c->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    c->launches += 2;
    *lane_ops_per_s = ops / (ms * 1e-3);
    return ZIPGPU_OK;
}
