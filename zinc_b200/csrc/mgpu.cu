// zinc_b200/csrc/mgpu.cu -- zipgpu_mgpu_*: ONE process driving n GPUs behind the commit API (include/zipgpu.h).
//
// The reference's caller is a single process (ZincProver: RaaCode::new -> setup -> commit, zinc/prover.rs:313-315) and
// every row of the evaluation matrix is encoded and Merkle-hashed independently (commit.rs:71-81), so a commit shards
// by contiguous row range and a batch_commit by polynomial.  This layer is built on the PUBLIC single-GPU API only:
// one zipgpu_ctx and one worker thread per device (CUDA calls of different devices never serialise on one host
// thread, and the host slices go over all PCIe links at once), and one zipgpu_peer_roots exchange per commitment size,
// connected through direct peer access, so that the kernel which produces a GPU's roots stores them into every
// other GPU's buffer (merkle.cu / peer_sync.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zipgpu.h"

namespace zipgpu {
void set_last_error(const std::string &msg);
int data_open_columns_wire_strided(const zipgpu_data *d, size_t num_cols, const uint32_t *columns, uint8_t *stream_out,
                                   size_t total_rows, size_t row_offset);
}  // namespace zipgpu

namespace {

// a worker thread bound to one device: runs one task at a time, keeps rc and message of the last one
class Worker {
public:
    Worker() : th_([this] { loop(); }) {}
    ~Worker() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        th_.join();
    }
    void submit(std::function<int()> fn) {
        std::lock_guard<std::mutex> lk(mu_);
        task_ = std::move(fn);
        busy_ = true;
        cv_.notify_all();
    }
    int wait(std::string *msg) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return !busy_; });
        if (msg) *msg = msg_;
        return rc_;
    }

private:
    void loop() {
        for (;;) {
            std::function<int()> fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || (busy_ && task_); });
                if (stop_) return;
                fn = std::move(task_);
                task_ = nullptr;
            }
            int rc;
            std::string msg;
            try {
                rc = fn();
                if (rc != 0) msg = zipgpu_last_error();  // this thread's message
            } catch (const std::exception &ex) {
                rc = ZIPGPU_ERR_NOMEM;
                msg = std::string("exception in device worker: ") + ex.what();
            }
            {
                std::lock_guard<std::mutex> lk(mu_);
                rc_ = rc;
                msg_ = msg;
                busy_ = false;
            }
            cv_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::function<int()> task_;
    bool busy_ = false, stop_ = false;
    int rc_ = 0;
    std::string msg_;
    std::thread th_;
};

int fail(int code, const std::string &msg) {
    zipgpu::set_last_error(msg);
    return code;
}

void shard_range(size_t n, int g, int world, size_t *begin, size_t *count) {
    const size_t base = n / world, rem = n % world;
    *begin = g * base + std::min<size_t>(g, rem);
    *count = base + ((size_t)g < rem ? 1 : 0);
}

}  // namespace

struct zipgpu_mgpu {
    int n = 0;
    std::vector<int> devices;
    std::vector<zipgpu_ctx *> ctx;
    std::vector<Worker *> workers;
    std::mutex mu;  // one multi-GPU call at a time
    std::map<size_t, std::vector<zipgpu_peer_roots *>> exchanges;  // by total_rows
};

struct zipgpu_mgpu_code {
    zipgpu_mgpu *m;
    std::vector<zipgpu_code *> code;
    size_t row_len, cw;
    int in_limbs, out_limbs, depth;
};

struct zipgpu_mgpu_data {
    zipgpu_mgpu *m;
    size_t num_rows, row_len;
    int out_limbs, depth;
    std::vector<zipgpu_data *> shard;
    std::vector<size_t> begin, count;
    size_t wire_col_bytes;
};

// run fn(g) on every device's worker, wait for all; first error wins (its message becomes the caller's last error)
static int run_all(zipgpu_mgpu *m, const std::function<int(int)> &fn) {
    for (int g = 0; g < m->n; g++) m->workers[g]->submit([&fn, g] { return fn(g); });
    int rc = 0;
    std::string msg;
    for (int g = 0; g < m->n; g++) {
        std::string mg;
        const int r = m->workers[g]->wait(&mg);
        if (r != 0 && rc == 0) {
            rc = r;
            msg = "device " + std::to_string(m->devices[g]) + ": " + mg;
        }
    }
    if (rc) zipgpu::set_last_error(msg);
    return rc;
}

extern "C" void zipgpu_mgpu_destroy(zipgpu_mgpu *m) {
    if (!m) return;
    for (auto &kv : m->exchanges)
        for (zipgpu_peer_roots *p : kv.second) zipgpu_peer_roots_destroy(p);
    for (Worker *w : m->workers) delete w;
    for (zipgpu_ctx *c : m->ctx) zipgpu_ctx_destroy(c);
    delete m;
}

extern "C" int zipgpu_mgpu_create(const int *devices, int n, zipgpu_mgpu **out) {
    if (!out) return fail(ZIPGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int visible = 0;
    int rc = zipgpu_device_count(&visible);
    if (rc) return rc;
    if (visible == 0) return fail(ZIPGPU_ERR_NO_DEVICE, "no CUDA device visible: libzipgpu has no CPU fallback");
    if (!devices && n <= 0) n = visible;
    if (n < 1 || n > 16) return fail(ZIPGPU_ERR_INVALID, "need 1..16 devices");
    zipgpu_mgpu *m = new (std::nothrow) zipgpu_mgpu();
    if (!m) return fail(ZIPGPU_ERR_NOMEM, "host allocation failed");
    m->n = n;
    for (int g = 0; g < n; g++) {
        const int dev = devices ? devices[g] : g;
        // (test knob: several ranks on ONE device exercise the whole multi-rank path on a single-GPU box; their kernels
        // wait for each other, which needs them co-resident -- fine for tests, not a production configuration)
        for (int q = 0; q < g && !getenv("ZIPGPU_MGPU_ALLOW_DUPLICATE"); q++)
            if (m->devices[q] == dev) {
                zipgpu_mgpu_destroy(m);
                return fail(ZIPGPU_ERR_INVALID, "device listed twice");
            }
        m->devices.push_back(dev);
    }
    for (int g = 0; g < n; g++) {
        zipgpu_ctx *c = nullptr;
        rc = zipgpu_ctx_create(m->devices[g], &c);
        if (rc) {
            zipgpu_mgpu_destroy(m);
            return rc;
        }
        m->ctx.push_back(c);
    }
    for (int g = 0; g < n; g++) m->workers.push_back(new Worker());
    *out = m;
    return ZIPGPU_OK;
}

extern "C" int zipgpu_mgpu_num_devices(const zipgpu_mgpu *m) { return m ? m->n : 0; }
extern "C" zipgpu_ctx *zipgpu_mgpu_ctx(zipgpu_mgpu *m, int index) {
    return (m && index >= 0 && index < m->n) ? m->ctx[index] : nullptr;
}
extern "C" uint64_t zipgpu_mgpu_launch_count(const zipgpu_mgpu *m) {
    uint64_t t = 0;
    if (m)
        for (zipgpu_ctx *c : m->ctx) t += zipgpu_ctx_launch_count(c);
    return t;
}

// the roots exchange for commitments of `total_rows` rows (created on first use, reused afterwards)
static int get_exchange(zipgpu_mgpu *m, size_t total_rows, std::vector<zipgpu_peer_roots *> **out) {
    auto it = m->exchanges.find(total_rows);
    if (it == m->exchanges.end()) {
        if (m->exchanges.size() >= 8) {  // keep the number of cached exchanges (and their buffers) bounded
            auto victim = m->exchanges.begin();
            for (zipgpu_peer_roots *p : victim->second) zipgpu_peer_roots_destroy(p);
            m->exchanges.erase(victim);
        }
        std::vector<zipgpu_peer_roots *> ex(m->n, nullptr);
        int rc = 0;
        for (int g = 0; g < m->n && rc == 0; g++) rc = zipgpu_peer_roots_create(m->ctx[g], total_rows, g, m->n, &ex[g], nullptr);
        if (rc == 0) rc = zipgpu_peer_roots_connect_local(ex.data(), m->n);
        if (rc) {
            for (zipgpu_peer_roots *p : ex) zipgpu_peer_roots_destroy(p);
            return rc;
        }
        it = m->exchanges.emplace(total_rows, std::move(ex)).first;
    }
    *out = &it->second;
    return ZIPGPU_OK;
}

// after a failed sharded call the ranks' step counters may disagree: drop the exchange, the next call builds a new one
static void drop_exchange(zipgpu_mgpu *m, size_t total_rows) {
    auto it = m->exchanges.find(total_rows);
    if (it == m->exchanges.end()) return;
    for (zipgpu_peer_roots *p : it->second) zipgpu_peer_roots_destroy(p);
    m->exchanges.erase(it);
}

// ------------------------------------------------------------------------------------------------------
// code
// ------------------------------------------------------------------------------------------------------
extern "C" void zipgpu_mgpu_code_destroy(zipgpu_mgpu_code *c) {
    if (!c) return;
    for (zipgpu_code *k : c->code) zipgpu_code_destroy(k);
    delete c;
}

extern "C" int zipgpu_mgpu_code_create(zipgpu_mgpu *m, size_t row_len, size_t rep, int in_limbs, int out_limbs,
                                       const uint32_t *perm1, const uint32_t *perm2, zipgpu_mgpu_code **out) {
    if (!out) return fail(ZIPGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!m) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(m->mu);
    zipgpu_mgpu_code *c = new (std::nothrow) zipgpu_mgpu_code();
    if (!c) return fail(ZIPGPU_ERR_NOMEM, "host allocation failed");
    c->m = m;
    c->code.assign(m->n, nullptr);
    const int rc = run_all(m, [&](int g) {
        return zipgpu_code_create(m->ctx[g], row_len, rep, in_limbs, out_limbs, perm1, perm2, &c->code[g]);
    });
    if (rc) {
        zipgpu_mgpu_code_destroy(c);
        return rc;
    }
    c->row_len = row_len;
    c->cw = zipgpu_code_codeword_len(c->code[0]);
    c->in_limbs = in_limbs;
    c->out_limbs = out_limbs;
    c->depth = (c->cw & (c->cw - 1)) == 0 ? zipgpu_code_merkle_depth(c->code[0]) : -1;
    *out = c;
    return ZIPGPU_OK;
}

extern "C" zipgpu_code *zipgpu_mgpu_code_device(zipgpu_mgpu_code *c, int index) {
    return (c && index >= 0 && index < c->m->n) ? c->code[index] : nullptr;
}

// ------------------------------------------------------------------------------------------------------
// encode_rows / commit / batch_commit with host buffers
// ------------------------------------------------------------------------------------------------------
extern "C" int zipgpu_mgpu_encode_rows(zipgpu_mgpu_code *c, size_t num_rows, const uint64_t *evals, uint64_t *rows_out) {
    if (!c || (num_rows && (!evals || !rows_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    zipgpu_mgpu *m = c->m;
    std::lock_guard<std::mutex> lk(m->mu);
    return run_all(m, [&](int g) {
        size_t b, n;
        shard_range(num_rows, g, m->n, &b, &n);
        return zipgpu_encode_rows(c->code[g], n, evals + b * c->row_len * c->in_limbs, rows_out + b * c->cw * c->out_limbs);
    });
}

extern "C" int zipgpu_mgpu_commit(zipgpu_mgpu_code *c, size_t num_rows, const uint64_t *evals, uint64_t *rows_out,
                                  uint8_t *layers_out, uint8_t *roots_out) {
    if (!c || (num_rows && (!evals || !roots_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (c->depth < 0) return fail(ZIPGPU_ERR_INVALID, "leaves.len().is_power_of_two(): codeword_len is not a power of two");
    zipgpu_mgpu *m = c->m;
    std::lock_guard<std::mutex> lk(m->mu);
    if (rows_out || layers_out) {
        // the host takes every output: each GPU copies its row range straight to its place, nothing to exchange
        const size_t lay_row = (((size_t)2 << c->depth) - 2) * 32;
        return run_all(m, [&](int g) {
            size_t b, n;
            shard_range(num_rows, g, m->n, &b, &n);
            return zipgpu_commit(c->code[g], n, evals + b * c->row_len * c->in_limbs,
                                 rows_out ? rows_out + b * c->cw * c->out_limbs : nullptr,
                                 layers_out ? layers_out + b * lay_row : nullptr, roots_out + b * 32);
        });
    }
    std::vector<zipgpu_peer_roots *> *ex = nullptr;
    int rc = get_exchange(m, num_rows, &ex);
    if (rc) return rc;
    rc = run_all(m, [&](int g) {
        size_t b, n;
        shard_range(num_rows, g, m->n, &b, &n);
        return zipgpu_commit_resident_sharded(c->code[g], (*ex)[g], b, n, evals + b * c->row_len * c->in_limbs,
                                              g == 0 ? roots_out : nullptr, nullptr);
    });
    if (rc) {
        const std::string msg = zipgpu_last_error();
        drop_exchange(m, num_rows);
        zipgpu::set_last_error(msg);
    }
    return rc;
}

extern "C" int zipgpu_mgpu_batch_commit(zipgpu_mgpu_code *c, size_t num_polys, size_t num_rows, const uint64_t *const *evals,
                                        uint64_t *const *rows_out, uint8_t *const *layers_out, uint8_t *const *roots_out) {
    if (!c || (num_polys && (!evals || !roots_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    zipgpu_mgpu *m = c->m;
    std::lock_guard<std::mutex> lk(m->mu);
    return run_all(m, [&](int g) {
        std::vector<const uint64_t *> ev;
        std::vector<uint64_t *> ro;
        std::vector<uint8_t *> la, rt;
        for (size_t p = (size_t)g; p < num_polys; p += (size_t)m->n) {  // polynomial p -> device p mod n
            ev.push_back(evals[p]);
            ro.push_back(rows_out ? rows_out[p] : nullptr);
            la.push_back(layers_out ? layers_out[p] : nullptr);
            rt.push_back(roots_out[p]);
        }
        if (ev.empty()) return (int)ZIPGPU_OK;
        return zipgpu_batch_commit(c->code[g], ev.size(), num_rows, ev.data(), rows_out ? ro.data() : nullptr,
                                   layers_out ? la.data() : nullptr, rt.data());
    });
}

// ------------------------------------------------------------------------------------------------------
// resident prover data
// ------------------------------------------------------------------------------------------------------
extern "C" void zipgpu_mgpu_data_free(zipgpu_mgpu_data *d) {
    if (!d) return;
    for (zipgpu_data *s : d->shard) zipgpu_data_free(s);
    delete d;
}

extern "C" int zipgpu_mgpu_commit_resident(zipgpu_mgpu_code *c, size_t num_rows, const uint64_t *evals, uint8_t *roots_out,
                                           zipgpu_mgpu_data **handle) {
    if (!c || !handle || (num_rows && (!evals || !roots_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    *handle = nullptr;
    if (c->depth < 0) return fail(ZIPGPU_ERR_INVALID, "leaves.len().is_power_of_two(): codeword_len is not a power of two");
    zipgpu_mgpu *m = c->m;
    std::lock_guard<std::mutex> lk(m->mu);
    std::vector<zipgpu_peer_roots *> *ex = nullptr;
    int rc = get_exchange(m, num_rows, &ex);
    if (rc) return rc;
    zipgpu_mgpu_data *d = new (std::nothrow) zipgpu_mgpu_data();
    if (!d) return fail(ZIPGPU_ERR_NOMEM, "host allocation failed");
    d->m = m;
    d->num_rows = num_rows;
    d->row_len = c->row_len;
    d->out_limbs = c->out_limbs;
    d->depth = c->depth;
    d->shard.assign(m->n, nullptr);
    d->begin.assign(m->n, 0);
    d->count.assign(m->n, 0);
    d->wire_col_bytes = num_rows * ((size_t)c->out_limbs * 8 + 8 + (size_t)c->depth * 32);
    for (int g = 0; g < m->n; g++) shard_range(num_rows, g, m->n, &d->begin[g], &d->count[g]);
    rc = run_all(m, [&](int g) {
        return zipgpu_commit_resident_sharded(c->code[g], (*ex)[g], d->begin[g], d->count[g],
                                              evals + d->begin[g] * c->row_len * c->in_limbs, g == 0 ? roots_out : nullptr,
                                              &d->shard[g]);
    });
    if (rc) {
        const std::string msg = zipgpu_last_error();
        zipgpu_mgpu_data_free(d);
        drop_exchange(m, num_rows);
        zipgpu::set_last_error(msg);
        return rc;
    }
    *handle = d;
    return ZIPGPU_OK;
}

extern "C" size_t zipgpu_mgpu_data_num_rows(const zipgpu_mgpu_data *d) { return d ? d->num_rows : 0; }
extern "C" zipgpu_data *zipgpu_mgpu_data_shard(const zipgpu_mgpu_data *d, int index, size_t *row_begin, size_t *row_count) {
    if (!d || index < 0 || index >= d->m->n) return nullptr;
    if (row_begin) *row_begin = d->begin[index];
    if (row_count) *row_count = d->count[index];
    return d->shard[index];
}

extern "C" int zipgpu_mgpu_data_open_columns(const zipgpu_mgpu_data *d, size_t num_cols, const uint32_t *columns,
                                             uint64_t *col_values_out, uint8_t *paths_out) {
    if (!d || (num_cols && (!columns || !col_values_out || (!paths_out && d->depth > 0))))
        return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    zipgpu_mgpu *m = d->m;
    std::lock_guard<std::mutex> lk(m->mu);
    return run_all(m, [&](int g) {
        if (!d->shard[g]) return (int)ZIPGPU_OK;
        return zipgpu_data_open_columns_strided(d->shard[g], num_cols, columns, col_values_out, paths_out, d->num_rows,
                                                d->begin[g]);
    });
}

extern "C" size_t zipgpu_mgpu_data_open_columns_wire_bytes(const zipgpu_mgpu_data *d) { return d ? d->wire_col_bytes : 0; }

extern "C" int zipgpu_mgpu_data_open_columns_wire(const zipgpu_mgpu_data *d, size_t num_cols, const uint32_t *columns,
                                                  uint8_t *stream_out) {
    if (!d || (num_cols && (!columns || !stream_out))) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    zipgpu_mgpu *m = d->m;
    std::lock_guard<std::mutex> lk(m->mu);
    return run_all(m, [&](int g) {
        if (!d->shard[g]) return (int)ZIPGPU_OK;
        return zipgpu::data_open_columns_wire_strided(d->shard[g], num_cols, columns, stream_out, d->num_rows, d->begin[g]);
    });
}

extern "C" int zipgpu_mgpu_data_combine_rows(const zipgpu_mgpu_data *d, const uint64_t *coeffs, int out_limbs,
                                             uint64_t *combined_out) {
    if (!d || !coeffs || !combined_out) return fail(ZIPGPU_ERR_INVALID, "NULL argument");
    if (out_limbs < 3 || out_limbs > 64) return fail(ZIPGPU_ERR_WIDTH, "combine_rows needs 3 <= out_limbs <= 64");
    zipgpu_mgpu *m = d->m;
    std::lock_guard<std::mutex> lk(m->mu);
    const size_t words = d->row_len * (size_t)out_limbs;
    std::vector<std::vector<uint64_t>> part(m->n);
    int rc = run_all(m, [&](int g) {
        if (!d->shard[g]) return (int)ZIPGPU_OK;
        part[g].resize(words);
        return zipgpu_data_combine_rows(d->shard[g], coeffs + d->begin[g], out_limbs, part[g].data());
    });
    if (rc) return rc;
    // exact sum of the per-GPU partial rows: out_limbs-limb two's-complement adds with carry
    std::memset(combined_out, 0, words * 8);
    for (int g = 0; g < m->n; g++) {
        if (part[g].empty()) continue;
        for (size_t i = 0; i < d->row_len; i++) {
            unsigned carry = 0;
            for (int l = 0; l < out_limbs; l++) {
                const uint64_t a = combined_out[i * out_limbs + l], b = part[g][i * out_limbs + l];
                const uint64_t s1 = a + b, s2 = s1 + carry;
                carry = (s1 < a) | (s2 < s1);
                combined_out[i * out_limbs + l] = s2;
            }
        }
    }
    return ZIPGPU_OK;
}

extern "C" const uint8_t *zipgpu_mgpu_data_roots_device(const zipgpu_mgpu_data *d, int index) {
    if (!d || index < 0 || index >= d->m->n || !d->shard[index]) return nullptr;
    return zipgpu_data_all_roots_device(d->shard[index]);
}
