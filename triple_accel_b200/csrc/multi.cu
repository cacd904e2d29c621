// multi.cu -- one ta_ctx over several GPUs of a box (include/triple_accel_b200.h: ta_init_multi).
//
// The path shards embarrassingly: a batch is cut into contiguous ranges of pairs / haystacks balanced by bytes, every
// range goes through the single-device entry point of its own sub-context (own streams, own staging buffers, own
// host thread), and results land in the caller's arrays at the range's offset.  No data-path collective.  The one
// exchange the path has is the needle of a search: it is uploaded to the first device and sent to the others with
// ncclBroadcast over a single-process communicator (ncclCommInitAll), as BASELINE.json's north_star asks; NCCL is
// loaded at run time (dlopen of libnccl.so.2), so single-GPU users do not need it.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <new>
#include <thread>
#include <vector>

#include "ta_common.cuh"

namespace {

// ---- a host thread per extra device -----------------------------------------------------------------------------
class Worker {
  public:
    Worker() : th_([this] { loop(); }) {}
    ~Worker() {
        {
            std::lock_guard<std::mutex> g(mu_);
            quit_ = true;
        }
        cv_.notify_all();
        th_.join();
    }
    void post(std::function<void()> job) {
        {
            std::lock_guard<std::mutex> g(mu_);
            job_ = std::move(job);
            state_ = 1;
        }
        cv_.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [this] { return state_ == 0; });
    }

  private:
    void loop() {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [this] { return quit_ || state_ == 1; });
                if (quit_) return;
                job = std::move(job_);
                state_ = 2;
            }
            job();
            {
                std::lock_guard<std::mutex> g(mu_);
                state_ = 0;
            }
            cv_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::function<void()> job_;
    int state_ = 0;  // 0 idle, 1 posted, 2 running
    bool quit_ = false;
    std::thread th_;  // last: started after the other members exist
};

// ---- NCCL, resolved at run time -----------------------------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
struct Nccl {
    void *so = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok() const { return CommInitAll && CommDestroy && GroupStart && GroupEnd && Broadcast; }
};
constexpr int NCCL_UINT8 = 1;  // ncclUint8 (nccl.h)

Nccl *load_nccl() {
    static Nccl lib;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = getenv("TA_NCCL_LIB");
        const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm || !*nm) continue;
            lib.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (lib.so) break;
        }
        if (!lib.so) return;
        lib.CommInitAll = (decltype(lib.CommInitAll))dlsym(lib.so, "ncclCommInitAll");
        lib.CommDestroy = (decltype(lib.CommDestroy))dlsym(lib.so, "ncclCommDestroy");
        lib.GroupStart = (decltype(lib.GroupStart))dlsym(lib.so, "ncclGroupStart");
        lib.GroupEnd = (decltype(lib.GroupEnd))dlsym(lib.so, "ncclGroupEnd");
        lib.Broadcast = (decltype(lib.Broadcast))dlsym(lib.so, "ncclBroadcast");
        lib.GetErrorString = (decltype(lib.GetErrorString))dlsym(lib.so, "ncclGetErrorString");
    });
    return lib.ok() ? &lib : nullptr;
}

}  // namespace

struct ta_multi {
    std::vector<ta_ctx *> subs;
    std::vector<Worker *> workers;  // workers[r] serves subs[r] for r >= 1 (shard 0 runs on the calling thread)
    Nccl *nccl = nullptr;
    std::vector<ncclComm_t> comms;
    uint64_t needle_bcasts = 0;  // broadcasts done through NCCL so far
    size_t min_shard_bytes = (size_t)4 << 20;
};

// contiguous ranges of units balanced by bytes: bound[r] .. bound[r + 1], r < parts
static void shard_bounds(const uint64_t *a_off, const uint64_t *b_off, size_t n, int parts, std::vector<size_t> &bound) {
    bound.assign(parts + 1, n);
    bound[0] = 0;
    auto bytes_before = [&](size_t i) {
        return (a_off[i] - a_off[0]) + (b_off ? b_off[i] - b_off[0] : 0);
    };
    const uint64_t total = bytes_before(n);
    for (int r = 1; r < parts; r++) {
        const uint64_t want = total / parts * r + total % parts * r / parts;
        size_t lo = bound[r - 1], hi = n;
        while (lo < hi) {  // first i with bytes_before(i) >= want
            const size_t mid = lo + (hi - lo) / 2;
            if (bytes_before(mid) < want)
                lo = mid + 1;
            else
                hi = mid;
        }
        bound[r] = lo;
    }
    if (total == 0)  // all-empty strings: balance by count
        for (int r = 1; r < parts; r++) bound[r] = n * (size_t)r / parts;
}

int ta_multi_parts(ta_ctx *ctx, uint64_t total_bytes, size_t n) {
    ta_multi *m = ctx->multi;
    const int nd = (int)m->subs.size();
    uint64_t by_bytes = m->min_shard_bytes ? total_bytes / m->min_shard_bytes : (uint64_t)nd;
    if (by_bytes < 1) by_bytes = 1;
    return (int)std::min<uint64_t>({(uint64_t)nd, by_bytes, (uint64_t)std::max<size_t>(n, 1)});
}

// runs fn(r) for r in [0, parts): r = 0 on the calling thread, the others on their device's worker; first error wins
int ta_multi_run(ta_ctx *ctx, int parts, const std::function<int(int)> &fn) {
    ta_multi *m = ctx->multi;
    std::vector<int> rc(parts, TA_OK);
    for (int r = 1; r < parts; r++) m->workers[r]->post([&, r] { rc[r] = fn(r); });
    rc[0] = fn(0);
    for (int r = 1; r < parts; r++) m->workers[r]->wait();
    for (int r = 0; r < parts; r++)
        if (rc[r] != TA_OK) {
            ctx->last_error = m->subs[r]->last_error;
            return rc[r];
        }
    uint64_t launches = 0;
    for (ta_ctx *s : m->subs) launches += s->launches;
    ctx->launches = launches;
    return TA_OK;
}

ta_ctx *ta_multi_sub(ta_ctx *ctx, int r) { return ctx->multi->subs[r]; }
int ta_multi_size(ta_ctx *ctx) { return ctx->multi ? (int)ctx->multi->subs.size() : 1; }
void ta_multi_bounds(const uint64_t *a_off, const uint64_t *b_off, size_t n, int parts, std::vector<size_t> &bound) {
    shard_bounds(a_off, b_off, n, parts, bound);
}

// Needle of a search: host -> device 0, then ncclBroadcast to the other devices' needle buffers (ctx->d_b[0] of each
// sub-context), each on its own compute stream so the shard's kernels are ordered behind it.  Falls back to one
// host-to-device copy per device when NCCL is not available (TA_NCCL=0, or no libnccl on the box).
int ta_multi_needle(ta_ctx *ctx, const uint8_t *needle, size_t needle_len, int parts) {
    ta_multi *m = ctx->multi;
    for (int r = 0; r < parts; r++) {
        ta_ctx *s = m->subs[r];
        TA_CUDA(s, cudaSetDevice(s->device));
        int rc = ta_dev_reserve(s, s->d_b[0], needle_len + 64);
        if (rc != TA_OK) return rc;
    }
    const bool use_nccl = m->nccl && (int)m->comms.size() == (int)m->subs.size() && parts == (int)m->subs.size() && parts > 1;
    ta_ctx *s0 = m->subs[0];
    TA_CUDA(s0, cudaSetDevice(s0->device));
    TA_CUDA(s0, cudaMemcpyAsync(s0->d_b[0].p, needle, needle_len, cudaMemcpyHostToDevice, s0->stream));
    if (use_nccl) {
        int e = m->nccl->GroupStart();
        for (int r = 0; r < parts && e == 0; r++) {
            ta_ctx *s = m->subs[r];
            cudaSetDevice(s->device);
            e = m->nccl->Broadcast(s0->d_b[0].p, s->d_b[0].p, needle_len, NCCL_UINT8, 0, m->comms[r], s->stream);
        }
        const int e2 = m->nccl->GroupEnd();
        if (e == 0) e = e2;
        if (e != 0) {
            char buf[256];
            snprintf(buf, sizeof buf, "ncclBroadcast(needle): %s", m->nccl->GetErrorString ? m->nccl->GetErrorString(e) : "?");
            ctx->last_error = buf;
            return TA_ERR_CUDA;
        }
        m->needle_bcasts++;
    } else {
        for (int r = 1; r < parts; r++) {
            ta_ctx *s = m->subs[r];
            TA_CUDA(s, cudaSetDevice(s->device));
            TA_CUDA(s, cudaMemcpyAsync(s->d_b[0].p, needle, needle_len, cudaMemcpyHostToDevice, s->stream));
        }
    }
    return TA_OK;
}

extern "C" {

int ta_init_multi(const int *devices, int n_devices, ta_ctx **out) {
    if (!out) return TA_ERR_BAD_ARG;
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > 64) return TA_ERR_BAD_ARG;
    for (int i = 0; i < n_devices; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) return TA_ERR_BAD_ARG;
    if (n_devices == 1) return ta_init(devices[0], out);
    ta_ctx *ctx = new (std::nothrow) ta_ctx();
    ta_multi *m = new (std::nothrow) ta_multi();
    if (!ctx || !m) {
        delete ctx;
        delete m;
        return TA_ERR_NOMEM;
    }
    ctx->multi = m;
    ctx->device = devices[0];
    int rc = TA_OK;
    for (int i = 0; i < n_devices && rc == TA_OK; i++) {
        ta_ctx *s = nullptr;
        rc = ta_init(devices[i], &s);
        if (rc == TA_OK) m->subs.push_back(s);
    }
    if (rc != TA_OK) {
        ta_shutdown(ctx);
        return rc;
    }
    ctx->sm_count = m->subs[0]->sm_count;
    ctx->smem_optin = m->subs[0]->smem_optin;
    m->workers.assign(n_devices, nullptr);
    for (int i = 1; i < n_devices; i++) m->workers[i] = new Worker();
    if (const char *e = getenv("TA_MULTI_MIN_BYTES")) m->min_shard_bytes = (size_t)strtoull(e, nullptr, 10);
    const char *use = getenv("TA_NCCL");
    if (!(use && use[0] == '0')) {
        m->nccl = load_nccl();
        if (m->nccl) {
            m->comms.assign(n_devices, nullptr);
            const int e = m->nccl->CommInitAll(m->comms.data(), n_devices, devices);
            if (e != 0) {
                fprintf(stderr, "triple_accel_b200: ncclCommInitAll failed (%s); the needle will be copied per device\n",
                        m->nccl->GetErrorString ? m->nccl->GetErrorString(e) : "?");
                m->comms.clear();
                m->nccl = nullptr;
            }
        }
    }
    *out = ctx;
    return TA_OK;
}

int ta_device_count(ta_ctx *ctx) { return ctx ? ta_multi_size(ctx) : 0; }

/* The split a multi-device context applies to a batch (pure host arithmetic, no device needed). */
int ta_shard_bounds(const uint64_t *a_off, const uint64_t *b_off, size_t n, int parts, uint64_t *bounds_out) {
    if (!a_off || !bounds_out || parts < 1) return TA_ERR_BAD_ARG;
    std::vector<size_t> bound;
    shard_bounds(a_off, b_off, n, parts, bound);
    for (int r = 0; r <= parts; r++) bounds_out[r] = bound[r];
    return TA_OK;
}

/* 1 when the needle of a multi-device search travels by ncclBroadcast, 0 when it is copied host-to-device per device */
int ta_multi_uses_nccl(ta_ctx *ctx) { return ctx && ctx->multi && ctx->multi->nccl && !ctx->multi->comms.empty() ? 1 : 0; }
uint64_t ta_multi_needle_broadcasts(ta_ctx *ctx) { return ctx && ctx->multi ? ctx->multi->needle_bcasts : 0; }

}  // extern "C"

void ta_multi_shutdown(ta_ctx *ctx) {
    ta_multi *m = ctx->multi;
    if (!m) return;
    for (Worker *w : m->workers) delete w;
    if (m->nccl)
        for (ncclComm_t c : m->comms)
            if (c) m->nccl->CommDestroy(c);
    for (ta_ctx *s : m->subs) ta_shutdown(s);
    delete m;
    ctx->multi = nullptr;
}
