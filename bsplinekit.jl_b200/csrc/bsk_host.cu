// bsk_host.cu — host-buffer entry points: chunked, multi-stream pipelines that overlap the
// host->device copy of chunk i+1, the kernel of chunk i and the device->host copy of chunk i-1.
// These are what a caller with ordinary (ideally pinned) host arrays uses; they are also the
// "e2e" leg of bench.py.  Staging buffers and streams are cached per device (grow-only) so that
// repeated calls do not pay cudaMalloc / cudaFree (which synchronises the device).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "handles.cuh"

using namespace bsk;

namespace {

constexpr int kSlots = 3;
constexpr int kBufs = 5;          // x + up to 4 outputs (eval); buffer 0 only (solve)
constexpr int kMaxDev = 16;

struct Pool {
    cudaStream_t st[kSlots] = {nullptr, nullptr, nullptr};
    void *buf[kSlots][kBufs] = {{nullptr}};
    size_t cap[kSlots][kBufs] = {{0}};
};
Pool g_pool[kMaxDev];
std::mutex g_pool_mu[kMaxDev];

cudaError_t pool_get(Pool &p, int slot, int b, size_t bytes, void **out) {
    if (p.cap[slot][b] < bytes) {
        if (p.buf[slot][b]) bsk::dev_free(p.buf[slot][b]);
        p.buf[slot][b] = nullptr;
        p.cap[slot][b] = 0;
        cudaError_t e = bsk::dev_malloc(&p.buf[slot][b], bytes);
        if (e != cudaSuccess) return e;
        p.cap[slot][b] = bytes;
    }
    *out = p.buf[slot][b];
    return cudaSuccess;
}

// Pinned staging for PAGEABLE caller arrays (what a Julia Vector or a numpy array is): the driver's own path for
// pageable memory stages through one thread at ~10 GB/s and blocks the caller for every chunk; here the caller's
// chunk is copied to / from page-locked buffers by a few host threads while the previous chunk is on the device.
struct PinPool {
    void *buf[kSlots][kBufs] = {{nullptr}};
    size_t cap[kSlots][kBufs] = {{0}};
};
PinPool g_pin[kMaxDev];

cudaError_t pin_get(PinPool &p, int slot, int b, size_t bytes, void **out) {
    if (p.cap[slot][b] < bytes) {
        if (p.buf[slot][b]) cudaFreeHost(p.buf[slot][b]);
        p.buf[slot][b] = nullptr;
        p.cap[slot][b] = 0;
        cudaError_t e = cudaHostAlloc(&p.buf[slot][b], bytes, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        p.cap[slot][b] = bytes;
    }
    *out = p.buf[slot][b];
    return cudaSuccess;
}

bool is_pageable(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

// A few persistent copy threads (created on first use, joined at exit): spawning threads per chunk costs more
// than the copy of a 1 MB piece.
class CopyPool {
  public:
    explicit CopyPool(unsigned n) {
        for (unsigned t = 0; t < n; ++t) th_.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> l(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    unsigned size() const { return (unsigned)th_.size(); }
    // copy `bytes` in `parts` pieces; the caller copies piece 0 itself and waits for the others
    void copy(void *dst, const void *src, size_t bytes, unsigned parts) {
        const size_t step = ((bytes + parts - 1) / parts + 63) & ~(size_t)63;
        {
            std::lock_guard<std::mutex> l(mu_);
            for (unsigned t = 1; t < parts; ++t) {
                const size_t o = std::min(bytes, (size_t)t * step), m = std::min(step, bytes - o);
                if (m) { jobs_.push_back({(char *)dst + o, (const char *)src + o, m}); ++pending_; }
            }
        }
        cv_.notify_all();
        memcpy(dst, src, std::min(step, bytes));
        std::unique_lock<std::mutex> l(mu_);
        done_.wait(l, [this] { return pending_ == 0; });
    }

  private:
    struct Job { char *d; const char *s; size_t n; };
    void run() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [this] { return stop_ || !jobs_.empty(); });
                if (stop_ && jobs_.empty()) return;
                j = jobs_.back();
                jobs_.pop_back();
            }
            memcpy(j.d, j.s, j.n);
            {
                std::lock_guard<std::mutex> l(mu_);
                --pending_;
            }
            done_.notify_all();
        }
    }
    std::vector<std::thread> th_;
    std::vector<Job> jobs_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    size_t pending_ = 0;
    bool stop_ = false;
};

// (called with the device's pool mutex held: one staged call per device at a time; calls on different devices
//  share the copy threads, serialised by copy_mu)
void parallel_memcpy(void *dst, const void *src, size_t bytes) {
    static const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    static const unsigned cap = getenv("BSK_HOST_THREADS") ? (unsigned)std::max(1, atoi(getenv("BSK_HOST_THREADS"))) : 8u;
    static std::mutex copy_mu;
    const size_t per = (size_t)1 << 20;                       // at least 1 MB per piece
    const unsigned nt = (unsigned)std::min<size_t>(std::min<unsigned>(cap, hw), (bytes + per - 1) / per);
    if (nt <= 1) { memcpy(dst, src, bytes); return; }
    std::lock_guard<std::mutex> l(copy_mu);
    static bool usable = true;
    if (usable) {
        try {                                                 // no exception may cross the C ABI: a process that
            static CopyPool pool(std::min<unsigned>(cap, hw) - 1);      // cannot start threads copies with one
            pool.copy(dst, src, bytes, std::min(nt, pool.size() + 1));
            return;
        } catch (...) {
            usable = false;
        }
    }
    memcpy(dst, src, bytes);
}

cudaError_t pool_stream(Pool &p, int slot, cudaStream_t *out) {
    if (!p.st[slot]) {
        cudaError_t e = cudaStreamCreateWithFlags(&p.st[slot], cudaStreamNonBlocking);
        if (e != cudaSuccess) return e;
    }
    *out = p.st[slot];
    return cudaSuccess;
}

}  // namespace

static bsk_status eval_host_impl(const bsk_spline *s, const void *h_x, int64_t n, int32_t nout,
                                 const int32_t *derivs, void *const *h_outs, int32_t algo, bool copy_only) {
    BSK_REQUIRE(s && (derivs || copy_only) && h_outs, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(nout >= 1 && nout <= 4, BSK_ERR_ARGUMENT, "nout must be in 1..4");
    if (n == 0) return BSK_OK;
    BSK_REQUIRE(h_x, BSK_ERR_ARGUMENT, "NULL points");
    const int device = s->basis->device;
    BSK_REQUIRE(device >= 0 && device < kMaxDev, BSK_ERR_UNSUPPORTED, "device index out of range");
    const size_t es = s->basis->dtype == BSK_F64 ? 8 : 4;
    BSK_CUDA(cudaSetDevice(device));
    std::lock_guard<std::mutex> lock(g_pool_mu[device]);
    Pool &pool = g_pool[device];
    // pageable caller arrays go through page-locked staging buffers in smaller chunks (see PinPool)
    bool staged = getenv("BSK_HOST_NO_STAGING") == nullptr && n >= (1 << 14) && is_pageable(h_x);
    for (int o = 0; o < nout && staged; ++o) staged = is_pageable(h_outs[o]);
    int chunk_log2 = staged ? (n >= ((int64_t)1 << 24) ? 22 : 20) : 23;      // 2^23 points per chunk (64 MB of Float64 points); staged: 2^20 .. 2^22
    if (const char *ev = getenv("BSK_HOST_CHUNK_LOG2")) chunk_log2 = std::max(16, std::min(27, atoi(ev)));
    int64_t chunk = std::min<int64_t>(n, (int64_t)1 << chunk_log2);
    if (staged && n > (1 << 16) && n < 3 * chunk) chunk = ((n + 2) / 3 + 1023) / 1024 * 1024;   // at least three chunks to overlap
    const int nslots = (int)std::min<int64_t>(kSlots, (n + chunk - 1) / chunk);
    cudaStream_t st[kSlots] = {nullptr, nullptr, nullptr};
    void *dx[kSlots] = {nullptr, nullptr, nullptr};
    void *dout[kSlots][4] = {{nullptr}};
    bsk_status rc = BSK_OK;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < nslots && e == cudaSuccess; ++i) {
        e = pool_stream(pool, i, &st[i]);
        if (e == cudaSuccess) e = pool_get(pool, i, 0, es * chunk, &dx[i]);
        for (int o = 0; o < nout && e == cudaSuccess; ++o) e = pool_get(pool, i, 1 + o, es * chunk, &dout[i][o]);
    }
    void *px[kSlots] = {nullptr, nullptr, nullptr};
    void *pout[kSlots][4] = {{nullptr}};
    if (staged) {
        PinPool &pin = g_pin[device];
        for (int i = 0; i < nslots && e == cudaSuccess; ++i) {
            e = pin_get(pin, i, 0, es * chunk, &px[i]);
            for (int o = 0; o < nout && e == cudaSuccess; ++o) e = pin_get(pin, i, 1 + o, es * chunk, &pout[i][o]);
        }
        if (e != cudaSuccess) { cudaGetLastError(); staged = false; e = cudaSuccess; }      // no page-locked memory: plain path
    }
    if (e == cudaSuccess && staged) {
        int64_t pend_off[kSlots] = {-1, -1, -1}, pend_m[kSlots] = {0, 0, 0};
        auto drain = [&](int slot) -> cudaError_t {           // results of the chunk last issued on `slot` -> caller
            if (pend_off[slot] < 0) return cudaSuccess;
            cudaError_t ee = cudaStreamSynchronize(st[slot]);
            if (ee != cudaSuccess) return ee;
            for (int o = 0; o < nout; ++o)
                parallel_memcpy((char *)h_outs[o] + es * pend_off[slot], pout[slot][o], es * pend_m[slot]);
            pend_off[slot] = -1;
            return cudaSuccess;
        };
        int64_t off = 0;
        for (int64_t ci = 0; off < n && rc == BSK_OK; ++ci, off += chunk) {
            const int slot = (int)(ci % nslots);
            const int64_t m = std::min(chunk, n - off);
            if ((e = drain(slot)) != cudaSuccess) break;
            parallel_memcpy(px[slot], (const char *)h_x + es * off, es * m);
            e = cudaMemcpyAsync(dx[slot], px[slot], es * m, cudaMemcpyHostToDevice, st[slot]);
            if (e != cudaSuccess) break;
            if (!copy_only) rc = bsk_spline_eval(s, dx[slot], m, nout, derivs, dout[slot], algo, (void *)st[slot]);
            if (rc != BSK_OK) break;
            for (int o = 0; o < nout && e == cudaSuccess; ++o)
                e = cudaMemcpyAsync(pout[slot][o], dout[slot][o], es * m, cudaMemcpyDeviceToHost, st[slot]);
            if (e != cudaSuccess) break;
            pend_off[slot] = off;
            pend_m[slot] = m;
        }
        // the chunks still in flight, oldest first
        if (e == cudaSuccess && rc == BSK_OK) {
            const int64_t nchunks = (n + chunk - 1) / chunk;
            for (int64_t ci = std::max<int64_t>(0, nchunks - nslots); ci < nchunks && e == cudaSuccess; ++ci)
                e = drain((int)(ci % nslots));
        }
    } else if (e == cudaSuccess) {
        int64_t off = 0;
        for (int64_t ci = 0; off < n && rc == BSK_OK; ++ci, off += chunk) {
            const int slot = (int)(ci % nslots);
            const int64_t m = std::min(chunk, n - off);
            // stream order on the slot guarantees the previous use of its buffers is finished
            e = cudaMemcpyAsync(dx[slot], (const char *)h_x + es * off, es * m, cudaMemcpyHostToDevice, st[slot]);
            if (e != cudaSuccess) break;
            if (!copy_only) rc = bsk_spline_eval(s, dx[slot], m, nout, derivs, dout[slot], algo, (void *)st[slot]);
            if (rc != BSK_OK) break;
            for (int o = 0; o < nout && e == cudaSuccess; ++o)
                e = cudaMemcpyAsync((char *)h_outs[o] + es * off, dout[slot][o], es * m, cudaMemcpyDeviceToHost, st[slot]);
            if (e != cudaSuccess) break;
        }
    }
    for (int i = 0; i < nslots; ++i)
        if (st[i]) {
            cudaError_t e2 = cudaStreamSynchronize(st[i]);
            if (e == cudaSuccess) e = e2;
        }
    if (rc != BSK_OK) return rc;
    if (e != cudaSuccess) { set_error("bsk_spline_eval_host failed: %s", cudaGetErrorString(e)); return BSK_ERR_CUDA; }
    return BSK_OK;
}

extern "C" {

bsk_status bsk_spline_eval_host(const bsk_spline *s, const void *h_x, int64_t n, int32_t nout,
                                const int32_t *derivs, void *const *h_outs, int32_t algo) {
    return eval_host_impl(s, h_x, n, nout, derivs, h_outs, algo, false);
}

bsk_status bsk_probe_copy_host(const bsk_spline *s, const void *h_x, int64_t n, int32_t nout, void *const *h_outs) {
    return eval_host_impl(s, h_x, n, nout, nullptr, h_outs, 0, true);
}

bsk_status bsk_banded_solve_host(const bsk_banded *h, double *h_rhs, int64_t nrhs, int32_t algo) {
    BSK_REQUIRE(h && (nrhs == 0 || h_rhs), BSK_ERR_ARGUMENT, "NULL argument");
    if (nrhs == 0) return BSK_OK;
    BSK_REQUIRE(h->device >= 0 && h->device < kMaxDev, BSK_ERR_UNSUPPORTED, "device index out of range");
    BSK_CUDA(cudaSetDevice(h->device));
    std::lock_guard<std::mutex> lock(g_pool_mu[h->device]);
    Pool &pool = g_pool[h->device];
    const int64_t n = h->n;
    // column blocks of ~128 MB, multiples of 32 columns
    int64_t cb = std::max<int64_t>(32, (((int64_t)128 << 20) / (8 * n)) / 32 * 32);
    cb = std::min(nrhs, cb);
    const int nslots = (int)std::min<int64_t>(kSlots, (nrhs + cb - 1) / cb);
    cudaStream_t st[kSlots] = {nullptr, nullptr, nullptr};
    double *d[kSlots] = {nullptr, nullptr, nullptr};
    cudaError_t e = cudaSuccess;
    bsk_status rc = BSK_OK;
    for (int i = 0; i < nslots && e == cudaSuccess; ++i) {
        e = pool_stream(pool, i, &st[i]);
        if (e == cudaSuccess) e = pool_get(pool, i, 0, 8 * (size_t)n * cb, (void **)&d[i]);
    }
    if (e == cudaSuccess) {
        int64_t c0 = 0;
        for (int64_t ci = 0; c0 < nrhs && rc == BSK_OK; ++ci, c0 += cb) {
            const int slot = (int)(ci % nslots);
            const int64_t w = std::min(cb, nrhs - c0);
            e = cudaMemcpy2DAsync(d[slot], 8 * w, h_rhs + c0, 8 * nrhs, 8 * w, n, cudaMemcpyHostToDevice, st[slot]);
            if (e != cudaSuccess) break;
            rc = bsk_banded_solve(h, d[slot], w, algo, (void *)st[slot]);
            if (rc != BSK_OK) break;
            e = cudaMemcpy2DAsync(h_rhs + c0, 8 * nrhs, d[slot], 8 * w, 8 * w, n, cudaMemcpyDeviceToHost, st[slot]);
            if (e != cudaSuccess) break;
        }
    }
    for (int i = 0; i < nslots; ++i)
        if (st[i]) {
            cudaError_t e2 = cudaStreamSynchronize(st[i]);
            if (e == cudaSuccess) e = e2;
        }
    if (rc != BSK_OK) return rc;
    if (e != cudaSuccess) { set_error("bsk_banded_solve_host failed: %s", cudaGetErrorString(e)); return BSK_ERR_CUDA; }
    return BSK_OK;
}

}  // extern "C"
