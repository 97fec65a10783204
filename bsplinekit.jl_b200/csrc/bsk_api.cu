// bsk_api.cu — library utilities, basis handles, search-LUT construction.
#include <algorithm>

#include "handles.cuh"

namespace bsk {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace bsk

using namespace bsk;

namespace bsk {

void ensure_pool(int device) {
    static bool done[64] = {false};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = 1ull << 30;     // keep up to 1 GiB of freed blocks instead of returning them at every sync
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done[device] = true;
}

cudaError_t dev_malloc(void **p, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    ensure_pool(dev);
    cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 1, (cudaStream_t)0);
    if (e == cudaSuccess) return cudaStreamSynchronize((cudaStream_t)0);   // usable from any stream from here on
    cudaGetLastError();
    return cudaMalloc(p, bytes ? bytes : 1);
}

cudaError_t dev_free(void *p) {
    if (!p) return cudaSuccess;
    cudaError_t e = cudaDeviceSynchronize();      // like cudaFree: nothing in flight may still use the block
    if (e != cudaSuccess) return e;
    e = cudaFreeAsync(p, (cudaStream_t)0);
    if (e == cudaSuccess) return e;
    cudaGetLastError();
    return cudaFree(p);
}

}  // namespace bsk

// 1 read : 2 writes per element, 16-byte vectors, one CTA of 1024 threads per SM, next vector prefetched:
// the memory behaviour of the fused value + derivative evaluation without its table lookups.
__global__ void __launch_bounds__(1024, 1)
k_probe_stream_1r2w(const double2 *__restrict__ in, double2 *__restrict__ o0, double2 *__restrict__ o1, int64_t nvec) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double2 nxt = make_double2(0.0, 0.0);
    if (v < nvec) nxt = in[v];
    for (; v < nvec; v += stride) {
        const double2 x = nxt;
        if (v + stride < nvec) nxt = in[v + stride];
        o0[v] = make_double2(x.x + 1.0, x.y + 1.0);
        o1[v] = make_double2(x.x * 0.5, x.y * 0.5);
    }
}

// The same stream through TMA: one elected thread per CTA issues cp.async.bulk global -> shared for the
// next tile of points (mbarrier completion), the tile is transformed in shared memory and leaves through
// cp.async.bulk shared -> global.  Answers whether a bulk-copy write path gives the 1 : 2 stream more
// HBM bandwidth than 128-bit LDG / STG (bsk_probe_stream_1r2w).
namespace {
constexpr int kTmaTile = 2048;              // doubles per tile: 16 KB in, 2 x 16 KB out, two stages
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
}  // namespace

__global__ void __launch_bounds__(1024, 1)
k_probe_stream_1r2w_tma(const double *__restrict__ in, double *__restrict__ o0, double *__restrict__ o1, int64_t ntiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    double *sin = (double *)smem;                                  // [2][kTmaTile]
    double *so0 = sin + 2 * kTmaTile;                              // [2][kTmaTile]
    double *so1 = so0 + 2 * kTmaTile;
    __shared__ uint64_t full[2];
    if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    int64_t tile = blockIdx.x;
    if (threadIdx.x == 0 && tile < ntiles) {
        mbar_expect_tx(&full[0], kTmaTile * 8);
        bulk_g2s(sin, in + tile * kTmaTile, kTmaTile * 8, &full[0]);
    }
    uint32_t phase[2] = {0, 0};
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const int st = it & 1;
        const int64_t nxt = tile + gridDim.x;
        if (threadIdx.x == 0) {
            // the stage we are about to refill / rewrite must have been drained by its bulk stores
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            if (nxt < ntiles) {
                mbar_expect_tx(&full[st ^ 1], kTmaTile * 8);
                bulk_g2s(sin + (st ^ 1) * kTmaTile, in + nxt * kTmaTile, kTmaTile * 8, &full[st ^ 1]);
            }
        }
        __syncthreads();                                           // stage st: outputs of two tiles ago are out
        mbar_wait(&full[st], phase[st]);
        phase[st] ^= 1;
        const double2 x = ((const double2 *)(sin + st * kTmaTile))[threadIdx.x];
        ((double2 *)(so0 + st * kTmaTile))[threadIdx.x] = make_double2(x.x + 1.0, x.y + 1.0);
        ((double2 *)(so1 + st * kTmaTile))[threadIdx.x] = make_double2(x.x * 0.5, x.y * 0.5);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_s2g(o0 + tile * kTmaTile, so0 + st * kTmaTile, kTmaTile * 8);
            bulk_s2g(o1 + tile * kTmaTile, so1 + st * kTmaTile, kTmaTile * 8);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Peak rate of the FP64 / FP32 FMA pipes: 8 independent fused multiply-add chains per thread, full
// occupancy, nothing else in the loop.
template <typename T>
__global__ void __launch_bounds__(256)
k_probe_fma(T *out, int iters, T a, T b) {
    T v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (T)(threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fma(v[j], a, b);
    }
    T s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    if (s == (T)-1.2345) out[0] = s;        // never true: keeps the chains alive
}

extern "C" {

const char *bsk_version(void) { return "bsplinekit_b200 0.1.0 (sm_100a)"; }

size_t bsk_last_error(char *buf, size_t buflen) {
    size_t n = strlen(g_err);
    if (buf && buflen) {
        size_t m = n < buflen - 1 ? n : buflen - 1;
        memcpy(buf, g_err, m);
        buf[m] = 0;
    }
    return n;
}

bsk_status bsk_device_count(int32_t *count) {
    BSK_REQUIRE(count, BSK_ERR_ARGUMENT, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
        return BSK_ERR_CUDA;
    }
    *count = n;
    return BSK_OK;
}

bsk_status bsk_probe_stream_1r2w(const double *d_in, double *d_out0, double *d_out1, int64_t n, void *stream) {
    BSK_REQUIRE(d_in && d_out0 && d_out1, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(n >= 0 && n % 2 == 0, BSK_ERR_ARGUMENT, "n must be even");
    BSK_REQUIRE((((uintptr_t)d_in | (uintptr_t)d_out0 | (uintptr_t)d_out1) & 15) == 0, BSK_ERR_ARGUMENT,
                "pointers must be 16-byte aligned");
    if (n == 0) return BSK_OK;
    int dev = 0;
    BSK_CUDA(cudaGetDevice(&dev));
    k_probe_stream_1r2w<<<bsk::sm_count(dev), 1024, 0, (cudaStream_t)stream>>>((const double2 *)d_in, (double2 *)d_out0,
                                                                              (double2 *)d_out1, n / 2);
    BSK_CUDA(cudaGetLastError());
    return BSK_OK;
}

bsk_status bsk_probe_stream_1r2w_tma(const double *d_in, double *d_out0, double *d_out1, int64_t n, void *stream) {
    BSK_REQUIRE(d_in && d_out0 && d_out1, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(n >= 0 && n % kTmaTile == 0, BSK_ERR_ARGUMENT, "n must be a multiple of %d", kTmaTile);
    BSK_REQUIRE((((uintptr_t)d_in | (uintptr_t)d_out0 | (uintptr_t)d_out1) & 15) == 0, BSK_ERR_ARGUMENT,
                "pointers must be 16-byte aligned");
    if (n == 0) return BSK_OK;
    int dev = 0;
    BSK_CUDA(cudaGetDevice(&dev));
    const int smem = 6 * kTmaTile * 8;
    BSK_CUDA(cudaFuncSetAttribute(k_probe_stream_1r2w_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_probe_stream_1r2w_tma<<<bsk::sm_count(dev), 1024, smem, (cudaStream_t)stream>>>(d_in, d_out0, d_out1, n / kTmaTile);
    BSK_CUDA(cudaGetLastError());
    return BSK_OK;
}

bsk_status bsk_probe_fma(int32_t dtype, int64_t iters, void *d_scratch, double *fma_count, void *stream) {
    BSK_REQUIRE(d_scratch && iters > 0 && iters < (1ll << 31), BSK_ERR_ARGUMENT, "bad argument");
    BSK_REQUIRE(dtype == BSK_F32 || dtype == BSK_F64, BSK_ERR_ARGUMENT, "bad dtype");
    int dev = 0;
    BSK_CUDA(cudaGetDevice(&dev));
    const int grid = bsk::sm_count(dev) * 8;
    if (dtype == BSK_F64) k_probe_fma<double><<<grid, 256, 0, (cudaStream_t)stream>>>((double *)d_scratch, (int)iters, 0.999999, 1e-7);
    else k_probe_fma<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float *)d_scratch, (int)iters, 0.99999f, 1e-4f);
    BSK_CUDA(cudaGetLastError());
    if (fma_count) *fma_count = 8.0 * (double)iters * 256.0 * (double)grid;
    return BSK_OK;
}

bsk_status bsk_malloc(int32_t device, size_t nbytes, void **d_ptr) {
    BSK_REQUIRE(d_ptr, BSK_ERR_ARGUMENT, "d_ptr is NULL");
    BSK_CUDA(cudaSetDevice(device));
    BSK_CUDA(cudaMalloc(d_ptr, nbytes ? nbytes : 1));
    return BSK_OK;
}

bsk_status bsk_free(int32_t device, void *d_ptr) {
    BSK_CUDA(cudaSetDevice(device));
    BSK_CUDA(cudaFree(d_ptr));
    return BSK_OK;
}

bsk_status bsk_memcpy_h2d(int32_t device, void *d_dst, const void *h_src, size_t nbytes, void *stream) {
    BSK_CUDA(cudaSetDevice(device));
    BSK_CUDA(cudaMemcpyAsync(d_dst, h_src, nbytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return BSK_OK;
}

bsk_status bsk_memcpy_d2h(int32_t device, void *h_dst, const void *d_src, size_t nbytes, void *stream) {
    BSK_CUDA(cudaSetDevice(device));
    BSK_CUDA(cudaMemcpyAsync(h_dst, d_src, nbytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return BSK_OK;
}

bsk_status bsk_stream_create(int32_t device, void **stream) {
    BSK_REQUIRE(stream, BSK_ERR_ARGUMENT, "stream is NULL");
    BSK_CUDA(cudaSetDevice(device));
    // A BLOCKING stream: the handle-building entry points (bsk_banded_create / _lu, bsk_spline_create,
    // bsk_spline_set_coefs_device, table builds) run on the legacy default stream, and a blocking stream
    // orders against it in both directions — work enqueued here before such a call is finished when the
    // call reads it, and the call's results are complete before later work on this stream starts.
    cudaStream_t s;
    BSK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamDefault));
    *stream = (void *)s;
    return BSK_OK;
}

bsk_status bsk_stream_destroy(int32_t device, void *stream) {
    BSK_CUDA(cudaSetDevice(device));
    BSK_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return BSK_OK;
}

bsk_status bsk_stream_sync(int32_t device, void *stream) {
    BSK_CUDA(cudaSetDevice(device));
    BSK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return BSK_OK;
}

bsk_status bsk_event_create(int32_t device, void **event) {
    BSK_REQUIRE(event, BSK_ERR_ARGUMENT, "event is NULL");
    BSK_CUDA(cudaSetDevice(device));
    cudaEvent_t e;
    BSK_CUDA(cudaEventCreate(&e));
    *event = (void *)e;
    return BSK_OK;
}

bsk_status bsk_event_record(void *event, void *stream) {
    BSK_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return BSK_OK;
}

bsk_status bsk_event_elapsed_ms(void *start, void *stop, float *ms) {
    BSK_REQUIRE(ms, BSK_ERR_ARGUMENT, "ms is NULL");
    BSK_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
    BSK_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return BSK_OK;
}

bsk_status bsk_event_destroy(void *event) {
    BSK_CUDA(cudaEventDestroy((cudaEvent_t)event));
    return BSK_OK;
}

}  // extern "C"

// ---- search LUT -----------------------------------------------------------------------
// For each of `ncell` uniform cells over [v[0], v[n-1]] store a bracket [lo, hi] that is
// guaranteed to contain searchsortedlast(v, x) for every x whose *computed* cell index is
// c; `margin` (in cells) absorbs the rounding of (x - v[0]) * scale in the knot dtype.
namespace bsk {

static void build_lut(const std::vector<double> &v, int dtype, std::vector<uint32_t> &lut,
                      int &ncell, double &scale) {
    const int n = (int)v.size();
    lut.clear();
    ncell = 0;
    scale = 0;
    double range = v[n - 1] - v[0];
    if (!(range > 0) || !std::isfinite(range) || n >= (1 << 24)) return;
    int nc = 64;
    while (nc < 4 * n && nc < 65536) nc <<= 1;
    double sc = (double)nc / range;
    if (dtype == BSK_F32) sc = (double)(float)sc;
    if (!std::isfinite(sc) || !(sc > 0)) return;
    const double margin = (dtype == BSK_F32) ? 0.05 : 1e-6;
    lut.resize(nc);
    for (int c = 0; c < nc; ++c) {
        double e_lo = v[0] + ((double)c - margin) / sc;
        double e_hi = v[0] + ((double)c + 1.0 + margin) / sc;
        // lo = #{v < e_lo}, hi = #{v <= e_hi}
        int lo = (int)(std::lower_bound(v.begin(), v.end(), e_lo) - v.begin());
        int hi = (int)(std::upper_bound(v.begin(), v.end(), e_hi) - v.begin());
        if (c == 0) lo = 0;
        if (c == nc - 1) hi = n;
        int w = hi - lo;
        if (w > 254) w = 255;
        lut[c] = (uint32_t)lo | ((uint32_t)w << 24);
    }
    ncell = nc;
    scale = sc;
}

}  // namespace bsk

extern "C" {

bsk_status bsk_basis_create(int32_t kind, int32_t k, int32_t dtype, const void *h_knots,
                            int64_t nknots, const bsk_recomb_desc *recomb, int32_t device,
                            bsk_basis **out) {
    BSK_REQUIRE(out, BSK_ERR_ARGUMENT, "out is NULL");
    *out = nullptr;
    BSK_REQUIRE(kind >= 0 && kind <= 2, BSK_ERR_ARGUMENT, "unknown basis kind %d", kind);
    BSK_REQUIRE(k >= 1, BSK_ERR_ARGUMENT, "B-spline order must be k >= 1");   // basis.jl:80-82
    BSK_REQUIRE(k <= BSK_MAXK, BSK_ERR_UNSUPPORTED, "orders 1..%d are supported (got %d)",
                BSK_MAXK, k);
    BSK_REQUIRE(dtype == BSK_F32 || dtype == BSK_F64, BSK_ERR_ARGUMENT, "bad dtype");
    BSK_REQUIRE(h_knots, BSK_ERR_ARGUMENT, "knots is NULL");
    if (kind == BSK_BASIS_PERIODIC)
        BSK_REQUIRE(nknots >= 2, BSK_ERR_ARGUMENT, "periodic basis needs >= 2 breakpoints");
    else
        BSK_REQUIRE(nknots >= 2 * (int64_t)k, BSK_ERR_ARGUMENT, "knot vector too short: %lld < 2k",
                    (long long)nknots);
    BSK_REQUIRE(nknots < (1ll << 30), BSK_ERR_UNSUPPORTED, "too many knots");
    BSK_REQUIRE((kind == BSK_BASIS_RECOMBINED) == (recomb != nullptr), BSK_ERR_ARGUMENT,
                "recombination descriptor must be given for (and only for) a recombined basis");

    bsk_basis *b = new bsk_basis();
    b->kind = kind; b->k = k; b->dtype = dtype; b->device = device;
    b->nt = nknots;
    b->N = (kind == BSK_BASIS_PERIODIC) ? nknots - 1 : nknots - k;
    b->len = b->N;
    b->h_knots.resize(nknots);
    for (int64_t i = 0; i < nknots; ++i)
        b->h_knots[i] = dtype == BSK_F64 ? ((const double *)h_knots)[i] : (double)((const float *)h_knots)[i];
    b->d_knots = nullptr; b->d_lut = nullptr; b->ncell = 0; b->lut_scale = 0;
    memset(&b->rv, 0, sizeof(b->rv));

    if (kind == BSK_BASIS_RECOMBINED) {
        const bsk_recomb_desc *r = recomb;
        bool ok = r->cl >= 0 && r->cr >= 0 && r->nl >= 0 && r->nr >= 0 && r->ul_rows * r->nl <= 36 &&
                  r->lr_rows * r->nr <= 36 && (r->nl == 0 || r->ul) && (r->nr == 0 || r->lr);
        if (!ok) { delete b; set_error("bad recombination descriptor"); return BSK_ERR_ARGUMENT; }
        b->rv.enabled = 1;
        b->rv.cl = r->cl; b->rv.cr = r->cr; b->rv.nl = r->nl; b->rv.nr = r->nr;
        b->rv.ul_rows = r->ul_rows; b->rv.lr_rows = r->lr_rows;
        b->rv.M = (int)(b->N - r->cl - r->cr);
        for (int i = 0; i < r->ul_rows * r->nl; ++i) b->rv.ul[i] = r->ul[i];
        for (int i = 0; i < r->lr_rows * r->nr; ++i) b->rv.lr[i] = r->lr[i];
        b->len = b->rv.M;
    }

    // walk-back target of find_knot_interval (evaluate_all.jl:108-115)
    {
        int64_t i = nknots;   // 1-based
        double tlast = b->h_knots[nknots - 1];
        do { i -= 1; } while (i > 1 && b->h_knots[i - 1] == tlast);
        b->ilast = (int)i;
    }

    cudaError_t e = cudaSetDevice(device);
    size_t es = dtype == BSK_F64 ? 8 : 4;
    if (e == cudaSuccess) e = bsk::dev_malloc(&b->d_knots, es * nknots);
    if (e == cudaSuccess) e = cudaMemcpy(b->d_knots, h_knots, es * nknots, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        std::vector<uint32_t> lut;
        build_lut(b->h_knots, dtype, lut, b->ncell, b->lut_scale);
        if (b->ncell > 0) {
            e = bsk::dev_malloc((void **)&b->d_lut, sizeof(uint32_t) * lut.size());
            if (e == cudaSuccess)
                e = cudaMemcpy(b->d_lut, lut.data(), sizeof(uint32_t) * lut.size(), cudaMemcpyHostToDevice);
        }
    }
    if (e != cudaSuccess) {
        set_error("basis upload failed: %s", cudaGetErrorString(e));
        if (b->d_knots) bsk::dev_free(b->d_knots);
        if (b->d_lut) bsk::dev_free(b->d_lut);
        delete b;
        return BSK_ERR_CUDA;
    }
    *out = b;
    return BSK_OK;
}

bsk_status bsk_basis_destroy(bsk_basis *b) {
    if (!b) return BSK_OK;
    cudaSetDevice(b->device);
    if (b->d_knots) bsk::dev_free(b->d_knots);
    if (b->d_lut) bsk::dev_free(b->d_lut);
    delete b;
    return BSK_OK;
}

bsk_status bsk_basis_length(const bsk_basis *b, int64_t *len) {
    BSK_REQUIRE(b && len, BSK_ERR_ARGUMENT, "NULL argument");
    *len = b->len;
    return BSK_OK;
}

}  // extern "C"
