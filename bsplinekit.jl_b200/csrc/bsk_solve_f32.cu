// bsk_solve_f32.cu — the interpolation path in Float32: collocation assembly, BANFAC, multi-RHS
// BANSLV and the periodic-cubic cyclic solve for Float32 bases, points and data.
//
// The reference code is generic in the element type: `interpolate(x::Vector{Float32}, y, ...)`
// builds a CollocationMatrix{Float32} (src/Collocation/Collocation.jl:165-198) and runs the same
// lu! / ldiv! statements in Float32 (src/Collocation/matrix.jl:132-201,218-271); periodic cubic
// bases go through CyclicTridiagonalMatrix{Float32} (src/Collocation/cyclic.jl:112-197).  The
// kernels here keep that operation order with non-contracting Float32 arithmetic, so assembly,
// factors (the BANFAC kernels of banfac.cuh, shared with Float64) and solutions are bit-identical to the Float32 restatement of those statements (checked in tests/).
// The Float64 path has its own, more elaborate kernels (bsk_solve.cu, bsk_solve_win.cu).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "banfac.cuh"
#include "banslv.cuh"
#include "handles.cuh"

using namespace bsk;

struct bsk_cyclic_f32 {
    int device;
    int64_t n;
    // factor-once Thomas elimination of T = A - u v' (cyclic.jl:121-146): row multipliers, pivots
    float *d_w;           // n: w_i = a_i / b'_{i-1}   (w_1 unused)
    float *d_binv;        // n: 1 / b'_i
    float *d_c;           // n: super-diagonal
    float *d_q;           // n: T^{-1} u
    float vn, inv_denom;
};

namespace {

constexpr int kThreads = 256;

KnotView<float> make_view_f(const bsk_basis *b) {
    KnotView<float> kv;
    kv.t = (const float *)b->d_knots;
    kv.nt = (int)b->nt;
    kv.N = (int)b->N;
    kv.k = b->k;
    kv.kind = b->kind == BSK_BASIS_PERIODIC ? 1 : 0;
    kv.ilast = b->ilast;
    kv.a = (float)b->h_knots.front();
    kv.b = (float)b->h_knots.back();
    kv.period = (float)((float)b->h_knots.back() - (float)b->h_knots.front());
    kv.lut = b->d_lut;
    kv.ncell = b->ncell;
    kv.lut_scale = (float)b->lut_scale;
    return kv;
}

#define BSK_DISPATCH_K(kval, ...)                                  \
    switch (kval) {                                                \
        case 2: { constexpr int K = 2; __VA_ARGS__; } break;       \
        case 3: { constexpr int K = 3; __VA_ARGS__; } break;       \
        case 4: { constexpr int K = 4; __VA_ARGS__; } break;       \
        case 5: { constexpr int K = 5; __VA_ARGS__; } break;       \
        case 6: { constexpr int K = 6; __VA_ARGS__; } break;       \
        case 7: { constexpr int K = 7; __VA_ARGS__; } break;       \
        case 8: { constexpr int K = 8; __VA_ARGS__; } break;       \
        default: break;                                            \
    }

// collocation_matrix! (Collocation.jl:209-251), one thread per row
template <int K>
__global__ void __launch_bounds__(kThreads)
k_collocation_f32(KnotView<float> kv, RecombView rv, float dom_a, float dom_b, const float *__restrict__ x,
                  int nx, int deriv, float clip, float *__restrict__ band, int ncols,
                  unsigned long long *__restrict__ n_outside) {
    const int l = K - 2, u = l;
    const int nroww = l + u + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x + 1; i <= nx; i += gridDim.x * blockDim.x) {
        const float xi = x[i - 1];
        if (!(xi >= dom_a && xi <= dom_b)) continue;          // isindomain (basis.jl:229-234)
        float bq[K];
        int jlast = eval_all_point<K, 0, float, float>(kv, kv.t, xi, deriv, 0, bq);
        if (rv.enabled) {
            float ph[K];
            jlast = recombine_point<K, float, float>(rv, jlast, bq, ph);
#pragma unroll
            for (int j = 0; j < K; ++j) bq[j] = ph[j];
        }
#pragma unroll
        for (int dj = 1; dj <= K; ++dj) {
            const int j = jlast + 1 - dj;
            const float bj = bq[dj - 1];
            if (fabsf(bj) <= clip) continue;                  // Collocation.jl:237
            if ((i > j && i - j > l) || (i < j && j - i > u) || j < 1 || j > ncols) {
                atomicAdd(n_outside, 1ull);                   // reference: @warn (Collocation.jl:239-244)
                continue;
            }
            band[(u + i - j) + (size_t)(j - 1) * nroww] = bj;  // matrix.jl:105-117
        }
    }
}

// CyclicTridiagonalMatrix{Float32} (cyclic.jl:21-33, setindex! :77-92)
template <int K>
__global__ void __launch_bounds__(kThreads)
k_collocation_cyclic_f32(KnotView<float> kv, const float *__restrict__ x, int nx, int deriv, float clip,
                         float *__restrict__ a, float *__restrict__ b, float *__restrict__ c,
                         unsigned long long *__restrict__ n_outside) {
    const int n = kv.N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x + 1; i <= nx; i += gridDim.x * blockDim.x) {
        float bq[K];
        int jlast = eval_all_point<K, 1, float, float>(kv, kv.t, x[i - 1], deriv, 0, bq);
        const int im = (i == 1) ? n : i - 1, ip = (i == n) ? 1 : i + 1;
#pragma unroll
        for (int dj = 1; dj <= K; ++dj) {
            int j = jlast + 1 - dj;
            while (j < 1) j += n;                             // basis_to_array_index (periodic.jl:245-253)
            while (j > n) j -= n;
            const float bj = bq[dj - 1];
            if (fabsf(bj) <= clip) continue;
            if (i == j) b[i - 1] = bj;
            else if (j == im) a[i - 1] = bj;
            else if (j == ip) c[i - 1] = bj;
            else atomicAdd(n_outside, 1ull);
        }
    }
}

__global__ void k_repack_f32(const float *__restrict__ w, int bw, int n, float *__restrict__ lrow,
                             float *__restrict__ urow, float *__restrict__ diag) {
    const int nroww = 2 * bw + 1, middle = bw + 1;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x + 1; r <= n; r += gridDim.x * blockDim.x) {
        for (int j = 1; j <= bw; ++j) {
            lrow[(size_t)(r - 1) * bw + j - 1] = (r - j >= 1) ? w[(middle + j - 1) + (size_t)(r - j - 1) * nroww] : 0.0f;
            urow[(size_t)(r - 1) * bw + j - 1] = (r + j <= n) ? w[(middle - j - 1) + (size_t)(r + j - 1) * nroww] : 0.0f;
        }
        diag[r - 1] = w[(middle - 1) + (size_t)(r - 1) * nroww];
    }
}

// float factors -> double copies for the decay analysis (shared with the Float64 path)
__global__ void k_f32_to_f64(const float *__restrict__ in, double *__restrict__ out, int64_t n, int invert) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = invert ? 1.0 / (double)in[i] : (double)in[i];
}

template <int BW>
bsk_status solve_f32(const bsk_banded_f32 *h, float *d_rhs, int64_t nrhs, cudaStream_t st) {
    // one thread per column walks all n rows: small CTAs spread the columns evenly over the SMs
    // (65 536 columns = 1024 CTAs of 64 threads = 6.9 per SM) and 32 rows of register prefetch keep
    // enough 128-byte requests in flight per warp
    constexpr int kBlock = 64;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nrhs + kBlock - 1) / kBlock, 148 * 32));
    if (BW > 0) k_banslv_sweep<BW, 32, +1, float><<<grid, kBlock, 0, st>>>(h->d_lrow, h->d_diag, d_rhs, nrhs, nrhs, (int)h->n);
    k_banslv_sweep<BW, 32, -1, float><<<grid, kBlock, 0, st>>>(h->d_urow, h->d_diag, d_rhs, nrhs, nrhs, (int)h->n);
    BSK_CUDA(cudaGetLastError());
    return BSK_OK;
}

// Periodic cubic: Thomas sweeps with the factor-once elimination of T, then the Sherman-Morrison
// correction x = y - q (y_1 + v_n y_n) / (1 + q_1 + v_n q_n)   (cyclic.jl:152-158).
__global__ void __launch_bounds__(kThreads)
k_cyclic_f32(const float *__restrict__ w, const float *__restrict__ binv, const float *__restrict__ cc,
             const float *__restrict__ q, float vn, float inv_denom, float *__restrict__ rhs, int64_t nrhs,
             int n) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nrhs; c += (int64_t)gridDim.x * blockDim.x) {
        float *p = rhs + c;
        float prev = p[0];
        for (int i = 1; i < n; ++i) {                           // d_i -= w_i d_{i-1}
            prev = fmaf(-__ldg(w + i), prev, p[(int64_t)i * nrhs]);
            p[(int64_t)i * nrhs] = prev;
        }
        const float yn = prev * __ldg(binv + n - 1);
        p[(int64_t)(n - 1) * nrhs] = yn;
        prev = yn;
        for (int i = n - 2; i >= 0; --i) {                      // y_i = (d_i - c_i y_{i+1}) / b'_i
            prev = fmaf(-__ldg(cc + i), prev, p[(int64_t)i * nrhs]) * __ldg(binv + i);
            p[(int64_t)i * nrhs] = prev;
        }
        const float alpha = (prev + vn * yn) * inv_denom;       // prev == y_1
        for (int i = 0; i < n; ++i) p[(int64_t)i * nrhs] = fmaf(-alpha, __ldg(q + i), p[(int64_t)i * nrhs]);
    }
}

}  // namespace

extern "C" {

bsk_status bsk_collocation_matrix_f32(const bsk_basis *b, const float *d_x, int64_t nx, int32_t deriv,
                                      float clip_threshold, float *d_band, int64_t *h_n_outside, void *stream) {
    BSK_REQUIRE(b && d_x && d_band, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(b->dtype == BSK_F32, BSK_ERR_ARGUMENT, "bsk_collocation_matrix_f32 needs a Float32 basis");
    BSK_REQUIRE(b->kind != BSK_BASIS_PERIODIC, BSK_ERR_UNSUPPORTED,
                "periodic bases: use bsk_collocation_cyclic_f32 (k = 4)");
    BSK_REQUIRE(nx == b->len, BSK_ERR_ARGUMENT, "wrong dimensions of collocation matrix");   // Collocation.jl:214-216
    BSK_REQUIRE(deriv >= 0, BSK_ERR_ARGUMENT, "derivative order must be >= 0");
    BSK_REQUIRE(b->k >= 2, BSK_ERR_UNSUPPORTED, "order must be >= 2");
    BSK_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int nroww = 2 * (b->k - 2) + 1;
    BSK_CUDA(cudaMemsetAsync(d_band, 0, sizeof(float) * (size_t)nroww * (size_t)b->len, st));
    unsigned long long *d_cnt = nullptr;
    BSK_CUDA(bsk::dev_malloc((void **)&d_cnt, 8));
    BSK_CUDA(cudaMemsetAsync(d_cnt, 0, 8, st));
    KnotView<float> kv = make_view_f(b);
    const float dom_a = (float)b->h_knots[b->k - 1], dom_b = (float)b->h_knots[b->N];
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nx + kThreads - 1) / kThreads, 148 * 8));
    BSK_DISPATCH_K(b->k, (k_collocation_f32<K><<<grid, kThreads, 0, st>>>(kv, b->rv, dom_a, dom_b, d_x, (int)nx, deriv,
                                                                          clip_threshold, d_band, (int)b->len, d_cnt)))
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && h_n_outside) {
        unsigned long long cnt = 0;
        e = cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        *h_n_outside = (int64_t)cnt;
    }
    cudaFreeAsync(d_cnt, st);
    if (e != cudaSuccess) { set_error("collocation kernel failed: %s", cudaGetErrorString(e)); return BSK_ERR_CUDA; }
    return BSK_OK;
}

bsk_status bsk_collocation_cyclic_f32(const bsk_basis *b, const float *d_x, int64_t nx, int32_t deriv,
                                      float clip_threshold, float *d_a, float *d_b, float *d_c,
                                      int64_t *h_n_outside, void *stream) {
    BSK_REQUIRE(b && d_x && d_a && d_b && d_c, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(b->dtype == BSK_F32, BSK_ERR_ARGUMENT, "bsk_collocation_cyclic_f32 needs a Float32 basis");
    BSK_REQUIRE(b->kind == BSK_BASIS_PERIODIC, BSK_ERR_ARGUMENT, "basis is not periodic");
    BSK_REQUIRE(nx == b->len, BSK_ERR_ARGUMENT, "wrong dimensions of collocation matrix");
    BSK_REQUIRE(b->N >= 3, BSK_ERR_ARGUMENT, "cyclic tridiagonal matrix needs n >= 3");
    BSK_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    BSK_CUDA(cudaMemsetAsync(d_a, 0, 4 * (size_t)nx, st));
    BSK_CUDA(cudaMemsetAsync(d_b, 0, 4 * (size_t)nx, st));
    BSK_CUDA(cudaMemsetAsync(d_c, 0, 4 * (size_t)nx, st));
    unsigned long long *d_cnt = nullptr;
    BSK_CUDA(bsk::dev_malloc((void **)&d_cnt, 8));
    BSK_CUDA(cudaMemsetAsync(d_cnt, 0, 8, st));
    KnotView<float> kv = make_view_f(b);
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nx + kThreads - 1) / kThreads, 148 * 8));
    BSK_DISPATCH_K(b->k, (k_collocation_cyclic_f32<K><<<grid, kThreads, 0, st>>>(kv, d_x, (int)nx, deriv, clip_threshold,
                                                                                 d_a, d_b, d_c, d_cnt)))
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && h_n_outside) {
        unsigned long long cnt = 0;
        e = cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        *h_n_outside = (int64_t)cnt;
    }
    cudaFreeAsync(d_cnt, st);
    if (e != cudaSuccess) { set_error("collocation kernel failed: %s", cudaGetErrorString(e)); return BSK_ERR_CUDA; }
    return BSK_OK;
}

bsk_status bsk_banded_f32_destroy(bsk_banded_f32 *h) {
    if (!h) return BSK_OK;
    cudaSetDevice(h->device);
    for (float *p : {h->d_w, h->d_lrow, h->d_urow, h->d_diag})
        if (p) bsk::dev_free(p);
    if (h->d_lrec) bsk::dev_free(h->d_lrec);
    if (h->d_urec) bsk::dev_free(h->d_urec);
    delete h;
    return BSK_OK;
}

static bsk_status banded_f32_alloc(int64_t n, int32_t l, int32_t u, int32_t device, bsk_banded_f32 **out) {
    BSK_REQUIRE(out, BSK_ERR_ARGUMENT, "out is NULL");
    *out = nullptr;
    BSK_REQUIRE(n >= 1, BSK_ERR_ARGUMENT, "matrix is empty");                       // matrix.jl:145
    BSK_REQUIRE(l == u, BSK_ERR_ARGUMENT, "expected same number of lower and upper bands");   // matrix.jl:58-60
    BSK_REQUIRE(l >= 0 && l <= BSK_MAXK - 2, BSK_ERR_UNSUPPORTED, "bandwidths 0..%d are supported", BSK_MAXK - 2);
    BSK_REQUIRE(n < (1ll << 30), BSK_ERR_UNSUPPORTED, "matrix too large");
    BSK_CUDA(cudaSetDevice(device));
    bsk::ensure_pool(device);
    bsk_banded_f32 *h = new bsk_banded_f32();
    h->device = device; h->n = n; h->l = l; h->u = u; h->factored = false;
    h->d_w = h->d_lrow = h->d_urow = h->d_diag = nullptr;
    h->d_lrec = h->d_urec = nullptr; h->n_pad = 0; h->halo = 0; h->halo_f = 0;
    const size_t nw = (size_t)(l + u + 1) * n, nb = (size_t)std::max(1, l) * n;
    cudaError_t e = bsk::dev_malloc((void **)&h->d_w, 4 * nw);
    if (e == cudaSuccess) e = bsk::dev_malloc((void **)&h->d_lrow, 4 * nb);
    if (e == cudaSuccess) e = bsk::dev_malloc((void **)&h->d_urow, 4 * nb);
    if (e == cudaSuccess) e = bsk::dev_malloc((void **)&h->d_diag, 4 * n);
    if (e != cudaSuccess) {
        set_error("cudaMalloc failed: %s", cudaGetErrorString(e));
        bsk_banded_f32_destroy(h);
        return BSK_ERR_ALLOC;
    }
    *out = h;
    return BSK_OK;
}

bsk_status bsk_banded_f32_create(const float *d_band, int64_t n, int32_t l, int32_t u, int32_t device,
                                 bsk_banded_f32 **out) {
    BSK_REQUIRE(d_band, BSK_ERR_ARGUMENT, "NULL band data");
    bsk_status rc = banded_f32_alloc(n, l, u, device, out);
    if (rc != BSK_OK) return rc;
    BSK_CUDA(cudaMemcpy((*out)->d_w, d_band, 4 * (size_t)(l + u + 1) * n, cudaMemcpyDeviceToDevice));
    return BSK_OK;
}

bsk_status bsk_banded_f32_create_host(const float *h_band, int64_t n, int32_t l, int32_t u, int32_t device,
                                      bsk_banded_f32 **out) {
    BSK_REQUIRE(h_band, BSK_ERR_ARGUMENT, "NULL band data");
    bsk_status rc = banded_f32_alloc(n, l, u, device, out);
    if (rc != BSK_OK) return rc;
    BSK_CUDA(cudaMemcpy((*out)->d_w, h_band, 4 * (size_t)(l + u + 1) * n, cudaMemcpyHostToDevice));
    return BSK_OK;
}

bsk_status bsk_banded_f32_lu(bsk_banded_f32 *h, int64_t *info_pivot) {
    BSK_REQUIRE(h, BSK_ERR_ARGUMENT, "NULL handle");
    BSK_REQUIRE(!h->factored, BSK_ERR_ARGUMENT, "matrix is already factorised");
    BSK_CUDA(cudaSetDevice(h->device));
    long long *d_info = nullptr;
    BSK_CUDA(bsk::dev_malloc((void **)&d_info, 8));
    long long info = 0;
    cudaError_t e = bsk::launch_banfac<float>(h->d_w, h->l, h->u, (int)h->n, d_info, false);
    if (e == cudaSuccess) e = cudaMemcpy(&info, d_info, 8, cudaMemcpyDeviceToHost);
    bsk::dev_free(d_info);
    if (e != cudaSuccess) { set_error("banded LU failed: %s", cudaGetErrorString(e)); return BSK_ERR_CUDA; }
    if (info_pivot) *info_pivot = info;
    if (info != 0) {
        set_error("ZeroPivotException(%lld)", info);
        return BSK_ERR_ZERO_PIVOT;
    }
    const int n = (int)h->n;
    const int grid = std::max(1, std::min((n + kThreads - 1) / kThreads, 148 * 4));
    k_repack_f32<<<grid, kThreads>>>(h->d_w, h->l, n, h->d_lrow, h->d_urow, h->d_diag);
    BSK_CUDA(cudaGetLastError());
    // Decay halos for the single-pass solve (same analysis as the Float64 path, on double copies of the
    // Float32 factors).  Tolerance: the start-state error of a restarted recurrence must stay below
    // 2^-34 (6e-11) of the solution's magnitude, i.e. 2^-10 of a Float32 ulp.
    h->halo = h->halo_f = 0;
    const int bw = h->l;
    if (bw == 0) {
        h->halo = h->halo_f = 8;
    } else if (getenv("BSK_F32_NO_WINDOWED") == nullptr) {
        double *d_u = nullptr, *d_l = nullptr, *d_dinv = nullptr;
        cudaError_t ce = bsk::dev_malloc((void **)&d_u, 8 * (size_t)bw * n);
        if (ce == cudaSuccess) ce = bsk::dev_malloc((void **)&d_l, 8 * (size_t)bw * n);
        if (ce == cudaSuccess) ce = bsk::dev_malloc((void **)&d_dinv, 8 * (size_t)n);
        if (ce == cudaSuccess) {
            const int g2 = std::max(1, std::min((bw * n + kThreads - 1) / kThreads, 148 * 8));
            k_f32_to_f64<<<g2, kThreads>>>(h->d_urow, d_u, (int64_t)bw * n, 0);
            k_f32_to_f64<<<g2, kThreads>>>(h->d_lrow, d_l, (int64_t)bw * n, 0);
            k_f32_to_f64<<<g2, kThreads>>>(h->d_diag, d_dinv, (int64_t)n, 1);
            h->halo = bsk::decay_halo_device(d_u, d_dinv, n, bw, 0, 6e-11, 8, 448, 0);
            if (h->halo > 0) h->halo_f = bsk::decay_halo_device(d_l, nullptr, n, bw, 1, 6e-11, 8, 448, n % 8);
        }
        for (double *p : {d_u, d_l, d_dinv})
            if (p) bsk::dev_free(p);
    }
    if (h->halo > 0) {
        bsk_status rc = bsk::windowed_prepare_f32(h);
        if (rc != BSK_OK) return rc;
    }
    BSK_CUDA(cudaDeviceSynchronize());
    h->factored = true;
    return BSK_OK;
}

bsk_status bsk_banded_f32_halo(const bsk_banded_f32 *h, int32_t *halo) {
    BSK_REQUIRE(h && halo, BSK_ERR_ARGUMENT, "NULL argument");
    *halo = h->halo;
    return BSK_OK;
}

bsk_status bsk_banded_f32_data(const bsk_banded_f32 *h, float *h_out) {
    BSK_REQUIRE(h && h_out, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_CUDA(cudaSetDevice(h->device));
    BSK_CUDA(cudaMemcpy(h_out, h->d_w, 4 * (size_t)(h->l + h->u + 1) * h->n, cudaMemcpyDeviceToHost));
    return BSK_OK;
}

bsk_status bsk_banded_f32_solve(const bsk_banded_f32 *h, float *d_rhs, int64_t nrhs, int32_t algo, void *stream) {
    BSK_REQUIRE(h && (nrhs == 0 || d_rhs), BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(h->factored, BSK_ERR_ARGUMENT, "matrix is not factorised: call bsk_banded_f32_lu first");
    BSK_REQUIRE(algo >= 0 && algo <= 2, BSK_ERR_ARGUMENT, "unknown algo %d", algo);
    BSK_REQUIRE(nrhs >= 0, BSK_ERR_DIMENSION, "negative number of right-hand sides");
    if (nrhs == 0) return BSK_OK;
    BSK_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (algo == BSK_SOLVE_WINDOWED && h->halo == 0) {
        set_error("windowed solve unavailable: the decay bound of U^{-1} was not met for this matrix");
        return BSK_ERR_UNSUPPORTED;
    }
    // Single pass (AUTO when the halo was validated): column PAIRS per lane, 64 columns per warp; the
    // block must be 128-byte aligned with rows a multiple of 16 bytes apart (TMA tiles: nrhs % 4 == 0); columns
    // beyond the last full group of 64 go through the BANSLV-order kernels below.
    if ((algo == BSK_SOLVE_WINDOWED || (algo == BSK_SOLVE_AUTO && h->halo > 0)) && nrhs >= 64 && nrhs % 4 == 0 &&
        (((uintptr_t)d_rhs) & 127) == 0) {
        const int64_t ngroups = nrhs / 64;
        bsk_status rc = bsk::windowed_solve_f32(h, d_rhs, ngroups, nrhs / 2, st);
        if (rc == BSK_OK) {
            const int64_t done = ngroups * 64, rest = nrhs - done;
            if (rest == 0) return BSK_OK;
            constexpr int kBlock = 64;
            const int grid = (int)((rest + kBlock - 1) / kBlock);
#define BSK_F32_REST(BWV)                                                                                                      \
    case BWV:                                                                                                                  \
        if (BWV > 0) k_banslv_sweep<BWV, 32, +1, float><<<grid, kBlock, 0, st>>>(h->d_lrow, h->d_diag, d_rhs + done, rest, nrhs, (int)h->n); \
        k_banslv_sweep<BWV, 32, -1, float><<<grid, kBlock, 0, st>>>(h->d_urow, h->d_diag, d_rhs + done, rest, nrhs, (int)h->n); \
        break;
            switch (h->l) { BSK_F32_REST(0) BSK_F32_REST(1) BSK_F32_REST(2) BSK_F32_REST(3) BSK_F32_REST(4) BSK_F32_REST(5) BSK_F32_REST(6) }
#undef BSK_F32_REST
            BSK_CUDA(cudaGetLastError());
            return BSK_OK;
        }
        if (algo == BSK_SOLVE_WINDOWED || rc != BSK_ERR_UNSUPPORTED) return rc;
    } else if (algo == BSK_SOLVE_WINDOWED) {
        set_error("windowed Float32 solve needs >= 64 columns, a column count that is a multiple of 4 and a 128-byte aligned block");
        return BSK_ERR_UNSUPPORTED;
    }
    switch (h->l) {
        case 0: return solve_f32<0>(h, d_rhs, nrhs, st);
        case 1: return solve_f32<1>(h, d_rhs, nrhs, st);
        case 2: return solve_f32<2>(h, d_rhs, nrhs, st);
        case 3: return solve_f32<3>(h, d_rhs, nrhs, st);
        case 4: return solve_f32<4>(h, d_rhs, nrhs, st);
        case 5: return solve_f32<5>(h, d_rhs, nrhs, st);
        case 6: return solve_f32<6>(h, d_rhs, nrhs, st);
        default: break;
    }
    return BSK_ERR_UNSUPPORTED;
}

bsk_status bsk_cyclic_f32_destroy(bsk_cyclic_f32 *h) {
    if (!h) return BSK_OK;
    cudaSetDevice(h->device);
    for (float *p : {h->d_w, h->d_binv, h->d_c, h->d_q})
        if (p) bsk::dev_free(p);
    delete h;
    return BSK_OK;
}

bsk_status bsk_cyclic_f32_create(const float *h_a, const float *h_b, const float *h_c, int64_t n, int32_t device,
                                 bsk_cyclic_f32 **out) {
    BSK_REQUIRE(out, BSK_ERR_ARGUMENT, "out is NULL");
    *out = nullptr;
    BSK_REQUIRE(h_a && h_b && h_c, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(n >= 3, BSK_ERR_ARGUMENT, "cyclic tridiagonal matrix needs n >= 3");
    BSK_REQUIRE(n < (1ll << 30), BSK_ERR_UNSUPPORTED, "matrix too large");
    // The elimination of the matrix (cyclic.jl:121-146 and the matrix part of solve_thomas!
    // :168-197) does not depend on the data: done once, here, in Float32 like the reference.
    const float a1 = h_a[0], cn = h_c[n - 1];
    const float ac = a1 * cn;
    BSK_REQUIRE(ac > 0.0f, BSK_ERR_ARGUMENT, "a[1] * c[n] must be positive (cyclic.jl:127)");
    const float gamma = std::sqrt(ac);
    std::vector<float> b(h_b, h_b + n), q(n, 0.0f), w(n, 0.0f), binv(n), c(h_c, h_c + n);
    q[0] = gamma;
    q[n - 1] = cn;
    const float vn = a1 / gamma;
    b[0] -= gamma;
    b[n - 1] -= ac / gamma;
    for (int64_t i = 1; i < n; ++i) {
        if (b[i - 1] == 0.0f) { set_error("ZeroPivotException(%lld)", (long long)i); return BSK_ERR_ZERO_PIVOT; }
        w[i] = h_a[i] / b[i - 1];
        b[i] = b[i] - w[i] * c[i - 1];
        q[i] = q[i] - w[i] * q[i - 1];
    }
    if (b[n - 1] == 0.0f) { set_error("ZeroPivotException(%lld)", (long long)n); return BSK_ERR_ZERO_PIVOT; }
    q[n - 1] = q[n - 1] / b[n - 1];
    for (int64_t i = n - 2; i >= 0; --i) q[i] = (q[i] - c[i] * q[i + 1]) / b[i];
    for (int64_t i = 0; i < n; ++i) binv[i] = 1.0f / b[i];
    const float denom = 1.0f + (q[0] + vn * q[n - 1]);
    BSK_REQUIRE(denom != 0.0f, BSK_ERR_ZERO_PIVOT, "singular cyclic matrix");
    BSK_CUDA(cudaSetDevice(device));
    bsk::ensure_pool(device);
    bsk_cyclic_f32 *h = new bsk_cyclic_f32();
    h->device = device; h->n = n; h->vn = vn; h->inv_denom = 1.0f / denom;
    h->d_w = h->d_binv = h->d_c = h->d_q = nullptr;
    cudaError_t e = cudaSuccess;
    for (float **p : {&h->d_w, &h->d_binv, &h->d_c, &h->d_q})
        if (e == cudaSuccess) e = bsk::dev_malloc((void **)p, 4 * (size_t)n);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_w, w.data(), 4 * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_binv, binv.data(), 4 * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_c, c.data(), 4 * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_q, q.data(), 4 * (size_t)n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("cyclic handle: %s", cudaGetErrorString(e));
        bsk_cyclic_f32_destroy(h);
        return BSK_ERR_CUDA;
    }
    *out = h;
    return BSK_OK;
}

bsk_status bsk_cyclic_f32_solve(const bsk_cyclic_f32 *h, float *d_rhs, int64_t nrhs, void *stream) {
    BSK_REQUIRE(h && (nrhs == 0 || d_rhs), BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(nrhs >= 0, BSK_ERR_DIMENSION, "negative number of right-hand sides");
    if (nrhs == 0) return BSK_OK;
    BSK_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nrhs + kThreads - 1) / kThreads, 148 * 8));
    k_cyclic_f32<<<grid, kThreads, 0, st>>>(h->d_w, h->d_binv, h->d_c, h->d_q, h->vn, h->inv_denom, d_rhs, nrhs,
                                            (int)h->n);
    BSK_CUDA(cudaGetLastError());
    return BSK_OK;
}

}  // extern "C"
