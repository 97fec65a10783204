// bsk_multi.cu — vector-valued spline evaluation.
#include <algorithm>

#include <cub/device/device_radix_sort.cuh>

#include "handles.cuh"

using namespace bsk;

#define BSK_DISPATCH_K(kval, ...)                                  \
    switch (kval) {                                                \
        case 1: { constexpr int K = 1; __VA_ARGS__; } break;       \
        case 2: { constexpr int K = 2; __VA_ARGS__; } break;       \
        case 3: { constexpr int K = 3; __VA_ARGS__; } break;       \
        case 4: { constexpr int K = 4; __VA_ARGS__; } break;       \
        case 5: { constexpr int K = 5; __VA_ARGS__; } break;       \
        case 6: { constexpr int K = 6; __VA_ARGS__; } break;       \
        case 7: { constexpr int K = 7; __VA_ARGS__; } break;       \
        case 8: { constexpr int K = 8; __VA_ARGS__; } break;       \
        default: break;                                            \
    }

// ============================================================================================
// Vector-valued splines: M splines sharing one basis (coefficients SVector{M,T}).
// Replaces (S::Spline)(x) with eltype(coefs) <: SVector (src/Splines/spline.jl:180-184: the de Boor
// kernel broadcasts over the coefficient elements; test/periodic.jl:8-23, test/interpolation.jl:57)
// broadcast over many points.  Two steps: K2 gives (i, b_i..b_{i-k+1}) per point, then a banded
// "SpMM": out[p][m] = sum_j bs[p][j] * C[i_p - j][m].  One thread per spline (column m, coalesced
// rows of C and of out), looping over a chunk of points with the k coefficients of the current
// knot interval cached in registers (re-read only when the interval changes).  The points are
// visited in the order of their knot interval (device radix sort of the interval indices, CUB), so
// that arbitrary point orders still re-use the cached coefficients; every output row is written
// where the caller expects it (out[p][:], a coalesced row), so no extra permutation pass is needed.
// ============================================================================================
namespace {

__global__ void k_iota(int32_t *p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (int32_t)i;
}

template <int K, int KIND, typename T>
__global__ void __launch_bounds__(256)
k_spline_multi(const int64_t *__restrict__ idx_sorted, const int32_t *__restrict__ perm,
               const T *__restrict__ bs, const T *__restrict__ coefs,
               int64_t ncoef, int64_t M, int64_t n, int64_t pts_per_block, T *__restrict__ out) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t p0 = (int64_t)blockIdx.y * pts_per_block;
    const int64_t p1 = min(n, p0 + pts_per_block);
    if (m >= M) return;
    T c[K];
    int64_t cur = INT64_MIN;
    for (int64_t ps = p0; ps < p1; ++ps) {
        const int64_t i = idx_sorted[ps];               // 1-based index of b_i (uniform over the block)
        const int64_t p = perm[ps];                     // original position of this point
        if (i != cur) {
#pragma unroll
            for (int j = 0; j < K; ++j) {
                int64_t r = i - j;                      // coefficient of b_{i-j}
                if (KIND == 1) {                        // PeriodicVector wrap (Splines/periodic.jl:48-62)
                    r = ((r - 1) % ncoef + ncoef) % ncoef + 1;
                }
                c[j] = (r >= 1 && r <= ncoef) ? coefs[(r - 1) * M + m] : (T)0;
            }
            cur = i;
        }
        T acc = (T)0;
#pragma unroll
        for (int j = 0; j < K; ++j) acc = fma(__ldg(bs + p * K + j), c[j], acc);
        out[p * M + m] = acc;
    }
}

template <typename T>
bsk_status spline_multi_t(const bsk_basis *b, const T *d_coefs, int64_t ncoef, int64_t M, const T *d_x, int64_t n,
                          T *d_out, cudaStream_t st) {
    BSK_REQUIRE(n < (1ll << 31), BSK_ERR_UNSUPPORTED, "at most 2^31 - 1 points per call");
    int64_t *d_idx = nullptr, *d_idx_s = nullptr;
    int32_t *d_perm = nullptr, *d_perm_s = nullptr;
    T *d_bs = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_idx, d_idx_s, d_perm, d_perm_s, (int)n, 0, 40, st);
    BSK_CUDA(cudaMallocAsync((void **)&d_idx, sizeof(int64_t) * n, st));
    BSK_CUDA(cudaMallocAsync((void **)&d_idx_s, sizeof(int64_t) * n, st));
    BSK_CUDA(cudaMallocAsync((void **)&d_perm, sizeof(int32_t) * n, st));
    BSK_CUDA(cudaMallocAsync((void **)&d_perm_s, sizeof(int32_t) * n, st));
    BSK_CUDA(cudaMallocAsync((void **)&d_bs, sizeof(T) * n * b->k, st));
    BSK_CUDA(cudaMallocAsync(&d_tmp, tmp_bytes ? tmp_bytes : 16, st));
    bsk_status rc = bsk_evaluate_all(b, d_x, n, 0, b->dtype, nullptr, d_idx, d_bs, (void *)st);
    if (rc == BSK_OK) {
        k_iota<<<(unsigned)std::min<int64_t>(1024, (n + 255) / 256), 256, 0, st>>>(d_perm, n);
        // interval indices fit in 40 bits (periodic indices can be negative: they sort after the
        // positive ones, which only affects the visiting order)
        cudaError_t ce = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_idx, d_idx_s, d_perm, d_perm_s, (int)n, 0, 40, st);
        if (ce != cudaSuccess) { set_error("radix sort failed: %s", cudaGetErrorString(ce)); rc = BSK_ERR_CUDA; }
    }
    if (rc == BSK_OK) {
        const int64_t col_blocks = (M + 255) / 256;
        // enough blocks to fill the machine, long enough point runs to reuse the cached coefficients
        int64_t want_y = std::max<int64_t>(1, (int64_t)sm_count(b->device) * 8 / col_blocks);
        int64_t ppb = std::max<int64_t>(32, (n + want_y - 1) / want_y);
        const int64_t gy = std::min<int64_t>(65535, (n + ppb - 1) / ppb);
        ppb = (n + gy - 1) / gy;
        dim3 grid((unsigned)col_blocks, (unsigned)gy);
        rc = BSK_ERR_UNSUPPORTED;
        if (b->kind == BSK_BASIS_PERIODIC) {
            BSK_DISPATCH_K(b->k, (k_spline_multi<K, 1, T><<<grid, 256, 0, st>>>(d_idx_s, d_perm_s, d_bs, d_coefs, ncoef, M, n, ppb, d_out), rc = BSK_OK))
        } else {
            BSK_DISPATCH_K(b->k, (k_spline_multi<K, 0, T><<<grid, 256, 0, st>>>(d_idx_s, d_perm_s, d_bs, d_coefs, ncoef, M, n, ppb, d_out), rc = BSK_OK))
        }
        if (rc == BSK_OK && cudaGetLastError() != cudaSuccess) { set_error("k_spline_multi launch failed"); rc = BSK_ERR_CUDA; }
    }
    cudaFreeAsync(d_idx, st);
    cudaFreeAsync(d_idx_s, st);
    cudaFreeAsync(d_perm, st);
    cudaFreeAsync(d_perm_s, st);
    cudaFreeAsync(d_bs, st);
    cudaFreeAsync(d_tmp, st);
    return rc;
}

}  // namespace

extern "C" bsk_status bsk_spline_eval_multi(const bsk_basis *b, const void *d_coefs, int64_t ncoef, int64_t M,
                                            const void *d_x, int64_t n, void *d_out, void *stream) {
    BSK_REQUIRE(b && d_coefs && (n == 0 || (d_x && d_out)), BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(b->kind != BSK_BASIS_RECOMBINED, BSK_ERR_UNSUPPORTED,
                "vector-valued evaluation needs a plain or periodic basis (convert to the parent basis first)");
    BSK_REQUIRE(ncoef == b->len, BSK_ERR_DIMENSION, "wrong number of coefficients: got %lld, basis has length %lld",
                (long long)ncoef, (long long)b->len);
    BSK_REQUIRE(M >= 1, BSK_ERR_ARGUMENT, "M must be >= 1");
    if (n == 0) return BSK_OK;
    BSK_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (b->dtype == BSK_F64)
        return spline_multi_t<double>(b, (const double *)d_coefs, ncoef, M, (const double *)d_x, n, (double *)d_out, st);
    return spline_multi_t<float>(b, (const float *)d_coefs, ncoef, M, (const float *)d_x, n, (float *)d_out, st);
}
