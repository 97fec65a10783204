// bsk_multi.cu — vector-valued spline evaluation.
#include <algorithm>

#include <cub/device/device_radix_sort.cuh>

#include "handles.cuh"
#include "tma.cuh"

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
// knot interval cached in registers; when the interval advances by d < k the window slides and only
// d new coefficient rows are read.  The points are
// visited in the order of their knot interval (device radix sort of the interval indices, CUB), so
// that arbitrary point orders still re-use the cached coefficients; every output row is written
// where the caller expects it (out[p][:], a coalesced row), so no extra permutation pass is needed.
// ============================================================================================
namespace {

__global__ void k_iota(int32_t *p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (int32_t)i;
}

// V = columns (splines) per thread: 2 doubles / 4 floats = one 16-byte access per row when the
// coefficient block allows it (M % V == 0, 16-byte aligned), else 1.
template <int K, int KIND, typename T, int V>
__global__ void __launch_bounds__(256)
k_spline_multi(const int64_t *__restrict__ idx_sorted, const int32_t *__restrict__ perm,
               const T *__restrict__ bs, const T *__restrict__ coefs,
               int64_t ncoef, int64_t M, int64_t n, int64_t pts_per_block, T *__restrict__ out) {
    struct alignas(sizeof(T) * V) Vec { T v[V]; };
    const int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    const int64_t p0 = (int64_t)blockIdx.y * pts_per_block;
    const int64_t p1 = min(n, p0 + pts_per_block);
    const bool active = m < M;                          // inactive threads still help staging
    Vec c[K];
    int64_t cur = INT64_MIN;
    auto load_row = [&](int64_t r) -> Vec {             // coefficients of b_r for this thread's splines
        if (KIND == 1) r = ((r - 1) % ncoef + ncoef) % ncoef + 1;   // PeriodicVector wrap (Splines/periodic.jl:48-62)
        Vec z;
#pragma unroll
        for (int q = 0; q < V; ++q) z.v[q] = (T)0;
        return (r >= 1 && r <= ncoef) ? *reinterpret_cast<const Vec *>(coefs + (r - 1) * M + m) : z;
    };
    // The coefficient rows are consumed in increasing order (sorted points) but each one only when its
    // interval comes up, one dependent DRAM round trip per new row: pull the rows kPF ahead into L2.
    constexpr int kPF = 8;
    auto prefetch_row = [&](int64_t r) {
        if (KIND == 1) r = ((r - 1) % ncoef + ncoef) % ncoef + 1;
        if (r >= 1 && r <= ncoef) asm volatile("prefetch.global.L2 [%0];" ::"l"(coefs + (r - 1) * M + m));
    };
    // Per-point metadata (interval index, original position, the k basis values) is the same for every
    // column: the block stages it in shared memory 256 points at a time, so that the column loop only
    // waits on its own coefficient rows (the dependent idx -> perm -> bs loads were the critical path).
    __shared__ int64_t s_i[256];
    __shared__ int32_t s_p[256];
    __shared__ T s_b[256][K];
    for (int64_t q0 = p0; q0 < p1; q0 += 256) {
        const int nq = (int)min((int64_t)256, p1 - q0);
        __syncthreads();
        if ((int)threadIdx.x < nq) {
            const int64_t ps = q0 + threadIdx.x;
            const int32_t pp = perm[ps];
            s_i[threadIdx.x] = idx_sorted[ps];
            s_p[threadIdx.x] = pp;
#pragma unroll
            for (int j = 0; j < K; ++j) s_b[threadIdx.x][j] = bs[(int64_t)pp * K + j];
        }
        __syncthreads();
        if (!active) continue;
        for (int q = 0; q < nq; ++q) {
            const int64_t i = s_i[q];                   // 1-based index of b_i (uniform over the block)
            const int64_t p = s_p[q];                   // original position of this point
            if (i != cur) {
                const int64_t d = i - cur;
                if (cur != INT64_MIN && d > 0 && d < K) {
                    // sorted points: the window of k coefficient rows slides by d rows, only d are new
                    for (int64_t sft = 1; sft <= d; ++sft) {
#pragma unroll
                        for (int j = K - 1; j >= 1; --j) c[j] = c[j - 1];
                        c[0] = load_row(cur + sft);
                        prefetch_row(cur + sft + kPF);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < K; ++j) c[j] = load_row(i - j);
                    for (int sft = 1; sft <= kPF; ++sft) prefetch_row(i + sft);
                }
                cur = i;
            }
            Vec acc;
#pragma unroll
            for (int v = 0; v < V; ++v) acc.v[v] = (T)0;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const T bj = s_b[q][j];
#pragma unroll
                for (int v = 0; v < V; ++v) acc.v[v] = fma(bj, c[j].v[v], acc.v[v]);
            }
            *reinterpret_cast<Vec *>(out + p * M + m) = acc;
        }
    }
}

// Plain bases, 16-byte column groups: the coefficient rows a warp needs form one increasing run
// [rlo, rhi] (sorted points), so they are STREAMED: an elected lane keeps kRing rows (one
// cp.async.bulk of 32 * 16 B each, SASS UBLKCP, completion on an mbarrier per slot) in flight ahead of
// the row being consumed, in a warp-private shared-memory ring.  The dependent DRAM round trip per new
// row (37 % + 12 % of the stall samples of the register-window kernel above) disappears from the
// threads' critical path.
constexpr int kRing = 8;

template <int K, typename T, int V>
__global__ void __launch_bounds__(256)
k_spline_multi_ring(const int64_t *__restrict__ idx_sorted, const int32_t *__restrict__ perm,
                    const T *__restrict__ bs, const T *__restrict__ coefs,
                    int64_t ncoef, int64_t M, int64_t n, int64_t pts_per_block, T *__restrict__ out) {
    struct alignas(sizeof(T) * V) Vec { T v[V]; };
    static_assert(sizeof(Vec) == 16, "one 16-byte access per lane and row");
    extern __shared__ __align__(128) unsigned char dyn[];       // [8 warps][kRing][32] Vec | [8][kRing] mbarrier
    __shared__ int64_t s_i[256];
    __shared__ int32_t s_p[256];
    __shared__ T s_b[256][K];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t p0 = (int64_t)blockIdx.y * pts_per_block;
    const int64_t p1 = min(n, p0 + pts_per_block);
    if (p0 >= p1) return;
    const int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    const int64_t m_w = ((int64_t)blockIdx.x * blockDim.x + warp * 32) * V;
    const bool active = m < M;
    const unsigned wbytes = m_w < M ? (unsigned)(min((int64_t)32 * V, M - m_w) * sizeof(T)) : 0u;
    Vec *const ring = (Vec *)dyn + (size_t)warp * kRing * 32;
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(dyn) + 8u * kRing * 32u * 16u + (unsigned)warp * kRing * 8u;
    const int64_t rlo = idx_sorted[p0] - K + 1, rhi = idx_sorted[p1 - 1];
    const int64_t nrows = rhi - rlo + 1;
    if (lane == 0) {
#pragma unroll
        for (int sidx = 0; sidx < kRing; ++sidx) mbar_init(bar_s + 8 * sidx, 1);
        fence_proxy_async();
    }
    __syncwarp();
    auto issue = [&](int64_t rr) {                       // rr: offset of the row in [0, nrows)
        const unsigned bar = bar_s + 8u * (unsigned)(rr % kRing);
        const int64_t r = rlo + rr;
        if (elect_one()) {
            if (r >= 1 && r <= ncoef && wbytes > 0) {
                fence_proxy_async();
                mbar_expect_tx(bar, wbytes);
                bulk_g2s(ring_s + (unsigned)(rr % kRing) * 32u * 16u, coefs + (r - 1) * M + m_w, wbytes, bar);
            } else {
                mbar_arrive(bar);
            }
        }
    };
    for (int64_t rr = 0; rr < min((int64_t)kRing, nrows); ++rr) issue(rr);

    Vec c[K];
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int v = 0; v < V; ++v) c[j].v[v] = (T)0;
    int64_t consumed = 0;                                // rows rlo .. rlo + consumed - 1 are in the window
    for (int64_t q0 = p0; q0 < p1; q0 += 256) {
        const int nq = (int)min((int64_t)256, p1 - q0);
        __syncthreads();
        if ((int)threadIdx.x < nq) {
            const int64_t ps = q0 + threadIdx.x;
            const int32_t pp = perm[ps];
            s_i[threadIdx.x] = idx_sorted[ps];
            s_p[threadIdx.x] = pp;
#pragma unroll
            for (int j = 0; j < K; ++j) s_b[threadIdx.x][j] = bs[(int64_t)pp * K + j];
        }
        __syncthreads();
        for (int q = 0; q < nq; ++q) {
            const int64_t target = s_i[q] - rlo + 1;     // rows to have consumed so that c[0] is row i
            while (consumed < target) {                  // warp-uniform
                const int slot = (int)(consumed % kRing);
                mbar_wait(bar_s + 8u * slot, (unsigned)((consumed / kRing) & 1));
                const int64_t r = rlo + consumed;
#pragma unroll
                for (int j = K - 1; j >= 1; --j) c[j] = c[j - 1];
                Vec z;
#pragma unroll
                for (int v = 0; v < V; ++v) z.v[v] = (T)0;
                c[0] = (active && r >= 1 && r <= ncoef) ? ring[slot * 32 + lane] : z;
                ++consumed;
                __syncwarp();                            // every lane has its copy: the slot can be refilled
                if (consumed - 1 + kRing < nrows) issue(consumed - 1 + kRing);
            }
            if (active) {
                const int64_t p = s_p[q];
                Vec acc;
#pragma unroll
                for (int v = 0; v < V; ++v) acc.v[v] = (T)0;
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const T bj = s_b[q][j];
#pragma unroll
                    for (int v = 0; v < V; ++v) acc.v[v] = fma(bj, c[j].v[v], acc.v[v]);
                }
                *reinterpret_cast<Vec *>(out + p * M + m) = acc;
            }
        }
    }
    // drain: every issued copy must have landed before the CTA exits
    const int64_t issued = min(nrows, consumed + kRing);
    for (int64_t rr = consumed; rr < issued; ++rr) mbar_wait(bar_s + 8u * (unsigned)(rr % kRing), (unsigned)((rr / kRing) & 1));
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
    // Sort keys = knot-interval indices.  Plain bases: 1 .. nt, so bit_length(nt) key bits (two 8-bit passes for
    // C3's 4102 knots instead of five).  Periodic indices of points far from the main period can be large or
    // negative: 40 bits (negative ones sort after the positive ones, which only affects the visiting order).
    int key_bits = 40;
    if (b->kind != BSK_BASIS_PERIODIC) {
        key_bits = 1;
        while ((1ll << key_bits) <= b->nt + 1 && key_bits < 40) ++key_bits;
    }
    cudaError_t ae = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_idx, d_idx_s, d_perm, d_perm_s, (int)n, 0, key_bits, st);
    // ONE stream-ordered allocation carved into the six scratch arrays (a call on few points is a handful of
    // microsecond kernels: six allocations and six frees were a visible part of it)
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_idx = 0, o_idx_s = o_idx + up(sizeof(int64_t) * n), o_perm = o_idx_s + up(sizeof(int64_t) * n),
                 o_perm_s = o_perm + up(sizeof(int32_t) * n), o_bs = o_perm_s + up(sizeof(int32_t) * n),
                 o_tmp = o_bs + up(sizeof(T) * n * b->k), total = o_tmp + up(tmp_bytes ? tmp_bytes : 16);
    unsigned char *arena = nullptr;
    if (ae == cudaSuccess) ae = cudaMallocAsync((void **)&arena, total, st);
    if (ae == cudaSuccess) {
        d_idx = (int64_t *)(arena + o_idx);
        d_idx_s = (int64_t *)(arena + o_idx_s);
        d_perm = (int32_t *)(arena + o_perm);
        d_perm_s = (int32_t *)(arena + o_perm_s);
        d_bs = (T *)(arena + o_bs);
        d_tmp = arena + o_tmp;
    }
    bsk_status rc = BSK_OK;
    if (ae != cudaSuccess) { set_error("vector-valued evaluation: scratch allocation failed: %s", cudaGetErrorString(ae)); rc = BSK_ERR_ALLOC; }
    if (rc == BSK_OK) rc = bsk_evaluate_all(b, d_x, n, 0, b->dtype, nullptr, d_idx, d_bs, (void *)st);
    if (rc == BSK_OK) {
        k_iota<<<(unsigned)std::min<int64_t>(1024, (n + 255) / 256), 256, 0, st>>>(d_perm, n);
        cudaError_t ce = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_idx, d_idx_s, d_perm, d_perm_s, (int)n, 0, key_bits, st);
        if (ce != cudaSuccess) { set_error("radix sort failed: %s", cudaGetErrorString(ce)); rc = BSK_ERR_CUDA; }
    }
    if (rc == BSK_OK) {
        constexpr int VMAX = 16 / (int)sizeof(T);
        const bool vec = M % VMAX == 0 && ((((uintptr_t)d_coefs) | ((uintptr_t)d_out)) & 15) == 0;
        const int V = vec ? VMAX : 1;
        const int64_t col_blocks = (M / V + 255) / 256;
        // enough blocks to fill the machine, long enough point runs to reuse the cached coefficients
        int64_t want_y = std::max<int64_t>(1, (int64_t)sm_count(b->device) * 8 / col_blocks);
        int64_t ppb = std::max<int64_t>(32, (n + want_y - 1) / want_y);
        const int64_t gy = std::min<int64_t>(65535, (n + ppb - 1) / ppb);
        ppb = (n + gy - 1) / gy;
        dim3 grid((unsigned)col_blocks, (unsigned)gy);
        rc = BSK_ERR_UNSUPPORTED;
        if (b->kind == BSK_BASIS_PERIODIC) {
            if (vec) { BSK_DISPATCH_K(b->k, (k_spline_multi<K, 1, T, VMAX><<<grid, 256, 0, st>>>(d_idx_s, d_perm_s, d_bs, d_coefs, ncoef, M, n, ppb, d_out), rc = BSK_OK)) }
            else { BSK_DISPATCH_K(b->k, (k_spline_multi<K, 1, T, 1><<<grid, 256, 0, st>>>(d_idx_s, d_perm_s, d_bs, d_coefs, ncoef, M, n, ppb, d_out), rc = BSK_OK)) }
        } else {
            if (vec && getenv("BSK_MULTI_NO_RING") == nullptr) {
                const size_t dyn = (size_t)8 * kRing * 32 * 16 + (size_t)8 * kRing * 8;
                BSK_DISPATCH_K(b->k, {
                    auto kern = k_spline_multi_ring<K, T, VMAX>;
                    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
                    kern<<<grid, 256, dyn, st>>>(d_idx_s, d_perm_s, d_bs, d_coefs, ncoef, M, n, ppb, d_out);
                    rc = BSK_OK;
                })
            } else if (vec) { BSK_DISPATCH_K(b->k, (k_spline_multi<K, 0, T, VMAX><<<grid, 256, 0, st>>>(d_idx_s, d_perm_s, d_bs, d_coefs, ncoef, M, n, ppb, d_out), rc = BSK_OK)) }
            else { BSK_DISPATCH_K(b->k, (k_spline_multi<K, 0, T, 1><<<grid, 256, 0, st>>>(d_idx_s, d_perm_s, d_bs, d_coefs, ncoef, M, n, ppb, d_out), rc = BSK_OK)) }
        }
        (void)V;
        if (rc == BSK_OK && cudaGetLastError() != cudaSuccess) { set_error("k_spline_multi launch failed"); rc = BSK_ERR_CUDA; }
    }
    if (arena) cudaFreeAsync(arena, st);
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
