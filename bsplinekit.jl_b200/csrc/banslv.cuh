// banslv.cuh — K6a: multi-RHS BANSLV (src/Collocation/matrix.jl:218-271) in the reference's operation
// order, shared by the Float64 (bsk_solve.cu) and Float32 (bsk_solve_f32.cu) paths.
#pragma once
#include "handles.cuh"

namespace bsk {

// BANSLV (matrix.jl:218-271), one thread per right-hand side, rhs[r * nrhs + c].  The reference
// updates x[i+j] -= x[i] * w[middle+j, i] for j = 1..bw right after x[i] is final; per element that
// is the sequence of subtractions below, in the same order (oldest multiplier first).  One routine
// serves both sweeps: DIR = +1 walks the rows downwards with the rows of L (unit diagonal), DIR = -1
// upwards with the rows of U and a true division by the pivot.  Rows are prefetched U at a time into
// registers; addresses advance by pointer increments and full batches run without predication (the
// per-row instruction count is what bounds a single thread walking all n rows).
template <int BW>
__device__ __forceinline__ void load_rec(const double *rp, double (&l)[BW > 0 ? BW : 1]) {
    if (BW % 2 == 0 && BW > 0) {
#pragma unroll
        for (int j = 0; j < BW; j += 2) {
            const double2 v = __ldg((const double2 *)(rp + j));
            l[j] = v.x; l[(j + 1) % (BW > 0 ? BW : 1)] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < BW; ++j) l[j] = __ldg(rp + j);
    }
}

template <int BW>
__device__ __forceinline__ void load_rec(const float *rp, float (&l)[BW > 0 ? BW : 1]) {
    if (BW == 4) {
        const float4 v = __ldg((const float4 *)rp);
        l[0] = v.x; l[1 % (BW > 0 ? BW : 1)] = v.y; l[2 % (BW > 0 ? BW : 1)] = v.z; l[3 % (BW > 0 ? BW : 1)] = v.w;
    } else if (BW == 2) {
        const float2 v = __ldg((const float2 *)rp);
        l[0] = v.x; l[1 % (BW > 0 ? BW : 1)] = v.y;
    } else {
#pragma unroll
        for (int j = 0; j < BW; ++j) l[j] = __ldg(rp + j);
    }
}

template <int BW, int DIR, typename T>
__device__ __forceinline__ T row_step(T v, T (&st)[BW > 0 ? BW : 1], const T *rp, const T *dp) {
    T l[BW > 0 ? BW : 1];
    load_rec<BW>(rp, l);
#pragma unroll
    for (int j = BW; j >= 1; --j) v = xsub(v, xmul(st[j - 1], l[j - 1]));
    if (DIR < 0) v = xdiv(v, __ldg(dp));
#pragma unroll
    for (int j = BW - 1; j >= 1; --j) st[j] = st[j - 1];
    if (BW > 0) st[0] = v;
    return v;
}

template <int BW, int U, int DIR, typename T>
__global__ void __launch_bounds__(256)
k_banslv_sweep(const T *__restrict__ rec, const T *__restrict__ diag, T *__restrict__ rhs, int64_t nrhs, int64_t ld,
               int n) {
    const int64_t sd = DIR > 0 ? ld : -ld;                  // element stride between consecutive steps
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nrhs; c += (int64_t)gridDim.x * blockDim.x) {
        T st[BW > 0 ? BW : 1];
#pragma unroll
        for (int j = 0; j < (BW > 0 ? BW : 1); ++j) st[j] = (T)0;
        const int64_t first = DIR > 0 ? 0 : (int64_t)(n - 1);
        T *ps = rhs + c + first * ld;                  // element of the current step
        const T *pl = ps;                                // element of the next step to load
        const T *rp = rec + first * BW;
        const T *dp = diag + first;
        T cur[U], nxt[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
            cur[q] = q < n ? *pl : (T)0;
            pl += sd;
        }
        for (int s0 = 0; s0 < n; s0 += U) {
            if (s0 + 2 * U <= n) {                           // next batch complete: no predication
#pragma unroll
                for (int q = 0; q < U; ++q) { nxt[q] = *pl; pl += sd; }
            } else {
#pragma unroll
                for (int q = 0; q < U; ++q) { nxt[q] = (s0 + U + q < n) ? *pl : (T)0; pl += sd; }
            }
            if (s0 + U <= n) {
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    *ps = row_step<BW, DIR, T>(cur[q], st, rp, dp);
                    ps += sd; rp += DIR * BW; dp += DIR;
                }
            } else {
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    if (s0 + q < n) {
                        *ps = row_step<BW, DIR, T>(cur[q], st, rp, dp);
                        ps += sd; rp += DIR * BW; dp += DIR;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < U; ++q) cur[q] = nxt[q];
        }
    }
}

}  // namespace bsk
