// core.cuh — numerics shared by every kernel and by the host-side setup code of
// libbsplinekit_b200.  Everything here is __host__ __device__ so that the handful of
// O(N) set-up computations (local-polynomial tables, derivative coefficients) run the
// very same code as the kernels.
//
// "Exact" arithmetic: the bit-faithful paths use the round-to-nearest intrinsics
// (__dadd_rn & co.), which nvcc never contracts into FMAs, so that they follow the
// reference's operation order (Julia does not contract a*b+c) to the last bit.  The
// fast paths use fma() explicitly.
#pragma once
#include <cstdint>
#include <cmath>

// Also included by plain C++ translation units (the host-side table builder is unit-tested on CPU).
#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#endif

#define BSK_HD __host__ __device__ __forceinline__
#define BSK_MAXK 8

namespace bsk {

// ---- layout of the uniform-cell tables (pptable.h builds them, bsk_eval_cell.cu reads them) ----
// Tables staged in shared memory use the "jump" encoding of cells that contain a knot for Float64
// splines of order >= 3.  Odd orders have a spare slot in the 16-byte-padded record, which carries the
// flag and the side index; even orders (>= 4) carry them in low mantissa bits of the coefficients.
__host__ __device__ constexpr bool cell_jump_scheme(int elt_bytes, int k) { return elt_bytes == 8 && k >= 3; }
__host__ __device__ constexpr bool cell_jump_pad(int k) { return (k & 1) != 0; }     // RECV = k + 1 for doubles
// Record stride (in elements) of a cell table that stays in global memory: records longer than one
// 32-byte sector are padded to 64 bytes, so that a record is two whole sectors of one 128-byte line,
// fetched by two 256-bit loads (bsk_eval_cell.cu).
__host__ __device__ constexpr int cell_global_stride(int elt_bytes, int recv) {
    return (recv * elt_bytes > 32 && (recv * elt_bytes) % 64 != 0) ? ((recv * elt_bytes + 63) / 64) * 64 / elt_bytes : recv;
}

// ---- non-contracting arithmetic ------------------------------------------------------
BSK_HD double xadd(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
BSK_HD double xsub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
BSK_HD double xmul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
BSK_HD double xdiv(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
BSK_HD float xadd(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
BSK_HD float xsub(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
BSK_HD float xmul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
BSK_HD float xdiv(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}

// ---- knot view -------------------------------------------------------------------------
// kind 0: plain knot vector t[1..nt]           (BSplines/basis.jl:70-90)
// kind 1: PeriodicKnots over data[1..N+1]      (BSplines/periodic.jl:37-55); nt = N + 1
// The optional LUT accelerates searchsortedlast: a uniform grid of `ncell` cells over
// [t[1], t[nt]]; entry = lo | (min(hi - lo, 255) << 24) brackets the answer, and the
// bracket is then resolved by comparing against the actual knots, so the result is
// identical to a plain binary search whatever the rounding of the cell index.
template <typename KT>
struct KnotView {
    const KT *t;
    int nt;
    int N;        // number of basis functions of the (parent) basis
    int k;
    int kind;
    int ilast;    // plain: index returned when searchsortedlast == nt (walk-back target)
    KT a, b;      // plain: t[1], t[nt];  periodic: boundaries
    KT period;
    const uint32_t *lut;
    int ncell;
    KT lut_scale;
};

// searchsortedlast restricted to a bracket: number of v[j] (0-based j in [0,n)) that are
// "not greater" than x in the isless sense, given that it lies in [lo, hi].
template <typename KT>
BSK_HD int count_le(const KT *v, int lo, int hi, KT x) {
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (x < v[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// Base.searchsortedlast(v, x) for 1-based result (0 if x < v[1]); NaN -> n.
template <typename KT>
BSK_HD int search_last(const KT *v, int n, const uint32_t *lut, int ncell, KT scale, KT x) {
    if (!(x == x)) return n;
    if (x < v[0]) return 0;
    if (!(x < v[n - 1])) return n;
    int lo = 0, hi = n;
    if (lut != nullptr) {
        int c = (int)((x - v[0]) * scale);
        c = c < 0 ? 0 : (c >= ncell ? ncell - 1 : c);
        uint32_t e = lut[c];
        lo = (int)(e & 0xFFFFFFu);
        int w = (int)(e >> 24);
        if (w != 255) hi = lo + w;
    }
    return count_le(v, lo, hi, x);
}

// Periodic bases follow the reference's wrap loops (repeated +-L, one rounding per step) for points
// and knot indices up to kFarPeriods periods away from the main one, so that results there are the
// reference's bit for bit.  Farther out the reference itself needs that many sequential steps per
// point; there the shift is formed by ONE multiplication (documented deviation, results then agree
// with the reference to rounding only), and beyond 2^30 knot indices x is treated like +-Inf.
#define BSK_FAR_PERIODS 4096

// getindex(ts, i), i 1-based  (periodic: BSplines/periodic.jl:102-115)
template <int KIND, typename KT>
BSK_HD KT knot_at(const KnotView<KT> &kv, const KT *t, int i) {
    if (KIND == 0) return t[i - 1];
    KT x = 0;
    i -= kv.k / 2;
    const long long far = (long long)BSK_FAR_PERIODS * kv.N;
    if ((long long)i < 1 - far || (long long)i > (long long)kv.nt + far) {
        int m = (i - 1) / kv.N;                       // floor division
        if ((i - 1) % kv.N < 0) m -= 1;
        x = xmul((KT)m, kv.period);
        i -= m * kv.N;
    }
    while (i < 1) { x = xsub(x, kv.period); i += kv.N; }
    while (i > kv.nt) { x = xadd(x, kv.period); i -= kv.N; }
    return xadd(x, t[i - 1]);
}

// find_knot_interval (BSplines/evaluate_all.jl:102-119; periodic.jl:85-100)
template <int KIND, typename KT>
BSK_HD int find_interval(const KnotView<KT> &kv, const KT *t, KT x, int &zone) {
    if (KIND == 1) {
        int i = kv.k / 2;
        zone = 0;
        if (!(x == x)) return i + kv.nt;       // NaN passes both wrap loops; searchsortedlast(data, NaN) = length(data)
        if (!(fabs((double)x) <= 1.7e308)) {   // +-Inf: the reference loops forever
            return i + kv.N;
        }
        KT xw = x;
        // The reference wraps by repeated +-L (periodic.jl:88-97).  Beyond BSK_FAR_PERIODS periods we
        // jump most of the way first (the remaining steps are still the reference's loop).
        KT far = (KT)BSK_FAR_PERIODS * kv.period;
        if (xw < kv.a - far || xw > kv.b + far) {
            double m = floor(((double)xw - (double)kv.a) / (double)kv.period);
            double mm = m > 0 ? m - 1 : m + 1;
            if (!(fabs(mm) * (double)kv.N < 1073741824.0)) return i + kv.N;      // index out of range: like Inf
            xw = (KT)((double)xw - mm * (double)kv.period);
            i += (int)mm * kv.N;
        }
        while (xw < kv.a) { xw = xadd(xw, kv.period); i -= kv.N; }
        while (xw >= kv.b) { xw = xsub(xw, kv.period); i += kv.N; }
        i += search_last(t, kv.nt, kv.lut, kv.ncell, kv.lut_scale, xw);
        return i;
    }
    if (x < kv.a) { zone = -1; return 1; }
    int i = search_last(t, kv.nt, kv.lut, kv.ncell, kv.lut_scale, x);
    if (i == kv.nt) {
        zone = (kv.b < x) ? 1 : 0;
        return kv.ilast;
    }
    zone = 0;
    return i;
}

template <int KIND, typename KT>
BSK_HD int knot_zone(const KnotView<KT> &kv, KT x) {   // evaluate_all.jl:126; periodic.jl:81
    if (KIND == 1) return 0;
    return (x < kv.a) ? -1 : (x > kv.b) ? 1 : 0;
}

// ---- evaluate_all: all K non-zero B-splines (and derivatives) --------------------------
// Mirrors _evaluate_all (BSplines/evaluate_all.jl:128-204, steps :271-314): knot
// differences in KT, rounded to VT, triangle in VT; bs in REVERSED order.
template <typename KT, typename VT> struct Promote { typedef VT type; };
template <> struct Promote<double, float> { typedef double type; };

#ifndef BSK_KNOT_WINDOW_MAXK
#define BSK_KNOT_WINDOW_MAXK 6      // measured: k = 4 +5 %, k = 6 +5 %, k = 7 / 8 -8 % with the window in registers
#endif
template <int K, int KIND, typename KT, typename VT>
BSK_HD void eval_all_core(const KnotView<KT> &kv, const KT *t, KT x, int i, int deriv, VT *bq) {
    typedef typename Promote<KT, VT>::type PT;
    const int m = K - deriv;   // order of the values stage
    VT Ds[BSK_MAXK];
    bq[0] = (VT)1;
    // The 2 K - 2 knots t[i-K+2 .. i+K-1] every stage below draws from, read ONCE: left in place, each (q, j) pair
    // re-reads its two knots (K (K - 1) loads instead of 2 K - 2; the compiler does not merge them across the
    // division sequences), and those bank-conflicted gathers were what kept the L1 / shared pipe at 88 %.
    // (orders above BSK_KNOT_WINDOW_MAXK keep the per-use reads: 2 K - 2 more live doubles cost them occupancy)
    constexpr bool PRE = K <= BSK_KNOT_WINDOW_MAXK;
    KT tw[PRE && 2 * K - 2 > 0 ? 2 * K - 2 : 1];
    if constexpr (PRE) {
#pragma unroll
        for (int w = 0; w < 2 * K - 2; ++w) tw[w] = knot_at<KIND>(kv, t, i - K + 2 + w);
    }
#pragma unroll
    for (int q = 2; q <= K; ++q) {
        if (q <= m) {
#pragma unroll
            for (int j = 1; j <= q - 1; ++j) {
                KT ti = PRE ? tw[PRE ? K - 1 - j : 0] : knot_at<KIND>(kv, t, i - j + 1);        // t[i - j + 1]
                KT tj = PRE ? tw[PRE ? K - 2 - j + q : 0] : knot_at<KIND>(kv, t, i - j + q);    // t[i - j + q]
                Ds[j - 1] = (VT)xdiv(xsub(x, ti), xsub(tj, ti));
            }
            VT bp = bq[0];
            VT Dp = Ds[0];
            bq[0] = xmul(Dp, bp);
#pragma unroll
            for (int j = 2; j <= q - 1; ++j) {
                VT bpp = bp; bp = bq[j - 1];
                VT Dpp = Dp; Dp = Ds[j - 1];
                bq[j - 1] = xadd(xmul(Dp, bp), xmul(xsub((VT)1, Dpp), bpp));
            }
            bq[q - 1] = xmul(xsub((VT)1, Dp), bp);
        }
    }
    // derivative steps (_evaluate_step_deriv, :290-314)
#pragma unroll
    for (int q = 2; q <= K; ++q) {
        if (q > m) {
            const int p = q - 1;
            VT us[BSK_MAXK];
#pragma unroll
            for (int dj = 1; dj <= p; ++dj) {
                KT dt = PRE ? xsub(tw[PRE ? K - 2 + q - dj : 0], tw[PRE ? K - 1 - dj : 0])      // t[i + q - dj] - t[i + 1 - dj]
                            : xsub(knot_at<KIND>(kv, t, i + q - dj), knot_at<KIND>(kv, t, i + 1 - dj));
                us[dj - 1] = (VT)xdiv((PT)bq[dj - 1], (PT)dt);
            }
            bq[0] = xmul((VT)p, us[0]);
#pragma unroll
            for (int j = 2; j <= q - 1; ++j) bq[j - 1] = xmul((VT)p, xsub(us[j - 1], us[j - 2]));
            bq[q - 1] = xmul(-(VT)p, us[p - 1]);
        }
    }
}

// Full evaluate_all for one point: search (or ileft), zone handling, core.
// Returns the 1-based index; bq[0..K-1].
template <int K, int KIND, typename KT, typename VT>
BSK_HD int eval_all_point(const KnotView<KT> &kv, const KT *t, KT x, int deriv, long long ileft,
                          VT *bq) {
#pragma unroll
    for (int j = 0; j < K; ++j) bq[j] = (VT)0;
    if (deriv >= K) return 1;                  // evaluate_all.jl:180-183
    int zone, i;
    if (ileft != 0) { i = (int)ileft; zone = knot_zone<KIND>(kv, x); }
    else i = find_interval<KIND>(kv, t, x, zone);
    if (zone != 0) return i;                   // :137-139
    eval_all_core<K, KIND, KT, VT>(kv, t, x, i, deriv, bq);
    return i;
}

// ---- recombination view (Recombinations/matrices.jl:70-110, 377-452) -------------------
struct RecombView {
    int enabled;
    int cl, cr, nl, nr, ul_rows, lr_rows;
    int M;                 // length of the recombined basis
    double ul[36];         // column-major ul_rows x nl
    double lr[36];
};

BSK_HD double recomb_get(const RecombView &A, int i, int j, int block) {
    if (block == 2) return (i == j + A.cl) ? 1.0 : 0.0;
    if (block == 1) {
        if (i <= A.ul_rows) return A.ul[(i - 1) + (j - 1) * A.ul_rows];
        return 0.0;
    }
    int h = A.M - A.nr;
    int ii = i - h - A.cl;
    if (ii >= 1 && ii <= A.lr_rows) return A.lr[(ii - 1) + (j - h - 1) * A.lr_rows];
    return 0.0;
}

// evaluate_all(R::RecombinedBSplineBasis, ...)  (Recombinations/bases.jl:386-420)
template <int K, typename KT, typename VT>
BSK_HD int recombine_point(const RecombView &A, int ilast, const VT *bs, VT *phis) {
    typedef typename Promote<KT, VT>::type PT;
    const int jlast = ilast - A.cl;
#pragma unroll
    for (int dj = 1; dj <= K; ++dj) {
        VT acc = (VT)0;
        int j = jlast + 1 - dj;
        if (j >= 1 && j <= A.M) {
            int block = (j <= A.nl) ? 1 : (j > A.M - A.nr) ? 3 : 2;
            if (block == 2) {
                acc = bs[dj - 1];
            } else {
#pragma unroll
                for (int di = 1; di <= K; ++di) {
                    int i = ilast + 1 - di;
                    if (i <= 0) continue;
                    KT a = (KT)recomb_get(A, i, j, block);
                    acc = (VT)xadd((PT)acc, xmul((PT)a, (PT)bs[di - 1]));
                }
            }
        }
        phis[dj - 1] = acc;
    }
    return jlast;
}

// ---- de Boor evaluation of a spline (Splines/spline.jl:174-209) -------------------------
// periodic coefficient wrap: PeriodicVector (Splines/periodic.jl:48-62)
template <int KIND, typename T>
BSK_HD T coef_at(const T *c, int ncoef, long long stride, int i) {
    if (KIND == 1 && (i < 1 || i > ncoef)) {        // PeriodicVector: wrap into 1..N
        i = (i - 1) % ncoef;
        if (i < 0) i += ncoef;
        i += 1;
    }
    return c[(long long)(i - 1) * stride];
}

template <int K, int KIND, typename T>
BSK_HD T deboor_point(const KnotView<T> &kv, const T *t, const T *c, int ncoef, long long cstride,
                      int n, T x) {
    T d[BSK_MAXK];
#pragma unroll
    for (int j = 1; j <= K; ++j) d[j - 1] = xadd((T)0, coef_at<KIND>(c, ncoef, cstride, j + n - K));
    constexpr bool PRE = K <= BSK_KNOT_WINDOW_MAXK;
    T tw[PRE && 2 * K - 2 > 0 ? 2 * K - 2 : 1];       // t[n-K+2 .. n+K-1], read once (see eval_all_core)
    if constexpr (PRE) {
#pragma unroll
        for (int w = 0; w < 2 * K - 2; ++w) tw[w] = knot_at<KIND>(kv, t, n - K + 2 + w);
    }
#pragma unroll
    for (int r = 2; r <= K; ++r) {
#pragma unroll
        for (int j = K; j >= r; --j) {
            T ti = PRE ? tw[PRE ? j - 2 : 0] : knot_at<KIND>(kv, t, j - K + n);
            T tj = PRE ? tw[PRE ? j - r + K - 1 : 0] : knot_at<KIND>(kv, t, j - r + n + 1);
            T alpha = xdiv(xsub(x, ti), xsub(tj, ti));
            d[j - 1] = xadd(xmul(xsub((T)1, alpha), d[j - 2]), xmul(alpha, d[j - 1]));
        }
    }
    return d[K - 1];
}

// Runtime-order de Boor (same operation order as deboor_point).  Used on rare slow paths
// inside kernels (periodic points outside the main period) and by host-side set-up code.
// kv.k must be the order `k` (periodic knot offset k/2).
template <int KIND, typename T>
__host__ __device__ inline T deboor_generic(const KnotView<T> &kv, const T *t, const T *c, int ncoef,
                                            long long cstride, int k, int n, T x) {
    T d[BSK_MAXK];
    for (int j = 1; j <= k; ++j) d[j - 1] = xadd((T)0, coef_at<KIND>(c, ncoef, cstride, j + n - k));
    for (int r = 2; r <= k; ++r) {
        for (int j = k; j >= r; --j) {
            T ti = knot_at<KIND>(kv, t, j - k + n);
            T tj = knot_at<KIND>(kv, t, j - r + n + 1);
            T alpha = xdiv(xsub(x, ti), xsub(tj, ti));
            d[j - 1] = xadd(xmul(xsub((T)1, alpha), d[j - 2]), xmul(alpha, d[j - 1]));
        }
    }
    return d[k - 1];
}

}  // namespace bsk
