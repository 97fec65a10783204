// pp.cuh — per-knot-interval local-polynomial ("pp-form") table shared by the evaluation
// kernels.  Record r (REC values of T, 16-byte aligned): { t_lo, t_hi, c_0 .. c_{K-1}, pad };
// on [t_lo, t_hi):  S(x) = sum_m c_m u^m,  u = x - (t_lo + t_hi)/2.
#pragma once
#include "handles.cuh"

namespace bsk {

constexpr int kMaxOut = 4;

struct EvalOuts {
    void *out[kMaxOut];
    const void *dcoefs[kMaxOut];   // coefficients of Derivative(deriv) * S (periodic slow path)
    int deriv[kMaxOut];
    int nout;
    int extrap;                    // BSK_EXTRAP_*: what to return outside the domain
};

template <int K, typename T> struct PPRec {
    static constexpr int VPA = 16 / (int)sizeof(T);                 // values per 16 bytes
    static constexpr int REC = ((K + 2 + VPA - 1) / VPA) * VPA;     // record length in values
};

template <typename T>
struct PPView {
    const T *rec;
    const uint32_t *lut;     // cell -> bracket of record indices (same packing as the knot LUT)
    int nrec, ncell;
    T scale;
};

template <typename T>
inline PPView<T> make_ppview(const bsk_spline *s) {
    PPView<T> pv;
    pv.rec = (const T *)s->d_pp;
    pv.lut = s->d_pp_lut;
    pv.nrec = (int)s->pp_nrec;
    pv.ncell = s->pp_ncell;
    pv.scale = (T)s->pp_scale;
    return pv;
}

// falling factorial m (m-1) ... (m-d+1)
template <typename T>
__host__ __device__ __forceinline__ T falling(int m, int d) {
    T ff = (T)1;
#pragma unroll
    for (int q = 0; q < BSK_MAXK; ++q)
        if (q < d) ff *= (T)(m - q);
    return ff;
}

// Rare path of Linear extrapolation (SplineExtrapolations.jl:40-52): slope of D^d S at the boundary
// from the boundary polynomial piece, times the distance to the boundary.
template <int K, typename T>
__device__ __noinline__ T extrap_linear_term(const T *cm, T u, int d, T dx) {
    T acc = (T)0;
#pragma unroll
    for (int m = K - 1; m >= 0; --m)
        if (m >= d + 1) acc = fma(acc, u, cm[m] * falling<T>(m, d + 1));
    return acc * dx;
}

// Locate the record of xs (which must lie inside the domain) and return its coefficients
// and the local abscissa u.  `rec` / `lut` may point to shared or global memory.
template <int K, typename T>
__device__ __forceinline__ void pp_lookup(const PPView<T> &pv, T xs, T *cm, T *u) {
    constexpr int REC = PPRec<K, T>::REC;
    constexpr int VPA = 16 / (int)sizeof(T);
    const T *rec = pv.rec;
    int c = (int)((xs - rec[0]) * pv.scale);
    c = c < 0 ? 0 : (c >= pv.ncell ? pv.ncell - 1 : c);
    const uint32_t e = pv.lut[c];
    int lo = (int)(e & 0xFFFFFFu);
    const int w = (int)(e >> 24);
    int hi = (w != 255) ? lo + w : pv.nrec;
    hi = hi > pv.nrec ? pv.nrec : hi;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (xs < rec[mid * REC]) hi = mid; else lo = mid + 1;
    }
    int r = lo - 1;
    r = r < 0 ? 0 : (r >= pv.nrec ? pv.nrec - 1 : r);
    T buf[REC];
    const int4 *R4 = (const int4 *)(rec + (size_t)r * REC);
#pragma unroll
    for (int v = 0; v < REC / VPA; ++v) {
        int4 raw = R4[v];
        memcpy(&buf[v * VPA], &raw, 16);
    }
#pragma unroll
    for (int m = 0; m < K; ++m) cm[m] = buf[2 + m];
    *u = xs - (T)0.5 * (buf[0] + buf[1]);
}

// implemented in bsk_eval_cell.cu (one translation unit per element type)
template <typename T>
bsk_status launch_cell(const bsk_spline *s, const KnotView<T> &kv, const T *d_x, int64_t nvec,
                       const EvalOuts &outs, cudaStream_t st);

}  // namespace bsk
