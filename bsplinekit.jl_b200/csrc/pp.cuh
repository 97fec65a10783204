// pp.cuh — per-knot-interval local-polynomial ("pp-form") table shared by the evaluation
// kernels (built on the host by pptable.h).  Record r (REC values of T, 16-byte aligned):
// { t_lo, mid, c_0 .. c_{K-1}, pad };  on [t_lo, t_hi):  S(x) = sum_m c_m u^m,  u = x - mid.
// The least significant mantissa bit of `mid` is a flag: set = the Horner evaluation of this record
// is NOT certified to the tolerance (pptable.h); such points are evaluated by the bit-faithful
// de Boor recurrence (deboor_exact below).
#pragma once
#include "handles.cuh"

namespace bsk {

constexpr int kMaxOut = 4;

struct EvalOuts {
    void *out[kMaxOut];
    const void *dcoefs[kMaxOut];   // coefficients of Derivative(deriv) * S (de Boor routes inside the kernels)
    int deriv[kMaxOut];
    int nout;
    int extrap;                    // BSK_EXTRAP_*: what to return outside the domain
    // Flat / Linear extrapolation (plain bases): D^deriv S and its slope at the two boundaries, formed on
    // the host by the reference's own recurrences (de Boor on the differenced coefficients)
    double bval[kMaxOut][2], bslope[kMaxOut][2];
    // What a NaN point returns: NaN, except for D^(k-1) S — piecewise constant, so the reference's evaluation
    // never touches x once searchsortedlast has placed NaN past the last knot, and the constant of that piece
    // comes back (Splines/spline.jl:144-209 with BSplines/evaluate_all.jl:102-119, periodic.jl:85-100)
    double nanv[kMaxOut];
    float nanvf[kMaxOut];
    template <typename T> __host__ __device__ T nan_value(int o) const {
        if (sizeof(T) == 4) return (T)nanvf[o];
        return (T)nanv[o];
    }
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

template <typename T> __device__ __forceinline__ bool lsb_set(T v);
template <> __device__ __forceinline__ bool lsb_set<double>(double v) { return (__double2loint(v) & 1) != 0; }
template <> __device__ __forceinline__ bool lsb_set<float>(float v) { return (__float_as_int(v) & 1) != 0; }

// Rare path: D^d S in the reference's own arithmetic (knot search + de Boor on the coefficients of
// Derivative(d) * S; src/Splines/spline.jl:155-209, 293-330), for points whose interval record is not
// certified.  `xs` selects the knot interval, `xq` is where the piece is evaluated (they differ only for
// Smooth extrapolation).  kv = view of the spline's own basis (order K).
template <int KIND, typename T>
__device__ __noinline__ T deboor_exact(KnotView<T> kv, int K, int d, const T *dcoefs, int ncoef, T xs, T xq) {
    kv.k = K - d;
    kv.lut = nullptr;
    kv.ncell = 0;
    if (KIND == 0) {                 // knots of the derivative basis: t[1+d : end-d] (BSplines/basis.jl:104-111)
        kv.t += d;
        kv.nt -= 2 * d;
        kv.N -= d;
        kv.ilast -= d;
        ncoef -= d;
    }
    int z;
    const int i = find_interval<KIND>(kv, kv.t, xs, z);
    return deboor_generic<KIND, T>(kv, kv.t, dcoefs, ncoef, 1, kv.k, i, xq);
}

// Locate the record of xs (which must lie inside the domain) and return its coefficients
// and the local abscissa u; the return value is the record's "not certified" flag.
// `rec` / `lut` may point to shared or global memory.
template <int K, typename T>
__device__ __forceinline__ bool pp_lookup(const PPView<T> &pv, T xs, T *cm, T *u) {
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
    *u = xs - buf[1];
    return lsb_set<T>(buf[1]);
}

// implemented in bsk_eval_cell.cu (one translation unit per element type)
template <typename T>
bsk_status launch_cell(const bsk_spline *s, const KnotView<T> &kv, const T *d_x, int64_t nvec,
                       const EvalOuts &outs, cudaStream_t st);

}  // namespace bsk
