// bsk_eval_cell.cu — K3c: spline evaluation from a uniform-cell local-polynomial table staged
// in shared memory (the headline cubic-evaluation kernel).
//
// The knot domain [a, b] is cut into `ncell` uniform cells (ncell ~ 4-5x the number of knot
// intervals, sized so that the whole table fits in the 227 KB of shared memory of one SM).
// Cell c stores the Taylor coefficients, about the cell centre, of the polynomial piece that
// covers the cell's left edge.  The cell index is ONE multiply-add and a float->int conversion:
// the O(1) arithmetic lookup of the uniform-grid case, extended to non-uniform knots.
// Cells that contain a knot carry a flag (the least significant mantissa bit of the last
// slot), the knot itself and the index of a side-table entry {left piece, right piece}; the
// knot is then compared exactly, so the selected knot interval is exactly the one `searchsortedlast` finds
// (src/BSplines/evaluate_all.jl:102-119) whatever the rounding of the cell index.  Cells with
// two or more knots (clustered knots) fall back to the per-interval table in global memory.
// Float64 splines of order >= 3 staged in shared memory use the "jump" encoding instead: the
// flagged record keeps the piece left of the (simple) knot xi, the side entry is {xi, J} and the
// piece right of the knot is  P_L + J (x - xi)^(k-1)  — one 16-byte gather on the slow path instead
// of two, and a 4x smaller side table (more cells, fewer flagged lanes).  Odd orders carry the flag
// and the side index in the spare slot of the padded record (coefficients untouched), even orders in
// bits 1..3 of the first four coefficients.
//
// Replaces the body of (S::Spline)(x) broadcast over many points
// (src/Splines/spline.jl:144-209) and of (Derivative(n) * S).(x) (spline.jl:272-330) fused in
// the same pass.  Result differs from the reference's de Boor triangle only by rounding
// (few ulp; tests/test_gpu_parity.py bounds it at 1e-13 relative).
#include <algorithm>

#include "handles.cuh"
#include "pp.cuh"

using namespace bsk;

namespace {

constexpr int kCellThreads = 1024;

template <typename T> struct Bits;
template <> struct Bits<double> {
    static __device__ __forceinline__ int flag(double v) { return __double2loint(v) & 1; }
    // flagged cells: last slot = (side index << 1) | 1 as raw bits
    static __device__ __forceinline__ unsigned index(double v) { return ((unsigned)__double2loint(v)) >> 1; }
};
template <> struct Bits<float> {
    static __device__ __forceinline__ int flag(float v) { return __float_as_int(v) & 1; }
    static __device__ __forceinline__ unsigned index(float v) { return ((unsigned)__float_as_int(v)) >> 1; }
};
constexpr unsigned kMultiKnotCell = 0x7FFFFFFFu;     // cell with two or more knots -> interval table
constexpr unsigned kJumpMultiCell = 4095u;           // same, "jump" encoding (12 index bits)

// "jump" encoding: side index = bits 1..3 of the first four coefficients of the record
__device__ __forceinline__ unsigned jump_index(const double *rec) {
    return ((unsigned)(__double2loint(rec[0]) >> 1) & 7u) | (((unsigned)(__double2loint(rec[1]) >> 1) & 7u) << 3) |
           (((unsigned)(__double2loint(rec[2]) >> 1) & 7u) << 6) | (((unsigned)(__double2loint(rec[3]) >> 1) & 7u) << 9);
}
__device__ __forceinline__ unsigned jump_index(const float *) { return kJumpMultiCell; }   // never used

// acc + D^d of  J dx^(K-1)  =  acc + J (K-1)!/(K-1-d)! dx^(K-1-d).  Explicitly rounded products and
// one explicit fma, so that the compile-time (DMASK) and run-time derivative lists give the same bits.
template <int K, typename T>
__device__ __forceinline__ T add_jump_term(T acc, T jump, T dx, int d) {
    T pw = jump;
#pragma unroll
    for (int q = 0; q < K - 1; ++q)
        if (q < K - 1 - d) pw = xmul(pw, dx);
    return d > K - 1 ? acc : fma(pw, falling<T>(K - 1, d), acc);
}

template <int K, typename T> struct CellRec {
    static constexpr int VPA = 16 / (int)sizeof(T);
    static constexpr int RECV = ((K + VPA - 1) / VPA) * VPA;      // values per cell record
};

// One 32-byte record of a table that stays in global memory = one L2 sector = ONE 256-bit load
// (SASS LDG.E.ENL2.256, sm_100+): a gather of 32 random records costs the L1 pipe 32 wavefronts
// instead of the 64 of two 128-bit loads, and that pipe is what bounds the global-table kernels.
__device__ __forceinline__ void ldg256(const void *p, int4 &lo, int4 &hi) {
    // L1::no_allocate: the table is far larger than L1 (hit rate 0.6 %); not allocating lines for the gathers is
    // worth 2 % on the bare access pattern (bsk_probe_gather, variants 0 / 2)
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
}

template <typename T>
struct CellView {
    // Shared-memory tables are stored CHUNK-MAJOR: the w-th 16-byte chunk of record c sits at
    // [(w * count + c) * VEC], so that one LDS.128 of a warp spreads over all eight 16-byte
    // bank groups (a 32-byte record stride would only ever touch every other group).  Tables
    // that stay in global memory are RECORD-MAJOR instead: a 32-byte record is one L2 sector.
    // Unflagged cell: record = Taylor coefficients about the cell centre.  Flagged cell (a knot
    // inside): slot XI = the knot, last slot = (side index << 1) | 1; the two polynomial pieces
    // (left / right of the knot, both about the cell centre) are side[0][j] and side[1][j].
    const T *rec;             // (RECV/VEC) x ncell x VEC
    const T *rc;              // 2 x (RECV/VEC) x nside x VEC
    int ncell, nside;
    T x0, scale, cw;
    T a, b;                   // domain
};

// Horner evaluation of D^d of  sum_m c_m u^m  for the derivative orders selected at compile
// time by DMASK (bit d set -> output), or at run time (DMASK == 0).
template <int K, typename T, int DMASK>
__device__ __forceinline__ void horner_outputs(const T *cm, T u, const EvalOuts &outs, T *res) {
    if (DMASK != 0) {
        int o = 0;
#pragma unroll
        for (int d = 0; d < K; ++d) {
            if (DMASK & (1 << d)) {
                // falling factorial m!/(m-d)! as a compile-time constant
                T acc = (T)0;
#pragma unroll
                for (int m = K - 1; m >= d; --m) {
                    T ff = (T)1;
#pragma unroll
                    for (int q = 0; q < d; ++q) ff *= (T)(m - q);
                    acc = fma(acc, u, d == 0 ? cm[m] : cm[m] * ff);
                }
                res[o++] = acc;
            }
        }
    } else {
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
            if (o < outs.nout) {
                const int d = outs.deriv[o];
                T acc = (T)0;
#pragma unroll
                for (int m = K - 1; m >= 0; --m)
                    if (m >= d) acc = fma(acc, u, d == 0 ? cm[m] : cm[m] * falling<T>(m, d));
                res[o] = acc;
            }
        }
    }
}

template <int DMASK> struct MaskCount {
    static constexpr int value = (DMASK & 1) + ((DMASK >> 1) & 1) + ((DMASK >> 2) & 1) + ((DMASK >> 3) & 1) +
                                 ((DMASK >> 4) & 1) + ((DMASK >> 5) & 1) + ((DMASK >> 6) & 1) + ((DMASK >> 7) & 1);
};

// Rare path, kept out of line: per-interval table in global memory (pp.cuh).  This kernel is only
// launched for splines all of whose interval records are certified (bsk_eval.cu: splines with
// uncertified records go through k_spline_pp, which carries the de Boor route), so the record's flag
// is not looked at here.
template <int K, typename T>
__device__ __noinline__ bool fallback_interval(const PPView<T> pv, T xs, T *cm, T *u) {
    return pp_lookup<K, T>(pv, xs, cm, u);
}

// UNC = 1 kernels only (splines with interval records that are not certified, e.g. two knots 1e-9 apart): the
// point is evaluated by the bit-faithful de Boor recurrence, out of line.  Such intervals always sit in cells that
// take the slow path above (the table builder sends every cell that touches one to the interval table), so this
// costs the kernels of ordinary splines nothing: they are the UNC = 0 instantiations.
template <int KIND, typename T>
__device__ __noinline__ T deboor_exact_ool(KnotView<T> kv, int K, int d, const T *dcoefs, int ncoef, T xs, T xq) {
    return deboor_exact<KIND, T>(kv, K, d, dcoefs, ncoef, xs, xq);
}

// EX = 1: extrapolation methods compiled in (SplineExtrapolations.jl:22-71); the plain kernels carry
// none of that code.
template <int K, int KIND, typename T, int DMASK, int STAGE, int EX, int UNC = 0>
__global__ void __launch_bounds__(kCellThreads, 1)
k_spline_cell(CellView<T> cv, PPView<T> pv, KnotView<T> kv, int ncoef, const T *__restrict__ x,
              int64_t nvec, EvalOuts outs) {
    constexpr int RECV = CellRec<K, T>::RECV;
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int NOUT = DMASK ? MaskCount<DMASK>::value : kMaxOut;
    // global-memory tables: record stride RS (records longer than a sector are padded to 64 bytes)
    constexpr int RS = STAGE ? RECV : bsk::cell_global_stride((int)sizeof(T), RECV);
    constexpr bool WIDE = !STAGE && (RS * sizeof(T)) % 32 == 0;       // 256-bit record gathers
    constexpr bool JUMP = STAGE && bsk::cell_jump_scheme((int)sizeof(T), K);
    constexpr bool JPAD = JUMP && bsk::cell_jump_pad(K);            // flag + side index in the spare slot
    extern __shared__ __align__(16) unsigned char smem[];
    // Programmatic dependent launch: the next launch on the stream may be scheduled as soon as SMs fall free, so
    // that ITS table staging overlaps OUR tail; it touches points and results only after griddepcontrol.wait below
    // (= this grid has completed and its writes are visible).  A no-op for launches without the attribute.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // ---- tables: staged in shared memory (STAGE = 1), or, when they are too large for it, read
    //      in place from global memory through L1/L2 by the very same code (STAGE = 0) ----------
    T *const d_rec = (T *)smem;
    T *const d_rc = d_rec + (size_t)cv.ncell * RECV;
    if (STAGE) {
        const int4 *src = (const int4 *)cv.rec;
        int4 *dst = (int4 *)d_rec;
        const int n16 = cv.ncell * RECV / VEC;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
        src = (const int4 *)cv.rc;
        dst = (int4 *)d_rc;
        const int m16 = JUMP ? cv.nside : 2 * cv.nside * RECV / VEC;
        for (int i = threadIdx.x; i < m16; i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    // the tables were uploaded by the host long before; everything below depends on the previous kernel
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const T *const s_rec = STAGE ? d_rec : cv.rec;
    const T *const s_rc = STAGE ? d_rc : cv.rc;
    constexpr int XI = (K == 1) ? 1 : 0;                 // slot of the knot in a flagged record

    const T xa = cv.a, xb = cv.b, x0 = cv.x0, scale = cv.scale, cw = cv.cw;
    const int E = EX ? outs.extrap : 0;
    const int ncell1 = cv.ncell - 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int4 *x4 = (const int4 *)x;
    int4 raw_next;
    if (v < nvec) raw_next = x4[v];
    for (; v < nvec; v += stride) {
        const int4 raw = raw_next;
        if (v + stride < nvec) raw_next = x4[v + stride];     // software prefetch of the next vector
        T xv[VEC];
        memcpy(xv, &raw, 16);
        T res[NOUT][VEC];
        // phase 1: all cell-record gathers of this vector are issued before any of them is used
        T xs_[VEC], rec_[VEC][RS];
        int c_[VEC];
        bool out_[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
            if (KIND == 1 && E == BSK_EXTRAP_FLAT)      // periodic + Flat: clamp(x, a, b), then as usual
                xv[q] = xv[q] < xa ? xa : (xv[q] > xb ? xb : xv[q]);
            const T xq = xv[q];
            bool outside;                  // plain: zone != 0; periodic: outside the main period
            if (KIND == 1) outside = !(xq >= xa && xq < xb);
            else outside = !(xq >= xa && xq <= xb);
            // outside: evaluate at a (no extrapolation: the result is discarded), or at the nearer
            // boundary (Flat / Linear / Smooth, SplineExtrapolations.jl:22-71)
            const T xs = outside ? ((KIND == 0 && E != 0 && xq > xa) ? xb : xa) : xq;
            int c = (int)((xs - x0) * scale);
            c = c < 0 ? 0 : (c > ncell1 ? ncell1 : c);
            xs_[q] = xs;
            c_[q] = c;
            out_[q] = outside;
            // STAGE: chunk-major (bank spread); global: record-major (one 32-byte sector)
            const int4 *R4 = (const int4 *)s_rec + (STAGE ? (size_t)c : (size_t)c * (RS / VEC));
            const int cs = STAGE ? cv.ncell : 1;
            if (WIDE) {
#pragma unroll
                for (int w = 0; w < RS / VEC; w += 2) {
                    int4 lo, hi;
                    ldg256(R4 + w, lo, hi);
                    memcpy(&rec_[q][w * VEC], &lo, 16);
                    memcpy(&rec_[q][(w + 1) * VEC], &hi, 16);
                }
            } else {
#pragma unroll
                for (int w = 0; w < RECV / VEC; ++w) {
                    int4 r = R4[w * cs];
                    memcpy(&rec_[q][w * VEC], &r, 16);
                }
            }
        }
        // phase 2: cells holding a knot -> exact comparison with the knot, gather of the piece on
        // the right side of it (again all gathers of the vector in flight together)
        bool multi_[VEC];
        T jump_[VEC], dxk_[VEC];           // "jump" encoding: J and x - knot right of the knot, else 0
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
            const T last = rec_[q][JPAD ? RECV - 1 : K - 1];
            multi_[q] = false;
            jump_[q] = (T)0;
            dxk_[q] = (T)0;
            if (JUMP) {
                if (Bits<T>::flag(last)) {
                    const unsigned j = JPAD ? Bits<T>::index(last) : jump_index(rec_[q]);
                    if (j == (JPAD ? kMultiKnotCell : kJumpMultiCell)) {
                        multi_[q] = true;
                    } else {
                        const int4 r = ((const int4 *)s_rc)[j];
                        T kj[VEC];
                        memcpy(kj, &r, 16);
                        if (!(xs_[q] < kj[0])) {            // at or right of the knot: searchsortedlast
                            jump_[q] = kj[1];
                            dxk_[q] = xs_[q] - kj[0];
                        }
                    }
                }
            } else if (Bits<T>::flag(last)) {
                const unsigned j = Bits<T>::index(last);
                if (j == kMultiKnotCell) {
                    multi_[q] = true;
                } else {
                    const int right = (xs_[q] < rec_[q][XI]) ? 0 : 1;
                    const int4 *R4 = (const int4 *)s_rc + (size_t)right * (RS / VEC) * cv.nside +
                                     (STAGE ? (size_t)j : (size_t)j * (RS / VEC));
                    const int cs = STAGE ? cv.nside : 1;
                    if (WIDE) {
#pragma unroll
                        for (int w = 0; w < RS / VEC; w += 2) {
                            int4 lo, hi;
                            ldg256(R4 + w, lo, hi);
                            memcpy(&rec_[q][w * VEC], &lo, 16);
                            memcpy(&rec_[q][(w + 1) * VEC], &hi, 16);
                        }
                    } else {
#pragma unroll
                        for (int w = 0; w < RECV / VEC; ++w) {
                            int4 r = R4[w * cs];
                            memcpy(&rec_[q][w * VEC], &r, 16);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
            const T xq = xv[q];
            const bool isnan_ = !(xq == xq);
            const bool outside = out_[q];
            const T xs = xs_[q];
            T cm[K];
#pragma unroll
            for (int m = 0; m < K; ++m) cm[m] = rec_[q][m];
            T u = xs - fma((T)c_[q] + (T)0.5, cw, x0);
            const bool smooth_out = KIND == 0 && E == BSK_EXTRAP_SMOOTH && outside;
            bool exact = false;
            if (multi_[q]) {
                T cm2[K], u2;              // separate locals: keep cm[] itself in registers
                const bool flagged = fallback_interval<K, T>(pv, xs, cm2, &u2);
                if (UNC) exact = flagged;
#pragma unroll
                for (int m = 0; m < K; ++m) cm[m] = cm2[m];
                u = u2;
            }
            if (smooth_out) u += xq - xs;                     // boundary piece continued
            T r1[NOUT];
            horner_outputs<K, T, DMASK>(cm, u, outs, r1);
            if (JUMP) {
                if (DMASK != 0) {
                    int o = 0;
#pragma unroll
                    for (int d = 0; d < K; ++d)
                        if (DMASK & (1 << d)) { r1[o] = add_jump_term<K, T>(r1[o], jump_[q], dxk_[q], d); ++o; }
                } else {
#pragma unroll
                    for (int o = 0; o < NOUT; ++o)
                        if (o < outs.nout) r1[o] = add_jump_term<K, T>(r1[o], jump_[q], dxk_[q], outs.deriv[o]);
                }
            }
            if (UNC) {
                if (exact) {
#pragma unroll
                    for (int o = 0; o < NOUT; ++o)
                        if (DMASK != 0 || o < outs.nout)
                            r1[o] = deboor_exact_ool<KIND, T>(kv, K, outs.deriv[o], (const T *)outs.dcoefs[o], ncoef, xs,
                                                              smooth_out ? xq : xs);
                }
            }
#pragma unroll
            for (int o = 0; o < NOUT; ++o) {
                T val = r1[o];
                if (KIND == 0 && outside) {
                    const int side = xq > xa ? 1 : 0;
                    if (E == 0) val = (T)0;                 // Splines/spline.jl:164-168
                    else if (E == BSK_EXTRAP_FLAT) val = (T)outs.bval[o][side];
                    else if (E == BSK_EXTRAP_LINEAR)        // S(a) + S'(a) (x - a), SplineExtrapolations.jl:40-52
                        val = xadd((T)outs.bval[o][side], xmul((T)outs.bslope[o][side], xsub(xq, xs)));
                }
                if (isnan_) val = outs.template nan_value<T>(o);
                res[o][q] = val;
            }
            if (KIND == 1 && outside && !isnan_) {
                // outside the main period: the reference's own arithmetic (knots shifted by
                // multiples of L, x not wrapped; BSplines/periodic.jl:85-115)
#pragma unroll
                for (int o = 0; o < NOUT; ++o) {
                    if (DMASK != 0 || o < outs.nout) {
                        KnotView<T> kd = kv;
                        kd.k = K - outs.deriv[o];
                        int z;
                        int i = find_interval<1>(kd, kd.t, xq, z);
                        res[o][q] = deboor_generic<1, T>(kd, kd.t, (const T *)outs.dcoefs[o], ncoef, 1, kd.k, i, xq);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
            if (DMASK != 0 || o < outs.nout) {
                int4 rawo;
                memcpy(&rawo, res[o], 16);
                ((int4 *)outs.out[o])[v] = rawo;
            }
        }
    }
}

template <int K, int KIND, typename T, int DMASK, int EX = 0>
bsk_status launch_cell_mask(const bsk_spline *s, const CellView<T> &cv, const PPView<T> &pv,
                            const KnotView<T> &kv, const T *d_x, int64_t nvec, const EvalOuts &outs,
                            cudaStream_t st) {
    const size_t smem = s->cell_smem;                    // 0: tables stay in global memory
    const int sms = sm_count(s->basis->device);
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nvec + kCellThreads - 1) / kCellThreads, sms));
    // launched with programmatic stream serialization (see the kernel's griddepcontrol pair)
    static const bool pdl = getenv("BSK_NO_PDL") == nullptr;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kCellThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const int ncoef = (int)s->ncoef;
    if constexpr (EX == 0) {
        if (s->pp_uncertified > 0) {                     // the variant that carries the de Boor route
            if (smem > 0) {
                auto kern = k_spline_cell<K, KIND, T, DMASK, 1, 0, 1>;
                BSK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                BSK_CUDA(cudaLaunchKernelEx(&cfg, kern, cv, pv, kv, ncoef, d_x, nvec, outs));
            } else {
                auto kern = k_spline_cell<K, KIND, T, DMASK, 0, 0, 1>;
                BSK_CUDA(cudaLaunchKernelEx(&cfg, kern, cv, pv, kv, ncoef, d_x, nvec, outs));
            }
            return BSK_OK;
        }
    }
    if (smem > 0) {
        auto kern = k_spline_cell<K, KIND, T, DMASK, 1, EX>;
        BSK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BSK_CUDA(cudaLaunchKernelEx(&cfg, kern, cv, pv, kv, ncoef, d_x, nvec, outs));
    } else {
        auto kern = k_spline_cell<K, KIND, T, DMASK, 0, EX>;
        BSK_CUDA(cudaLaunchKernelEx(&cfg, kern, cv, pv, kv, ncoef, d_x, nvec, outs));
    }
    return BSK_OK;
}

template <int K, int KIND, typename T>
bsk_status launch_cell_k(const bsk_spline *s, const KnotView<T> &kv, const T *d_x, int64_t nvec,
                         const EvalOuts &outs, cudaStream_t st) {
    constexpr int VEC = 16 / (int)sizeof(T);
    CellView<T> cv;
    cv.rec = (const T *)s->d_cell_rec;
    cv.rc = (const T *)s->d_cell_rc;
    cv.ncell = s->cell_ncell;
    cv.nside = s->cell_nside;
    cv.x0 = (T)s->cell_x0;
    cv.scale = (T)s->cell_scale;
    cv.cw = (T)s->cell_cw;
    cv.a = (T)s->cell_x0;
    cv.b = (T)s->cell_b;
    (void)VEC;
    PPView<T> pv = make_ppview<T>(s);
    // derivative mask: outputs must be listed in increasing order for the compiled masks
    int mask = 0;
    bool increasing = true;
    for (int o = 0; o < outs.nout; ++o) {
        if (o > 0 && outs.deriv[o] <= outs.deriv[o - 1]) increasing = false;
        mask |= 1 << outs.deriv[o];
    }
    if (outs.extrap != 0) return launch_cell_mask<K, KIND, T, 0, 1>(s, cv, pv, kv, d_x, nvec, outs, st);
    if (increasing && mask == 1) return launch_cell_mask<K, KIND, T, 1>(s, cv, pv, kv, d_x, nvec, outs, st);
    if (increasing && mask == 3 && K >= 2) return launch_cell_mask<K, KIND, T, (K >= 2 ? 3 : 1)>(s, cv, pv, kv, d_x, nvec, outs, st);
    return launch_cell_mask<K, KIND, T, 0>(s, cv, pv, kv, d_x, nvec, outs, st);
}

}  // namespace

namespace bsk {

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

// Launch on `nvec` 16-byte vectors of points (d_x and every output 16-byte aligned).
template <typename T>
bsk_status launch_cell(const bsk_spline *s, const KnotView<T> &kv, const T *d_x, int64_t nvec,
                       const EvalOuts &outs, cudaStream_t st) {
    bsk_status rc = BSK_ERR_UNSUPPORTED;
    if (s->basis->kind == BSK_BASIS_PERIODIC) {
        BSK_DISPATCH_K(s->basis->k, rc = (launch_cell_k<K, 1, T>(s, kv, d_x, nvec, outs, st)))
    } else {
        BSK_DISPATCH_K(s->basis->k, rc = (launch_cell_k<K, 0, T>(s, kv, d_x, nvec, outs, st)))
    }
    return rc;
}

#ifdef BSK_CELL_F32
template bsk_status launch_cell<float>(const bsk_spline *, const KnotView<float> &, const float *, int64_t,
                                       const EvalOuts &, cudaStream_t);
#else
template bsk_status launch_cell<double>(const bsk_spline *, const KnotView<double> &, const double *, int64_t,
                                        const EvalOuts &, cudaStream_t);
#endif

}  // namespace bsk
