// bsk_eval.cu — K1 knot_interval, K2 evaluate_all, K3 spline_eval (de Boor and
// local-polynomial table), spline handles, Derivative(n) * S.
//
// Reference bodies replaced (paths relative to the BSplineKit.jl tree):
//   find_knot_interval      src/BSplines/evaluate_all.jl:102-126, periodic.jl:85-100
//   _evaluate_all           src/BSplines/evaluate_all.jl:128-204, 271-314
//   evaluate_all(::Recombined...) src/Recombinations/bases.jl:386-420
//   (S::Spline)(x)          src/Splines/spline.jl:144-209
//   Derivative(n) * S       src/Splines/spline.jl:272-330, Splines/periodic.jl:76-117
#include <algorithm>
#include <cstdlib>

#include "handles.cuh"
#include "pp.cuh"

using namespace bsk;

namespace {

constexpr int kThreads = 256;
constexpr size_t kSmemStageLimit = 200 * 1024;   // per-CTA budget for staged tables

template <typename KT>
KnotView<KT> make_view(const bsk_basis *b) {
    KnotView<KT> kv;
    kv.t = (const KT *)b->d_knots;
    kv.nt = (int)b->nt;
    kv.N = (int)b->N;
    kv.k = b->k;
    kv.kind = b->kind == BSK_BASIS_PERIODIC ? 1 : 0;
    kv.ilast = b->ilast;
    kv.a = (KT)b->h_knots.front();
    kv.b = (KT)b->h_knots.back();
    kv.period = (KT)((KT)b->h_knots.back() - (KT)b->h_knots.front());
    kv.lut = b->d_lut;
    kv.ncell = b->ncell;
    kv.lut_scale = (KT)b->lut_scale;
    return kv;
}

inline int grid_for(int64_t n, int device, int ctas_per_sm) {
    int64_t want = (n + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)sm_count(device) * ctas_per_sm;
    return (int)std::max<int64_t>(1, std::min(want, cap));
}

// Stage knots (+ LUT) into dynamic shared memory; returns the knot pointer to use.
template <typename KT>
__device__ __forceinline__ const KT *stage_knots(KnotView<KT> &kv, unsigned char *smem, int stage,
                                                 size_t &off) {
    if (!stage) return kv.t;
    KT *st = (KT *)(smem + off);
    for (int i = threadIdx.x; i < kv.nt; i += blockDim.x) st[i] = kv.t[i];
    off += ((size_t)kv.nt * sizeof(KT) + 15) & ~(size_t)15;
    if (kv.lut) {
        uint32_t *sl = (uint32_t *)(smem + off);
        for (int i = threadIdx.x; i < kv.ncell; i += blockDim.x) sl[i] = kv.lut[i];
        off += (size_t)kv.ncell * 4;
        kv.lut = sl;
    }
    return st;
}

template <typename KT>
size_t knots_smem_bytes(const bsk_basis *b) {
    return (((size_t)b->nt * sizeof(KT) + 15) & ~(size_t)15) + (size_t)b->ncell * 4;
}

// ---- K1 ---------------------------------------------------------------------------------
template <int KIND, typename KT>
__global__ void __launch_bounds__(kThreads)
k_find_interval(KnotView<KT> kv, const KT *__restrict__ x, int64_t n, int64_t *__restrict__ idx,
                int8_t *__restrict__ zone, int stage) {
    extern __shared__ __align__(16) unsigned char smem[];
    size_t off = 0;
    const KT *t = stage_knots(kv, smem, stage, off);
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (int64_t)gridDim.x * blockDim.x) {
        int z;
        int i = find_interval<KIND>(kv, t, x[p], z);
        idx[p] = i;
        if (zone) zone[p] = (int8_t)z;
    }
}

// ---- K2 ---------------------------------------------------------------------------------
template <int K, int KIND, typename KT, typename VT>
__global__ void __launch_bounds__(kThreads)
k_evaluate_all(KnotView<KT> kv, RecombView rv, const KT *__restrict__ x, int64_t n, int deriv,
               const int64_t *__restrict__ ileft, int64_t *__restrict__ idx, VT *__restrict__ bs,
               int stage) {
    extern __shared__ __align__(16) unsigned char smem[];
    size_t off = 0;
    const KT *t = stage_knots(kv, smem, stage, off);
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (int64_t)gridDim.x * blockDim.x) {
        VT bq[K];
        long long il = ileft ? ileft[p] : 0;
        if (rv.enabled && il != 0) il += rv.cl;          // bases.jl:393-395
        int i = eval_all_point<K, KIND, KT, VT>(kv, t, x[p], deriv, il, bq);
        if (rv.enabled) {
            VT ph[K];
            i = recombine_point<K, KT, VT>(rv, i, bq, ph);
#pragma unroll
            for (int j = 0; j < K; ++j) bq[j] = ph[j];
        }
        idx[p] = i;
#pragma unroll
        for (int j = 0; j < K; ++j) bs[p * K + j] = bq[j];
    }
}

// ---- K3a: de Boor, reference operation order ------------------------------------------------
template <int K, int KIND, typename T>
__global__ void __launch_bounds__(kThreads)
k_spline_deboor(KnotView<T> kv, const T *__restrict__ coefs, int ncoef, const T *__restrict__ x,
                int64_t n, T *__restrict__ out, int stage, int extrap) {
    extern __shared__ __align__(16) unsigned char smem[];
    size_t off = 0;
    const T *t = stage_knots(kv, smem, stage, off);
    const T *c = coefs;
    if (stage) {
        T *sc = (T *)(smem + off);
        for (int i = threadIdx.x; i < ncoef; i += blockDim.x) sc[i] = coefs[i];
        c = sc;
    }
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (int64_t)gridDim.x * blockDim.x) {
        T xv = x[p];
        if (extrap == BSK_EXTRAP_FLAT)                     // S(clamp(x, a, b)), SplineExtrapolations.jl:29-33
            xv = xv < kv.a ? kv.a : (xv > kv.b ? kv.b : xv);
        int z;
        int i = find_interval<KIND>(kv, t, xv, z);
        if (KIND == 0 && extrap == BSK_EXTRAP_SMOOTH && z != 0) {   // interval clamped, SplineExtrapolations.jl:60-71
            i = i < K ? K : (i > ncoef ? ncoef : i);
            z = 0;
        }
        T r = (T)0;
        // i < K can only happen for knot vectors without full end multiplicity, where the
        // reference reads coefficients out of bounds; we return 0 there.
        if (z == 0 && (KIND == 1 || (i >= K && i <= ncoef)))
            r = deboor_point<K, KIND, T>(kv, t, c, ncoef, 1, i, xv);
        out[p] = r;
    }
}

// ---- K3b: per-interval local-polynomial (pp) table (pp.cuh) -----------------------------------
// Used when the uniform-cell table (bsk_eval_cell.cu) is not available: staged in shared memory
// if it fits, otherwise served from global memory through L1/L2.
template <int K, int KIND, typename T, int PTS>
__global__ void __launch_bounds__(kThreads)
k_spline_pp(KnotView<T> kv, int ncoef, PPView<T> pv, const T *__restrict__ x, int64_t n,
            EvalOuts outs, int stage) {
    constexpr int REC = PPRec<K, T>::REC;
    extern __shared__ __align__(16) unsigned char smem[];
    if (stage) {
        T *sr = (T *)smem;
        const int4 *src = (const int4 *)pv.rec;
        int4 *dst = (int4 *)sr;
        const int n16 = pv.nrec * REC * (int)sizeof(T) / 16;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
        uint32_t *sl = (uint32_t *)(smem + (size_t)n16 * 16);
        for (int i = threadIdx.x; i < pv.ncell; i += blockDim.x) sl[i] = pv.lut[i];
        pv.rec = sr;
        pv.lut = sl;
    }
    __syncthreads();
    const T xa = kv.a, xb = kv.b;
    const int E = outs.extrap;

    const int64_t stride = (int64_t)gridDim.x * blockDim.x * PTS;
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PTS; base < n; base += stride) {
        T xv[PTS];
        const bool full = base + PTS <= n;
        if (full && PTS * sizeof(T) == 16) {
            int4 raw = *(const int4 *)(x + base);
            memcpy(xv, &raw, 16);
        } else {
#pragma unroll
            for (int q = 0; q < PTS; ++q) xv[q] = base + q < n ? x[base + q] : xa;
        }
        T res[kMaxOut][PTS];
#pragma unroll
        for (int q = 0; q < PTS; ++q) {
            if (KIND == 1 && E == BSK_EXTRAP_FLAT)
                xv[q] = xv[q] < xa ? xa : (xv[q] > xb ? xb : xv[q]);
            const T xq = xv[q];
            const bool isnan_ = !(xq == xq);
            bool outside;
            if (KIND == 1) outside = !(xq >= xa && xq < xb);
            else outside = !(xq >= xa && xq <= xb);
            const T xs = outside ? ((KIND == 0 && E != 0 && xq > xa) ? xb : xa) : xq;
            T cm[K];
            T u;
            pp_lookup<K, T>(pv, xs, cm, &u);
            if (KIND == 0 && E == BSK_EXTRAP_SMOOTH && outside) u += xq - xs;
#pragma unroll
            for (int o = 0; o < kMaxOut; ++o) {
                if (o < outs.nout) {
                    const int d = outs.deriv[o];
                    T acc = (T)0;
#pragma unroll
                    for (int m = K - 1; m >= 0; --m)
                        if (m >= d) acc = fma(acc, u, d == 0 ? cm[m] : cm[m] * falling<T>(m, d));
                    if (KIND == 0 && outside) {
                        if (E == 0) acc = (T)0;                 // Splines/spline.jl:164-168
                        else if (E == BSK_EXTRAP_LINEAR) {
                            T cml[K];                       // separate copy: cm[] itself stays in registers
#pragma unroll
                            for (int m = 0; m < K; ++m) cml[m] = cm[m];
                            acc += extrap_linear_term<K, T>(cml, u, d, xq - xs);
                        }
                    }
                    if (isnan_) acc = xq;
                    res[o][q] = acc;
                }
            }
            if (KIND == 1 && outside && !isnan_) {
                // Outside the main period: follow the reference exactly (knots shifted by
                // multiples of L, x NOT wrapped; BSplines/periodic.jl:85-115), using the
                // coefficient-differenced derivative splines.
#pragma unroll
                for (int o = 0; o < kMaxOut; ++o) {
                    if (o < outs.nout) {
                        KnotView<T> kd = kv;
                        kd.k = K - outs.deriv[o];
                        int z;
                        int i = find_interval<1>(kd, kd.t, xq, z);
                        res[o][q] = deboor_generic<1, T>(kd, kd.t, (const T *)outs.dcoefs[o], ncoef, 1,
                                                         kd.k, i, xq);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
            if (o < outs.nout) {
                T *op = (T *)outs.out[o];
                if (full && PTS * sizeof(T) == 16) {
                    int4 raw;
                    memcpy(&raw, res[o], 16);
                    *(int4 *)(op + base) = raw;
                } else {
#pragma unroll
                    for (int q = 0; q < PTS; ++q)
                        if (base + q < n) op[base + q] = res[o][q];
                }
            }
        }
    }
}

// ---- host-side helpers ---------------------------------------------------------------------

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

template <typename F>
cudaError_t set_smem(F func, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// Derivative coefficients, _derivative (Splines/spline.jl:293-330; periodic: Splines/periodic.jl:76-117)
// computed in T (the reference differentiates in the coefficient type).  t = stored knots
// (plain: full vector; periodic: breakpoints incl. endpoint).  Output: plain ncoef - nd, periodic ncoef.
template <typename T>
void derivative_coefs(int kind, int k, const std::vector<double> &tk, const std::vector<double> &c,
                      int nd, std::vector<double> &out) {
    const int64_t n = (int64_t)c.size();
    std::vector<T> du(n);
    if (kind == 0) {
        for (int64_t i = 0; i < n; ++i) du[i] = (T)c[i];
        for (int m = 1; m <= nd; ++m)
            for (int64_t i = n; i >= 1; --i) {
                T dt = (T)tk[i + k - m - 1] - (T)tk[i - 1];
                if (dt == (T)0 || i == 1) du[i - 1] = (T)0;
                else du[i - 1] = (T)(k - m) * (du[i - 1] - du[i - 2]) / dt;
            }
        out.resize(n - nd);
        for (int64_t i = 0; i < n - nd; ++i) out[i] = (double)du[i + nd];
    } else {
        // knots(B') : same breakpoints, order k - nd (BSplines/periodic.jl:255-260)
        std::vector<T> tt(tk.size());
        for (size_t i = 0; i < tk.size(); ++i) tt[i] = (T)tk[i];
        KnotView<T> kd;
        kd.t = tt.data(); kd.nt = (int)tt.size(); kd.N = (int)tt.size() - 1; kd.k = k - nd; kd.kind = 1;
        kd.ilast = 0; kd.a = tt.front(); kd.b = tt.back(); kd.period = tt.back() - tt.front();
        kd.lut = nullptr; kd.ncell = 0; kd.lut_scale = 0;
        const int64_t N = n;
        const int delta = k / 2 - (k - nd) / 2;
        for (int64_t i = 1; i <= N; ++i) {
            int64_t j = i - delta;
            while (j < 1) j += N;
            while (j > N) j -= N;
            du[j - 1] = (T)c[i - 1];
        }
        for (int m = 1; m <= nd; ++m) {
            T du0 = du[N - 1];
            for (int64_t i = N; i >= 1; --i) {
                T dt = knot_at<1>(kd, kd.t, (int)(i + k - m)) - knot_at<1>(kd, kd.t, (int)i);
                T prev = (i == 1) ? du0 : du[i - 2];
                du[i - 1] = (T)(k - m) * (du[i - 1] - prev) / dt;
            }
        }
        out.resize(N);
        for (int64_t i = 0; i < N; ++i) out[i] = (double)du[i];
    }
}

bsk_status clone_basis(const bsk_basis *b, int kind, int k, const std::vector<double> &knots,
                       bsk_basis **out) {
    if (b->dtype == BSK_F64)
        return bsk_basis_create(kind, k, b->dtype, knots.data(), (int64_t)knots.size(), nullptr, b->device, out);
    std::vector<float> kf(knots.begin(), knots.end());
    return bsk_basis_create(kind, k, b->dtype, kf.data(), (int64_t)kf.size(), nullptr, b->device, out);
}

bsk_status upload_coefs(bsk_spline *s) {
    const bsk_basis *b = s->basis;
    size_t es = b->dtype == BSK_F64 ? 8 : 4;
    BSK_CUDA(cudaSetDevice(b->device));
    if (!s->d_coefs) BSK_CUDA(bsk::dev_malloc(&s->d_coefs, es * std::max<int64_t>(1, s->ncoef)));
    if (b->dtype == BSK_F64) {
        BSK_CUDA(cudaMemcpy(s->d_coefs, s->h_coefs.data(), es * s->ncoef, cudaMemcpyHostToDevice));
    } else {
        std::vector<float> cf(s->h_coefs.begin(), s->h_coefs.end());
        BSK_CUDA(cudaMemcpy(s->d_coefs, cf.data(), es * s->ncoef, cudaMemcpyHostToDevice));
    }
    return BSK_OK;
}

void invalidate(bsk_spline *s) {
    for (bsk_spline *d : s->derivs)
        if (d) bsk_spline_destroy(d);
    s->derivs.clear();
    s->pp_valid = false;
    s->cell_valid = false;
}

// parent_coefficients (Recombinations/splines.jl:32-56): y = [ul * x[1:nl]; x[nl+1:H]; lr * x[H+1:N]]
template <typename T>
void parent_coefficients(const RecombView &rv, const std::vector<double> &x, std::vector<double> &y) {
    const int N = (int)x.size();
    const int Na = rv.nl, Nc = rv.nr, H = N - Nc;
    y.clear();
    for (int i = 0; i < rv.ul_rows; ++i) {
        T acc = (T)0;
        for (int j = 0; j < Na; ++j) {
            T term = (T)rv.ul[i + j * rv.ul_rows] * (T)x[j];
            acc = (j == 0) ? term : acc + term;
        }
        y.push_back((double)acc);
    }
    for (int i = Na; i < H; ++i) y.push_back(x[i]);
    for (int i = 0; i < rv.lr_rows; ++i) {
        T acc = (T)0;
        for (int j = 0; j < Nc; ++j) {
            T term = (T)rv.lr[i + j * rv.lr_rows] * (T)x[H + j];
            acc = (j == 0) ? term : acc + term;
        }
        y.push_back((double)acc);
    }
}

// set host coefficients given in the USER basis (recombined -> parent)
bsk_status set_host_coefs(bsk_spline *s, const bsk_basis *user_basis, const std::vector<double> &c) {
    if (user_basis->kind == BSK_BASIS_RECOMBINED) {
        if (user_basis->dtype == BSK_F64) parent_coefficients<double>(user_basis->rv, c, s->h_coefs);
        else parent_coefficients<float>(user_basis->rv, c, s->h_coefs);
    } else {
        s->h_coefs = c;
    }
    BSK_REQUIRE((int64_t)s->h_coefs.size() == s->ncoef, BSK_ERR_DIMENSION,
                "internal: parent coefficient count %lld != %lld", (long long)s->h_coefs.size(),
                (long long)s->ncoef);
    invalidate(s);
    return upload_coefs(s);
}

// ---- local-polynomial table ------------------------------------------------------------------
bool pp_eligible(const bsk_spline *s) {
    const bsk_basis *b = s->basis;
    if (b->kind == BSK_BASIS_PERIODIC) return b->N >= 1;
    // plain: need full multiplicity k at both ends so that [t1, tend] == [t_k, t_{N+1}]
    const auto &t = b->h_knots;
    return t[0] == t[b->k - 1] && t[b->N] == t[b->nt - 1] && t[0] < t[b->nt - 1];
}

// Taylor coefficients of one polynomial piece of the spline about an arbitrary point, from the
// de Boor evaluation of the coefficient-differenced derivative splines (all in double).
struct PieceTaylor {
    const bsk_basis *b;
    int k, kind;
    std::vector<std::vector<double>> dc;     // dc[m] = coefficients of D^m S
    explicit PieceTaylor(const bsk_spline *s) : b(s->basis), k(s->basis->k),
                                                kind(s->basis->kind == BSK_BASIS_PERIODIC ? 1 : 0), dc(s->basis->k) {
        dc[0] = s->h_coefs;
        for (int m = 1; m < k; ++m) derivative_coefs<double>(kind, k, b->h_knots, s->h_coefs, m, dc[m]);
    }
    // piece = 1-based knot-interval index i (plain) or breakpoint index s (periodic)
    void operator()(int piece, double xpt, double *c) const {
        double fact = 1.0;
        for (int m = 0; m < k; ++m) {
            if (m > 0) fact *= (double)m;
            KnotView<double> kd;
            kd.kind = kind; kd.k = k - m; kd.lut = nullptr; kd.ncell = 0; kd.lut_scale = 0; kd.ilast = 0;
            double val;
            if (kind == 0) {
                // knots of the D^m basis: t[1+m : end-m]; interval index i - m (basis.jl:104-111)
                kd.t = b->h_knots.data() + m; kd.nt = (int)b->nt - 2 * m; kd.N = kd.nt - kd.k;
                kd.a = kd.t[0]; kd.b = kd.t[kd.nt - 1]; kd.period = 0;
                val = deboor_generic<0, double>(kd, kd.t, dc[m].data(), (int)dc[m].size(), 1, kd.k, piece - m, xpt);
            } else {
                kd.t = b->h_knots.data(); kd.nt = (int)b->nt; kd.N = (int)b->N;
                kd.a = kd.t[0]; kd.b = kd.t[kd.nt - 1]; kd.period = kd.b - kd.a;
                val = deboor_generic<1, double>(kd, kd.t, dc[m].data(), (int)dc[m].size(), 1, kd.k,
                                                piece + kd.k / 2, xpt);
            }
            c[m] = val / fact;
        }
    }
};

// non-empty polynomial pieces of the spline: left ends, right ends, piece ids
void list_pieces(const bsk_basis *b, std::vector<double> &lo, std::vector<double> &hi, std::vector<int> &id) {
    const auto &t = b->h_knots;
    if (b->kind != BSK_BASIS_PERIODIC) {
        for (int64_t i = b->k; i <= b->N; ++i)
            if (t[i - 1] < t[i]) { lo.push_back(t[i - 1]); hi.push_back(t[i]); id.push_back((int)i); }
    } else {
        for (int64_t sidx = 1; sidx <= b->N; ++sidx)
            if (t[sidx - 1] < t[sidx]) { lo.push_back(t[sidx - 1]); hi.push_back(t[sidx]); id.push_back((int)sidx); }
    }
}

template <typename T>
bsk_status build_pp(bsk_spline *s) {
    bsk_basis *b = s->basis;
    const int k = b->k;
    const int VPA = 16 / (int)sizeof(T);
    const int REC = ((k + 2 + VPA - 1) / VPA) * VPA;
    PieceTaylor taylor(s);
    std::vector<double> plo, phi;
    std::vector<int> pid;
    list_pieces(b, plo, phi, pid);
    const int64_t nrec = (int64_t)plo.size();
    BSK_REQUIRE(nrec >= 1, BSK_ERR_ARGUMENT, "spline has no non-empty knot interval");
    std::vector<T> recs((size_t)nrec * REC, (T)0);
    std::vector<double> xi(nrec + 1);   // record left ends + right end of the domain
    for (int64_t r = 0; r < nrec; ++r) {
        T tlo = (T)plo[r], thi = (T)phi[r];
        T midT = (T)0.5 * (tlo + thi);
        double c[BSK_MAXK];
        taylor(pid[r], (double)midT, c);
        recs[r * REC] = tlo; recs[r * REC + 1] = thi;
        for (int m = 0; m < k; ++m) recs[r * REC + 2 + m] = (T)c[m];
        xi[r] = (double)tlo;
    }
    xi[nrec] = b->h_knots[b->nt - 1];
    // LUT over the record left ends (same bracket construction as the knot LUT, bsk_api.cu)
    std::vector<uint32_t> lut;
    int ncell = 0;
    double scale = 0;
    {
        const int n = (int)xi.size();
        double range = xi[n - 1] - xi[0];
        int nc = 64;
        while (nc < 4 * n && nc < 65536) nc <<= 1;
        double sc = (double)nc / range;
        if (sizeof(T) == 4) sc = (double)(float)sc;
        const double margin = sizeof(T) == 4 ? 0.05 : 1e-6;
        lut.resize(nc);
        for (int c = 0; c < nc; ++c) {
            double e_lo = xi[0] + ((double)c - margin) / sc;
            double e_hi = xi[0] + ((double)c + 1.0 + margin) / sc;
            int lo = (int)(std::lower_bound(xi.begin(), xi.end(), e_lo) - xi.begin());
            int hi = (int)(std::upper_bound(xi.begin(), xi.end(), e_hi) - xi.begin());
            if (c == 0) lo = 0;
            if (c == nc - 1) hi = n;
            int w = hi - lo;
            if (w > 254) w = 255;
            lut[c] = (uint32_t)lo | ((uint32_t)w << 24);
        }
        ncell = nc;
        scale = sc;
    }
    BSK_CUDA(cudaSetDevice(b->device));
    if (s->d_pp) { bsk::dev_free(s->d_pp); s->d_pp = nullptr; }
    if (s->d_pp_lut) { bsk::dev_free(s->d_pp_lut); s->d_pp_lut = nullptr; }
    BSK_CUDA(bsk::dev_malloc(&s->d_pp, recs.size() * sizeof(T)));
    BSK_CUDA(cudaMemcpy(s->d_pp, recs.data(), recs.size() * sizeof(T), cudaMemcpyHostToDevice));
    BSK_CUDA(bsk::dev_malloc((void **)&s->d_pp_lut, lut.size() * 4));
    BSK_CUDA(cudaMemcpy(s->d_pp_lut, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice));
    s->pp_nrec = nrec;
    s->pp_ncell = ncell;
    s->pp_scale = scale;
    s->pp_rec = REC;
    s->pp_valid = true;

    // ---- uniform-cell table (bsk_eval_cell.cu) -------------------------------------------------
    s->cell_valid = false;
    {
        const int RECV = ((k + VPA - 1) / VPA) * VPA;
        const int NCH = RECV / VPA;
        const int XI = (k == 1) ? 1 : 0;
        const size_t budget = 214 * 1024;
        const int nbp = (int)nrec - 1;                         // interior breakpoints
        const double a = plo[0], bb = phi[nrec - 1];
        // raw-bit helpers: flag = LSB of the last coefficient; flagged: last slot = (index << 1) | 1
        auto setflag = [](T v, int f) {
            if (sizeof(T) == 8) {
                uint64_t u; double d = (double)v; memcpy(&u, &d, 8);
                u = (u & ~1ull) | (uint64_t)f; memcpy(&d, &u, 8); return (T)d;
            }
            uint32_t u; float g = (float)v; memcpy(&u, &g, 4);
            u = (u & ~1u) | (uint32_t)f; memcpy(&g, &u, 4); return (T)g;
        };
        auto index_bits = [](uint32_t idx) {
            const uint32_t raw = (idx << 1) | 1u;
            T out;
            if (sizeof(T) == 8) { uint64_t u = raw; double d; memcpy(&d, &u, 8); out = (T)d; }
            else { float g; memcpy(&g, &raw, 4); out = (T)g; }
            return out;
        };
        const uint32_t kMulti = 0x7FFFFFFFu;
        // "Jump" scheme (shared-memory tables, Float64, k >= 4; cell_jump_scheme in handles.cuh): a cell
        // with one SIMPLE knot xi keeps the Taylor coefficients of the piece LEFT of the knot in its
        // record, and the piece right of it is  P_L + J (x - xi)^(k-1)  (the spline is C^(k-2) at a
        // simple knot: only the leading coefficient jumps).  The side entry is just {xi, J} (16 bytes
        // instead of two full pieces), its index sits in bits 1..3 of the first four coefficients
        // (<= 8 ulp of each), so a flagged lane needs ONE extra 16-byte gather instead of two and the
        // smaller side table leaves room for ~40 % more cells (fewer flagged lanes).
        const bool jump_ok = bsk::cell_jump_scheme((int)sizeof(T), k);
        const uint32_t kJumpMulti = 4095u;
        const size_t side_bytes = jump_ok ? 16 : (size_t)2 * RECV * sizeof(T);
        // first guess: one side entry per interior breakpoint; shrink if it overflows
        int64_t nc = (int64_t)((budget - std::min<size_t>(budget, (size_t)(nbp + 8) * side_bytes)) /
                               (RECV * sizeof(T)));
        nc = std::min<int64_t>(nc, 16 * nrec);
        nc = std::min<int64_t>(nc, 60000);
        // Too many intervals for shared memory: keep the same tables in global memory, record-major
        // (one 32-byte L2 sector per cubic F64 / order-8 F32 record) and L2-resident.
        const bool in_global = nc < 3 * nrec && nrec <= 400000 && getenv("BSK_NO_GLOBAL_CELL") == nullptr;
        const bool jump = jump_ok && !in_global;
        const bool jpad = jump && bsk::cell_jump_pad(k);      // odd order: flag + index in the spare slot
        double dscale[BSK_MAXK];                               // magnitude of D^d S: max |coefficient|
        for (int d = 0; d < k; ++d) {
            dscale[d] = 0.0;
            for (double v : taylor.dc[d]) dscale[d] = std::max(dscale[d], std::fabs(v));
        }
        auto put_index_bits = [](T *rec, uint32_t idx) {      // bits 1..3 of c_0 .. c_3
            for (int m = 0; m < 4; ++m) {
                uint64_t u; double d = (double)rec[m]; memcpy(&u, &d, 8);
                u = (u & ~0xEull) | ((uint64_t)((idx >> (3 * m)) & 7u) << 1);
                memcpy(&d, &u, 8); rec[m] = (T)d;
            }
        };
        if (in_global) {
            const char *e = getenv("BSK_GLOBAL_CELL_MULT2");        // cells per interval, x2
            const int m2 = e ? std::max(6, atoi(e)) : 64;
            nc = (nrec * m2) / 2;                                   // 32 cells per interval ...
            const int64_t cap = (32ll << 20) / (int64_t)(bsk::cell_global_stride((int)sizeof(T), RECV) * sizeof(T));   // ... within 32 MB (L2-resident)
            nc = std::max<int64_t>(std::min(nc, cap), 3 * nrec);
        }
        for (int attempt = 0; attempt < 6 && nc >= 3 * nrec && nc >= 8; ++attempt, nc = nc * 9 / 10) {
            const T x0 = (T)a;
            const T cwT = (T)((bb - a) / (double)nc);
            const T scT = (T)((double)nc / (bb - a));
            const double cw = (double)cwT;
            const double margin = (sizeof(T) == 4 ? 0.02 : 1e-6) * cw;
            std::vector<T> crec((size_t)nc * RECV, (T)0), side_l, side_r;     // cell-major while building
            int nmulti = 0;
            for (int64_t c = 0; c < nc; ++c) {
                const double e0 = a + (double)c * cw, e1 = a + (double)(c + 1) * cw;
                const T centerT = std::fma((T)c + (T)0.5, cwT, x0);     // exactly the device's formula
                // piece at the left edge: number of interior breakpoints strictly below e0 - margin
                int pl = (int)(std::lower_bound(plo.begin() + 1, plo.end(), e0 - margin) - (plo.begin() + 1));
                int nb = 0;
                while (pl + 1 + nb < (int)nrec && plo[pl + 1 + nb] <= e1 + margin) ++nb;
                if (c == nc - 1) nb = (int)nrec - 1 - pl;        // everything up to the right end
                double cf[BSK_MAXK];
                taylor(pid[pl], (double)centerT, cf);
                if (nb == 0) {
                    for (int m = 0; m < k; ++m) crec[c * RECV + m] = (T)cf[m];
                    if (!jpad) crec[c * RECV + k - 1] = setflag(crec[c * RECV + k - 1], 0);
                } else if (nb == 1 && jump) {
                    // simple knot <=> the two pieces are consecutive knot intervals; the first and the
                    // last cell are kept unflagged-or-multi (boundary pieces feed the extrapolations)
                    const bool simple = pid[pl + 1] == pid[pl] + 1 && c != 0 && c != nc - 1;
                    const uint32_t j = (uint32_t)(side_l.size() / 2);
                    for (int m = 0; m < k; ++m) crec[c * RECV + m] = (T)cf[m];
                    double cr[BSK_MAXK];
                    bool accurate = simple;
                    if (simple) {
                        // Right of the knot the kernel forms P_L + J dx^(k-1) with P_L continued beyond its
                        // own interval.  With very uneven neighbouring intervals the two terms can be much
                        // larger than the spline and cancel: accept the cell only if the sum of the absolute
                        // terms stays within 16x the magnitude of D^d S for every derivative order (rounding
                        // error <= ~4e-15 of the output scale); otherwise the cell takes the interval table.
                        taylor(pid[pl + 1], (double)centerT, cr);
                        const double jmp = cr[k - 1] - cf[k - 1];
                        const double h = e1 - (double)centerT, dxm = e1 - plo[pl + 1];
                        for (int d = 0; d < k && accurate; ++d) {
                            double sum = 0.0, hp = 1.0;
                            for (int m = d; m < k; ++m) {
                                double ff = 1.0;
                                for (int q = 0; q < d; ++q) ff *= (double)(m - q);
                                sum += std::fabs(cf[m]) * ff * hp;
                                hp *= std::fabs(h);
                            }
                            double ffj = 1.0;
                            for (int q = 0; q < d; ++q) ffj *= (double)(k - 1 - q);
                            sum += std::fabs(jmp) * ffj * std::pow(std::fabs(dxm), k - 1 - d);
                            if (!(sum <= 16.0 * dscale[d] + 1e-300)) accurate = false;
                        }
                    }
                    if (accurate && (jpad || j < kJumpMulti)) {
                        side_l.push_back((T)plo[pl + 1]);
                        side_l.push_back((T)(cr[k - 1] - cf[k - 1]));
                        if (jpad) crec[c * RECV + RECV - 1] = index_bits(j);
                        else put_index_bits(&crec[c * RECV], j);
                    } else {
                        if (jpad) crec[c * RECV + RECV - 1] = index_bits(kMulti);
                        else put_index_bits(&crec[c * RECV], kJumpMulti);
                        ++nmulti;
                    }
                    if (!jpad) crec[c * RECV + k - 1] = setflag(crec[c * RECV + k - 1], 1);
                } else if (nb >= 2 && jump) {
                    if (jpad) crec[c * RECV + RECV - 1] = index_bits(kMulti);
                    else {
                        put_index_bits(&crec[c * RECV], kJumpMulti);
                        crec[c * RECV + k - 1] = setflag(crec[c * RECV + k - 1], 1);
                    }
                    ++nmulti;
                } else if (nb == 1) {
                    const uint32_t j = (uint32_t)(side_l.size() / RECV);
                    crec[c * RECV + XI] = (T)plo[pl + 1];
                    crec[c * RECV + k - 1] = index_bits(j);
                    size_t base = side_l.size();
                    side_l.resize(base + RECV, (T)0);
                    side_r.resize(base + RECV, (T)0);
                    for (int m = 0; m < k; ++m) side_l[base + m] = (T)cf[m];
                    taylor(pid[pl + 1], (double)centerT, cf);
                    for (int m = 0; m < k; ++m) side_r[base + m] = (T)cf[m];
                } else {
                    crec[c * RECV + k - 1] = index_bits(kMulti);
                    ++nmulti;
                }
            }
            if (jump && side_l.empty()) side_l.resize(2, (T)0);
            if (side_l.empty()) { side_l.resize(RECV, (T)0); side_r.resize(RECV, (T)0); }
            const int nside = jump ? (int)(side_l.size() / 2) : (int)(side_l.size() / RECV);
            const size_t smem = jump ? ((size_t)nc * RECV + 2 * (size_t)nside) * sizeof(T) + 16
                                     : ((size_t)nc + 2 * (size_t)nside) * RECV * sizeof(T) + 16;
            if (!in_global && smem > 226 * 1024) continue;      // retry with fewer cells
            if (nmulti * 50 > nc) break;                        // clustered knots: not worth it
            // device layout: chunk-major for shared memory, record-major for global (CellView)
            const int RS = in_global ? bsk::cell_global_stride((int)sizeof(T), RECV) : RECV;   // record stride on the device
            std::vector<T> crec_cm((size_t)nc * RS, (T)0), rc_cm(jump ? side_l.size() : 2 * (size_t)nside * RS, (T)0);
            if (jump) rc_cm = side_l;                           // {knot, jump} pairs, 16 bytes each
            for (int64_t c = 0; c < nc; ++c)
                for (int w = 0; w < NCH; ++w)
                    for (int e = 0; e < VPA; ++e)
                        crec_cm[in_global ? (size_t)c * RS + w * VPA + e : ((size_t)w * nc + c) * VPA + e] =
                            crec[c * RECV + w * VPA + e];
            for (int side = 0; side < 2 && !jump; ++side) {
                const std::vector<T> &src = side ? side_r : side_l;
                for (int64_t j = 0; j < nside; ++j)
                    for (int w = 0; w < NCH; ++w)
                        for (int e = 0; e < VPA; ++e)
                            rc_cm[in_global ? ((size_t)side * nside + j) * RS + w * VPA + e
                                            : (((size_t)side * NCH + w) * nside + j) * VPA + e] = src[j * RECV + w * VPA + e];
            }
            for (void **p : {&s->d_cell_rec, &s->d_cell_rc})
                if (*p) { bsk::dev_free(*p); *p = nullptr; }
            BSK_CUDA(bsk::dev_malloc(&s->d_cell_rec, crec_cm.size() * sizeof(T)));
            BSK_CUDA(cudaMemcpy(s->d_cell_rec, crec_cm.data(), crec_cm.size() * sizeof(T), cudaMemcpyHostToDevice));
            BSK_CUDA(bsk::dev_malloc(&s->d_cell_rc, rc_cm.size() * sizeof(T)));
            BSK_CUDA(cudaMemcpy(s->d_cell_rc, rc_cm.data(), rc_cm.size() * sizeof(T), cudaMemcpyHostToDevice));
            s->cell_ncell = (int)nc;
            s->cell_nside = nside;
            s->cell_x0 = (double)x0;
            s->cell_b = (double)(T)bb;
            s->cell_scale = (double)scT;
            s->cell_cw = (double)cwT;
            s->cell_smem = in_global ? 0 : smem;
            s->cell_valid = true;
            break;
        }
    }
    return BSK_OK;
}

bsk_status get_derivative(const bsk_spline *s_const, int nd, bsk_spline **out) {
    bsk_spline *s = const_cast<bsk_spline *>(s_const);
    if ((int)s->derivs.size() <= nd) s->derivs.resize(nd + 1, nullptr);
    if (!s->derivs[nd]) {
        bsk_spline *d = nullptr;
        bsk_status st = bsk_spline_derivative(s, nd, &d);
        if (st != BSK_OK) return st;
        s->derivs[nd] = d;
    }
    *out = s->derivs[nd];
    return BSK_OK;
}

template <int K, int KIND, typename T>
bsk_status launch_deboor(const bsk_spline *s, const T *d_x, int64_t n, T *d_out, int extrap, cudaStream_t st) {
    const bsk_basis *b = s->basis;
    KnotView<T> kv = make_view<T>(b);
    size_t smem = knots_smem_bytes<T>(b) + (size_t)s->ncoef * sizeof(T);
    int stage = smem <= kSmemStageLimit ? 1 : 0;
    if (!stage) smem = 0;
    auto kern = k_spline_deboor<K, KIND, T>;
    BSK_CUDA(set_smem(kern, smem));
    int ctas = stage ? std::max<int>(1, (int)std::min<size_t>(8, (220 * 1024) / std::max<size_t>(smem, 1))) : 8;
    int grid = grid_for(n, b->device, ctas);
    kern<<<grid, kThreads, smem, st>>>(kv, (const T *)s->d_coefs, (int)s->ncoef, d_x, n, d_out, stage, extrap);
    BSK_CUDA(cudaGetLastError());
    return BSK_OK;
}

template <typename T>
bsk_status eval_deboor_t(const bsk_spline *s, const T *d_x, int64_t n, T *d_out, int extrap, cudaStream_t st) {
    const bsk_basis *b = s->basis;
    bsk_status rc = BSK_ERR_UNSUPPORTED;
    if (b->kind == BSK_BASIS_PERIODIC) {
        BSK_DISPATCH_K(b->k, rc = (launch_deboor<K, 1, T>(s, d_x, n, d_out, extrap, st)))
    } else {
        BSK_DISPATCH_K(b->k, rc = (launch_deboor<K, 0, T>(s, d_x, n, d_out, extrap, st)))
    }
    if (rc == BSK_ERR_UNSUPPORTED) set_error("unsupported order %d", b->k);
    return rc;
}

template <int K, int KIND, typename T>
bsk_status launch_pp(const bsk_spline *s, const T *d_x, int64_t n, const EvalOuts &outs, bool force,
                     bool scalar_only, cudaStream_t st, bool *used) {
    constexpr int REC = PPRec<K, T>::REC;
    constexpr int PTS = 16 / (int)sizeof(T);
    const bsk_basis *b = s->basis;
    size_t smem = (size_t)s->pp_nrec * REC * sizeof(T) + (size_t)s->pp_ncell * 4;
    int stage = smem <= kSmemStageLimit ? 1 : 0;
    if (!stage && !force) { *used = false; return BSK_OK; }
    if (scalar_only && n < 4096) stage = 0;          // tiny head/tail launches: skip the staging
    if (!stage) smem = 0;
    *used = true;
    KnotView<T> kv = make_view<T>(b);
    PPView<T> pv = make_ppview<T>(s);
    bool aligned = !scalar_only && (((uintptr_t)d_x) & 15) == 0;
    for (int o = 0; o < outs.nout; ++o) aligned = aligned && ((((uintptr_t)outs.out[o]) & 15) == 0);
    int ctas = stage ? std::max<int>(1, (int)std::min<size_t>(4, (220 * 1024) / std::max<size_t>(smem, 1))) : 4;
    if (aligned) {
        auto kern = k_spline_pp<K, KIND, T, PTS>;
        BSK_CUDA(set_smem(kern, smem));
        int grid = grid_for((n + PTS - 1) / PTS, b->device, ctas);
        kern<<<grid, kThreads, smem, st>>>(kv, (int)s->ncoef, pv, d_x, n, outs, stage);
    } else {
        auto kern = k_spline_pp<K, KIND, T, 1>;
        BSK_CUDA(set_smem(kern, smem));
        int grid = grid_for(n, b->device, ctas);
        kern<<<grid, kThreads, smem, st>>>(kv, (int)s->ncoef, pv, d_x, n, outs, stage);
    }
    BSK_CUDA(cudaGetLastError());
    return BSK_OK;
}

template <typename T>
bsk_status dispatch_pp(const bsk_spline *s, const T *d_x, int64_t n, const EvalOuts &outs, bool force,
                       bool scalar_only, cudaStream_t st, bool *used) {
    bsk_status rc = BSK_ERR_UNSUPPORTED;
    if (s->basis->kind == BSK_BASIS_PERIODIC) {
        BSK_DISPATCH_K(s->basis->k, rc = (launch_pp<K, 1, T>(s, d_x, n, outs, force, scalar_only, st, used)))
    } else {
        BSK_DISPATCH_K(s->basis->k, rc = (launch_pp<K, 0, T>(s, d_x, n, outs, force, scalar_only, st, used)))
    }
    return rc;
}

template <typename T>
bsk_status spline_eval_t(const bsk_spline *s, const T *d_x, int64_t n, int nout, const int32_t *derivs,
                         void *const *d_outs, int algo, int extrap, cudaStream_t st) {
    const bsk_basis *b = s->basis;
    if (b->kind == BSK_BASIS_PERIODIC) {
        if (extrap == BSK_EXTRAP_SMOOTH) extrap = BSK_EXTRAP_NONE;      // zone is always 0: nothing to do
        if (extrap == BSK_EXTRAP_LINEAR) {
            set_error("Linear extrapolation of a periodic spline is not supported");
            return BSK_ERR_UNSUPPORTED;
        }
    }
    bool try_pp = (algo == BSK_EVAL_AUTO || algo == BSK_EVAL_PP) && pp_eligible(s);
    if (algo == BSK_EVAL_PP && !try_pp) {
        set_error("local-polynomial path needs full end-knot multiplicity");
        return BSK_ERR_UNSUPPORTED;
    }
    // lazily built state (tables, derivative splines) is created under the handle's mutex; the
    // kernels themselves only read it
    std::unique_lock<std::mutex> lock(const_cast<bsk_spline *>(s)->mu);
    if (try_pp) {
        bsk_spline *sm = const_cast<bsk_spline *>(s);
        if (!s->pp_valid) {
            bsk_status rc = build_pp<T>(sm);
            if (rc != BSK_OK) return rc;
        }
        bsk_status rc = BSK_OK;
        EvalOuts outs;
        memset(&outs, 0, sizeof(outs));
        outs.nout = nout;
        outs.extrap = extrap;
        for (int o = 0; o < nout; ++o) {
            outs.out[o] = d_outs[o];
            outs.deriv[o] = derivs[o];
            outs.dcoefs[o] = s->d_coefs;
            if (b->kind == BSK_BASIS_PERIODIC && derivs[o] > 0) {
                bsk_spline *ds = nullptr;
                rc = get_derivative(s, derivs[o], &ds);
                if (rc != BSK_OK) return rc;
                outs.dcoefs[o] = ds->d_coefs;
            }
        }
        // 1) uniform-cell table in shared memory (bsk_eval_cell.cu): needs every pointer to share
        //    the same 16-byte phase; the unaligned head / tail go through the scalar pp kernel.
        constexpr int VEC = 16 / (int)sizeof(T);
        const bool want_cell = s->cell_valid && getenv("BSK_NO_CELL") == nullptr;
        if (want_cell) {
            const uintptr_t phase = ((uintptr_t)d_x) & 15;
            bool same = (phase % sizeof(T)) == 0;
            for (int o = 0; o < nout; ++o) same = same && ((((uintptr_t)d_outs[o]) & 15) == phase);
            if (same) {
                const int64_t head = std::min<int64_t>(n, phase ? (int64_t)((16 - phase) / sizeof(T)) : 0);
                const int64_t nvec = (n - head) / VEC;
                const int64_t tail = n - head - nvec * VEC;
                bool used = false;
                auto shifted = [&](int64_t off) {
                    EvalOuts e = outs;
                    for (int o = 0; o < nout; ++o) e.out[o] = (T *)outs.out[o] + off;
                    return e;
                };
                if (head > 0) {
                    rc = dispatch_pp<T>(s, d_x, head, outs, true, true, st, &used);
                    if (rc != BSK_OK) return rc;
                }
                if (nvec > 0) {
                    KnotView<T> kv = make_view<T>(b);
                    rc = launch_cell<T>(s, kv, d_x + head, nvec, shifted(head), st);
                    if (rc != BSK_OK) return rc;
                }
                if (tail > 0) {
                    rc = dispatch_pp<T>(s, d_x + head + nvec * VEC, tail, shifted(head + nvec * VEC), true, true, st, &used);
                    if (rc != BSK_OK) return rc;
                }
                return BSK_OK;
            }
        }
        // 2) per-interval table: shared memory if it fits, else served from global memory through
        //    L1/L2 (measured 2.3x faster than the de Boor kernel on 10^4-knot bases)
        bool used = false;
        rc = dispatch_pp<T>(s, d_x, n, outs, true, false, st, &used);
        if (rc != BSK_OK) return rc;
        if (used) return BSK_OK;
    }
    // de Boor in the reference's operation order, one pass per requested derivative
    if (extrap == BSK_EXTRAP_LINEAR) {
        set_error("Linear extrapolation needs the local-polynomial path (BSK_EVAL_AUTO / BSK_EVAL_PP on a "
                  "basis with full end-knot multiplicity)");
        return BSK_ERR_UNSUPPORTED;
    }
    for (int o = 0; o < nout; ++o) {
        const bsk_spline *ds = s;
        if (derivs[o] > 0) {
            bsk_spline *tmp = nullptr;
            bsk_status rc = get_derivative(s, derivs[o], &tmp);
            if (rc != BSK_OK) return rc;
            ds = tmp;
        }
        bsk_status rc = eval_deboor_t<T>(ds, d_x, n, (T *)d_outs[o], extrap, st);
        if (rc != BSK_OK) return rc;
    }
    return BSK_OK;
}

}  // namespace

// ============================================================================================
extern "C" {

bsk_status bsk_find_knot_interval(const bsk_basis *b, const void *d_x, int64_t n, int64_t *d_idx,
                                  int8_t *d_zone, void *stream) {
    BSK_REQUIRE(b && (n == 0 || (d_x && d_idx)), BSK_ERR_ARGUMENT, "NULL argument");
    if (n == 0) return BSK_OK;
    BSK_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int periodic = b->kind == BSK_BASIS_PERIODIC;
    if (b->dtype == BSK_F64) {
        KnotView<double> kv = make_view<double>(b);
        size_t smem = knots_smem_bytes<double>(b);
        int stage = smem <= kSmemStageLimit;
        if (!stage) smem = 0;
        int grid = grid_for(n, b->device, 8);
        if (periodic) {
            BSK_CUDA(set_smem(k_find_interval<1, double>, smem));
            k_find_interval<1, double><<<grid, kThreads, smem, st>>>(kv, (const double *)d_x, n, d_idx, d_zone, stage);
        } else {
            BSK_CUDA(set_smem(k_find_interval<0, double>, smem));
            k_find_interval<0, double><<<grid, kThreads, smem, st>>>(kv, (const double *)d_x, n, d_idx, d_zone, stage);
        }
    } else {
        KnotView<float> kv = make_view<float>(b);
        size_t smem = knots_smem_bytes<float>(b);
        int stage = smem <= kSmemStageLimit;
        if (!stage) smem = 0;
        int grid = grid_for(n, b->device, 8);
        if (periodic) {
            BSK_CUDA(set_smem(k_find_interval<1, float>, smem));
            k_find_interval<1, float><<<grid, kThreads, smem, st>>>(kv, (const float *)d_x, n, d_idx, d_zone, stage);
        } else {
            BSK_CUDA(set_smem(k_find_interval<0, float>, smem));
            k_find_interval<0, float><<<grid, kThreads, smem, st>>>(kv, (const float *)d_x, n, d_idx, d_zone, stage);
        }
    }
    BSK_CUDA(cudaGetLastError());
    return BSK_OK;
}

}  // extern "C"

namespace {

template <int K, int KIND, typename KT, typename VT>
bsk_status launch_eval_all(const bsk_basis *b, const KT *d_x, int64_t n, int deriv, const int64_t *d_ileft,
                           int64_t *d_idx, VT *d_bs, cudaStream_t st) {
    KnotView<KT> kv = make_view<KT>(b);
    size_t smem = knots_smem_bytes<KT>(b);
    int stage = smem <= kSmemStageLimit;
    if (!stage) smem = 0;
    auto kern = k_evaluate_all<K, KIND, KT, VT>;
    BSK_CUDA(set_smem(kern, smem));
    int grid = grid_for(n, b->device, 4);
    kern<<<grid, kThreads, smem, st>>>(kv, b->rv, d_x, n, deriv, d_ileft, d_idx, d_bs, stage);
    BSK_CUDA(cudaGetLastError());
    return BSK_OK;
}

template <typename KT, typename VT>
bsk_status eval_all_t(const bsk_basis *b, const void *d_x, int64_t n, int deriv, const int64_t *d_ileft,
                      int64_t *d_idx, void *d_bs, cudaStream_t st) {
    bsk_status rc = BSK_ERR_UNSUPPORTED;
    if (b->kind == BSK_BASIS_PERIODIC) {
        BSK_DISPATCH_K(b->k, rc = (launch_eval_all<K, 1, KT, VT>(b, (const KT *)d_x, n, deriv, d_ileft, d_idx, (VT *)d_bs, st)))
    } else {
        BSK_DISPATCH_K(b->k, rc = (launch_eval_all<K, 0, KT, VT>(b, (const KT *)d_x, n, deriv, d_ileft, d_idx, (VT *)d_bs, st)))
    }
    return rc;
}

}  // namespace

extern "C" {

bsk_status bsk_evaluate_all(const bsk_basis *b, const void *d_x, int64_t n, int32_t deriv, int32_t out_dtype,
                            const int64_t *d_ileft, int64_t *d_idx, void *d_bs, void *stream) {
    BSK_REQUIRE(b && (n == 0 || (d_x && d_idx && d_bs)), BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(deriv >= 0, BSK_ERR_ARGUMENT, "derivative order must be >= 0");
    BSK_REQUIRE(out_dtype == BSK_F32 || out_dtype == BSK_F64, BSK_ERR_ARGUMENT, "bad out_dtype");
    BSK_REQUIRE(!(b->dtype == BSK_F32 && out_dtype == BSK_F64), BSK_ERR_UNSUPPORTED,
                "Float64 output from a Float32 basis is not supported");
    if (n == 0) return BSK_OK;
    BSK_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (b->dtype == BSK_F64 && out_dtype == BSK_F64) return eval_all_t<double, double>(b, d_x, n, deriv, d_ileft, d_idx, d_bs, st);
    if (b->dtype == BSK_F64 && out_dtype == BSK_F32) return eval_all_t<double, float>(b, d_x, n, deriv, d_ileft, d_idx, d_bs, st);
    return eval_all_t<float, float>(b, d_x, n, deriv, d_ileft, d_idx, d_bs, st);
}

bsk_status bsk_spline_create(const bsk_basis *b, const void *h_coefs, int64_t ncoef, bsk_spline **out) {
    BSK_REQUIRE(b && h_coefs && out, BSK_ERR_ARGUMENT, "NULL argument");
    *out = nullptr;
    BSK_REQUIRE(ncoef == b->len, BSK_ERR_DIMENSION,
                "wrong number of coefficients: got %lld, basis has length %lld",   // spline.jl:52-56
                (long long)ncoef, (long long)b->len);
    bsk_spline *s = new bsk_spline();
    s->basis = nullptr; s->d_coefs = nullptr; s->owns_coefs = true; s->d_pp = nullptr;
    s->d_pp_lut = nullptr; s->pp_valid = false; s->pp_nrec = 0; s->pp_ncell = 0; s->pp_scale = 0; s->pp_rec = 0;
    s->d_cell_rec = s->d_cell_rc = nullptr; s->cell_valid = false;
    int ekind = b->kind == BSK_BASIS_PERIODIC ? BSK_BASIS_PERIODIC : BSK_BASIS_PLAIN;
    bsk_status rc = clone_basis(b, ekind, b->k, b->h_knots, &s->basis);
    if (rc != BSK_OK) { delete s; return rc; }
    s->ncoef = s->basis->N;
    std::vector<double> c(ncoef);
    for (int64_t i = 0; i < ncoef; ++i)
        c[i] = b->dtype == BSK_F64 ? ((const double *)h_coefs)[i] : (double)((const float *)h_coefs)[i];
    s->user_basis_kind = b->kind;
    s->user_rv = b->rv;
    s->user_len = b->len;
    rc = set_host_coefs(s, b, c);
    if (rc != BSK_OK) { bsk_spline_destroy(s); return rc; }
    *out = s;
    return BSK_OK;
}

bsk_status bsk_spline_set_coefs_device(bsk_spline *s, const void *d_coefs, int64_t ncoef, int64_t stride,
                                       void *stream) {
    BSK_REQUIRE(s && d_coefs, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(ncoef == s->user_len, BSK_ERR_DIMENSION, "wrong number of coefficients: got %lld, expected %lld",
                (long long)ncoef, (long long)s->user_len);
    BSK_REQUIRE(stride >= 1, BSK_ERR_ARGUMENT, "stride must be >= 1");
    const bsk_basis *b = s->basis;
    size_t es = b->dtype == BSK_F64 ? 8 : 4;
    BSK_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<unsigned char> raw(es * ncoef);
    BSK_CUDA(cudaMemcpy2DAsync(raw.data(), es, d_coefs, es * stride, es, ncoef, cudaMemcpyDeviceToHost, st));
    BSK_CUDA(cudaStreamSynchronize(st));
    std::vector<double> c(ncoef);
    for (int64_t i = 0; i < ncoef; ++i)
        c[i] = b->dtype == BSK_F64 ? ((const double *)raw.data())[i] : (double)((const float *)raw.data())[i];
    bsk_basis ub;   // minimal view of the user basis for the recombination step
    ub.kind = s->user_basis_kind; ub.dtype = b->dtype; ub.rv = s->user_rv;
    return set_host_coefs(s, &ub, c);
}

bsk_status bsk_spline_destroy(bsk_spline *s) {
    if (!s) return BSK_OK;
    if (s->basis) cudaSetDevice(s->basis->device);
    for (bsk_spline *d : s->derivs)
        if (d) bsk_spline_destroy(d);
    if (s->d_coefs && s->owns_coefs) bsk::dev_free(s->d_coefs);
    if (s->d_pp) bsk::dev_free(s->d_pp);
    if (s->d_pp_lut) bsk::dev_free(s->d_pp_lut);
    if (s->d_cell_rec) bsk::dev_free(s->d_cell_rec);
    if (s->d_cell_rc) bsk::dev_free(s->d_cell_rc);
    if (s->basis) bsk_basis_destroy(s->basis);
    delete s;
    return BSK_OK;
}

bsk_status bsk_spline_coefficients(const bsk_spline *s, void *h_out, int64_t ncoef) {
    BSK_REQUIRE(s && h_out, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(ncoef == s->ncoef, BSK_ERR_DIMENSION, "expected %lld coefficients", (long long)s->ncoef);
    for (int64_t i = 0; i < ncoef; ++i) {
        if (s->basis->dtype == BSK_F64) ((double *)h_out)[i] = s->h_coefs[i];
        else ((float *)h_out)[i] = (float)s->h_coefs[i];
    }
    return BSK_OK;
}

bsk_status bsk_spline_order(const bsk_spline *s, int32_t *k, int64_t *ncoef, int64_t *nknots) {
    BSK_REQUIRE(s, BSK_ERR_ARGUMENT, "NULL argument");
    if (k) *k = s->basis->k;
    if (ncoef) *ncoef = s->ncoef;
    if (nknots) *nknots = s->basis->nt;
    return BSK_OK;
}

bsk_status bsk_spline_knots(const bsk_spline *s, void *h_out, int64_t nknots) {
    BSK_REQUIRE(s && h_out, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(nknots == s->basis->nt, BSK_ERR_DIMENSION, "expected %lld knots", (long long)s->basis->nt);
    for (int64_t i = 0; i < nknots; ++i) {
        if (s->basis->dtype == BSK_F64) ((double *)h_out)[i] = s->basis->h_knots[i];
        else ((float *)h_out)[i] = (float)s->basis->h_knots[i];
    }
    return BSK_OK;
}

bsk_status bsk_spline_derivative(const bsk_spline *s, int32_t nd, bsk_spline **out) {
    BSK_REQUIRE(s && out, BSK_ERR_ARGUMENT, "NULL argument");
    *out = nullptr;
    const bsk_basis *b = s->basis;
    BSK_REQUIRE(nd >= 1, BSK_ERR_ARGUMENT, "derivative order must be >= 1");
    BSK_REQUIRE(nd < b->k, BSK_ERR_ARGUMENT, "cannot differentiate order %d spline %d times!", b->k, nd);  // spline.jl:303-306
    const int kind = b->kind == BSK_BASIS_PERIODIC ? 1 : 0;
    std::vector<double> dc;
    if (b->dtype == BSK_F64) derivative_coefs<double>(kind, b->k, b->h_knots, s->h_coefs, nd, dc);
    else derivative_coefs<float>(kind, b->k, b->h_knots, s->h_coefs, nd, dc);
    std::vector<double> tk;
    if (kind == 0) tk.assign(b->h_knots.begin() + nd, b->h_knots.end() - nd);   // basis.jl:104-111
    else tk = b->h_knots;
    bsk_spline *d = new bsk_spline();
    d->basis = nullptr; d->d_coefs = nullptr; d->owns_coefs = true; d->d_pp = nullptr;
    d->d_pp_lut = nullptr; d->pp_valid = false; d->pp_nrec = 0; d->pp_ncell = 0; d->pp_scale = 0; d->pp_rec = 0;
    d->d_cell_rec = d->d_cell_rc = nullptr; d->cell_valid = false;
    bsk_status rc = clone_basis(b, kind ? BSK_BASIS_PERIODIC : BSK_BASIS_PLAIN, b->k - nd, tk, &d->basis);
    if (rc != BSK_OK) { delete d; return rc; }
    d->ncoef = (int64_t)dc.size();
    d->user_basis_kind = d->basis->kind;
    memset(&d->user_rv, 0, sizeof(d->user_rv));
    d->user_len = d->ncoef;
    d->h_coefs = dc;
    rc = upload_coefs(d);
    if (rc != BSK_OK) { bsk_spline_destroy(d); return rc; }
    *out = d;
    return BSK_OK;
}

bsk_status bsk_spline_eval(const bsk_spline *s, const void *d_x, int64_t n, int32_t nout, const int32_t *derivs,
                           void *const *d_outs, int32_t algo, void *stream) {
    return bsk_spline_eval_extrap(s, d_x, n, nout, derivs, d_outs, algo, BSK_EXTRAP_NONE, stream);
}

bsk_status bsk_spline_eval_extrap(const bsk_spline *s, const void *d_x, int64_t n, int32_t nout,
                                  const int32_t *derivs, void *const *d_outs, int32_t algo, int32_t extrap,
                                  void *stream) {
    BSK_REQUIRE(s && derivs && d_outs, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(extrap >= 0 && extrap <= 3, BSK_ERR_ARGUMENT, "unknown extrapolation method %d", extrap);
    BSK_REQUIRE(nout >= 1 && nout <= kMaxOut, BSK_ERR_ARGUMENT, "nout must be in 1..%d", kMaxOut);
    BSK_REQUIRE(algo >= 0 && algo <= 2, BSK_ERR_ARGUMENT, "unknown algo %d", algo);
    for (int o = 0; o < nout; ++o) {
        BSK_REQUIRE(derivs[o] >= 0, BSK_ERR_ARGUMENT, "derivative order must be >= 0");
        BSK_REQUIRE(derivs[o] < s->basis->k, BSK_ERR_ARGUMENT, "cannot differentiate order %d spline %d times!",
                    s->basis->k, derivs[o]);
        BSK_REQUIRE(n == 0 || d_outs[o], BSK_ERR_ARGUMENT, "NULL output");
    }
    if (n == 0) return BSK_OK;
    BSK_REQUIRE(d_x, BSK_ERR_ARGUMENT, "NULL points");
    BSK_CUDA(cudaSetDevice(s->basis->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (s->basis->dtype == BSK_F64)
        return spline_eval_t<double>(s, (const double *)d_x, n, nout, derivs, d_outs, algo, extrap, st);
    return spline_eval_t<float>(s, (const float *)d_x, n, nout, derivs, d_outs, algo, extrap, st);
}

}  // extern "C"
