/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement of the BSplineKit.jl (v0.19.2) hot-path algorithms, in the
 * reference's operation order (compile with -ffp-contract=off: Julia never
 * contracts a*b+c).  This file is a "template": it is included several times by
 * bsk_oracle.c with
 *     KT   knot / abscissa type       (double | float)
 *     VT   value type of the result   (double | float)
 *     SFX  symbol suffix              (dd | df | ff)
 * Rule reproduced (src/BSplines/evaluate_all.jl:152-155,295-298): knot
 * differences are formed in KT, *then* rounded to VT, and the triangle runs in VT.
 *
 * All indices are 1-based like the reference (Julia); C arrays are accessed
 * through the K() accessor below.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

/* ---- knot-vector abstraction --------------------------------------------------
 * kind 0: plain knot vector t[1..nt]                 (BSplineBasis, basis.jl:70-90)
 * kind 1: PeriodicKnots over data[1..N+1]            (periodic.jl:37-55)
 *         index_offset = k/2 (periodic.jl:62); ts[i] = data[i - k/2] +/- m L
 */
typedef struct {
    int kind;
    int k;
    int64_t nt;   /* plain: number of knots; periodic: N + 1 (length of data) */
    const KT *t;
    int64_t N;    /* periodic: period length (number of basis functions) */
    KT period, a, b;
} FN(oknots);

static FN(oknots) FN(make_knots_view)(int kind, int k, const KT *t, int64_t nt) {
    FN(oknots) ts;
    ts.kind = kind; ts.k = k; ts.nt = nt; ts.t = t;
    if (kind == 1) {
        ts.N = nt - 1;
        ts.a = t[0]; ts.b = t[nt - 1];
        ts.period = ts.b - ts.a;           /* periodic.jl:46-51 */
    } else {
        ts.N = nt - k;
        ts.a = t[0]; ts.b = t[nt - 1];
        ts.period = 0;
    }
    return ts;
}

/* getindex: plain vector, or PeriodicKnots (periodic.jl:102-115) */
static inline KT FN(K)(const FN(oknots) *ts, int64_t i) {
    if (ts->kind == 0) return ts->t[i - 1];
    KT x = 0;
    i -= ts->k / 2;
    while (i < 1) { x -= ts->period; i += ts->N; }
    while (i > ts->nt) { x += ts->period; i -= ts->N; }
    return x + ts->t[i - 1];
}

/* Base.searchsortedlast(v, x) with isless ordering restricted to non-signed-zero,
 * i.e. last i with !(x < v[i]); 0 if none.  NaN x: never "less" -> n. */
static int64_t FN(searchsortedlast)(const KT *v, int64_t n, KT x) {
    int64_t lo = 0, hi = n + 1;
    while (lo < hi - 1) {
        int64_t m = lo + ((hi - lo) >> 1);
        if (x < v[m - 1]) hi = m; else lo = m;
    }
    return lo;
}

/* find_knot_interval (evaluate_all.jl:102-119; periodic.jl:85-100) */
static int64_t FN(find_knot_interval)(const FN(oknots) *ts, KT x, int *zone) {
    if (ts->kind == 1) {
        int64_t i = ts->k / 2;
        while (x < ts->a) { x += ts->period; i -= ts->N; }
        while (x >= ts->b) { x -= ts->period; i += ts->N; }
        i += FN(searchsortedlast)(ts->t, ts->nt, x);
        *zone = 0;
        return i;
    }
    if (x < ts->t[0]) { *zone = -1; return 1; }
    int64_t i = FN(searchsortedlast)(ts->t, ts->nt, x);
    int64_t Nt = ts->nt;
    if (i == Nt) {
        KT tlast = ts->t[Nt - 1];
        for (;;) {
            i -= 1;
            if (ts->t[i - 1] != tlast) break;
        }
        *zone = (tlast < x) ? 1 : 0;
        return i;
    }
    *zone = 0;
    return i;
}

/* _knot_zone (evaluate_all.jl:126; periodic.jl:81) */
static int FN(knot_zone)(const FN(oknots) *ts, KT x) {
    if (ts->kind == 1) return 0;
    return (x < ts->t[0]) ? -1 : (x > ts->t[ts->nt - 1]) ? 1 : 0;
}

/* _evaluate_all_alt, Derivative{0} (evaluate_all.jl:207-235) with
 * _knotdiff (evaluate_all.jl:66-74).  bs has k entries, reversed order. */
static int64_t FN(eval_all_values)(const FN(oknots) *ts, KT x, int k, int64_t ileft,
                                   VT *bq, int *zone_out) {
    int zone;
    int64_t i;
    if (ileft != 0) { i = ileft; zone = FN(knot_zone)(ts, x); }   /* :121-124 */
    else i = FN(find_knot_interval)(ts, x, &zone);
    if (zone_out) *zone_out = zone;
    if (zone != 0) {
        for (int j = 0; j < k; ++j) bq[j] = 0;
        return i;
    }
    VT Ds[16];
    bq[0] = 1;
    for (int q = 2; q <= k; ++q) {
        for (int j = 1; j <= q - 1; ++j) {
            KT ti = FN(K)(ts, i - j + 1);
            KT tj = FN(K)(ts, i - j + 1 + (q - 1));
            KT d = (x - ti) / (tj - ti);
            Ds[j - 1] = (VT)d;
        }
        VT bp = bq[0];
        VT Dp = Ds[0];
        bq[0] = Dp * bp;
        for (int j = 2; j <= q - 1; ++j) {
            VT bpp = bp; bp = bq[j - 1];
            VT Dpp = Dp; Dp = Ds[j - 1];
            VT t1 = Dp * bp;
            VT t2 = ((VT)1 - Dpp) * bpp;
            bq[j - 1] = t1 + t2;
        }
        bq[q - 1] = ((VT)1 - Dp) * bp;
    }
    return i;
}

/* _evaluate_all_alt, Derivative{n>=1} (evaluate_all.jl:238-269) */
static int64_t FN(eval_all)(const FN(oknots) *ts, KT x, int k, int nderiv, int64_t ileft,
                            VT *bq) {
    if (nderiv == 0) return FN(eval_all_values)(ts, x, k, ileft, bq, NULL);
    int m = k - nderiv;
    if (m < 1) {
        for (int j = 0; j < k; ++j) bq[j] = 0;
        return 1;                                   /* firstindex(ts) */
    }
    for (int j = 0; j < k; ++j) bq[j] = 0;
    int zone;
    int64_t i = FN(eval_all_values)(ts, x, m, ileft, bq, &zone);
    /* Outside the knot domain the order-m values are zeros and the reference's
     * differencing keeps them at zero (0/dt); for zone == -1 it indexes t[<=0]
     * under @inbounds (undefined), so the defined behaviour here is: zeros. */
    if (zone != 0) return i;
    VT us[16];
    for (int q = m + 1; q <= k; ++q) {
        int p = q - 1;
        for (int dj = 1; dj <= p; ++dj) {
            KT dt = FN(K)(ts, i + q - dj) - FN(K)(ts, i + 1 - dj);
            /* bq is VT, dt is KT: division in promote type, rounded to VT (:295-298) */
#if KT_IS_WIDER
            us[dj - 1] = (VT)((KT)bq[dj - 1] / dt);
#else
            us[dj - 1] = bq[dj - 1] / (VT)dt;
#endif
        }
        bq[0] = (VT)p * us[0];
        for (int j = 2; j <= q - 1; ++j) bq[j - 1] = (VT)p * (us[j - 1] - us[j - 2]);
        bq[q - 1] = -(VT)p * us[p - 1];
    }
    return i;
}

/* ---- recombination (src/Recombinations) -------------------------------------- */
typedef struct {
    int cl, cr;          /* number of constraints left / right            */
    int nl, nr;          /* number of recombined functions left / right   */
    int ul_rows, lr_rows;/* corner blocks: ul is ul_rows x nl, lr is lr_rows x nr, column-major */
    const double *ul, *lr;
    int64_t Nparent;     /* length of the parent basis */
} FN(orecomb);

/* getindex(A::RecombineMatrix, i, j; block)  (matrices.jl:416-452) */
static double FN(recomb_get)(const FN(orecomb) *A, int64_t i, int64_t j, int block) {
    int64_t M = A->Nparent - A->cl - A->cr;
    if (block == 2) return (i == j + A->cl) ? 1.0 : 0.0;
    if (block == 1) {
        if (i <= A->ul_rows) return A->ul[(i - 1) + (j - 1) * (int64_t)A->ul_rows];
        return 0.0;
    }
    int64_t h = M - A->nr;
    int64_t ii = i - h - A->cl;
    if (ii >= 1 && ii <= A->lr_rows) return A->lr[(ii - 1) + (j - h - 1) * (int64_t)A->lr_rows];
    return 0.0;
}

/* evaluate_all(R::RecombinedBSplineBasis, ...)  (bases.jl:386-420) */
static int64_t FN(eval_all_recombined)(const FN(oknots) *ts, const FN(orecomb) *A, KT x, int k,
                                       int nderiv, int64_t ileft, VT *phis) {
    VT bs[16];
    int off = A->cl;
    if (ileft != 0) ileft += off;
    int64_t ilast = FN(eval_all)(ts, x, k, nderiv, ileft, bs);
    int64_t jlast = ilast - off;
    int64_t N = A->Nparent - A->cl - A->cr;
    for (int dj = 1; dj <= k; ++dj) phis[dj - 1] = 0;
    for (int dj = 1; dj <= k; ++dj) {
        int64_t j = jlast + 1 - dj;
        if (j < 1 || j > N) continue;
        int block = (j <= A->nl) ? 1 : (j > N - A->nr) ? 3 : 2;  /* matrices.jl:377-384 */
        if (block == 2) {
            phis[dj - 1] = bs[dj - 1];
        } else {
            for (int di = 1; di <= k; ++di) {
                int64_t i = ilast + 1 - di;
                if (i <= 0) continue;
                /* eltype(A) = eltype(knots) = KT; the product/sum run in
                 * promote(KT, VT) and are rounded to VT on store. */
                KT a = (KT)FN(recomb_get)(A, i, j, block);
#if KT_IS_WIDER
                phis[dj - 1] = (VT)((KT)phis[dj - 1] + a * (KT)bs[di - 1]);
#else
                phis[dj - 1] += (VT)a * bs[di - 1];
#endif
            }
        }
    }
    return jlast;
}

/* ---- spline evaluation: spline_kernel_alt (spline.jl:211-233) ------------------
 * coefficients have type VT here (result type = promote(VT, KT) = VT for the
 * dd / ff instantiations; the df instantiation is not used for splines). */
static inline VT FN(coef_get)(const FN(oknots) *ts, const VT *c, int64_t ncoef, int64_t i) {
    if (ts->kind == 1) {                 /* PeriodicVector (Splines/periodic.jl:48-62) */
        while (i < 1) i += ncoef;
        while (i > ncoef) i -= ncoef;
    }
    return c[i - 1];
}

static VT FN(spline_kernel)(const FN(oknots) *ts, const VT *c, int64_t ncoef, int64_t n, KT x,
                            int k) {
    VT d[16];
    for (int j = 1; j <= k; ++j) d[j - 1] = (VT)0 + FN(coef_get)(ts, c, ncoef, j + n - k);
    for (int r = 2; r <= k; ++r) {
        VT dprev = d[r - 2];
        for (int j = r; j <= k; ++j) {
            int64_t jn = j + n;
            KT ti = FN(K)(ts, jn - k);
            KT tj = FN(K)(ts, jn - r + 1);
            VT alpha = (VT)((x - ti) / (tj - ti));
            VT dtmp = dprev;
            dprev = d[j - 1];
            VT t1 = ((VT)1 - alpha) * dtmp;
            VT t2 = alpha * dprev;
            d[j - 1] = t1 + t2;
        }
    }
    return d[k - 1];
}

/* _evaluate(S, x)  (spline.jl:155-172) */
static VT FN(spline_eval1)(const FN(oknots) *ts, const VT *c, int64_t ncoef, KT x, int k) {
    int zone;
    int64_t n = FN(find_knot_interval)(ts, x, &zone);
    if (zone != 0) return (VT)0;
    return FN(spline_kernel)(ts, c, ncoef, n, x, k);
}

/* ---- exported batch drivers ----------------------------------------------------- */
void FN(orc_find_knot_interval)(int kind, int k, const KT *t, int64_t nt, const KT *x, int64_t n,
                                int64_t *idx, int8_t *zone) {
    FN(oknots) ts = FN(make_knots_view)(kind, k, t, nt);
    for (int64_t p = 0; p < n; ++p) {
        int z;
        idx[p] = FN(find_knot_interval)(&ts, x[p], &z);
        if (zone) zone[p] = (int8_t)z;
    }
}

void FN(orc_evaluate_all)(int kind, int k, const KT *t, int64_t nt, const KT *x, int64_t n,
                          int nderiv, const int64_t *ileft, int64_t *idx, VT *bs) {
    FN(oknots) ts = FN(make_knots_view)(kind, k, t, nt);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; ++p)
        idx[p] = FN(eval_all)(&ts, x[p], k, nderiv, ileft ? ileft[p] : 0, bs + p * k);
}

void FN(orc_evaluate_all_recombined)(int k, const KT *t, int64_t nt, int cl, int cr, int nl, int nr,
                                     int ul_rows, const double *ul, int lr_rows, const double *lr,
                                     const KT *x, int64_t n, int nderiv, const int64_t *ileft,
                                     int64_t *idx, VT *bs) {
    FN(oknots) ts = FN(make_knots_view)(0, k, t, nt);
    FN(orecomb) A = {cl, cr, nl, nr, ul_rows, lr_rows, ul, lr, nt - k};
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; ++p)
        idx[p] = FN(eval_all_recombined)(&ts, &A, x[p], k, nderiv, ileft ? ileft[p] : 0,
                                         bs + p * k);
}

#if !KT_IS_WIDER
void FN(orc_spline_eval)(int kind, int k, const KT *t, int64_t nt, const VT *c, int64_t ncoef,
                         const KT *x, int64_t n, VT *out) {
    FN(oknots) ts = FN(make_knots_view)(kind, k, t, nt);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; ++p) out[p] = FN(spline_eval1)(&ts, c, ncoef, x[p], k);
}

/* extrapolate(S, Smooth()) (src/SplineExtrapolations/SplineExtrapolations.jl:60-71): the usual
 * evaluation with the knot interval clamped to [k, N + k] whatever the zone. */
void FN(orc_spline_eval_smooth)(int k, const KT *t, int64_t nt, const VT *c, int64_t ncoef,
                                const KT *x, int64_t n, VT *out) {
    FN(oknots) ts = FN(make_knots_view)(0, k, t, nt);
    for (int64_t p = 0; p < n; ++p) {
        int zone;
        int64_t i = FN(find_knot_interval)(&ts, x[p], &zone);
        if (i < k) i = k;
        if (i > ncoef + k) i = ncoef + k;
        out[p] = FN(spline_kernel)(&ts, c, ncoef, i, x[p], k);
    }
}

/* Derivative(n) * S: _derivative (spline.jl:293-330; periodic: Splines/periodic.jl:76-117).
 * plain:    du_out has ncoef - nd entries (view du[1+nd : N]); new knots = t[1+nd : end-nd].
 * periodic: du_out has ncoef entries; new basis = same breakpoints, order k - nd. */
int FN(orc_spline_derivative)(int kind, int k, const KT *t, int64_t nt, const VT *c, int64_t ncoef,
                              int nd, VT *du_out) {
    if (nd >= k || nd < 1) return -1;
    VT *du = (VT *)malloc(sizeof(VT) * (size_t)ncoef);
    if (kind == 0) {
        for (int64_t i = 0; i < ncoef; ++i) du[i] = c[i];
        for (int m = 1; m <= nd; ++m)
            for (int64_t i = ncoef; i >= 1; --i) {
                KT dt = t[i + k - m - 1] - t[i - 1];
                if (dt == 0 || i == 1) du[i - 1] = 0;
                else du[i - 1] = (VT)(k - m) * (du[i - 1] - du[i - 2]) / (VT)dt;
            }
        for (int64_t i = 0; i < ncoef - nd; ++i) du_out[i] = du[i + nd];
    } else {
        FN(oknots) tsd = FN(make_knots_view)(1, k - nd, t, nt);   /* knots(B') */
        int64_t N = ncoef;
        int delta = k / 2 - (k - nd) / 2;
        for (int64_t i = 1; i <= N; ++i) {
            int64_t j = i - delta;
            while (j < 1) j += N;
            while (j > N) j -= N;
            du[j - 1] = c[i - 1];
        }
        for (int m = 1; m <= nd; ++m) {
            VT du0 = du[N - 1];
            for (int64_t i = N; i >= 1; --i) {
                KT dt = FN(K)(&tsd, i + k - m) - FN(K)(&tsd, i);
                VT prev = (i == 1) ? du0 : du[i - 2];
                du[i - 1] = (VT)(k - m) * (du[i - 1] - prev) / (VT)dt;
            }
        }
        for (int64_t i = 0; i < N; ++i) du_out[i] = du[i];
    }
    free(du);
    return 0;
}
#endif

#undef CAT_
#undef CAT
#undef FN
