/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement of the BSplineKit.jl v0.19.2 hot path (SURVEY.md §8a rows
 * a1..a17), written in the reference's operation order so that class-A golden
 * vectors (docstring values with decimal inputs) reproduce bit-for-bit.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The Julia reference itself cannot run in this
 * image (no julia binary), so this restatement is pinned against the reference's
 * docstring goldens and restated property tests (tests/test_oracle_*.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 * All citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- templated part: (knot type, value type) instantiations -------------------- */
#define KT double
#define VT double
#define SFX dd
#define KT_IS_WIDER 0
#include "bsk_oracle_tmpl.h"
#undef KT
#undef VT
#undef SFX
#undef KT_IS_WIDER

#define KT double
#define VT float
#define SFX df
#define KT_IS_WIDER 1
#include "bsk_oracle_tmpl.h"
#undef KT
#undef VT
#undef SFX
#undef KT_IS_WIDER

#define KT float
#define VT float
#define SFX ff
#define KT_IS_WIDER 0
#include "bsk_oracle_tmpl.h"
#undef KT
#undef VT
#undef SFX
#undef KT_IS_WIDER

/* ---- Float32 collocation / BANFAC / BANSLV / cyclic (the reference code is generic in T) ---- */
#define RT float
#define RSFX f32
#define EFN(name) name##_ff
#include "bsk_oracle_band_tmpl.h"
#undef RT
#undef RSFX
#undef EFN

/* ================================================================================
 * Single basis function, recursive Cox-de Boor (src/BSplines/basis_function.jl)
 * ================================================================================ */

/* in_local_support (basis_function.jl:259) */
static int in_local_support(double x, const double *t, int64_t nt, double ta, double tb) {
    return (ta <= x && x < tb) || (x == tb && tb == t[nt - 1]);
}

/* _evaluate(::BSplineOrder{k}, t, i, x, T)  (basis_function.jl:200-256); i is 1-based */
static double bfun_value(int k, const double *t, int64_t nt, int64_t i, double x) {
    if (k == 1) return in_local_support(x, t, nt, t[i - 1], t[i]) ? 1.0 : 0.0;
    double ta = t[i - 1], tb = t[i + k - 1];
    if (!in_local_support(x, t, nt, ta, tb)) return 0.0;
    double ta1 = t[i], tb1 = t[i + k - 2];
    double y = 0.0;
    if (tb1 != ta) {
        double dy = bfun_value(k - 1, t, nt, i, x) * (x - ta) / (tb1 - ta);
        y += dy;
    }
    if (ta1 != tb) {
        double dy = bfun_value(k - 1, t, nt, i + 1, x) * (tb - x) / (tb - ta1);
        y += dy;
    }
    return y;
}

/* evaluate_diff(::Derivative{N}, ...)  (basis_function.jl:138-170) */
static double bfun_diff(int nd, int k, const double *t, int64_t nt, int64_t i, double x) {
    if (nd == 0) return bfun_value(k, t, nt, i, x);
    double y = 0.0;
    double dt = t[i + k - 2] - t[i - 1];
    if (dt != 0.0) {
        double dy = bfun_diff(nd - 1, k - 1, t, nt, i, x) / dt;
        y += dy;
    }
    dt = t[i + k - 1] - t[i];
    if (dt != 0.0) {
        double dy = bfun_diff(nd - 1, k - 1, t, nt, i + 1, x) / dt;
        y -= dy;
    }
    return y * (double)(k - 1);
}

/* B[i](x, Derivative(nd)) for a plain knot vector */
double orc_basis_function(int k, const double *t, int64_t nt, int64_t i, double x, int nd) {
    return bfun_diff(nd, k, t, nt, i, x);
}

/* ================================================================================
 * Collocation matrix assembly (src/Collocation/Collocation.jl:209-251) into the
 * BandedMatrices layout data[u + 1 + i - j, j] (matrix.jl:105-117), (2k-3) x nx
 * column-major.  kind 0 plain / 2 recombined (knots = parent's).  Returns the
 * number of non-zeros found outside the bands (the reference @warns for each and
 * then the banded setindex! would throw; we count and skip the store).
 * ================================================================================ */
int64_t orc_collocation_matrix(int kind, int k, const double *t, int64_t nt,
                               int cl, int cr, int nl, int nr, int ul_rows, const double *ul,
                               int lr_rows, const double *lr,
                               const double *x, int64_t nx, int nderiv, double clip,
                               double *band /* (2k-3) x ncols */, int64_t ncols) {
    oknots_dd ts = make_knots_view_dd(0, k, t, nt);
    orecomb_dd A = {cl, cr, nl, nr, ul_rows, lr_rows, ul, lr, nt - k};
    int l = k - 2, u = k - 2;
    int nroww = l + u + 1;
    int64_t N = nt - k;                     /* parent length */
    double a = t[k - 1], b = t[N];          /* boundaries (basis.jl:217-222) */
    memset(band, 0, sizeof(double) * (size_t)nroww * (size_t)ncols);
    int64_t outside = 0;
    for (int64_t i = 1; i <= nx; ++i) {
        double xi = x[i - 1];
        if (!(xi >= a && xi <= b)) continue; /* isindomain (basis.jl:229-234); NaN is outside */
        double bs[16];
        int64_t jlast;
        if (kind == 2) jlast = eval_all_recombined_dd(&ts, &A, xi, k, nderiv, 0, bs);
        else jlast = eval_all_dd(&ts, xi, k, nderiv, 0, bs);
        for (int dj = 1; dj <= k; ++dj) {
            int64_t j = jlast + 1 - dj;
            double bj = bs[dj - 1];
            if (fabs(bj) <= clip) continue;
            if ((i > j && i - j > l) || (i < j && j - i > u) || j < 1 || j > ncols) {
                outside++;
                continue;
            }
            band[(u + i - j) + (j - 1) * (int64_t)nroww] = bj;
        }
    }
    return outside;
}

/* Periodic cubic case: CyclicTridiagonalMatrix (cyclic.jl:21-99 + Collocation.jl:209-251
 * with basis_to_array_index wrap, BSplines/periodic.jl:245-253).  a = sub-diagonal
 * (a[1] couples row 1 with column n), b = diagonal, c = super-diagonal (c[n] couples
 * row n with column 1).  Returns number of entries that fall outside the three
 * cyclic diagonals (reference: DomainError). */
int64_t orc_collocation_cyclic(int k, const double *data, int64_t ndata, const double *x,
                               int64_t nx, int nderiv, double clip, double *a, double *b,
                               double *c) {
    oknots_dd ts = make_knots_view_dd(1, k, data, ndata);
    int64_t n = ts.N;
    int64_t outside = 0;
    for (int64_t i = 0; i < n; ++i) a[i] = b[i] = c[i] = 0.0;
    for (int64_t i = 1; i <= nx; ++i) {
        double bs[16];
        int64_t jlast = eval_all_dd(&ts, x[i - 1], k, nderiv, 0, bs);
        for (int dj = 1; dj <= k; ++dj) {
            int64_t j = jlast + 1 - dj;
            while (j < 1) j += n;
            while (j > n) j -= n;
            double bj = bs[dj - 1];
            if (fabs(bj) <= clip) continue;
            int64_t im = (i == 1) ? n : i - 1, ip = (i == n) ? 1 : i + 1;
            if (i == j) b[i - 1] = bj;
            else if (j == im) a[i - 1] = bj;
            else if (j == ip) c[i - 1] = bj;
            else outside++;
        }
    }
    return outside;
}

/* ================================================================================
 * Banded LU without pivoting = BANFAC (src/Collocation/matrix.jl:132-201).
 * w is nroww x nrow column-major, nroww = l + u + 1, middle = u + 1.
 * Returns 0, or the 1-based index of the zero pivot (ZeroPivotException(i)).
 * ================================================================================ */
#define W(r, c) w[((r) - 1) + ((c) - 1) * (int64_t)nroww]
int64_t orc_banded_lu(double *w, int nbandl, int nbandu, int64_t nrow) {
    int nroww = nbandl + nbandu + 1;
    int middle = nbandu + 1;
    if (nrow == 1) return W(middle, nrow) == 0.0 ? 1 : 0;
    if (nbandl == 0) {
        for (int64_t i = 1; i <= nrow; ++i)
            if (W(middle, i) == 0.0) return i;
        return 0;
    }
    if (nbandu == 0) {
        for (int64_t i = 1; i <= nrow - 1; ++i) {
            double pivot = W(middle, i);
            if (pivot == 0.0) return i;
            double ipiv = 1.0 / pivot;
            int64_t jmax = nbandl < nrow - i ? nbandl : nrow - i;
            for (int64_t j = 1; j <= jmax; ++j) W(middle + j, i) *= ipiv;
        }
        return 0;
    }
    for (int64_t i = 1; i <= nrow - 1; ++i) {
        double pivot = W(middle, i);
        if (pivot == 0.0) return i;
        double ipiv = 1.0 / pivot;
        int64_t jmax = nbandl < nrow - i ? nbandl : nrow - i;
        for (int64_t j = 1; j <= jmax; ++j) W(middle + j, i) *= ipiv;
        int64_t kmax = nbandu < nrow - i ? nbandu : nrow - i;
        for (int64_t kk = 1; kk <= kmax; ++kk) {
            double factor = W(middle - kk, i + kk);
            for (int64_t j = 1; j <= jmax; ++j)
                W(middle - kk + j, i + kk) -= factor * W(middle + j, i);
        }
    }
    if (W(middle, nrow) == 0.0) return nrow;
    return 0;
}

/* BANSLV (src/Collocation/matrix.jl:218-271) applied to nrhs right-hand sides stored
 * interleaved: rhs[(i-1)*nrhs + c]  (== Vector{SVector{nrhs}} in the reference,
 * test/interpolation.jl:57).  In place. */
void orc_banded_solve(const double *w, int nbandl, int nbandu, int64_t nrow, double *rhs,
                      int64_t nrhs) {
    int nroww = nbandl + nbandu + 1;
    int middle = nbandu + 1;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < nrhs; ++c) {
        double *x = rhs + c;
#define X(i) x[((i) - 1) * nrhs]
        if (nrow == 1) { X(1) /= W(middle, 1); continue; }
        if (nbandl != 0) {
            for (int64_t i = 1; i <= nrow - 1; ++i) {
                int64_t jmax = nbandl < nrow - i ? nbandl : nrow - i;
                for (int64_t j = 1; j <= jmax; ++j) X(i + j) -= X(i) * W(middle + j, i);
            }
        }
        for (int64_t i = nrow; i >= 2; --i) {
            X(i) /= W(middle, i);
            int64_t jmax = nbandu < i - 1 ? nbandu : i - 1;
            for (int64_t j = 1; j <= jmax; ++j) X(i - j) -= X(i) * W(middle - j, i);
        }
        X(1) /= W(middle, 1);
#undef X
    }
}
#undef W

/* C * u for the (unfactorised) band storage; used by tests for residuals
 * (the reference routes this through BandedMatrices gbmv, matrix.jl:91-94). */
void orc_banded_matvec(const double *w, int nbandl, int nbandu, int64_t n, const double *xin,
                       double *yout, int64_t nrhs) {
    int nroww = nbandl + nbandu + 1;
    for (int64_t c = 0; c < nrhs; ++c)
        for (int64_t i = 1; i <= n; ++i) {
            double s = 0.0;
            int64_t j0 = i - nbandl < 1 ? 1 : i - nbandl;
            int64_t j1 = i + nbandu > n ? n : i + nbandu;
            for (int64_t j = j0; j <= j1; ++j)
                s += w[(nbandu + i - j) + (j - 1) * (int64_t)nroww] * xin[(j - 1) * nrhs + c];
            yout[(i - 1) * nrhs + c] = s;
        }
}

/* ================================================================================
 * Cyclic tridiagonal solve, Sherman-Morrison around Thomas
 * (src/Collocation/cyclic.jl:112-161, solve_thomas! :168-197).  The reference
 * destroys b and d and re-copies the matrix before every solve
 * (SplineInterpolations.jl:147-149); here a, b, c are const and copied internally.
 * d, x: interleaved n x nrhs (rhs index fastest).  d is used as workspace (destroyed).
 * solve_thomas! is @fastmath in the reference, i.e. not bitwise-defined; this
 * follows the written operation order.
 * ================================================================================ */
void orc_cyclic_solve(const double *a, const double *b_in, const double *c, int64_t n, double *d,
                      double *x, int64_t nrhs) {
    double *b = (double *)malloc(sizeof(double) * (size_t)n);
    double *q = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(b, b_in, sizeof(double) * (size_t)n);
    double a1 = a[0], cn = c[n - 1];
    double ac = a1 * cn;
    double gamma = sqrt(ac);
    for (int64_t i = 0; i < n; ++i) q[i] = 0.0;
    q[0] = gamma;
    q[n - 1] = cn;
    double v1 = 1.0, vn = a1 / gamma;
    b[0] -= gamma;
    b[n - 1] -= ac / gamma;
    /* solve_thomas!((y, q), (a, b, c), (d, u)) */
    for (int64_t i = 2; i <= n; ++i) {
        double w = a[i - 1] / b[i - 2];
        b[i - 1] = b[i - 1] - w * c[i - 2];
        for (int64_t r = 0; r < nrhs; ++r) d[(i - 1) * nrhs + r] = d[(i - 1) * nrhs + r] - w * d[(i - 2) * nrhs + r];
        q[i - 1] = q[i - 1] - w * q[i - 2];
    }
    for (int64_t r = 0; r < nrhs; ++r) d[(n - 1) * nrhs + r] = d[(n - 1) * nrhs + r] / b[n - 1];
    q[n - 1] = q[n - 1] / b[n - 1];
    for (int64_t i = n - 1; i >= 1; --i) {
        for (int64_t r = 0; r < nrhs; ++r)
            d[(i - 1) * nrhs + r] = (d[(i - 1) * nrhs + r] - c[i - 1] * d[i * nrhs + r]) / b[i - 1];
        q[i - 1] = (q[i - 1] - c[i - 1] * q[i]) / b[i - 1];
    }
    double vq = v1 * q[0] + vn * q[n - 1];
    for (int64_t r = 0; r < nrhs; ++r) {
        double vy = v1 * d[r] + vn * d[(n - 1) * nrhs + r];
        double alpha = vy / (1.0 + vq);
        for (int64_t i = 0; i < n; ++i) x[i * nrhs + r] = d[i * nrhs + r] - alpha * q[i];
    }
    free(b);
    free(q);
}

/* ================================================================================
 * Knot placement for interpolation: make_knots
 * (src/SplineInterpolations/SplineInterpolations.jl:302-326 and :333-371)
 * ================================================================================ */
void orc_make_knots(const double *x, int64_t N, int k, double *t /* N + k */) {
    for (int i = 0; i < k; ++i) { t[i] = x[0]; t[N + i] = x[N - 1]; }
    double kinv = 1.0 / (double)(k - 1);
    for (int64_t i = 1; i <= N - k; ++i) {
        double ti = 0.0;
        for (int j = 1; j <= k - 1; ++j) ti += x[i + j - 1];
        t[k + i - 1] = ti * kinv;
    }
}

void orc_make_knots_periodic(const double *xs, int64_t N, int k, double L, double *ts /* N+1 */) {
    if (k % 2 == 0) {
        for (int64_t i = 0; i < N; ++i) ts[i] = xs[i];
        ts[N] = xs[0] + L;
        return;
    }
    double x0 = xs[0], xn = x0 + L, xprev = x0;
    int64_t i = 1;
    while (i < N) {
        double xx = xs[i];
        ts[i - 1] = (xprev + xx) / 2;
        xprev = xx;
        i += 1;
    }
    ts[N - 1] = (xprev + xn) / 2;
    ts[0] = x0;
    ts[N] = xn;
}

/* ================================================================================
 * Greville sites with the two-step Kahan sliding sum
 * (src/Collocation/points.jl:56-126).  kind 0 plain (endpoints forced to a, b),
 * kind 2 recombined (skip cl boundary points, no forcing).  N = number of points.
 * ================================================================================ */
void orc_greville(int kind, int k, const double *t, int64_t nt, int cl, int64_t N, double *out) {
    int64_t Np = nt - k;
    double a = t[k - 1], b = t[Np];
    double tsum = 0.0, comp = 0.0;
    for (int64_t i = 0; i < N; ++i) {
        int64_t ii = i + cl;
        if (i == 0) {
            tsum = 0.0;
            for (int j = 2; j <= k; ++j) tsum += t[ii + j - 1];
        } else {
            double tsum_prev = tsum;
            double y = t[ii + k - 1] - comp;
            tsum = tsum_prev + y;
            comp = (tsum - tsum_prev) - y;
            tsum_prev = tsum;
            y = -t[ii] - comp;
            tsum = tsum_prev + y;
            comp = (tsum - tsum_prev) - y;
        }
        double xx = tsum / (double)(k - 1);
        if (kind == 0) {
            if (i == 0) xx = a;
            else if (i == N - 1) xx = b;
        }
        if (xx < a) xx = a;
        if (xx > b) xx = b;
        out[i] = xx;
    }
}

/* ================================================================================
 * Robin-type corner blocks of the recombination matrix for a single operator
 * op = sum_m alpha[m] D^m  (matrices.jl:253-313 with nc = 1, ndrop = 0, q = n).
 * Left: ul is (q+1) x q; right: lr is (q+1) x q, column-major.
 * ================================================================================ */
static double op_eval(int k, const double *t, int64_t nt, int64_t i, double x, const double *alpha,
                      int maxord, int left) {
    /* DifferentialOpSum evaluates each term then sums (basis_function.jl:184-191);
     * dot(op, LeftNormal()) flips the sign of odd derivatives (DifferentialOps.jl:128-131,160) */
    double s = 0.0;
    int first = 1;
    for (int m = 0; m <= maxord; ++m) {
        if (alpha[m] == 0.0) continue;
        double am = alpha[m];
        if (left && (m & 1)) am = -am;
        double v = am * bfun_diff(m, k, t, nt, i, x);
        if (first) { s = v; first = 0; } else s = s + v;
    }
    return s;
}

void orc_robin_corners(int k, const double *t, int64_t nt, const double *alpha, int maxord,
                       double *ul, double *lr) {
    int64_t N = nt - k;
    int q = maxord, nc = 1, ndrop = 0;
    int rows = q + nc;
    double xa = t[k - 1], xb = t[N];
    double bs[16];
    for (int i = 0; i < rows * q; ++i) ul[i] = lr[i] = 0.0;
    for (int d = 1; d <= q + 1; ++d) bs[d - 1] = op_eval(k, t, nt, ndrop + d, xa, alpha, maxord, 1);
    for (int l = 1; l <= q; ++l) {
        double b0 = bs[l - 1], b1 = bs[l];
        double r = 2 / (b0 - b1);
        ul[(l + nc - 1 - 1) + (l - 1) * rows] = -b1 * r;
        ul[(l + nc - 1) + (l - 1) * rows] = b0 * r;
    }
    int64_t M = N - ndrop;
    for (int d = 1; d <= q + 1; ++d) bs[d - 1] = op_eval(k, t, nt, M - d + 1, xb, alpha, maxord, 0);
    for (int l = 1; l <= q; ++l) {
        double b0 = bs[q - l + 1], b1 = bs[q - l];
        double r = 2 / (b0 - b1);
        lr[(l - 1) + (l - 1) * rows] = -b1 * r;
        lr[(l) + (l - 1) * rows] = b0 * r;
    }
}
