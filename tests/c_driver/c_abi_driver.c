/* c_abi_driver.c — a plain C program that drives the hot path through include/bsk.h only (no
 * Python, no torch): what a C / C++ / Julia-ccall caller links against.  It is TEST code: it also
 * links the CPU oracle (oracle/libbsk_oracle.so) as the checker.
 *
 *   C2-shaped evaluation: k = 4, 1000 non-uniform breakpoints, value + 1st derivative at N points
 *   C1-shaped interpolation: xdata = (0:10).^2, make_knots + collocation + LU + solve on the device
 *
 * Prints one line per check and "C-ABI DRIVER PASS" at the end; exit status 0 on success. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bsk.h"

/* oracle (checker only) */
void orc_spline_eval_dd(int kind, int k, const double *t, int64_t nt, const double *c, int64_t ncoef,
                        const double *x, int64_t n, double *out);
int orc_spline_derivative_dd(int kind, int k, const double *t, int64_t nt, const double *c, int64_t ncoef,
                             int nd, double *du_out);   /* new knots = t[1+nd : end-nd] */
int64_t orc_collocation_matrix(int kind, int k, const double *t, int64_t nt, int cl, int cr, int nl, int nr,
                               int ul_rows, const double *ul, int lr_rows, const double *lr, const double *x,
                               int64_t nx, int nderiv, double clip, double *band, int64_t ncols);

static double u01(uint64_t seed, uint64_t i) {          /* synth.py: splitmix64 */
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + i + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * 0x1.0p-53;
}

#define CHECK(call)                                                            \
    do {                                                                       \
        bsk_status st_ = (call);                                               \
        if (st_ != BSK_OK) {                                                   \
            char msg[512];                                                     \
            bsk_last_error(msg, sizeof msg);                                   \
            fprintf(stderr, "%s -> %d: %s\n", #call, (int)st_, msg);           \
            return 1;                                                          \
        }                                                                      \
    } while (0)

int main(int argc, char **argv) {
    const int64_t N = argc > 1 ? atoll(argv[1]) : (1ll << 24);
    int32_t ndev = 0;
    CHECK(bsk_device_count(&ndev));
    if (ndev < 1) { fprintf(stderr, "no CUDA device\n"); return 2; }
    printf("%s, %d device(s)\n", bsk_version(), ndev);

    /* ---- C2-shaped evaluation ------------------------------------------------------------ */
    enum { K = 4, NB = 1000 };
    const int64_t nt = NB + 2 * (K - 1), ncoef = nt - K;
    double *t = malloc(sizeof(double) * nt), *c = malloc(sizeof(double) * ncoef);
    for (int i = 0; i < NB; ++i) {
        double xi = ((double)i + 0.4 * (2.0 * u01(3, i) - 1.0)) / (NB - 1);
        if (i == 0) xi = 0.0;
        if (i == NB - 1) xi = 1.0;
        t[K - 1 + i] = xi;
    }
    for (int i = 0; i < K - 1; ++i) { t[i] = 0.0; t[nt - 1 - i] = 1.0; }
    for (int64_t j = 0; j < ncoef; ++j) c[j] = 2.0 * u01(4, j) - 1.0;
    double *x = malloc(sizeof(double) * N);
    for (int64_t j = 0; j < N; ++j) x[j] = u01(5, j);

    bsk_basis *B = NULL;
    bsk_spline *S = NULL;
    CHECK(bsk_basis_create(BSK_BASIS_PLAIN, K, BSK_F64, t, nt, NULL, 0, &B));
    CHECK(bsk_spline_create(B, c, ncoef, &S));
    void *dx = NULL, *dv = NULL, *dd = NULL, *e0 = NULL, *e1 = NULL;
    CHECK(bsk_malloc(0, sizeof(double) * N, &dx));
    CHECK(bsk_malloc(0, sizeof(double) * N, &dv));
    CHECK(bsk_malloc(0, sizeof(double) * N, &dd));
    CHECK(bsk_memcpy_h2d(0, dx, x, sizeof(double) * N, NULL));
    CHECK(bsk_event_create(0, &e0));
    CHECK(bsk_event_create(0, &e1));
    int32_t derivs[2] = {0, 1};
    void *outs[2] = {dv, dd};
    CHECK(bsk_spline_eval(S, dx, N, 2, derivs, outs, BSK_EVAL_AUTO, NULL));      /* warm-up: builds the tables */
    CHECK(bsk_event_record(e0, NULL));
    for (int r = 0; r < 5; ++r) CHECK(bsk_spline_eval(S, dx, N, 2, derivs, outs, BSK_EVAL_AUTO, NULL));
    CHECK(bsk_event_record(e1, NULL));
    float ms = 0.f;
    CHECK(bsk_event_elapsed_ms(e0, e1, &ms));
    printf("eval: %lld points, value + derivative: %.3f ms per pass, %.1f Gpoints/s\n", (long long)N, ms / 5,
           (double)N / (ms / 5) / 1e6);
    const int64_t NC = N < 65536 ? N : 65536;
    double *hv = malloc(sizeof(double) * NC), *hd = malloc(sizeof(double) * NC);
    double *rv = malloc(sizeof(double) * NC), *rd = malloc(sizeof(double) * NC);
    CHECK(bsk_memcpy_d2h(0, hv, dv, sizeof(double) * NC, NULL));
    CHECK(bsk_memcpy_d2h(0, hd, dd, sizeof(double) * NC, NULL));
    CHECK(bsk_stream_sync(0, NULL));
    orc_spline_eval_dd(0, K, t, nt, c, ncoef, x, NC, rv);
    double *c1 = malloc(sizeof(double) * ncoef);
    orc_spline_derivative_dd(0, K, t, nt, c, ncoef, 1, c1);
    orc_spline_eval_dd(0, K - 1, t + 1, nt - 2, c1, ncoef - 1, x, NC, rd);
    double ev = 0, ed = 0, mv = 0, md = 0;
    for (int64_t j = 0; j < NC; ++j) {
        ev = fmax(ev, fabs(hv[j] - rv[j])); mv = fmax(mv, fabs(rv[j]));
        ed = fmax(ed, fabs(hd[j] - rd[j])); md = fmax(md, fabs(rd[j]));
    }
    printf("parity vs oracle on %lld points: value %.2e, derivative %.2e (relative to max)\n", (long long)NC, ev / mv, ed / md);
    if (!(ev / mv < 1e-13 && ed / md < 1e-13)) { fprintf(stderr, "evaluation parity FAILED\n"); return 1; }
    /* the bit-faithful path must agree bit for bit */
    CHECK(bsk_spline_eval(S, dx, NC, 1, derivs, outs, BSK_EVAL_DEBOOR, NULL));
    CHECK(bsk_memcpy_d2h(0, hv, dv, sizeof(double) * NC, NULL));
    CHECK(bsk_stream_sync(0, NULL));
    if (memcmp(hv, rv, sizeof(double) * NC) != 0) { fprintf(stderr, "de Boor path is not bit-exact\n"); return 1; }
    printf("de Boor path: bit-exact on %lld points\n", (long long)NC);

    /* ---- C1-shaped interpolation: xdata = (0:10).^2, k = 4 ----------------------------------- */
    enum { NI = 11 };
    double xd[NI], yd[NI], ti[NI + K];
    for (int i = 0; i < NI; ++i) { xd[i] = (double)(i * i); yd[i] = u01(1, i); }
    for (int i = 0; i < K; ++i) { ti[i] = xd[0]; ti[NI + i] = xd[NI - 1]; }
    for (int i = 1; i <= NI - K; ++i) {                 /* make_knots, SplineInterpolations.jl:302-326 */
        double s = 0.0;
        for (int j = 1; j < K; ++j) s += xd[i + j - 1];
        ti[K + i - 1] = s * (1.0 / (K - 1));
    }
    bsk_basis *Bi = NULL;
    bsk_banded *F = NULL;
    CHECK(bsk_basis_create(BSK_BASIS_PLAIN, K, BSK_F64, ti, NI + K, NULL, 0, &Bi));
    void *dxd = NULL, *dband = NULL;
    const int nbands = 2 * K - 3;
    CHECK(bsk_malloc(0, sizeof xd, &dxd));
    CHECK(bsk_malloc(0, sizeof(double) * nbands * NI, &dband));
    CHECK(bsk_memcpy_h2d(0, dxd, xd, sizeof xd, NULL));
    int64_t nout = 0, info = 0;
    CHECK(bsk_collocation_matrix(Bi, dxd, NI, 0, 2.220446049250313e-16, dband, &nout, NULL));
    CHECK(bsk_banded_create(dband, NI, K - 2, K - 2, 0, &F));
    double band[5 * NI], ref[5 * NI];
    CHECK(bsk_banded_data(F, band));
    memset(ref, 0, sizeof ref);
    orc_collocation_matrix(0, K, ti, NI + K, 0, 0, 0, 0, 0, NULL, 0, NULL, xd, NI, 0, 2.220446049250313e-16, ref, NI);
    if (memcmp(band, ref, sizeof ref) != 0) { fprintf(stderr, "collocation matrix is not bit-exact\n"); return 1; }
    CHECK(bsk_banded_lu(F, &info));
    double cs[NI];
    memcpy(cs, yd, sizeof yd);
    CHECK(bsk_banded_solve_host(F, cs, 1, BSK_SOLVE_AUTO));
    bsk_spline *Si = NULL;
    CHECK(bsk_spline_create(Bi, cs, NI, &Si));
    double back[NI];
    void *o1[1] = {back};
    CHECK(bsk_spline_eval_host(Si, xd, NI, 1, derivs, o1, BSK_EVAL_AUTO));
    double ei = 0;
    for (int i = 0; i < NI; ++i) ei = fmax(ei, fabs(back[i] - yd[i]));
    printf("interpolation (0:10).^2: collocation bit-exact, max |S(x_i) - y_i| = %.2e\n", ei);
    if (!(ei < 1e-13)) { fprintf(stderr, "interpolation FAILED\n"); return 1; }

    bsk_spline_destroy(Si); bsk_banded_destroy(F); bsk_basis_destroy(Bi);
    bsk_spline_destroy(S); bsk_basis_destroy(B);
    bsk_free(0, dx); bsk_free(0, dv); bsk_free(0, dd); bsk_free(0, dxd); bsk_free(0, dband);
    bsk_event_destroy(e0); bsk_event_destroy(e1);
    printf("C-ABI DRIVER PASS\n");
    return 0;
}
