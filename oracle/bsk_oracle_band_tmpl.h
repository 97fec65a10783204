/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Collocation assembly, BANFAC, BANSLV and the cyclic tridiagonal solve for an element type
 * other than double (the double versions are written out in bsk_oracle.c).  Included by
 * bsk_oracle.c with
 *     RT    element type of knots, points, matrix and right-hand sides (float)
 *     RSFX  symbol suffix (f32);  EFN(name) = the evaluate_all instantiation to use (…_ff)
 * Same operation order as the reference: Collocation.jl:209-251, matrix.jl:132-201,218-271,
 * cyclic.jl:112-197 — the reference code is generic in T, so Float32 data runs the very same
 * statements in Float32.
 */
#define RCAT_(a, b) a##_##b
#define RCAT(a, b) RCAT_(a, b)
#define RFN(name) RCAT(name, RSFX)

int64_t RFN(orc_collocation_matrix)(int kind, int k, const RT *t, int64_t nt, int cl, int cr, int nl,
                                    int nr, int ul_rows, const double *ul, int lr_rows,
                                    const double *lr, const RT *x, int64_t nx, int nderiv, RT clip,
                                    RT *band /* (2k-3) x ncols */, int64_t ncols) {
    EFN(oknots) ts = EFN(make_knots_view)(0, k, t, nt);
    EFN(orecomb) A = {cl, cr, nl, nr, ul_rows, lr_rows, ul, lr, nt - k};
    int l = k - 2, u = k - 2;
    int nroww = l + u + 1;
    int64_t N = nt - k;
    RT a = t[k - 1], b = t[N];
    memset(band, 0, sizeof(RT) * (size_t)nroww * (size_t)ncols);
    int64_t outside = 0;
    for (int64_t i = 1; i <= nx; ++i) {
        RT xi = x[i - 1];
        if (!(xi >= a && xi <= b)) continue;
        RT bs[16];
        int64_t jlast;
        if (kind == 2) jlast = EFN(eval_all_recombined)(&ts, &A, xi, k, nderiv, 0, bs);
        else jlast = EFN(eval_all)(&ts, xi, k, nderiv, 0, bs);
        for (int dj = 1; dj <= k; ++dj) {
            int64_t j = jlast + 1 - dj;
            RT bj = bs[dj - 1];
            if ((bj < 0 ? -bj : bj) <= clip) continue;
            if ((i > j && i - j > l) || (i < j && j - i > u) || j < 1 || j > ncols) {
                outside++;
                continue;
            }
            band[(u + i - j) + (j - 1) * (int64_t)nroww] = bj;
        }
    }
    return outside;
}

int64_t RFN(orc_collocation_cyclic)(int k, const RT *data, int64_t ndata, const RT *x, int64_t nx,
                                    int nderiv, RT clip, RT *a, RT *b, RT *c) {
    EFN(oknots) ts = EFN(make_knots_view)(1, k, data, ndata);
    int64_t n = ts.N;
    int64_t outside = 0;
    for (int64_t i = 0; i < n; ++i) a[i] = b[i] = c[i] = 0;
    for (int64_t i = 1; i <= nx; ++i) {
        RT bs[16];
        int64_t jlast = EFN(eval_all)(&ts, x[i - 1], k, nderiv, 0, bs);
        for (int dj = 1; dj <= k; ++dj) {
            int64_t j = jlast + 1 - dj;
            while (j < 1) j += n;
            while (j > n) j -= n;
            RT bj = bs[dj - 1];
            if ((bj < 0 ? -bj : bj) <= clip) continue;
            int64_t im = (i == 1) ? n : i - 1, ip = (i == n) ? 1 : i + 1;
            if (i == j) b[i - 1] = bj;
            else if (j == im) a[i - 1] = bj;
            else if (j == ip) c[i - 1] = bj;
            else outside++;
        }
    }
    return outside;
}

#define RW(r, c) w[((r) - 1) + ((c) - 1) * (int64_t)nroww]
int64_t RFN(orc_banded_lu)(RT *w, int nbandl, int nbandu, int64_t nrow) {
    int nroww = nbandl + nbandu + 1;
    int middle = nbandu + 1;
    if (nrow == 1) return RW(middle, nrow) == 0 ? 1 : 0;
    if (nbandl == 0) {
        for (int64_t i = 1; i <= nrow; ++i)
            if (RW(middle, i) == 0) return i;
        return 0;
    }
    for (int64_t i = 1; i <= nrow - 1; ++i) {
        RT pivot = RW(middle, i);
        if (pivot == 0) return i;
        RT ipiv = (RT)1 / pivot;
        int64_t jmax = nbandl < nrow - i ? nbandl : nrow - i;
        for (int64_t j = 1; j <= jmax; ++j) RW(middle + j, i) *= ipiv;
        int64_t kmax = nbandu < nrow - i ? nbandu : nrow - i;
        for (int64_t kk = 1; kk <= kmax; ++kk) {
            RT factor = RW(middle - kk, i + kk);
            for (int64_t j = 1; j <= jmax; ++j)
                RW(middle - kk + j, i + kk) -= factor * RW(middle + j, i);
        }
    }
    if (RW(middle, nrow) == 0) return nrow;
    return 0;
}

void RFN(orc_banded_solve)(const RT *w, int nbandl, int nbandu, int64_t nrow, RT *rhs, int64_t nrhs) {
    int nroww = nbandl + nbandu + 1;
    int middle = nbandu + 1;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < nrhs; ++c) {
        RT *x = rhs + c;
#define RX(i) x[((i) - 1) * nrhs]
        if (nrow == 1) { RX(1) /= RW(middle, 1); continue; }
        if (nbandl != 0) {
            for (int64_t i = 1; i <= nrow - 1; ++i) {
                int64_t jmax = nbandl < nrow - i ? nbandl : nrow - i;
                for (int64_t j = 1; j <= jmax; ++j) RX(i + j) -= RX(i) * RW(middle + j, i);
            }
        }
        for (int64_t i = nrow; i >= 2; --i) {
            RX(i) /= RW(middle, i);
            int64_t jmax = nbandu < i - 1 ? nbandu : i - 1;
            for (int64_t j = 1; j <= jmax; ++j) RX(i - j) -= RX(i) * RW(middle - j, i);
        }
        RX(1) /= RW(middle, 1);
#undef RX
    }
}
#undef RW

/* cyclic.jl:112-197 in RT (written operation order; the reference's @fastmath makes the exact
 * bits implementation-defined). */
void RFN(orc_cyclic_solve)(const RT *a, const RT *b_in, const RT *c, int64_t n, RT *d, RT *x,
                           int64_t nrhs) {
    RT *b = (RT *)malloc(sizeof(RT) * (size_t)n);
    RT *q = (RT *)malloc(sizeof(RT) * (size_t)n);
    memcpy(b, b_in, sizeof(RT) * (size_t)n);
    RT a1 = a[0], cn = c[n - 1];
    RT ac = a1 * cn;
    RT gamma = (RT)sqrt((double)ac);         /* correctly rounded sqrt in RT */
    for (int64_t i = 0; i < n; ++i) q[i] = 0;
    q[0] = gamma;
    q[n - 1] = cn;
    RT vn = a1 / gamma;
    b[0] -= gamma;
    b[n - 1] -= ac / gamma;
    for (int64_t i = 2; i <= n; ++i) {
        RT w = a[i - 1] / b[i - 2];
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
    RT vq = q[0] + vn * q[n - 1];
    for (int64_t r = 0; r < nrhs; ++r) {
        RT vy = d[r] + vn * d[(n - 1) * nrhs + r];
        RT alpha = vy / ((RT)1 + vq);
        for (int64_t i = 0; i < n; ++i) x[i * nrhs + r] = d[i * nrhs + r] - alpha * q[i];
    }
    free(b);
    free(q);
}

#undef RFN
#undef RCAT
#undef RCAT_
