// bsk_galerkin.cu — K10: Galerkin matrices and projections.
//
// galerkin_matrix!(M, B, (Derivative(m), Derivative(n)))  — src/Galerkin/linear.jl:220-311
// galerkin_projection!(f, phi, B, Derivative(n))          — src/Galerkin/projection.jl:44-99
//
// The reference loops over the knot segments, evaluates all k non-zero B-splines at the k
// Gauss-Legendre nodes of each segment (evaluate_all with a known `ileft`) and accumulates the outer
// products into the banded matrix.  Here the nodes of ALL segments go through one batched
// evaluate_all (K2, bit-exact), and a gather kernel with one thread per matrix entry adds the
// contributions of the segments in the reference's own order (segment by segment, node by node, the
// products formed left to right, no FMA contraction) — same result as the sequential loop, no atomics.
#include <algorithm>

#include "handles.cuh"

using namespace bsk;

namespace {

struct Segments {
    std::vector<double> x, y;          // nodes, alpha * w        (nseg * nq)
    std::vector<int64_t> ileft;        // n - cl per node
    std::vector<int32_t> seg_of_n;     // knot index n (1-based) -> segment slot or -1
    int64_t nseg = 0;
};

// PeriodicKnots ts[1 .. N + k] materialised with the reference's own arithmetic (BSplines/periodic.jl:62,102-115):
// ts[i] = data[i - k / 2] shifted by repeated -+ L.
std::vector<double> periodic_knot_vector(const bsk_basis *b) {
    const int k = b->k;
    const int64_t N = b->N;
    const std::vector<double> &data = b->h_knots;             // N + 1 breakpoints
    const double L = data.back() - data.front();
    std::vector<double> ts((size_t)(N + k));
    for (int64_t i = 1; i <= N + k; ++i) {
        double x = 0.0;
        int64_t j = i - k / 2;
        while (j < 1) { x -= L; j += N; }
        while (j > N + 1) { x += L; j -= N; }
        ts[(size_t)i - 1] = x + data[(size_t)j - 1];
    }
    return ts;
}

// Knot segments and their quadrature nodes (linear.jl:268-286, quadratures.jl:52-65).  Plain / recombined bases:
// the segments inside the domain.  Periodic bases: ALL segments of ts[1 .. N + k] — the projection loop has no
// boundary test (projection.jl:72-96), the matrix loop keeps the N segments of the main period (linear.jl:277-279),
// which the periodic gather kernel selects by index.
bsk_status build_segments(const bsk_basis *b, const double *qx, const double *qw, int nq, Segments &sg) {
    BSK_REQUIRE(b->dtype == BSK_F64, BSK_ERR_UNSUPPORTED, "Galerkin assembly needs a Float64 basis");
    const int k = b->k;
    if (b->kind == BSK_BASIS_PERIODIC) {
        const std::vector<double> ts = periodic_knot_vector(b);
        const int64_t nt = (int64_t)ts.size();
        sg.seg_of_n.assign((size_t)nt + 1, -1);
        for (int64_t n = 1; n < nt; ++n) {
            const double tn = ts[n - 1], tn1 = ts[n];
            if (tn1 == tn) continue;
            const double alpha = (tn1 - tn) / 2, beta = (tn + tn1) / 2;
            sg.seg_of_n[n] = (int32_t)sg.nseg++;
            for (int q = 0; q < nq; ++q) {
                sg.x.push_back(alpha * qx[q] + beta);
                sg.y.push_back(alpha * qw[q]);
                sg.ileft.push_back(n);
            }
        }
        return BSK_OK;
    }
    const int64_t nt = b->nt, N = b->N;
    const std::vector<double> &t = b->h_knots;
    const double xa = t[k - 1], xb = t[N];                 // boundaries(B)
    const int off = b->rv.enabled ? b->rv.cl : 0;
    sg.seg_of_n.assign((size_t)nt + 1, -1);
    for (int64_t n = 1; n < nt; ++n) {
        const double tn = t[n - 1], tn1 = t[n];
        if (tn1 == tn) continue;
        if (tn1 <= xa || tn >= xb) continue;
        const double alpha = (tn1 - tn) / 2, beta = (tn + tn1) / 2;
        sg.seg_of_n[n] = (int32_t)sg.nseg++;
        for (int q = 0; q < nq; ++q) {
            sg.x.push_back(alpha * qx[q] + beta);
            sg.y.push_back(alpha * qw[q]);
            sg.ileft.push_back(n - off);
        }
    }
    return BSK_OK;
}

template <typename V>
bsk_status upload(void **d, const std::vector<V> &h) {
    BSK_CUDA(bsk::dev_malloc(d, std::max<size_t>(1, h.size()) * sizeof(V)));
    BSK_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(V), cudaMemcpyHostToDevice));
    return BSK_OK;
}

// One thread per band entry.  Band store data[bw + i - j, j], bw = k - 1.
__global__ void __launch_bounds__(128)
k_galerkin_gather(const double *__restrict__ bs1, const double *__restrict__ bs2, const int32_t *__restrict__ seg_of_n,
                  const double *__restrict__ y, const double *__restrict__ fx, int nq, int k, int off, int64_t nt,
                  int64_t N, double *__restrict__ band) {
    const int bw = k - 1, nb = 2 * bw + 1;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N * nb; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = e / nb;
        const int r = (int)(e - j * nb);
        const int64_t i = j + r - bw;
        double acc = 0.0;
        if (i >= 0 && i < N) {
            const int64_t i1 = i + 1, j1 = j + 1;
            const int64_t lo = i1 > j1 ? i1 : j1, hi = (i1 < j1 ? i1 : j1) + k - 1;
            for (int64_t ilast = lo; ilast <= hi; ++ilast) {          // segments in increasing order
                const int64_t n = ilast + off;
                if (n < 1 || n >= nt) continue;
                const int s = seg_of_n[n];
                if (s < 0) continue;
                const int di = (int)(ilast + 1 - i1), dj = (int)(ilast + 1 - j1);
                for (int q = 0; q < nq; ++q) {
                    const int64_t p = (int64_t)s * nq + q;
                    const double bi = bs1[p * k + di - 1], bj = bs2[p * k + dj - 1];
                    double term = xmul(xmul(y[p], bi), bj);            // A[i, j] += y * bi * bj * f(x)   (linear.jl:305)
                    if (fx) term = xmul(term, fx[p]);
                    acc = xadd(acc, term);
                }
            }
        }
        band[e] = acc;
    }
}

// Two bases derived from the same parent (e.g. (R, B) or two different recombinations; linear.jl:134-199): the
// matrix is N1 x N2 with l = (k - 1) + dl sub- and u = (k - 1) - dl super-diagonals, dl = cl2 - cl1; one thread per
// entry of its band store data[u + i - j, j].  Basis 1 indexes the segment's functions from n - off1, basis 2 from
// n - off2; segments in increasing order, products left to right, like the sequential loop.
__global__ void __launch_bounds__(128)
k_galerkin_gather_pair(const double *__restrict__ bs1, const double *__restrict__ bs2, const int32_t *__restrict__ seg_of_n,
                       const double *__restrict__ y, const double *__restrict__ fx, int nq, int k, int off1, int off2,
                       int64_t nt, int64_t N1, int64_t N2, int l, int u, double *__restrict__ band) {
    const int nb = l + u + 1;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N2 * nb; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = e / nb;
        const int r = (int)(e - j * nb);
        const int64_t i = j + r - u;
        double acc = 0.0;
        if (i >= 0 && i < N1) {
            const int64_t i1 = i + 1, j1 = j + 1;                  // 1-based indices in their own bases
            const int64_t a1 = i1 + off1, a2 = j1 + off2;          // first knot segment each function lives on
            const int64_t lo = a1 > a2 ? a1 : a2, hi = (a1 < a2 ? a1 : a2) + k - 1;
            for (int64_t n = lo; n <= hi; ++n) {                   // knot segments in increasing order
                if (n < 1 || n >= nt) continue;
                const int s = seg_of_n[n];
                if (s < 0) continue;
                const int di = (int)(n - off1 + 1 - i1), dj = (int)(n - off2 + 1 - j1);
                for (int q = 0; q < nq; ++q) {
                    const int64_t p = (int64_t)s * nq + q;
                    double term = xmul(xmul(y[p], bs1[p * k + di - 1]), bs2[p * k + dj - 1]);
                    if (fx) term = xmul(term, fx[p]);
                    acc = xadd(acc, term);
                }
            }
        }
        band[e] = acc;
    }
}

// Periodic bases: one thread per entry of the CYCLIC band store, data[bw + d, j] = A[(j + d) mod N, j], |d| <= bw =
// k - 1 (N >= 2 k - 1, so that the 2 k - 1 offsets are distinct entries).  Entry (I, J) collects, for every pair
// (delta_i, delta_j) with delta_j - delta_i = d, the main-period segment n in h + 1 .. h + N with n + 1 - delta_j = J
// (mod N); the reference visits the segments in increasing n (linear.jl:268), hence the wrapped ones first.
__global__ void __launch_bounds__(128)
k_galerkin_gather_periodic(const double *__restrict__ bs1, const double *__restrict__ bs2,
                           const int32_t *__restrict__ seg_of_n, const double *__restrict__ y,
                           const double *__restrict__ fx, int nq, int k, int64_t N, double *__restrict__ band) {
    const int bw = k - 1, nb = 2 * bw + 1, h = k / 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N * nb; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = e / nb;
        const int d = (int)(e - j * nb) - bw;                  // i - j (cyclic)
        const int64_t J = j + 1;
        const int dj_lo = d > 0 ? 1 + d : 1, dj_hi = d < 0 ? k + d : k;
        // n(delta_j) = h + 1 + ((J - 1 + delta_j) - (h + 1) mod N): increasing in delta_j up to one wrap
        auto seg_n = [&](int dj) -> int64_t {
            int64_t m = (J - 1 + dj - (h + 1)) % N;
            if (m < 0) m += N;
            return h + 1 + m;
        };
        const int64_t n_first = seg_n(dj_lo);
        double acc = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            for (int dj = dj_lo; dj <= dj_hi; ++dj) {
                const int64_t n = seg_n(dj);
                if ((n < n_first) != (pass == 0)) continue;        // pass 0: the wrapped (smaller) n, pass 1: the rest
                const int s = seg_of_n[n];
                if (s < 0) continue;
                const int di = dj - d;
                for (int q = 0; q < nq; ++q) {
                    const int64_t p = (int64_t)s * nq + q;
                    const double bi = bs1[p * k + di - 1], bj = bs2[p * k + dj - 1];
                    double term = xmul(xmul(y[p], bi), bj);
                    if (fx) term = xmul(term, fx[p]);
                    acc = xadd(acc, term);
                }
            }
        }
        band[e] = acc;
    }
}

__global__ void __launch_bounds__(128)
k_galerkin_project(const double *__restrict__ bs, const int32_t *__restrict__ seg_of_n, const double *__restrict__ y,
                   const double *__restrict__ fx, int nq, int k, int off, int64_t nt, int64_t N,
                   double *__restrict__ phi) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i1 = i + 1;
        double acc = 0.0;
        for (int64_t ilast = i1; ilast <= i1 + k - 1; ++ilast) {
            const int64_t n = ilast + off;
            if (n < 1 || n >= nt) continue;
            const int s = seg_of_n[n];
            if (s < 0) continue;
            const int di = (int)(ilast + 1 - i1);
            for (int q = 0; q < nq; ++q) {
                const int64_t p = (int64_t)s * nq + q;
                acc = xadd(acc, xmul(xmul(y[p], fx[p]), bs[p * k + di - 1]));   // y = alpha * w * f(x); phi[i] += y * bi
            }
        }
        phi[i] = acc;
    }
}

struct DevBufs {
    void *x = nullptr, *y = nullptr, *il = nullptr, *seg = nullptr, *idx = nullptr, *bs1 = nullptr, *bs2 = nullptr;
    ~DevBufs() {
        for (void *p : {x, y, il, seg, idx, bs1, bs2})
            if (p) bsk::dev_free(p);
    }
};

}  // namespace

extern "C" {

// Number of quadrature points (nseg * nq) and, if h_x != NULL, their abscissae in the reference's
// order (segment by segment) — where the caller evaluates f for bsk_galerkin_projection.
bsk_status bsk_galerkin_nodes(const bsk_basis *b, const double *h_qx, int32_t nq, double *h_x, int64_t capacity,
                              int64_t *npts) {
    BSK_REQUIRE(b && h_qx && npts && nq >= 1, BSK_ERR_ARGUMENT, "NULL argument");
    Segments sg;
    std::vector<double> w(nq, 0.0);
    bsk_status rc = build_segments(b, h_qx, w.data(), nq, sg);
    if (rc != BSK_OK) return rc;
    *npts = (int64_t)sg.x.size();
    if (h_x) {
        BSK_REQUIRE(capacity >= *npts, BSK_ERR_DIMENSION, "node buffer too small (%lld < %lld)", (long long)capacity,
                    (long long)*npts);
        std::copy(sg.x.begin(), sg.x.end(), h_x);
    }
    return BSK_OK;
}

// d_band: (2 (k - 1) + 1) x length(B) column-major band store, data[(k - 1) + i - j, j], fully overwritten.
// Periodic bases: the same store read cyclically, data[(k - 1) + d, j] = A[(j + d) mod N, j] (the SparseMatrixCSC of
// linear.jl:144 has exactly these entries).
bsk_status bsk_galerkin_matrix(const bsk_basis *b, int32_t deriv1, int32_t deriv2, const double *h_qx,
                               const double *h_qw, int32_t nq, double *d_band, void *stream) {
    return bsk_galerkin_matrix_weighted(b, deriv1, deriv2, h_qx, h_qw, nq, nullptr, 0, d_band, stream);
}

// galerkin_matrix(f, B, ...) (linear.jl:60-96,305): the integrand carries the weight f(x); d_fx = f at the nodes
// bsk_galerkin_nodes lists (NULL: f = 1).
bsk_status bsk_galerkin_matrix_weighted(const bsk_basis *b, int32_t deriv1, int32_t deriv2, const double *h_qx,
                                        const double *h_qw, int32_t nq, const double *d_fx, int64_t nfx,
                                        double *d_band, void *stream) {
    BSK_REQUIRE(b && h_qx && h_qw && d_band && nq >= 1, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(deriv1 >= 0 && deriv2 >= 0, BSK_ERR_ARGUMENT, "derivative order must be >= 0");
    BSK_REQUIRE(b->kind != BSK_BASIS_PERIODIC || b->len >= 2 * b->k - 1, BSK_ERR_UNSUPPORTED,
                "periodic Galerkin matrix: the cyclic band store needs length(B) >= 2 k - 1");
    Segments sg;
    bsk_status rc = build_segments(b, h_qx, h_qw, nq, sg);
    if (rc != BSK_OK) return rc;
    BSK_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int k = b->k;
    const int64_t npts = (int64_t)sg.x.size();
    BSK_REQUIRE(d_fx == nullptr || nfx == npts, BSK_ERR_DIMENSION, "expected the weight at %lld quadrature nodes, got %lld",
                (long long)npts, (long long)nfx);
    DevBufs d;
    if ((rc = upload(&d.x, sg.x)) != BSK_OK || (rc = upload(&d.y, sg.y)) != BSK_OK ||
        (rc = upload(&d.il, sg.ileft)) != BSK_OK || (rc = upload(&d.seg, sg.seg_of_n)) != BSK_OK)
        return rc;
    BSK_CUDA(bsk::dev_malloc(&d.idx, std::max<int64_t>(1, npts) * 8));
    BSK_CUDA(bsk::dev_malloc(&d.bs1, std::max<int64_t>(1, npts) * k * 8));
    rc = bsk_evaluate_all(b, d.x, npts, deriv1, BSK_F64, (const int64_t *)d.il, (int64_t *)d.idx, d.bs1, stream);
    if (rc != BSK_OK) return rc;
    const double *bs2 = (const double *)d.bs1;
    if (deriv2 != deriv1) {
        BSK_CUDA(bsk::dev_malloc(&d.bs2, std::max<int64_t>(1, npts) * k * 8));
        rc = bsk_evaluate_all(b, d.x, npts, deriv2, BSK_F64, (const int64_t *)d.il, (int64_t *)d.idx, d.bs2, stream);
        if (rc != BSK_OK) return rc;
        bs2 = (const double *)d.bs2;
    }
    const int off = b->rv.enabled ? b->rv.cl : 0;
    const int64_t nent = b->len * (2 * (k - 1) + 1);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nent + 127) / 128, 148 * 16));
    if (b->kind == BSK_BASIS_PERIODIC)
        k_galerkin_gather_periodic<<<grid, 128, 0, st>>>((const double *)d.bs1, bs2, (const int32_t *)d.seg,
                                                         (const double *)d.y, d_fx, nq, k, b->len, d_band);
    else
        k_galerkin_gather<<<grid, 128, 0, st>>>((const double *)d.bs1, bs2, (const int32_t *)d.seg, (const double *)d.y, d_fx, nq, k,
                                                off, b->nt, b->len, d_band);
    BSK_CUDA(cudaGetLastError());
    BSK_CUDA(cudaStreamSynchronize(st));                     // temporaries are released on return
    return BSK_OK;
}

// galerkin_matrix((B1, B2), (Derivative(m), Derivative(n))) for two bases that share their parent B-spline basis
// (linear.jl:134-199,220-311): N1 x N2, band store data[u + i - j, j] with (l + u + 1) x N2 entries, l = (k - 1) + dl,
// u = (k - 1) - dl, dl = (left constraints of B2) - (left constraints of B1); *h_l / *h_u return them.
bsk_status bsk_galerkin_matrix_pair(const bsk_basis *b1, const bsk_basis *b2, int32_t deriv1, int32_t deriv2,
                                    const double *h_qx, const double *h_qw, int32_t nq, const double *d_fx, int64_t nfx,
                                    double *d_band, int32_t *h_l, int32_t *h_u, void *stream) {
    BSK_REQUIRE(b1 && b2 && h_qx && h_qw && nq >= 1, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(deriv1 >= 0 && deriv2 >= 0, BSK_ERR_ARGUMENT, "derivative order must be >= 0");
    BSK_REQUIRE(b1->kind != BSK_BASIS_PERIODIC && b2->kind != BSK_BASIS_PERIODIC, BSK_ERR_UNSUPPORTED,
                "pairs of periodic bases are not built");
    BSK_REQUIRE(b1->k == b2->k && b1->h_knots == b2->h_knots && b1->device == b2->device, BSK_ERR_ARGUMENT,
                "all bases must share the same parent B-spline basis");           // linear.jl:146-152
    const int k = b1->k;
    const int off1 = b1->rv.enabled ? b1->rv.cl : 0, off2 = b2->rv.enabled ? b2->rv.cl : 0;
    const int cr1 = b1->rv.enabled ? b1->rv.cr : 0, cr2 = b2->rv.enabled ? b2->rv.cr : 0;
    const int dl = off2 - off1;
    BSK_REQUIRE(b1->len == b2->len + dl + (cr2 - cr1), BSK_ERR_DIMENSION, "inconsistent basis lengths");   // linear.jl:192
    const int l = (k - 1) + dl, u = (k - 1) - dl;
    BSK_REQUIRE(l >= 0 && u >= 0, BSK_ERR_UNSUPPORTED, "the bases differ by more constraints than the bandwidth");
    if (h_l) *h_l = l;
    if (h_u) *h_u = u;
    if (!d_band) return BSK_OK;                              // size query
    Segments sg;
    bsk_status rc = build_segments(b1, h_qx, h_qw, nq, sg);
    if (rc != BSK_OK) return rc;
    const int64_t npts = (int64_t)sg.x.size();
    BSK_REQUIRE(d_fx == nullptr || nfx == npts, BSK_ERR_DIMENSION, "expected the weight at %lld quadrature nodes, got %lld",
                (long long)npts, (long long)nfx);
    std::vector<int64_t> il2(sg.ileft);
    for (auto &v : il2) v += off1 - off2;                    // ileft of basis 2 at the same segment
    BSK_CUDA(cudaSetDevice(b1->device));
    cudaStream_t st = (cudaStream_t)stream;
    DevBufs d;
    void *d_il2 = nullptr;
    if ((rc = upload(&d.x, sg.x)) != BSK_OK || (rc = upload(&d.y, sg.y)) != BSK_OK ||
        (rc = upload(&d.il, sg.ileft)) != BSK_OK || (rc = upload(&d.seg, sg.seg_of_n)) != BSK_OK ||
        (rc = upload(&d_il2, il2)) != BSK_OK) {
        if (d_il2) bsk::dev_free(d_il2);
        return rc;
    }
    cudaError_t ce = bsk::dev_malloc(&d.idx, std::max<int64_t>(1, npts) * 8);
    if (ce == cudaSuccess) ce = bsk::dev_malloc(&d.bs1, std::max<int64_t>(1, npts) * k * 8);
    if (ce == cudaSuccess) ce = bsk::dev_malloc(&d.bs2, std::max<int64_t>(1, npts) * k * 8);
    if (ce != cudaSuccess) { bsk::dev_free(d_il2); BSK_CUDA(ce); }
    rc = bsk_evaluate_all(b1, d.x, npts, deriv1, BSK_F64, (const int64_t *)d.il, (int64_t *)d.idx, d.bs1, stream);
    if (rc == BSK_OK)
        rc = bsk_evaluate_all(b2, d.x, npts, deriv2, BSK_F64, (const int64_t *)d_il2, (int64_t *)d.idx, d.bs2, stream);
    if (rc == BSK_OK) {
        const int64_t nent = b2->len * (l + u + 1);
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nent + 127) / 128, 148 * 16));
        k_galerkin_gather_pair<<<grid, 128, 0, st>>>((const double *)d.bs1, (const double *)d.bs2, (const int32_t *)d.seg,
                                                     (const double *)d.y, d_fx, nq, k, off1, off2, b1->nt, b1->len, b2->len,
                                                     l, u, d_band);
        cudaError_t le = cudaGetLastError();
        if (le == cudaSuccess) le = cudaStreamSynchronize(st);
        if (le != cudaSuccess) { set_error("pair Galerkin assembly failed: %s", cudaGetErrorString(le)); rc = BSK_ERR_CUDA; }
    }
    bsk::dev_free(d_il2);
    return rc;
}

// phi[i] = sum over segments and nodes of (alpha w f(x)) b_i(x); d_fx = f at the nodes of bsk_galerkin_nodes.
bsk_status bsk_galerkin_projection(const bsk_basis *b, int32_t deriv, const double *h_qx, const double *h_qw, int32_t nq,
                                   const double *d_fx, int64_t nfx, double *d_phi, void *stream) {
    BSK_REQUIRE(b && h_qx && h_qw && d_fx && d_phi && nq >= 1, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(deriv >= 0, BSK_ERR_ARGUMENT, "derivative order must be >= 0");
    Segments sg;
    bsk_status rc = build_segments(b, h_qx, h_qw, nq, sg);
    if (rc != BSK_OK) return rc;
    const int64_t npts = (int64_t)sg.x.size();
    BSK_REQUIRE(nfx == npts, BSK_ERR_DIMENSION, "expected f at %lld quadrature nodes, got %lld", (long long)npts,
                (long long)nfx);
    BSK_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int k = b->k;
    DevBufs d;
    if ((rc = upload(&d.x, sg.x)) != BSK_OK || (rc = upload(&d.y, sg.y)) != BSK_OK ||
        (rc = upload(&d.il, sg.ileft)) != BSK_OK || (rc = upload(&d.seg, sg.seg_of_n)) != BSK_OK)
        return rc;
    BSK_CUDA(bsk::dev_malloc(&d.idx, std::max<int64_t>(1, npts) * 8));
    BSK_CUDA(bsk::dev_malloc(&d.bs1, std::max<int64_t>(1, npts) * k * 8));
    rc = bsk_evaluate_all(b, d.x, npts, deriv, BSK_F64, (const int64_t *)d.il, (int64_t *)d.idx, d.bs1, stream);
    if (rc != BSK_OK) return rc;
    const int off = b->rv.enabled ? b->rv.cl : 0;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((b->len + 127) / 128, 148 * 16));
    const int64_t nt_seg = b->kind == BSK_BASIS_PERIODIC ? b->len + k : b->nt;      // length of the knot vector looped over
    k_galerkin_project<<<grid, 128, 0, st>>>((const double *)d.bs1, (const int32_t *)d.seg, (const double *)d.y, d_fx, nq, k,
                                             off, nt_seg, b->len, d_phi);
    BSK_CUDA(cudaGetLastError());
    BSK_CUDA(cudaStreamSynchronize(st));
    return BSK_OK;
}

}  // extern "C"
