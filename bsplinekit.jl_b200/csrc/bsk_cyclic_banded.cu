// bsk_cyclic_banded.cu — cyclic banded systems: periodic spline interpolation of any order.
//
// The collocation matrix of a PeriodicBSplineBasis is banded with wrap-around corners.  The
// reference stores it as SparseMatrixCSC and factorises it with UMFPACK for every order except
// k = 4 (src/Collocation/Collocation.jl:141-142, src/SplineInterpolations/SplineInterpolations.jl:109).
// Here:  A = B + U V^T  with B the plain banded part (factored once by the banded LU of
// bsk_solve.cu, no pivoting: B-spline collocation matrices are totally positive) and U V^T the two
// bw x bw corner blocks (U = unit vectors of the first / last bw rows).  Woodbury:
//     x = y - Z (I + V^T Z)^{-1} V^T y,    y = B^{-1} d,   Z = B^{-1} U   (n x 2bw, once)
// V^T y only involves the first and last bw entries of y, so  x = y - K ye  with
// ye = (y_1..y_bw, y_{n-bw+1}..y_n) and K = Z (I + V^T Z)^{-1} P precomputed (n x 2bw).  The columns
// of Z decay away from the two ends, so K is only applied on `zone` rows at each end.
// Per data set: one pass of the multi-RHS banded solve (K6) + a correction of 2 * zone rows.
#include <algorithm>
#include <cmath>

#include "handles.cuh"

using namespace bsk;

struct bsk_cyclic_banded {
    int device;
    int64_t n;
    int bw, ntop, nbot;   // correction rows: 0..ntop-1 and n-nbot..n-1
    int shift;            // rows of the system are rotated by `shift` so that the band is centred
    bsk_banded *B;
    double *d_K;          // (ntop + nbot) x (2 bw), row-major: top rows then bottom rows
};

namespace {

template <int R>
__global__ void __launch_bounds__(128)
k_cyclic_banded_correct(const double *__restrict__ K, double *__restrict__ rhs, int64_t nrhs, int n, int ntop,
                        int nbot) {
    constexpr int BW = R / 2;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nrhs;
         c += (int64_t)gridDim.x * blockDim.x) {
        double *p = rhs + c;
        double ye[R];
#pragma unroll
        for (int m = 0; m < BW; ++m) {
            ye[m] = p[(int64_t)m * nrhs];
            ye[BW + m] = p[(int64_t)(n - BW + m) * nrhs];
        }
        for (int z = 0; z < ntop + nbot; ++z) {
            const int row = z < ntop ? z : n - nbot + (z - ntop);
            const double *kr = K + (size_t)z * R;
            double acc = 0.0;
#pragma unroll
            for (int m = 0; m < R; ++m) acc = fma(__ldg(kr + m), ye[m], acc);
            p[(int64_t)row * nrhs] -= acc;
        }
    }
}

// d'[i] = d[(i + s) mod n] in place, one thread per column (|s| small).  Reads always run ahead of
// writes, so rows can be moved in batches of 8 for memory-level parallelism.
__global__ void __launch_bounds__(128)
k_rotate_rows(double *__restrict__ rhs, int64_t nrhs, int n, int s) {
    constexpr int MAXS = BSK_MAXK, U = 8;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nrhs;
         c += (int64_t)gridDim.x * blockDim.x) {
        double *p = rhs + c;
        double save[MAXS];
        const int a = s > 0 ? s : -s;
        if (s > 0) {
            for (int i = 0; i < a; ++i) save[i] = p[(int64_t)i * nrhs];
            int i = 0;
            for (; i + U <= n - a; i += U) {
                double v[U];
#pragma unroll
                for (int q = 0; q < U; ++q) v[q] = p[(int64_t)(i + q + a) * nrhs];
#pragma unroll
                for (int q = 0; q < U; ++q) p[(int64_t)(i + q) * nrhs] = v[q];
            }
            for (; i < n - a; ++i) p[(int64_t)i * nrhs] = p[(int64_t)(i + a) * nrhs];
            for (int q = 0; q < a; ++q) p[(int64_t)(n - a + q) * nrhs] = save[q];
        } else if (s < 0) {
            for (int i = 0; i < a; ++i) save[i] = p[(int64_t)(n - a + i) * nrhs];
            int i = n - 1;
            for (; i - U + 1 >= a; i -= U) {
                double v[U];
#pragma unroll
                for (int q = 0; q < U; ++q) v[q] = p[(int64_t)(i - q - a) * nrhs];
#pragma unroll
                for (int q = 0; q < U; ++q) p[(int64_t)(i - q) * nrhs] = v[q];
            }
            for (; i >= a; --i) p[(int64_t)i * nrhs] = p[(int64_t)(i - a) * nrhs];
            for (int q = 0; q < a; ++q) p[(int64_t)q * nrhs] = save[q];
        }
    }
}

// dense r x r solve with partial pivoting: X = S^{-1} Rhs (r x r), row-major
bool small_solve(std::vector<double> S, std::vector<double> &X, int r) {
    for (int c = 0; c < r; ++c) {
        int piv = c;
        for (int i = c + 1; i < r; ++i)
            if (std::fabs(S[i * r + c]) > std::fabs(S[piv * r + c])) piv = i;
        if (S[piv * r + c] == 0.0) return false;
        if (piv != c)
            for (int j = 0; j < r; ++j) {
                std::swap(S[c * r + j], S[piv * r + j]);
                std::swap(X[c * r + j], X[piv * r + j]);
            }
        for (int i = c + 1; i < r; ++i) {
            const double f = S[i * r + c] / S[c * r + c];
            if (f == 0.0) continue;
            for (int j = c; j < r; ++j) S[i * r + j] -= f * S[c * r + j];
            for (int j = 0; j < r; ++j) X[i * r + j] -= f * X[c * r + j];
        }
    }
    for (int c = r - 1; c >= 0; --c)
        for (int j = 0; j < r; ++j) {
            double v = X[c * r + j];
            for (int q = c + 1; q < r; ++q) v -= S[c * r + q] * X[q * r + j];
            X[c * r + j] = v / S[c * r + c];
        }
    return true;
}

}  // namespace

extern "C" {

// h_cband: (2 bw + 1) x n column-major "cyclic band store": A(i, j) with cyclic offset
// d = i - j (mod n) in [shift - bw, shift + bw] sits at h_cband[(bw + d - shift) + (j - 1) (2 bw + 1)]
// (1-based i, j) — the band store of Collocation/Collocation.jl:165-198 with indices wrapped.  `shift`
// is the offset of the band centre (odd-order periodic bases: b_j peaks at x_{j+1}, shift = 1); the
// system is solved as (R A) x = R d with R the rotation of the rows by `shift`.
bsk_status bsk_cyclic_banded_create(const double *h_cband, int64_t n, int32_t bw, int32_t shift, int32_t device,
                                    bsk_cyclic_banded **out) {
    BSK_REQUIRE(h_cband && out, BSK_ERR_ARGUMENT, "NULL argument");
    *out = nullptr;
    BSK_REQUIRE(bw >= 1 && bw <= BSK_MAXK - 2, BSK_ERR_UNSUPPORTED, "bandwidths 1..%d are supported", BSK_MAXK - 2);
    BSK_REQUIRE(n >= 4 * (int64_t)bw + 1, BSK_ERR_UNSUPPORTED, "cyclic banded matrix needs n >= 4 bw + 1");
    BSK_REQUIRE(shift > -BSK_MAXK && shift < BSK_MAXK, BSK_ERR_ARGUMENT, "band shift out of range");
    const int nb = 2 * bw + 1, r = 2 * bw;
    // from here on A is the ROTATED matrix: row i of it is row i + shift of the caller's
    auto A = [&](int64_t i, int64_t j) -> double {          // 0-based, any |i - j| (cyclic)
        int64_t d = (i - j) % n;
        if (d > n / 2) d -= n;
        if (d < -(n / 2)) d += n;
        if (d < -bw || d > bw) return 0.0;
        return h_cband[(size_t)(bw + d) + (size_t)j * nb];
    };
    // banded part B (no wrap): data[u + i - j, j]
    std::vector<double> band((size_t)nb * n, 0.0);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = std::max<int64_t>(0, j - bw); i <= std::min<int64_t>(n - 1, j + bw); ++i)
            band[(size_t)(bw + i - j) + (size_t)j * nb] = A(i, j);
    bsk_banded *B = nullptr;
    bsk_status rc = bsk_banded_create_host(band.data(), n, bw, bw, device, &B);
    if (rc != BSK_OK) return rc;
    int64_t info = 0;
    rc = bsk_banded_lu(B, &info);
    if (rc != BSK_OK) { bsk_banded_destroy(B); return rc; }
    // Z = B^{-1} U, U = [e_0 .. e_{bw-1}, e_{n-bw} .. e_{n-1}]  (n x r, interleaved)
    std::vector<double> Z((size_t)n * r, 0.0);
    for (int m = 0; m < bw; ++m) {
        Z[(size_t)m * r + m] = 1.0;
        Z[(size_t)(n - bw + m) * r + bw + m] = 1.0;
    }
    rc = bsk_banded_solve_host(B, Z.data(), r, BSK_SOLVE_TWOPASS);
    if (rc != BSK_OK) { bsk_banded_destroy(B); return rc; }
    // P (r x r): V^T y = P ye, ye = (y_0..y_{bw-1}, y_{n-bw}..y_{n-1});  row m < bw of V^T is row m of
    // the top-right corner (columns n-bw..n-1), row bw+m is row n-bw+m of the bottom-left corner
    std::vector<double> P((size_t)r * r, 0.0), S((size_t)r * r, 0.0);
    for (int m = 0; m < bw; ++m)
        for (int q = 0; q < bw; ++q) {
            P[(size_t)m * r + bw + q] = A(m, n - bw + q);
            P[(size_t)(bw + m) * r + q] = A(n - bw + m, q);
        }
    // S = I + V^T Z = I + P Ze, Ze = rows {0..bw-1, n-bw..n-1} of Z
    auto Ze = [&](int e, int c) { return Z[(size_t)(e < bw ? e : n - 2 * bw + e) * r + c]; };
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < r; ++j) {
            double v = i == j ? 1.0 : 0.0;
            for (int e = 0; e < r; ++e) v += P[(size_t)i * r + e] * Ze(e, j);
            S[(size_t)i * r + j] = v;
        }
    std::vector<double> SiP(P);                               // S^{-1} P
    if (!small_solve(S, SiP, r)) {
        bsk_banded_destroy(B);
        set_error("cyclic banded matrix is singular (capacitance matrix)");
        return BSK_ERR_ZERO_PIVOT;
    }
    // K = Z S^{-1} P, kept on the rows where it matters
    std::vector<double> K((size_t)n * r, 0.0);
    double kmax = 0.0;
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < r; ++j) {
            double v = 0.0;
            for (int e = 0; e < r; ++e) v += Z[(size_t)i * r + e] * SiP[(size_t)e * r + j];
            K[(size_t)i * r + j] = v;
            kmax = std::max(kmax, std::fabs(v));
        }
    int64_t zone = bw;
    for (int64_t i = 0; i < n; ++i) {
        double rowmax = 0.0;
        for (int j = 0; j < r; ++j) rowmax = std::max(rowmax, std::fabs(K[(size_t)i * r + j]));
        if (rowmax > 1e-19 * kmax) zone = std::max<int64_t>(zone, std::min<int64_t>(i + 1, n - i));
    }
    int64_t ntop = zone, nbot = zone;
    if (2 * zone >= n) { ntop = n; nbot = 0; }               // no decay: correct every row
    std::vector<double> Kz((size_t)(ntop + nbot) * r);
    for (int64_t z = 0; z < ntop + nbot; ++z) {
        const int64_t row = z < ntop ? z : n - nbot + (z - ntop);
        for (int j = 0; j < r; ++j) Kz[(size_t)z * r + j] = K[(size_t)row * r + j];
    }
    BSK_CUDA(cudaSetDevice(device));
    bsk_cyclic_banded *h = new bsk_cyclic_banded();
    h->device = device; h->n = n; h->bw = bw; h->ntop = (int)ntop; h->nbot = (int)nbot; h->B = B; h->shift = shift; h->d_K = nullptr;
    cudaError_t e = bsk::dev_malloc((void **)&h->d_K, Kz.size() * 8);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_K, Kz.data(), Kz.size() * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("cyclic banded upload failed: %s", cudaGetErrorString(e));
        bsk_cyclic_banded_destroy(h);
        return BSK_ERR_CUDA;
    }
    *out = h;
    return BSK_OK;
}

// ldiv!(x, F, y) for the cyclic banded matrix, nrhs interleaved right-hand sides, in place.
bsk_status bsk_cyclic_banded_solve(const bsk_cyclic_banded *h, double *d_rhs, int64_t nrhs, void *stream) {
    BSK_REQUIRE(h && (nrhs == 0 || d_rhs), BSK_ERR_ARGUMENT, "NULL argument");
    if (nrhs == 0) return BSK_OK;
    BSK_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int bs = 128;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nrhs + bs - 1) / bs, 148 * 16));
    if (h->shift != 0) {
        k_rotate_rows<<<grid, bs, 0, st>>>(d_rhs, nrhs, (int)h->n, h->shift);
        BSK_CUDA(cudaGetLastError());
    }
    bsk_status rc = bsk_banded_solve(h->B, d_rhs, nrhs, BSK_SOLVE_AUTO, stream);   // B y = R d
    if (rc != BSK_OK) return rc;
    switch (h->bw) {
        case 1: k_cyclic_banded_correct<2><<<grid, bs, 0, st>>>(h->d_K, d_rhs, nrhs, (int)h->n, h->ntop, h->nbot); break;
        case 2: k_cyclic_banded_correct<4><<<grid, bs, 0, st>>>(h->d_K, d_rhs, nrhs, (int)h->n, h->ntop, h->nbot); break;
        case 3: k_cyclic_banded_correct<6><<<grid, bs, 0, st>>>(h->d_K, d_rhs, nrhs, (int)h->n, h->ntop, h->nbot); break;
        case 4: k_cyclic_banded_correct<8><<<grid, bs, 0, st>>>(h->d_K, d_rhs, nrhs, (int)h->n, h->ntop, h->nbot); break;
        case 5: k_cyclic_banded_correct<10><<<grid, bs, 0, st>>>(h->d_K, d_rhs, nrhs, (int)h->n, h->ntop, h->nbot); break;
        case 6: k_cyclic_banded_correct<12><<<grid, bs, 0, st>>>(h->d_K, d_rhs, nrhs, (int)h->n, h->ntop, h->nbot); break;
        default: set_error("unsupported bandwidth %d", h->bw); return BSK_ERR_UNSUPPORTED;
    }
    BSK_CUDA(cudaGetLastError());
    return BSK_OK;
}

bsk_status bsk_cyclic_banded_destroy(bsk_cyclic_banded *h) {
    if (!h) return BSK_OK;
    cudaSetDevice(h->device);
    if (h->d_K) bsk::dev_free(h->d_K);
    if (h->B) bsk_banded_destroy(h->B);
    delete h;
    return BSK_OK;
}

}  // extern "C"
