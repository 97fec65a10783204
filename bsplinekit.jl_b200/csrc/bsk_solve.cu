// bsk_solve.cu — K4 collocation assembly, K5 banded LU (BANFAC), K6 multi-RHS banded
// solve (BANSLV order two-pass, and single-pass windowed), K7 cyclic tridiagonal.
//
// Reference bodies replaced (paths relative to the BSplineKit.jl tree):
//   collocation_matrix!   src/Collocation/Collocation.jl:209-251
//   lu!(::CollocationMatrix)  src/Collocation/matrix.jl:132-201
//   ldiv!(x, F, y)        src/Collocation/matrix.jl:218-271
//   ldiv!(x, ::CyclicTridiagonalMatrix, d)  src/Collocation/cyclic.jl:112-197
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "banfac.cuh"
#include "banslv.cuh"
#include "handles.cuh"

using namespace bsk;

namespace {

constexpr int kThreads = 256;

template <typename KT>
KnotView<KT> make_view_s(const bsk_basis *b) {
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

// ---- K4: collocation assembly ----------------------------------------------------------
// One thread per row (collocation point).  The band store must be zero-filled before.
template <int K>
__global__ void __launch_bounds__(kThreads)
k_collocation(KnotView<double> kv, RecombView rv, double dom_a, double dom_b,
              const double *__restrict__ x, int nx, int deriv, double clip, double *__restrict__ band,
              int ncols, unsigned long long *__restrict__ n_outside) {
    const int l = K - 2 < 0 ? 0 : K - 2, u = l;
    const int nroww = l + u + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x + 1; i <= nx; i += gridDim.x * blockDim.x) {
        const double xi = x[i - 1];
        if (!(xi >= dom_a && xi <= dom_b)) continue;          // isindomain (basis.jl:229-234)
        double bq[K];
        int jlast = eval_all_point<K, 0, double, double>(kv, kv.t, xi, deriv, 0, bq);
        if (rv.enabled) {
            double ph[K];
            jlast = recombine_point<K, double, double>(rv, jlast, bq, ph);
#pragma unroll
            for (int j = 0; j < K; ++j) bq[j] = ph[j];
        }
#pragma unroll
        for (int dj = 1; dj <= K; ++dj) {
            const int j = jlast + 1 - dj;
            const double bj = bq[dj - 1];
            if (fabs(bj) <= clip) continue;                   // Collocation.jl:237
            if ((i > j && i - j > l) || (i < j && j - i > u) || j < 1 || j > ncols) {
                atomicAdd(n_outside, 1ull);                   // reference: @warn (Collocation.jl:239-244)
                continue;
            }
            band[(u + i - j) + (size_t)(j - 1) * nroww] = bj;  // matrix.jl:105-117
        }
    }
}

// Any number of points (least-squares fitting has more points than basis functions): the matrix row by row,
// exactly k slots per row — col[i][dj] = 1-based column (0 = entry clipped or row outside the domain),
// val[i][dj] = D^n b_col(x_i), slots in the reference's order dj = 1..k (column jlast + 1 - dj).  This is the
// SparseMatrixCSC / Matrix target of collocation_matrix! (Collocation.jl:209-251) with the column mapping of
// basis_to_array_index (periodic wrap, BSplines/periodic.jl:245-253).  `dense` != nullptr additionally
// scatters into a zero-filled column-major nx x ncols matrix.
template <int K, int KIND>
__global__ void __launch_bounds__(kThreads)
k_collocation_rows(KnotView<double> kv, RecombView rv, double dom_a, double dom_b, const double *__restrict__ x,
                   int64_t nx, int deriv, double clip, int64_t *__restrict__ col, double *__restrict__ val,
                   double *__restrict__ dense, int ncols) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += (int64_t)gridDim.x * blockDim.x) {
        const double xi = x[i];
        double bq[K];
#pragma unroll
        for (int j = 0; j < K; ++j) bq[j] = 0.0;
        int jlast = 0;
        const bool in = KIND == 1 || (xi >= dom_a && xi <= dom_b);        // isindomain (basis.jl:229-234)
        if (in) {
            jlast = eval_all_point<K, KIND, double, double>(kv, kv.t, xi, deriv, 0, bq);
            if (rv.enabled) {
                double ph[K];
                jlast = recombine_point<K, double, double>(rv, jlast, bq, ph);
#pragma unroll
                for (int j = 0; j < K; ++j) bq[j] = ph[j];
            }
        }
#pragma unroll
        for (int dj = 1; dj <= K; ++dj) {
            int j = jlast + 1 - dj;
            if (KIND == 1) { j = (j - 1) % ncols; if (j < 0) j += ncols; j += 1; }
            const double bj = bq[dj - 1];
            const bool keep = in && !(fabs(bj) <= clip) && j >= 1 && j <= ncols;
            if (col) col[i * K + dj - 1] = keep ? j : 0;
            if (val) val[i * K + dj - 1] = keep ? bj : 0.0;
            if (dense && keep) dense[i + (int64_t)(j - 1) * nx] = bj;
        }
    }
}

// Periodic cubic: CyclicTridiagonalMatrix (cyclic.jl:21-33, setindex! :77-92)
template <int K>
__global__ void __launch_bounds__(kThreads)
k_collocation_cyclic(KnotView<double> kv, const double *__restrict__ x, int nx, int deriv, double clip,
                     double *__restrict__ a, double *__restrict__ b, double *__restrict__ c,
                     unsigned long long *__restrict__ n_outside) {
    const int n = kv.N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x + 1; i <= nx; i += gridDim.x * blockDim.x) {
        double bq[K];
        int jlast = eval_all_point<K, 1, double, double>(kv, kv.t, x[i - 1], deriv, 0, bq);
        const int im = (i == 1) ? n : i - 1, ip = (i == n) ? 1 : i + 1;
#pragma unroll
        for (int dj = 1; dj <= K; ++dj) {
            int j = jlast + 1 - dj;
            while (j < 1) j += n;                             // basis_to_array_index (periodic.jl:245-253)
            while (j > n) j -= n;
            const double bj = bq[dj - 1];
            if (fabs(bj) <= clip) continue;
            if (i == j) b[i - 1] = bj;
            else if (j == im) a[i - 1] = bj;
            else if (j == ip) c[i - 1] = bj;
            else atomicAdd(n_outside, 1ull);                  // reference: DomainError
        }
    }
}

// Repack the factors row-wise for the solve kernels:
//   lrow[r][j-1] = L(r, r-j) = w[mid + j, r - j],  urow[r][j-1] = U(r, r+j) = w[mid - j, r + j]
__global__ void k_repack(const double *__restrict__ w, int bw, int n, double *__restrict__ lrow,
                         double *__restrict__ urow, double *__restrict__ diag, double *__restrict__ dinv) {
    const int nroww = 2 * bw + 1, middle = bw + 1;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x + 1; r <= n; r += gridDim.x * blockDim.x) {
        for (int j = 1; j <= bw; ++j) {
            lrow[(size_t)(r - 1) * bw + j - 1] = (r - j >= 1) ? w[(middle + j - 1) + (size_t)(r - j - 1) * nroww] : 0.0;
            urow[(size_t)(r - 1) * bw + j - 1] = (r + j <= n) ? w[(middle - j - 1) + (size_t)(r + j - 1) * nroww] : 0.0;
        }
        const double d = w[(middle - 1) + (size_t)(r - 1) * nroww];
        diag[r - 1] = d;
        dinv[r - 1] = 1.0 / d;
    }
}

// ---- K6a (BANSLV order, two passes over HBM, one thread per right-hand side): banslv.cuh -------

// ---- K6b (single-pass windowed solve) lives in bsk_solve_win.cu --------------------------------

// ---- K7: cyclic tridiagonal, factor-once Sherman-Morrison -------------------------------------
// A = T + u v^T with T tridiagonal (Collocation/cyclic.jl:121-146).  T is LU-factored once
// (a bsk_banded with l = u = 1) and T y = d goes through the banded multi-RHS solve; the
// rank-one correction x = y - q (y_1 + v_n y_n) / (1 + q_1 + v_n q_n)  (cyclic.jl:152-158) only
// touches the rows where q = T^{-1} u is not negligible, i.e. a few rows at both ends.
__global__ void __launch_bounds__(kThreads)
k_cyclic_correct(const double *__restrict__ q, double *__restrict__ rhs, int64_t nrhs, int n, int zone,
                 double vn, double inv_denom) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nrhs;
         c += (int64_t)gridDim.x * blockDim.x) {
        double *p = rhs + c;
        const double alpha = (p[0] + vn * p[(int64_t)(n - 1) * nrhs]) * inv_denom;
        if (2 * zone >= n) {
            for (int r = 0; r < n; ++r) p[(int64_t)r * nrhs] -= alpha * __ldg(q + r);
        } else {
            for (int r = 0; r < zone; ++r) p[(int64_t)r * nrhs] -= alpha * __ldg(q + r);
            for (int r = n - zone; r < n; ++r) p[(int64_t)r * nrhs] -= alpha * __ldg(q + r);
        }
    }
}

// Decay of the backward recurrence.  A backward sweep restarted at row s with a wrong state
// (errors e_{s+1..s+bw}) propagates  e_r = -(sum_j U(r,r+j) e_{r+j}) / U(r,r)  downwards; the error
// that reaches row s - h is T_{s,h} e with a bw x bw transfer matrix T.  We carry T explicitly
// (signed, exact recurrence in double) and bound |T e| <= max_row sum_col |T| for |e|_inf <= 1.
// Returns the smallest multiple of `step` (<= hmax) such that the bound is <= tol for every
// start row, or 0 when the recurrence does not decay (windowed solve unavailable).
// The same analysis on the device, one thread per start row (they are independent): rows of the
// recurrence are rec[r][j-1] * scale[r] (scale == nullptr: unit diagonal); `reversed` walks the
// matrix bottom-up (forward recurrence of L seen as a backward one).  need = max over start rows of
// the rows walked; fail = some start row did not decay within hmax rows.
template <int BW>
__global__ void __launch_bounds__(128)
k_decay_need(const double *__restrict__ rec, const double *__restrict__ scale, int n, int reversed, double tol,
             int step, int hmax, int phase, int *need, int *fail) {
    const int ncand = (n - 1 + step) / step + 1;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ncand; t += gridDim.x * blockDim.x) {
        // start rows s in [1, n-1] with (s + 1 - phase) % step == 0
        const int s = phase - 1 + t * step;
        if (s < 1 || s > n - 1) continue;
        double win[BW][BW];
#pragma unroll
        for (int j = 0; j < BW; ++j)
#pragma unroll
            for (int c = 0; c < BW; ++c) win[j][c] = j == c ? 1.0 : 0.0;
        int h = 0;
        bool ok = false, bad = false;
        for (int r = s; r >= 0 && h < hmax; --r) {
            const int rr = reversed ? n - 1 - r : r;
            const double sc = scale ? scale[rr] : 1.0;
            double acc[BW];
#pragma unroll
            for (int c = 0; c < BW; ++c) {
                double a = 0.0;
#pragma unroll
                for (int j = 1; j <= BW; ++j) a -= rec[(size_t)rr * BW + j - 1] * win[j - 1][c];
                acc[c] = a * sc;
            }
#pragma unroll
            for (int j = BW - 1; j >= 1; --j)
#pragma unroll
                for (int c = 0; c < BW; ++c) win[j][c] = win[j - 1][c];
#pragma unroll
            for (int c = 0; c < BW; ++c) win[0][c] = acc[c];
            ++h;
            double mx = 0.0;
#pragma unroll
            for (int j = 0; j < BW; ++j) {
                double rs = 0.0;
#pragma unroll
                for (int c = 0; c < BW; ++c) rs += fabs(win[j][c]);
                mx = fmax(mx, rs);
                if (!(rs == rs)) bad = true;
            }
            if (bad) break;
            if (mx <= tol) { ok = true; break; }
        }
        if (bad || (!ok && h >= hmax)) atomicExch(fail, 1);
        else atomicMax(need, h);     // ok: decayed after h rows; !ok: ran off row 0 after s + 1 rows (see decay_halo)
    }
}

// Launch the analysis on the default stream; d_res = {need, fail} (zeroed by the caller).
void decay_halo_launch(const double *d_rec, const double *d_scale, int n, int bw, int reversed, double tol, int step,
                       int hmax, int phase, int *d_res) {
    if (bw == 0 || n <= 2 * step) return;
    const int ncand = (n - 1 + step) / step + 1;
    const int grid = std::max(1, std::min((ncand + 127) / 128, 148 * 8));
    switch (bw) {
        case 1: k_decay_need<1><<<grid, 128>>>(d_rec, d_scale, n, reversed, tol, step, hmax, phase, d_res, d_res + 1); break;
        case 2: k_decay_need<2><<<grid, 128>>>(d_rec, d_scale, n, reversed, tol, step, hmax, phase, d_res, d_res + 1); break;
        case 3: k_decay_need<3><<<grid, 128>>>(d_rec, d_scale, n, reversed, tol, step, hmax, phase, d_res, d_res + 1); break;
        case 4: k_decay_need<4><<<grid, 128>>>(d_rec, d_scale, n, reversed, tol, step, hmax, phase, d_res, d_res + 1); break;
        case 5: k_decay_need<5><<<grid, 128>>>(d_rec, d_scale, n, reversed, tol, step, hmax, phase, d_res, d_res + 1); break;
        case 6: k_decay_need<6><<<grid, 128>>>(d_rec, d_scale, n, reversed, tol, step, hmax, phase, d_res, d_res + 1); break;
        default: break;
    }
}
// ... and turn {need, fail} (copied back to the host) into the halo
int decay_halo_finish(int need, int fail, int n, int bw, int step, int hmax) {
    if (bw == 0) return step;
    if (n <= 2 * step || bw > 6 || fail != 0) return 0;
    int H = ((need + step - 1) / step) * step;
    if (H == 0) H = step;
    return H <= hmax ? H : 0;
}

int decay_halo_device_impl(const double *d_rec, const double *d_scale, int n, int bw, int reversed, double tol, int step,
                           int hmax, int phase) {
    if (bw == 0) return step;
    if (n <= 2 * step) return 0;
    int *d_res = nullptr;
    if (bsk::dev_malloc((void **)&d_res, 8) != cudaSuccess) return 0;
    cudaMemset(d_res, 0, 8);
    decay_halo_launch(d_rec, d_scale, n, bw, reversed, tol, step, hmax, phase, d_res);
    int res[2] = {0, 1};
    cudaError_t e = cudaMemcpy(res, d_res, 8, cudaMemcpyDeviceToHost);
    bsk::dev_free(d_res);
    if (e != cudaSuccess) return 0;
    return decay_halo_finish(res[0], res[1], n, bw, step, hmax);
}

// Host version of the same analysis (kept as the readable statement of the bound; BSK_HALO_HOST=1).
// Restarts only happen at chunk boundaries: start rows s with (s + 1 - phase) % step == 0.
int decay_halo(const std::vector<double> &urow, const std::vector<double> &dinv, int n, int bw,
               double tol, int step, int hmax, int phase = 0) {
    if (bw == 0) return step;
    if (n <= 2 * step) return 0;
    int need = 0;
    std::vector<double> win((size_t)bw * bw), acc(bw);      // win[j*bw + c]: e_{r+1+j} = sum_c win * e0_c
    for (int s = n - 1; s >= 1; --s) {
        if (((s + 1 - phase) % step + step) % step != 0) continue;
        std::fill(win.begin(), win.end(), 0.0);
        for (int j = 0; j < bw; ++j) win[(size_t)j * bw + j] = 1.0;
        int h = 0;
        bool ok = false;
        for (int r = s; r >= 0 && h < hmax; --r) {
            for (int c = 0; c < bw; ++c) {
                double a = 0.0;
                for (int j = 1; j <= bw; ++j) a -= urow[(size_t)r * bw + j - 1] * win[(size_t)(j - 1) * bw + c];
                acc[c] = a * dinv[r];
            }
            for (int j = bw - 1; j >= 1; --j)
                for (int c = 0; c < bw; ++c) win[(size_t)j * bw + c] = win[(size_t)(j - 1) * bw + c];
            for (int c = 0; c < bw; ++c) win[c] = acc[c];
            ++h;
            double mx = 0.0;
            for (int j = 0; j < bw; ++j) {
                double rs = 0.0;
                for (int c = 0; c < bw; ++c) rs += std::fabs(win[(size_t)j * bw + c]);
                mx = std::max(mx, rs);
            }
            if (!(mx == mx)) return 0;
            if (mx <= tol) { ok = true; break; }
        }
        if (!ok && h >= hmax) return 0;          // no decay within hmax rows
        // ok: decayed after h rows.  !ok: ran off row 0 after h = s + 1 rows without decaying — a halo
        // of s + 1 rows makes this restart keep no rows at all, anything shorter would keep polluted ones.
        need = std::max(need, h);
    }
    int H = ((need + step - 1) / step) * step;
    if (H == 0) H = step;
    return H <= hmax ? H : 0;
}

// Copy `w` columns starting at column c0 of an (n x ld) block into a dense (n x W) block, zero-padded
// (dir = 0), or back (dir = 1).
__global__ void __launch_bounds__(256)
k_pack_cols(double *__restrict__ rhs, int64_t ld, int64_t c0, int64_t w, double *__restrict__ tmp, int64_t W, int64_t n,
            int dir) {
    const int64_t total = n * W;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / W, c = e - r * W;
        if (dir == 0) tmp[e] = c < w ? rhs[r * ld + c0 + c] : 0.0;
        else if (c < w) rhs[r * ld + c0 + c] = tmp[e];
    }
}

template <int BW>
bsk_status solve_dispatch(const bsk_banded *h, double *d_rhs, int64_t nrhs, int algo, cudaStream_t st) {
    const int n = (int)h->n;
    // The single-pass kernel also wins for FEW right-hand sides: its row segments run in parallel where
    // the BANSLV-order kernels walk all n rows in one thread (n = 4096: 0.11 ms vs 2.0 ms for 1..2048
    // columns).  BANSLV order is kept for matrices without a validated halo and on request (bit-exact).
    const bool windowed = (algo == BSK_SOLVE_WINDOWED) || (algo == BSK_SOLVE_AUTO && h->halo > 0);
    if (algo == BSK_SOLVE_WINDOWED && h->halo == 0) {
        set_error("windowed solve unavailable: the decay bound of U^{-1} was not met for this matrix");
        return BSK_ERR_UNSUPPORTED;
    }
    int64_t done = 0;                                         // columns handled in place by the windowed kernel
    if (windowed && (((uintptr_t)d_rhs) & 15) == 0 && nrhs % 2 == 0 && nrhs >= 32) {
        const int64_t ngroups = nrhs / 32;
        bsk_status rc = windowed_solve(h, d_rhs, ngroups, nrhs, st);
        if (rc != BSK_OK) return rc;
        done = ngroups * 32;
    }
    const int64_t rest = nrhs - done;
    if (rest > 0 && windowed) {
        // remaining (or unaligned / odd-pitch) columns: through a zero-padded block of 32-column groups
        const int64_t W = (rest + 31) / 32 * 32;
        double *tmp = nullptr;
        ensure_pool(h->device);
        BSK_CUDA(cudaMallocAsync((void **)&tmp, sizeof(double) * (size_t)n * W, st));
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)n * W + 255) / 256, 148 * 16));
        k_pack_cols<<<grid, 256, 0, st>>>(d_rhs, nrhs, done, rest, tmp, W, n, 0);
        bsk_status rc = windowed_solve(h, tmp, W / 32, W, st);
        if (rc == BSK_OK) k_pack_cols<<<grid, 256, 0, st>>>(d_rhs, nrhs, done, rest, tmp, W, n, 1);
        cudaError_t e = cudaGetLastError();
        cudaFreeAsync(tmp, st);
        if (rc != BSK_OK) return rc;
        BSK_CUDA(e);
    } else if (rest > 0) {                                    // BANSLV-order two-pass
        const int bs = rest >= 148 * 2 * 256 ? 128 : 64;
        int grid = (int)std::max<int64_t>(1, (rest + bs - 1) / bs);
        if (BW > 0) k_banslv_sweep<BW, 16, +1, double><<<grid, bs, 0, st>>>(h->d_lrow, h->d_diag, d_rhs + done, rest, nrhs, n);
        BSK_CUDA(cudaGetLastError());
        k_banslv_sweep<BW, 16, -1, double><<<grid, bs, 0, st>>>(h->d_urow, h->d_diag, d_rhs + done, rest, nrhs, n);
        BSK_CUDA(cudaGetLastError());
    }
    return BSK_OK;
}

}  // namespace

namespace bsk {
int decay_halo_device(const double *d_rec, const double *d_scale, int n, int bw, int reversed, double tol, int step,
                      int hmax, int phase) {
    return decay_halo_device_impl(d_rec, d_scale, n, bw, reversed, tol, step, hmax, phase);
}
}  // namespace bsk

extern "C" {

bsk_status bsk_collocation_matrix(const bsk_basis *b, const double *d_x, int64_t nx, int32_t deriv,
                                  double clip_threshold, double *d_band, int64_t *h_n_outside, void *stream) {
    BSK_REQUIRE(b && d_x && d_band, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(b->dtype == BSK_F64, BSK_ERR_UNSUPPORTED, "collocation assembly needs a Float64 basis");
    BSK_REQUIRE(b->kind != BSK_BASIS_PERIODIC, BSK_ERR_UNSUPPORTED,
                "periodic bases: use bsk_collocation_cyclic (k = 4)");
    BSK_REQUIRE(nx == b->len, BSK_ERR_ARGUMENT, "wrong dimensions of collocation matrix");   // Collocation.jl:214-216
    BSK_REQUIRE(deriv >= 0, BSK_ERR_ARGUMENT, "derivative order must be >= 0");
    BSK_REQUIRE(b->k >= 2, BSK_ERR_UNSUPPORTED, "order must be >= 2");
    BSK_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int nroww = 2 * (b->k - 2) + 1;
    BSK_CUDA(cudaMemsetAsync(d_band, 0, sizeof(double) * (size_t)nroww * (size_t)b->len, st));
    unsigned long long *d_cnt = nullptr;
    BSK_CUDA(bsk::dev_malloc((void **)&d_cnt, 8));
    BSK_CUDA(cudaMemsetAsync(d_cnt, 0, 8, st));
    KnotView<double> kv = make_view_s<double>(b);
    const double dom_a = b->h_knots[b->k - 1], dom_b = b->h_knots[b->N];   // boundaries (basis.jl:217-222)
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nx + kThreads - 1) / kThreads, 148 * 8));
    BSK_DISPATCH_K(b->k, (k_collocation<K><<<grid, kThreads, 0, st>>>(kv, b->rv, dom_a, dom_b, d_x, (int)nx, deriv,
                                                                      clip_threshold, d_band, (int)b->len, d_cnt)))
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && h_n_outside) {
        unsigned long long cnt = 0;
        e = cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        *h_n_outside = (int64_t)cnt;
    }
    cudaFreeAsync(d_cnt, st);
    if (e != cudaSuccess) { set_error("collocation kernel failed: %s", cudaGetErrorString(e)); return BSK_ERR_CUDA; }
    return BSK_OK;
}

bsk_status bsk_collocation_rows(const bsk_basis *b, const double *d_x, int64_t nx, int32_t deriv, double clip_threshold,
                                int64_t *d_col, double *d_val, double *d_dense, void *stream) {
    BSK_REQUIRE(b && (nx == 0 || d_x) && (d_col || d_val || d_dense), BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(b->dtype == BSK_F64, BSK_ERR_UNSUPPORTED, "collocation assembly needs a Float64 basis");
    BSK_REQUIRE(deriv >= 0, BSK_ERR_ARGUMENT, "derivative order must be >= 0");
    BSK_REQUIRE(nx >= 0, BSK_ERR_DIMENSION, "negative number of points");
    if (nx == 0) return BSK_OK;
    BSK_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (d_dense) BSK_CUDA(cudaMemsetAsync(d_dense, 0, sizeof(double) * (size_t)nx * (size_t)b->len, st));
    KnotView<double> kv = make_view_s<double>(b);
    const bool periodic = b->kind == BSK_BASIS_PERIODIC;
    const double dom_a = periodic ? b->h_knots.front() : b->h_knots[b->k - 1];
    const double dom_b = periodic ? b->h_knots.back() : b->h_knots[b->N];
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nx + kThreads - 1) / kThreads, 148 * 8));
    if (periodic) {
        BSK_DISPATCH_K(b->k, (k_collocation_rows<K, 1><<<grid, kThreads, 0, st>>>(kv, b->rv, dom_a, dom_b, d_x, nx, deriv, clip_threshold,
                                                                                 d_col, d_val, d_dense, (int)b->len)))
    } else {
        BSK_DISPATCH_K(b->k, (k_collocation_rows<K, 0><<<grid, kThreads, 0, st>>>(kv, b->rv, dom_a, dom_b, d_x, nx, deriv, clip_threshold,
                                                                                 d_col, d_val, d_dense, (int)b->len)))
    }
    BSK_CUDA(cudaGetLastError());
    return BSK_OK;
}

bsk_status bsk_collocation_cyclic(const bsk_basis *b, const double *d_x, int64_t nx, int32_t deriv,
                                  double clip_threshold, double *d_a, double *d_b, double *d_c,
                                  int64_t *h_n_outside, void *stream) {
    BSK_REQUIRE(b && d_x && d_a && d_b && d_c, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(b->dtype == BSK_F64, BSK_ERR_UNSUPPORTED, "collocation assembly needs a Float64 basis");
    BSK_REQUIRE(b->kind == BSK_BASIS_PERIODIC, BSK_ERR_ARGUMENT, "basis is not periodic");
    BSK_REQUIRE(nx == b->len, BSK_ERR_ARGUMENT, "wrong dimensions of collocation matrix");
    BSK_REQUIRE(b->N >= 3, BSK_ERR_ARGUMENT, "cyclic tridiagonal matrix needs n >= 3");
    BSK_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    BSK_CUDA(cudaMemsetAsync(d_a, 0, 8 * (size_t)nx, st));
    BSK_CUDA(cudaMemsetAsync(d_b, 0, 8 * (size_t)nx, st));
    BSK_CUDA(cudaMemsetAsync(d_c, 0, 8 * (size_t)nx, st));
    unsigned long long *d_cnt = nullptr;
    BSK_CUDA(bsk::dev_malloc((void **)&d_cnt, 8));
    BSK_CUDA(cudaMemsetAsync(d_cnt, 0, 8, st));
    KnotView<double> kv = make_view_s<double>(b);
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nx + kThreads - 1) / kThreads, 148 * 8));
    BSK_DISPATCH_K(b->k, (k_collocation_cyclic<K><<<grid, kThreads, 0, st>>>(kv, d_x, (int)nx, deriv, clip_threshold,
                                                                             d_a, d_b, d_c, d_cnt)))
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && h_n_outside) {
        unsigned long long cnt = 0;
        e = cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        *h_n_outside = (int64_t)cnt;
    }
    cudaFreeAsync(d_cnt, st);
    if (e != cudaSuccess) { set_error("collocation kernel failed: %s", cudaGetErrorString(e)); return BSK_ERR_CUDA; }
    return BSK_OK;
}

// Partitioned factorisation (banfac.cuh): used from this many rows on; partition length and warm-up columns
constexpr int kPartMinRows = 1024;
constexpr int kPartWarmup = 128;
inline int part_length(int64_t n) { return (int)std::max<int64_t>(256, (n + 148 * 2 - 1) / (148 * 2)); }

static bsk_status banded_alloc(int64_t n, int32_t l, int32_t u, int32_t device, bsk_banded **out) {
    BSK_REQUIRE(out, BSK_ERR_ARGUMENT, "out is NULL");
    *out = nullptr;
    BSK_REQUIRE(n >= 1, BSK_ERR_ARGUMENT, "matrix is empty");                       // matrix.jl:145
    BSK_REQUIRE(l == u, BSK_ERR_ARGUMENT, "expected same number of lower and upper bands");   // matrix.jl:58-60
    BSK_REQUIRE(l >= 0 && l <= BSK_MAXK - 2, BSK_ERR_UNSUPPORTED, "bandwidths 0..%d are supported", BSK_MAXK - 2);
    BSK_REQUIRE(n < (1ll << 30), BSK_ERR_UNSUPPORTED, "matrix too large");
    BSK_CUDA(cudaSetDevice(device));
    bsk_banded *h = new bsk_banded();
    h->device = device; h->n = n; h->l = l; h->u = u; h->factored = false; h->halo = 0; h->halo_f = 0;
    h->d_mrec = nullptr; h->mrec_bw = -1; h->mrec_npad = 0;
    h->d_w = h->d_lrow = h->d_urow = h->d_dinv = h->d_diag = h->d_lrec = h->d_urec = nullptr;
    h->arena = nullptr; h->d_w2 = h->d_ver = nullptr; h->d_status = nullptr;
    h->n_pad = (int)((n + 7) / 8 * 8);
    // ONE device allocation for everything the handle will ever hold (every allocation synchronises)
    const bool part = l >= 1 && n >= kPartMinRows;
    const size_t nw = (size_t)(l + u + 1) * n, nb = (size_t)std::max(1, l) * n;
    const size_t nparts = part ? (size_t)bsk::banfac_part_count((int)n, part_length(n)) : 0;
    const size_t sizes[10] = {nw, part ? nw : 0, nb, nb, (size_t)n, (size_t)n, (size_t)h->n_pad * bsk::win_lw(l),
                              (size_t)h->n_pad * bsk::win_uw(l), nparts * (l + 1) * (2 * l + 1), 16};
    size_t offs[10], total = 0;
    for (int i = 0; i < 10; ++i) { offs[i] = total; total += (sizes[i] * 8 + 255) / 256 * 256; }
    cudaError_t e = bsk::dev_malloc(&h->arena, total);
    if (e != cudaSuccess) {
        set_error("cudaMalloc failed: %s", cudaGetErrorString(e));
        delete h;
        return BSK_ERR_ALLOC;
    }
    char *base = (char *)h->arena;
    h->d_w = (double *)(base + offs[0]);
    h->d_w2 = part ? (double *)(base + offs[1]) : nullptr;
    h->d_lrow = (double *)(base + offs[2]);
    h->d_urow = (double *)(base + offs[3]);
    h->d_dinv = (double *)(base + offs[4]);
    h->d_diag = (double *)(base + offs[5]);
    h->d_lrec = (double *)(base + offs[6]);
    h->d_urec = (double *)(base + offs[7]);
    h->d_ver = part ? (double *)(base + offs[8]) : nullptr;
    h->d_status = (int *)(base + offs[9]);
    *out = h;
    return BSK_OK;
}

bsk_status bsk_banded_create(const double *d_band, int64_t n, int32_t l, int32_t u, int32_t device,
                             bsk_banded **out) {
    BSK_REQUIRE(d_band, BSK_ERR_ARGUMENT, "NULL band data");
    bsk_status rc = banded_alloc(n, l, u, device, out);
    if (rc != BSK_OK) return rc;
    const cudaError_t ce = cudaMemcpy((*out)->d_w, d_band, 8 * (size_t)(l + u + 1) * n, cudaMemcpyDeviceToDevice);
    if (ce != cudaSuccess) {
        bsk_banded_destroy(*out);
        *out = nullptr;
        BSK_CUDA(ce);
    }
    return BSK_OK;
}

bsk_status bsk_banded_create_host(const double *h_band, int64_t n, int32_t l, int32_t u, int32_t device,
                                  bsk_banded **out) {
    BSK_REQUIRE(h_band, BSK_ERR_ARGUMENT, "NULL band data");
    bsk_status rc = banded_alloc(n, l, u, device, out);
    if (rc != BSK_OK) return rc;
    const cudaError_t ce = cudaMemcpy((*out)->d_w, h_band, 8 * (size_t)(l + u + 1) * n, cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
        bsk_banded_destroy(*out);
        *out = nullptr;
        BSK_CUDA(ce);
    }
    return BSK_OK;
}

// lu!(C).  mode BSK_LU_EXACT: BANFAC's own statements, serial in the column index (bit-identical factors).
// mode BSK_LU_FAST: from kPartMinRows rows on, the partitioned factorisation of banfac.cuh (factors within
// ~1e-16 relative of BANFAC's, certified at every partition boundary); anything it cannot certify — and every
// zero pivot — goes through the serial kernel, which reports the pivot row.  Everything after the elimination
// (row repack, decay analysis for the single-pass solve, solve records) is enqueued behind it and ONE small
// device-to-host copy fetches all the results.
static bsk_status banded_lu_impl(bsk_banded *h, int64_t *info_pivot, bool fast) {
    BSK_REQUIRE(h, BSK_ERR_ARGUMENT, "NULL handle");
    BSK_REQUIRE(!h->factored, BSK_ERR_ARGUMENT, "matrix is already factorised");
    BSK_CUDA(cudaSetDevice(h->device));
    const bool timing = getenv("BSK_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    const int n = (int)h->n, bw = h->l;
    // status words: [0] partition failure, [1..2] backward halo {need, fail}, [3..4] forward halo, [8..9] BANFAC info
    int st[16];
    bool used_part = fast && h->d_w2 != nullptr && getenv("BSK_LU_SERIAL") == nullptr && getenv("BSK_LU_NO_PART") == nullptr;
    const bool halo_host = getenv("BSK_HALO_HOST") != nullptr;
    for (int attempt = 0; attempt < 2; ++attempt) {
        BSK_CUDA(cudaMemsetAsync(h->d_status, 0, 64, 0));
        const double *wf = h->d_w;
        cudaError_t e;
        if (used_part) {
            e = bsk::launch_banfac_part<double>(h->d_w, h->d_w2, h->d_ver, bw, n, part_length(n), kPartWarmup, 1e-14,
                                                h->d_status, 0);
            wf = h->d_w2;
        } else {
            e = bsk::launch_banfac<double>(h->d_w, h->l, h->u, n, (long long *)(h->d_status + 8), getenv("BSK_LU_SERIAL") != nullptr);
        }
        if (e != cudaSuccess) { set_error("banded LU failed: %s", cudaGetErrorString(e)); return BSK_ERR_CUDA; }
        int grid = std::max(1, std::min((n + kThreads - 1) / kThreads, 148 * 4));
        k_repack<<<grid, kThreads>>>(wf, bw, n, h->d_lrow, h->d_urow, h->d_diag, h->d_dinv);
        BSK_CUDA(cudaGetLastError());
        if (!halo_host) {
            // decay bounds for the windowed solve: backward recurrence of U, and the forward recurrence of the
            // unit-diagonal L walked bottom-up (forward sweeps start at rows f0 = 0 mod 8)
            decay_halo_launch(h->d_urow, h->d_dinv, n, bw, 0, 1e-17, 8, 448, 0, h->d_status + 1);
            decay_halo_launch(h->d_lrow, nullptr, n, bw, 1, 1e-17, 8, 448, n % 8, h->d_status + 3);
            BSK_CUDA(cudaGetLastError());
        }
        bsk_status rc = windowed_prepare(h, wf);
        if (rc != BSK_OK) return rc;
        BSK_CUDA(cudaMemcpy(st, h->d_status, 64, cudaMemcpyDeviceToHost));
        if (used_part && st[0] != 0) { used_part = false; continue; }   // not certified: serial kernel on the intact original
        break;
    }
    const double t_lu = now();
    long long info = 0;
    memcpy(&info, &st[8], 8);
    if (used_part) info = 0;
    if (info_pivot) *info_pivot = info;
    if (info != 0) {
        set_error("ZeroPivotException(%lld)", info);
        return BSK_ERR_ZERO_PIVOT;
    }
    if (used_part) std::swap(h->d_w, h->d_w2);                 // the factors live in the second buffer
    h->halo_f = 0;
    if (!halo_host) {
        h->halo = decay_halo_finish(st[1], st[2], n, bw, 8, 448);
        if (h->halo > 0) h->halo_f = decay_halo_finish(st[3], st[4], n, bw, 8, 448);
    } else {
        std::vector<double> urow((size_t)std::max(1, bw) * n), dinv(n);
        BSK_CUDA(cudaMemcpy(urow.data(), h->d_urow, 8 * (size_t)std::max(1, bw) * n, cudaMemcpyDeviceToHost));
        BSK_CUDA(cudaMemcpy(dinv.data(), h->d_dinv, 8 * (size_t)n, cudaMemcpyDeviceToHost));
        h->halo = decay_halo(urow, dinv, n, bw, 1e-17, 8, 448);
        if (h->halo > 0 && bw > 0) {
            // same bound for the forward recurrence (unit-diagonal L), by reversing the row order
            std::vector<double> lrow((size_t)bw * n), lrev((size_t)bw * n), ones(n, 1.0);
            BSK_CUDA(cudaMemcpy(lrow.data(), h->d_lrow, 8 * (size_t)bw * n, cudaMemcpyDeviceToHost));
            for (int r = 0; r < n; ++r)
                for (int j = 0; j < bw; ++j) lrev[(size_t)(n - 1 - r) * bw + j] = lrow[(size_t)r * bw + j];
            h->halo_f = decay_halo(lrev, ones, n, bw, 1e-17, 8, 448, n % 8);
        }
    }
    if (bw == 0) h->halo_f = 8;
    if (timing)
        fprintf(stderr, "bsk_banded_lu n=%d bw=%d %s: elimination + repack + decay analysis + records %.3f ms, halo %d / %d\n",
                n, bw, used_part ? "partitioned" : "serial", t_lu - t_start, h->halo, h->halo_f);
    h->factored = true;
    return BSK_OK;
}

bsk_status bsk_banded_lu(bsk_banded *h, int64_t *info_pivot) { return banded_lu_impl(h, info_pivot, false); }

bsk_status bsk_banded_lu_mode(bsk_banded *h, int32_t mode, int64_t *info_pivot) {
    BSK_REQUIRE(mode == BSK_LU_EXACT || mode == BSK_LU_FAST, BSK_ERR_ARGUMENT, "unknown LU mode %d", mode);
    return banded_lu_impl(h, info_pivot, mode == BSK_LU_FAST);
}

bsk_status bsk_banded_data(const bsk_banded *h, double *h_out) {
    BSK_REQUIRE(h && h_out, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_CUDA(cudaSetDevice(h->device));
    BSK_CUDA(cudaMemcpy(h_out, h->d_w, 8 * (size_t)(h->l + h->u + 1) * h->n, cudaMemcpyDeviceToHost));
    return BSK_OK;
}

bsk_status bsk_banded_halo(const bsk_banded *h, int32_t *halo) {
    BSK_REQUIRE(h && halo, BSK_ERR_ARGUMENT, "NULL argument");
    *halo = h->halo;
    return BSK_OK;
}

bsk_status bsk_banded_solve(const bsk_banded *h, double *d_rhs, int64_t nrhs, int32_t algo, void *stream) {
    BSK_REQUIRE(h && (nrhs == 0 || d_rhs), BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(h->factored, BSK_ERR_ARGUMENT, "matrix is not factorised: call bsk_banded_lu first");
    BSK_REQUIRE(algo >= 0 && algo <= 2, BSK_ERR_ARGUMENT, "unknown algo %d", algo);
    BSK_REQUIRE(nrhs >= 0, BSK_ERR_DIMENSION, "negative number of right-hand sides");
    if (nrhs == 0) return BSK_OK;
    BSK_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    switch (h->l) {
        case 0: return solve_dispatch<0>(h, d_rhs, nrhs, algo, st);
        case 1: return solve_dispatch<1>(h, d_rhs, nrhs, algo, st);
        case 2: return solve_dispatch<2>(h, d_rhs, nrhs, algo, st);
        case 3: return solve_dispatch<3>(h, d_rhs, nrhs, algo, st);
        case 4: return solve_dispatch<4>(h, d_rhs, nrhs, algo, st);
        case 5: return solve_dispatch<5>(h, d_rhs, nrhs, algo, st);
        case 6: return solve_dispatch<6>(h, d_rhs, nrhs, algo, st);
    }
    set_error("unsupported bandwidth %d", h->l);
    return BSK_ERR_UNSUPPORTED;
}

// du = F \ (Lm * u) in one pass (the mul! + ldiv! pair of a collocation time step,
// examples/heat.jl:314-328).  Falls back to bsk_banded_matvec + bsk_banded_solve when the fused
// kernel does not apply (bandwidth > 4, odd shapes, no validated halo).
bsk_status bsk_banded_matvec_solve(const bsk_banded *lu, const bsk_banded *Lm, const double *d_u, double *d_out,
                                   int64_t nrhs, void *stream) {
    BSK_REQUIRE(lu && Lm && (nrhs == 0 || (d_u && d_out)), BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(lu->factored, BSK_ERR_ARGUMENT, "matrix is not factorised: call bsk_banded_lu first");
    BSK_REQUIRE(!Lm->factored, BSK_ERR_ARGUMENT, "the matrix to apply holds LU factors (lu! works in place)");
    BSK_REQUIRE(lu->n == Lm->n, BSK_ERR_DIMENSION, "matrices have different sizes (%lld, %lld)", (long long)lu->n,
                (long long)Lm->n);
    BSK_REQUIRE(lu->device == Lm->device, BSK_ERR_ARGUMENT, "matrices live on different devices");
    BSK_REQUIRE(d_u != d_out, BSK_ERR_ARGUMENT, "u and du must not alias");
    if (nrhs == 0) return BSK_OK;
    BSK_CUDA(cudaSetDevice(lu->device));
    cudaStream_t st = (cudaStream_t)stream;
    const bool shape_ok = lu->halo > 0 && lu->l <= 4 && Lm->l <= lu->l && nrhs % 32 == 0 &&
                          (((uintptr_t)d_u) & 127) == 0 && (((uintptr_t)d_out) & 15) == 0 &&
                          getenv("BSK_NO_FUSED") == nullptr;
    if (shape_ok) {
        bsk_banded *L = const_cast<bsk_banded *>(Lm);
        static std::mutex mu;
        {
            std::lock_guard<std::mutex> lock(mu);
            if (!L->d_mrec || L->mrec_bw != lu->l || L->mrec_npad != lu->n_pad) {
                if (L->d_mrec) { bsk::dev_free(L->d_mrec); L->d_mrec = nullptr; }
                bsk_status rc = windowed_matrec(L, lu->l, lu->n_pad, &L->d_mrec);
                if (rc != BSK_OK) return rc;
                L->mrec_bw = lu->l;
                L->mrec_npad = lu->n_pad;
            }
        }
        bsk_status rc = windowed_solve_fused(lu, L->d_mrec, d_u, d_out, nrhs / 32, nrhs, st);
        if (rc != BSK_ERR_UNSUPPORTED) return rc;
    }
    bsk_status rc = bsk_banded_matvec(Lm, d_u, d_out, nrhs, stream);
    if (rc != BSK_OK) return rc;
    return bsk_banded_solve(lu, d_out, nrhs, BSK_SOLVE_AUTO, stream);
}

bsk_status bsk_banded_destroy(bsk_banded *h) {
    if (!h) return BSK_OK;
    cudaSetDevice(h->device);
    if (h->arena) bsk::dev_free(h->arena);                    // every array of the handle lives in it
    if (h->d_mrec) bsk::dev_free(h->d_mrec);
    delete h;
    return BSK_OK;
}

// ---- cyclic -----------------------------------------------------------------------------------
static bsk_status cyclic_build(const std::vector<double> &a, const std::vector<double> &b_in,
                               const std::vector<double> &c, int32_t device, bsk_cyclic **out) {
    const int64_t n = (int64_t)a.size();
    BSK_REQUIRE(n >= 3, BSK_ERR_ARGUMENT, "cyclic tridiagonal matrix needs n >= 3");
    // Collocation/cyclic.jl:121-146
    std::vector<double> b(b_in);
    const double a1 = a[0], cn = c[n - 1];
    const double ac = a1 * cn;
    const double gamma = std::sqrt(ac);
    BSK_REQUIRE(gamma == gamma && gamma != 0.0, BSK_ERR_ARGUMENT,
                "cyclic solve needs a[1]*c[n] > 0 (got %g)", ac);
    const double vn = a1 / gamma;
    b[0] -= gamma;
    b[n - 1] -= ac / gamma;
    // band storage of the (non-cyclic) tridiagonal T: data[u + 1 + i - j, j], l = u = 1
    std::vector<double> band((size_t)3 * n, 0.0);
    for (int64_t j = 0; j < n; ++j) {
        if (j >= 1) band[0 + 3 * j] = c[j - 1];      // T(j-1, j)
        band[1 + 3 * j] = b[j];                      // T(j, j)
        if (j + 1 < n) band[2 + 3 * j] = a[j + 1];   // T(j+1, j)
    }
    bsk_banded *tri = nullptr;
    bsk_status rc = bsk_banded_create_host(band.data(), n, 1, 1, device, &tri);
    if (rc != BSK_OK) return rc;
    int64_t info = 0;
    rc = bsk_banded_lu(tri, &info);
    if (rc != BSK_OK) { bsk_banded_destroy(tri); return rc; }
    // q = T^{-1} u, u = gamma e_1 + c_n e_n (cyclic.jl:131-135), through the same factorisation
    std::vector<double> q(n, 0.0);
    q[0] = gamma;
    q[n - 1] = cn;
    rc = bsk_banded_solve_host(tri, q.data(), 1, BSK_SOLVE_TWOPASS);   // exact decay of q: the correction zone is read off its tail
    if (rc != BSK_OK) { bsk_banded_destroy(tri); return rc; }
    const double denom = 1.0 + (q[0] + vn * q[n - 1]);
    double qmax = 0.0;
    for (int64_t i = 0; i < n; ++i) qmax = std::max(qmax, std::fabs(q[i]));
    int zone = 0;          // rows from each end where q is not negligible
    for (int64_t i = 0; i < n; ++i)
        if (std::fabs(q[i]) > 1e-19 * qmax) zone = std::max<int>(zone, (int)std::min<int64_t>(i + 1, n - i));
    BSK_CUDA(cudaSetDevice(device));
    bsk_cyclic *h = new bsk_cyclic();
    h->device = device; h->n = n; h->vn = vn; h->denom = denom; h->tri = tri; h->zone = zone; h->d_q = nullptr;
    cudaError_t e = bsk::dev_malloc((void **)&h->d_q, 8 * (size_t)n);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_q, q.data(), 8 * (size_t)n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("cyclic upload failed: %s", cudaGetErrorString(e));
        bsk_cyclic_destroy(h);
        return BSK_ERR_CUDA;
    }
    *out = h;
    return BSK_OK;
}

bsk_status bsk_cyclic_create(const double *h_a, const double *h_b, const double *h_c, int64_t n, int32_t device,
                             bsk_cyclic **out) {
    BSK_REQUIRE(h_a && h_b && h_c && out, BSK_ERR_ARGUMENT, "NULL argument");
    *out = nullptr;
    BSK_REQUIRE(n >= 3, BSK_ERR_ARGUMENT, "cyclic tridiagonal matrix needs n >= 3");
    std::vector<double> a(h_a, h_a + n), b(h_b, h_b + n), c(h_c, h_c + n);
    return cyclic_build(a, b, c, device, out);
}

bsk_status bsk_cyclic_create_device(const double *d_a, const double *d_b, const double *d_c, int64_t n,
                                    int32_t device, bsk_cyclic **out) {
    BSK_REQUIRE(d_a && d_b && d_c && out, BSK_ERR_ARGUMENT, "NULL argument");
    *out = nullptr;
    BSK_REQUIRE(n >= 3, BSK_ERR_ARGUMENT, "cyclic tridiagonal matrix needs n >= 3");
    BSK_CUDA(cudaSetDevice(device));
    std::vector<double> a(n), b(n), c(n);
    BSK_CUDA(cudaMemcpy(a.data(), d_a, 8 * (size_t)n, cudaMemcpyDeviceToHost));
    BSK_CUDA(cudaMemcpy(b.data(), d_b, 8 * (size_t)n, cudaMemcpyDeviceToHost));
    BSK_CUDA(cudaMemcpy(c.data(), d_c, 8 * (size_t)n, cudaMemcpyDeviceToHost));
    return cyclic_build(a, b, c, device, out);
}

bsk_status bsk_cyclic_solve(const bsk_cyclic *h, double *d_rhs, int64_t nrhs, void *stream) {
    BSK_REQUIRE(h && (nrhs == 0 || d_rhs), BSK_ERR_ARGUMENT, "NULL argument");
    if (nrhs == 0) return BSK_OK;
    BSK_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    bsk_status rc = bsk_banded_solve(h->tri, d_rhs, nrhs, BSK_SOLVE_AUTO, stream);   // T y = d
    if (rc != BSK_OK) return rc;
    const int bs = nrhs >= 148 * 2 * 256 ? 256 : 64;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nrhs + bs - 1) / bs, 148 * 16));
    k_cyclic_correct<<<grid, bs, 0, st>>>(h->d_q, d_rhs, nrhs, (int)h->n, h->zone, h->vn, 1.0 / h->denom);
    BSK_CUDA(cudaGetLastError());
    return BSK_OK;
}

bsk_status bsk_cyclic_destroy(bsk_cyclic *h) {
    if (!h) return BSK_OK;
    cudaSetDevice(h->device);
    if (h->d_q) bsk::dev_free(h->d_q);
    if (h->tri) bsk_banded_destroy(h->tri);
    delete h;
    return BSK_OK;
}

}  // extern "C"

// ---- batched banded mat-vec: Y = C * X for interleaved right-hand sides ---------------------------
// Replaces mul!(y, C::CollocationMatrix, x) (src/Collocation/matrix.jl:91-94, which goes through
// BandedMatrices' gbmv) for many vectors at once; used e.g. for residuals ||C c - y|| and for the
// "RHS per time step" pattern  du = A \ (L * u)  (examples/heat.jl:314-328).
namespace {

template <int BW, int U>
__global__ void __launch_bounds__(kThreads)
k_banded_matvec(const double *__restrict__ w, const double *__restrict__ x, double *__restrict__ y,
                int64_t nrhs, int n) {
    constexpr int NB = 2 * BW + 1;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nrhs;
         c += (int64_t)gridDim.x * blockDim.x) {
        const double *px = x + c;
        double *py = y + c;
        double win[NB];                      // win[d + BW] = x[r + d]
#pragma unroll
        for (int d = 0; d < NB; ++d) win[d] = 0.0;
#pragma unroll
        for (int d = 0; d < BW; ++d) win[BW + 1 + d] = (d < n) ? px[(int64_t)d * nrhs] : 0.0;   // x[0..BW-1] -> slots for r = -1
        for (int r0 = 0; r0 < n; r0 += U) {
            double nx[U];
#pragma unroll
            for (int q = 0; q < U; ++q) nx[q] = (r0 + q + BW < n) ? px[(int64_t)(r0 + q + BW) * nrhs] : 0.0;
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int r = r0 + q;
                if (r < n) {
#pragma unroll
                    for (int d = 0; d < NB - 1; ++d) win[d] = win[d + 1];
                    win[NB - 1] = nx[q];
                    double acc = 0.0;
#pragma unroll
                    for (int d = -BW; d <= BW; ++d) {
                        const int col = r + d;
                        const double a = (col >= 0 && col < n) ? __ldg(w + (BW - d) + (size_t)col * NB) : 0.0;
                        acc = fma(a, win[d + BW], acc);
                    }
                    py[(int64_t)r * nrhs] = acc;
                }
            }
        }
    }
}

template <int BW>
bsk_status matvec_t(const bsk_banded *h, const double *d_x, double *d_y, int64_t nrhs, cudaStream_t st) {
    const int bs = nrhs >= 148 * 2 * 256 ? 128 : 64;
    int grid = (int)std::max<int64_t>(1, (nrhs + bs - 1) / bs);
    k_banded_matvec<BW, 8><<<grid, bs, 0, st>>>(h->d_w, d_x, d_y, nrhs, (int)h->n);
    BSK_CUDA(cudaGetLastError());
    return BSK_OK;
}

}  // namespace

extern "C" bsk_status bsk_banded_matvec(const bsk_banded *h, const double *d_x, double *d_y, int64_t nrhs,
                                        void *stream) {
    BSK_REQUIRE(h && (nrhs == 0 || (d_x && d_y)), BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(!h->factored, BSK_ERR_ARGUMENT,
                "the band data now holds the LU factors (lu! works in place): apply C before factorising, or keep a copy");
    BSK_REQUIRE(d_x != d_y, BSK_ERR_ARGUMENT, "x and y must not alias");
    if (nrhs == 0) return BSK_OK;
    BSK_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    switch (h->l) {
        case 0: return matvec_t<0>(h, d_x, d_y, nrhs, st);
        case 1: return matvec_t<1>(h, d_x, d_y, nrhs, st);
        case 2: return matvec_t<2>(h, d_x, d_y, nrhs, st);
        case 3: return matvec_t<3>(h, d_x, d_y, nrhs, st);
        case 4: return matvec_t<4>(h, d_x, d_y, nrhs, st);
        case 5: return matvec_t<5>(h, d_x, d_y, nrhs, st);
        case 6: return matvec_t<6>(h, d_x, d_y, nrhs, st);
    }
    set_error("unsupported bandwidth %d", h->l);
    return BSK_ERR_UNSUPPORTED;
}
