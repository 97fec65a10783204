// handles.cuh — opaque handle layouts, error plumbing and launch helpers.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bsk.h"
#include "core.cuh"

namespace bsk {

void set_error(const char *fmt, ...);
// Device memory of the handles comes from the stream-ordered pool of the current device: plain
// cudaMalloc / cudaFree cost tens of milliseconds each once the process maps many GB (measured:
// 60 ms per band handle next to a 17 GB torch cache, 1 ms alone); pool allocations are served from
// memory the pool retains (release threshold 1 GiB).
void ensure_pool(int device);
cudaError_t dev_malloc(void **p, size_t bytes);
cudaError_t dev_free(void *p);
template <typename P> inline cudaError_t dev_malloc(P **p, size_t bytes) { return dev_malloc((void **)p, bytes); }

#define BSK_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            bsk::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                           __LINE__);                                                        \
            return BSK_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

#define BSK_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            bsk::set_error(__VA_ARGS__);  \
            return (code);                \
        }                                 \
    } while (0)

inline int sm_count(int device) {
    static int cached[64] = {0};
    if (device < 0 || device >= 64) return 148;
    if (cached[device] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0)
            n = 148;
        cached[device] = n;
    }
    return cached[device];
}

}  // namespace bsk

struct bsk_basis {
    int kind;       // BSK_BASIS_*
    int k;
    int dtype;      // BSK_F32 / BSK_F64
    int device;
    int64_t nt;     // stored knots: plain N+k, periodic N+1
    int64_t N;      // length of the (parent) basis
    int64_t len;    // length(B) as the user sees it (recombined: N - cl - cr)
    std::vector<double> h_knots;   // always kept in double on the host (exact for f32 input)
    void *d_knots;                 // device copy in `dtype`
    int ilast;                     // walk-back target (plain)
    // search LUT
    uint32_t *d_lut;
    int ncell;
    double lut_scale;
    // recombination
    bsk::RecombView rv;
};

struct bsk_spline {
    bsk_basis *basis;     // owned evaluation basis (the PARENT basis for a recombined spline)
    int64_t ncoef;        // number of coefficients in the evaluation basis
    std::vector<double> h_coefs;   // host mirror (double; exact for f32)
    void *d_coefs;        // device, basis dtype, contiguous
    bool owns_coefs;
    // what the user handed us (needed again by bsk_spline_set_coefs_device)
    int user_basis_kind;
    bsk::RecombView user_rv;
    int64_t user_len;
    // local-polynomial table (built lazily): per non-empty knot interval one record
    //   { t_lo, t_hi, c_0 .. c_{k-1}, pad }  with S(x) = sum_m c_m (x - mid)^m on [t_lo, t_hi)
    void *d_pp;
    uint32_t *d_pp_lut;   // search LUT over the record left ends
    int64_t pp_nrec;
    int pp_ncell;
    double pp_scale;
    int pp_rec;
    int64_t pp_uncertified;   // interval records flagged for the de Boor route (pptable.h)
    int range_safe_d = 8;     // largest derivative order the tables can form inside the element type's range (pptable.h)
    bool pp_valid;
    // uniform-cell table (bsk_eval_cell.cu)
    void *d_cell_rec, *d_cell_rc;
    int cell_ncell, cell_nside;
    double cell_x0, cell_b, cell_scale, cell_cw;
    size_t cell_smem;
    int64_t cell_slow;        // cells routed to the interval table
    bool cell_valid;
    std::vector<bsk_spline *> derivs;   // cached Derivative(n) * S, index n (0 unused)
    std::mutex mu;        // guards the lazily built tables / derivative cache (evaluation is otherwise read-only)
};

struct bsk_banded {
    int device;
    int64_t n;
    int l, u;
    double *d_w;          // (l+u+1) x n column-major: raw matrix, LU factors after bsk_banded_lu
    bool factored;
    // solve-ready repack (built by bsk_banded_lu): row-major per-row multipliers
    double *d_lrow;       // n x l : lrow[i][j-1] = L(i, i-j)      (multiplier of x[i-j] in row i)
    double *d_urow;       // n x u : urow[i][j-1] = U(i, i+j)
    double *d_dinv;       // n     : 1 / U(i,i)
    double *d_diag;       // n     : U(i,i)
    double *d_lrec, *d_urec;   // padded per-row records of the windowed solve (bsk_solve_win.cu)
    int n_pad;
    // one device allocation holds every array above (created once in bsk_banded_create*), plus the second
    // band buffer, the boundary columns and the status words of the partitioned factorisation (banfac.cuh)
    void *arena;
    double *d_w2, *d_ver;
    int *d_status;
    int halo_f;           // decay halo of the forward recurrence (row-segmented windowed solve)
    int halo;             // decay halo (rows) validated for the windowed solve, 0 = not usable
    // unfactored matrices used as the Lm of bsk_banded_matvec_solve: cached row records
    double *d_mrec;
    int mrec_bw, mrec_npad;
};

struct bsk_banded_f32 {
    int device;
    int64_t n;
    int l, u;
    float *d_w;           // (l+u+1) x n column-major: matrix, LU factors after bsk_banded_f32_lu
    float *d_lrow;        // n x l: lrow[i][j-1] = L(i, i-j)
    float *d_urow;        // n x u: urow[i][j-1] = U(i, i+j)
    float *d_diag;        // n:     U(i, i)
    bool factored;
    // single-pass (TMEM) solve, two right-hand sides per lane (bsk_solve_win.cu)
    double *d_lrec, *d_urec;   // records of packed, negated float pairs
    int n_pad;
    int halo, halo_f;     // decay halos validated at factor time (0 = single pass unavailable)
};

namespace bsk {
// record widths of the single-pass solve (bsk_solve_win.cu: RecW)
constexpr int win_lw(int bw) { return ((bw + 2) / 2) * 2; }
constexpr int win_uw(int bw) { return bw == 0 ? 2 : ((bw + 1) / 2) * 2; }
// decay bound of a triangular recurrence (bsk_solve.cu): rows rec[r][j-1] * scale[r]
int decay_halo_device(const double *d_rec, const double *d_scale, int n, int bw, int reversed, double tol, int step,
                      int hmax, int phase);
bsk_status windowed_prepare_f32(bsk_banded_f32 *h);
bsk_status windowed_solve_f32(const bsk_banded_f32 *h, float *d_rhs, int64_t ngroups, int64_t ld_pairs, cudaStream_t st);
bsk_status windowed_prepare(bsk_banded *h, const double *w);
bsk_status windowed_solve(const bsk_banded *h, double *d_rhs, int64_t ngroups, int64_t ld, cudaStream_t st);
bsk_status windowed_matrec(const bsk_banded *Lm, int bw, int n_pad, double **out);
bsk_status windowed_solve_fused(const bsk_banded *h, const double *mrec, const double *d_u, double *d_out,
                                int64_t ngroups, int64_t ld, cudaStream_t st);
}  // namespace bsk

struct bsk_cyclic {
    int device;
    int64_t n;
    // factor-once form of the Sherman-Morrison solve (Collocation/cyclic.jl:112-161)
    bsk_banded *tri;  // LU of the tridiagonal part T (l = u = 1)
    double *d_q;      // n: T^{-1} u
    int zone;         // rows from each end where q is not negligible
    double vn;        // a1 / gamma
    double denom;     // 1 + v.q
};
