/*
 * bsk.h — C ABI of libbsplinekit_b200.so (B200 / sm_100a).
 *
 * Drop-in boundary for the two data-parallel hot paths of BSplineKit.jl v0.19.2:
 * many-point spline evaluation and shared-xdata batched interpolation.  The
 * reference has no FFI of its own (it is pure Julia, extended by multiple dispatch),
 * so every entry point below names the reference *method* whose body it replaces
 * (file:line relative to the BSplineKit.jl source tree); the Julia-side `ccall`
 * overloads that bind them are in bsplinekit.jl_b200/julia/BSplineKitB200.jl and
 * are walked through in INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only; every function returns bsk_status (0 = ok, < 0 = error,
 *    text via bsk_last_error, thread-local).
 *  - all indices that cross the ABI are 1-based int64, exactly what the Julia
 *    reference returns.
 *  - pointers named d_* are DEVICE pointers on the handle's device; h_* are host
 *    pointers.  *_create copies what it needs; host arrays stay owned by the caller.
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Kernels
 *    are enqueued and the call returns without synchronising, except the *_host
 *    convenience calls, which stage through device memory and synchronise.
 *  - the handle-building calls (*_create, bsk_banded_lu[_mode], bsk_spline_set_coefs_device and the
 *    lazy table build inside the first bsk_spline_eval of a spline) take no stream: they run on the
 *    legacy default stream and return with their device work complete.  Streams from
 *    bsk_stream_create are BLOCKING streams, which order against the default stream; work a caller
 *    enqueues on its own non-blocking stream must be synchronised before such a call reads it.
 *  - the evaluation kernels of bsk_spline_eval* are launched with programmatic stream serialization: the
 *    staging of their (read-only, host-built) tables may overlap the tail of the previous kernel on the
 *    stream; points are read and results written only after that kernel has completed and flushed, so
 *    stream order is what the caller sees.
 *  - there is no CPU fallback: with no CUDA device every compute call fails with
 *    BSK_ERR_CUDA.
 */
#ifndef BSK_H
#define BSK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t bsk_status;

enum {
    BSK_OK = 0,
    BSK_ERR_ARGUMENT = -1,   /* -> Julia ArgumentError    (e.g. BSplines/basis.jl:80-82)     */
    BSK_ERR_DIMENSION = -2,  /* -> Julia DimensionMismatch (Collocation/matrix.jl:134-138)   */
    BSK_ERR_CUDA = -3,       /* -> ErrorException(bsk_last_error())                          */
    BSK_ERR_ZERO_PIVOT = -4, /* -> ZeroPivotException(i)   (Collocation/matrix.jl:150-198)   */
    BSK_ERR_ALLOC = -5,
    BSK_ERR_UNSUPPORTED = -6
};

enum { BSK_F32 = 0, BSK_F64 = 1 };

/* basis kinds */
enum {
    BSK_BASIS_PLAIN = 0,      /* BSplineBasis           (BSplines/basis.jl:70-90)       */
    BSK_BASIS_PERIODIC = 1,   /* PeriodicBSplineBasis   (BSplines/periodic.jl:181-205)  */
    BSK_BASIS_RECOMBINED = 2  /* RecombinedBSplineBasis (Recombinations/bases.jl:223-245) */
};

/* evaluation algorithm selector for bsk_spline_eval */
enum {
    BSK_EVAL_AUTO = 0,    /* fastest path whose result is within the stated tolerance       */
    BSK_EVAL_DEBOOR = 1,  /* de Boor triangle in the reference's operation order (bit-faithful)*/
    BSK_EVAL_PP = 2       /* per-interval local-polynomial table staged in shared memory    */
};

/* extrapolation method of bsk_spline_eval_extrap (SplineExtrapolations.jl:22-71) */
enum {
    BSK_EXTRAP_NONE = 0,    /* 0 outside the domain (Splines/spline.jl:164-168)              */
    BSK_EXTRAP_FLAT = 1,    /* Flat():   S(clamp(x, a, b))                                   */
    BSK_EXTRAP_LINEAR = 2,  /* Linear(): S(a) + S'(a) (x - a), S(b) + S'(b) (x - b)           */
    BSK_EXTRAP_SMOOTH = 3   /* Smooth(): the boundary polynomial pieces continued           */
};

/* banded LU mode selector for bsk_banded_lu_mode */
enum {
    BSK_LU_EXACT = 0,  /* BANFAC's statements in BANFAC's order, serial in the column index: bit-identical factors */
    BSK_LU_FAST = 1    /* partitioned elimination with validated overlap (n >= 1024), factors within ~1e-16
                          relative of BANFAC's; falls back to EXACT when a boundary check or a pivot fails  */
};

/* banded multi-RHS solve algorithm selector */
enum {
    BSK_SOLVE_AUTO = 0,
    BSK_SOLVE_TWOPASS = 1,  /* BANSLV order, forward sweep then backward sweep over HBM      */
    BSK_SOLVE_WINDOWED = 2  /* single pass: forward sweep + windowed backward sweep on chip  */
};

typedef struct bsk_basis bsk_basis;
typedef struct bsk_spline bsk_spline;
typedef struct bsk_banded bsk_banded;
typedef struct bsk_cyclic bsk_cyclic;
typedef struct bsk_cyclic_banded bsk_cyclic_banded;
typedef struct bsk_banded_f32 bsk_banded_f32;
typedef struct bsk_cyclic_f32 bsk_cyclic_f32;

/* Corner blocks of a RecombineMatrix (Recombinations/matrices.jl:70-110):
 * ul is ul_rows x nl, lr is lr_rows x nr, column-major doubles. */
typedef struct {
    int32_t cl, cr;           /* number of constraints left / right  (num_constraints) */
    int32_t nl, nr;           /* number of recombined functions       (num_recombined)  */
    int32_t ul_rows, lr_rows;
    const double *ul, *lr;
} bsk_recomb_desc;

/* ---- library / device utilities -------------------------------------------------- */
const char *bsk_version(void);
size_t bsk_last_error(char *buf, size_t buflen);
bsk_status bsk_device_count(int32_t *count);
/* Measurement aid (no reference counterpart): a pure streaming kernel with the access mix of the fused
 * value + derivative evaluation — n doubles read, 2 n doubles written, 16-byte vectors, same launch shape —
 * so that the HBM roofline of that 1 : 2 read : write stream can be measured beside the evaluation kernel. */
bsk_status bsk_probe_stream_1r2w(const double *d_in, double *d_out0, double *d_out1, int64_t n, void *stream);
/* The same stream moved by TMA: cp.async.bulk global -> shared for the point tiles, cp.async.bulk shared ->
 * global for the two result tiles (n a multiple of 2048).  Measures what a bulk-copy write path would give. */
bsk_status bsk_probe_stream_1r2w_tma(const double *d_in, double *d_out0, double *d_out1, int64_t n, void *stream);
/* Peak rate of the FP64 (dtype BSK_F64) or FP32 FMA pipe: `iters` x 8 independent fused multiply-adds per
 * thread at full occupancy; *fma_count = fused multiply-adds executed by the launch (time it with events).
 * The denominators of the pipe fractions bench.py reports beside the HBM fractions. */
bsk_status bsk_probe_fma(int32_t dtype, int64_t iters, void *d_scratch, double *fma_count, void *stream);
/* The memory behaviour of the global-table evaluation kernels (C4, C5) and nothing else: n doubles streamed in,
 * for each ONE 32-byte record gathered from a table of nrec records that lives in L2 (index (int)(x * nrec), x in
 * [0, 1)), a cubic Horner, n doubles streamed out.  variant 0 = the product kernel's access scheme (LDG.128 point
 * vectors with software prefetch, 256-bit gathers); the other variants move the point stream by per-warp TMA rings
 * of several shapes or change cache hints (csrc/bsk_probe.cu).  States the bound of those evaluations. */
bsk_status bsk_probe_gather(const double *d_x, double *d_out, int64_t n, const double *d_table, int64_t nrec,
                            int32_t variant, void *stream);
/* Copy-only twin of bsk_spline_eval_host: the same chunks on the same streams, host -> device copy of the
 * points and device -> host copy of nout result buffers, NO kernel — what the platform (PCIe, host memory)
 * gives the host-buffer call.  h_outs are overwritten with whatever the staging buffers hold. */
bsk_status bsk_probe_copy_host(const bsk_spline *s, const void *h_x, int64_t n, int32_t nout, void *const *h_outs);
bsk_status bsk_malloc(int32_t device, size_t nbytes, void **d_ptr);
bsk_status bsk_free(int32_t device, void *d_ptr);
bsk_status bsk_memcpy_h2d(int32_t device, void *d_dst, const void *h_src, size_t nbytes, void *stream);
bsk_status bsk_memcpy_d2h(int32_t device, void *h_dst, const void *d_src, size_t nbytes, void *stream);
bsk_status bsk_stream_create(int32_t device, void **stream);
bsk_status bsk_stream_destroy(int32_t device, void *stream);
bsk_status bsk_stream_sync(int32_t device, void *stream);

/* ---- bases ------------------------------------------------------------------------
 * kind PLAIN / RECOMBINED: h_knots = full knot vector t[1..N+k] (already augmented,
 *   BSplines/knots.jl:68-84), dtype BSK_F32/BSK_F64.  RECOMBINED additionally takes the
 *   corner blocks; knots are the parent's (Recombinations/bases.jl:300).
 * kind PERIODIC: h_knots = breakpoints in one period INCLUDING the endpoint
 *   (length N+1; BSplines/periodic.jl:37-55).
 * Replaces: the BSplineBasis / PeriodicBSplineBasis / RecombinedBSplineBasis structs as
 * seen by evaluate_all / find_knot_interval / Spline evaluation. */
bsk_status bsk_basis_create(int32_t kind, int32_t k, int32_t dtype, const void *h_knots,
                            int64_t nknots, const bsk_recomb_desc *recomb, int32_t device,
                            bsk_basis **out);
bsk_status bsk_basis_destroy(bsk_basis *b);
bsk_status bsk_basis_length(const bsk_basis *b, int64_t *len); /* length(B) */

/* find_knot_interval(ts, x)  — BSplines/evaluate_all.jl:102-119, periodic.jl:85-100,
 * knots.jl:32-41.  d_x: n abscissae of the basis dtype; d_idx: n int64 (1-based);
 * d_zone: n int8 in {-1,0,1} or NULL. */
bsk_status bsk_find_knot_interval(const bsk_basis *b, const void *d_x, int64_t n,
                                  int64_t *d_idx, int8_t *d_zone, void *stream);

/* evaluate_all(B, x, Derivative(deriv), T; ileft) — BSplines/evaluate_all.jl:53-64,
 * 128-204; periodic.jl:228-236; Recombinations/bases.jl:386-420.
 * out_dtype = T (BSK_F32 allowed with an F64 basis: knot differences are formed in
 * F64 then rounded, evaluate_all.jl:152-155).  d_bs is [n][k] with k fastest, i.e.
 * Julia's NTuple{k,T} per point, in the reference's REVERSED order (b_i .. b_{i-k+1}).
 * d_ileft: optional n int64 known interval indices (NULL = search). */
bsk_status bsk_evaluate_all(const bsk_basis *b, const void *d_x, int64_t n, int32_t deriv,
                            int32_t out_dtype, const int64_t *d_ileft, int64_t *d_idx,
                            void *d_bs, void *stream);

/* ---- splines ----------------------------------------------------------------------
 * Spline(B, coefs) — Splines/spline.jl:41-60.  h_coefs: length(B) values of the basis
 * dtype (for a recombined basis they are converted to parent coefficients once here,
 * Recombinations/splines.jl:32-65).  bsk_spline_set_coefs_device replaces the coefficients by
 * device-resident ones (e.g. one column of a bsk_banded_solve result with element stride
 * `stride`): they are copied (strided, synchronising `stream`), converted to the parent basis for a
 * recombined spline and re-uploaded; tables and cached derivative splines are rebuilt on the next
 * evaluation.  Not concurrent with evaluations of the same handle (it takes the handle's lock). */
bsk_status bsk_spline_create(const bsk_basis *b, const void *h_coefs, int64_t ncoef,
                             bsk_spline **out);
bsk_status bsk_spline_set_coefs_device(bsk_spline *s, const void *d_coefs, int64_t ncoef,
                                       int64_t stride, void *stream);
bsk_status bsk_spline_destroy(bsk_spline *s);
/* copy of coefficients(S) in the spline's own basis (parent basis for recombined) */
bsk_status bsk_spline_coefficients(const bsk_spline *s, void *h_out, int64_t ncoef);
bsk_status bsk_spline_order(const bsk_spline *s, int32_t *k, int64_t *ncoef, int64_t *nknots);
bsk_status bsk_spline_knots(const bsk_spline *s, void *h_out, int64_t nknots);

/* Derivative(n) * S — Splines/spline.jl:272-330, Splines/periodic.jl:76-117,
 * BSplines/basis.jl:104-111.  Returns a NEW spline of order k - n. */
bsk_status bsk_spline_derivative(const bsk_spline *s, int32_t n, bsk_spline **out);

/* (S::Spline)(x) broadcast over a device vector — Splines/spline.jl:144-209.
 * Evaluates nout derivative orders derivs[j] (0 = value) of S at the n points in d_x in
 * ONE pass over the points; d_outs[j] receives n values of the spline dtype.  Points
 * outside the knot domain give 0 (spline.jl:164-168); periodic bases wrap.
 * algo: BSK_EVAL_*. */
bsk_status bsk_spline_eval(const bsk_spline *s, const void *d_x, int64_t n, int32_t nout,
                           const int32_t *derivs, void *const *d_outs, int32_t algo,
                           void *stream);
/* extrapolate(S, method).(x) — SplineExtrapolation (src/SplineExtrapolations/SplineExtrapolations.jl:29-71),
 * fused into the same evaluation pass.  Every requested output D^d S is extrapolated as its own spline
 * (Linear uses D^{d+1} S at the boundary).  extrap: BSK_EXTRAP_*.  Periodic bases: Smooth is the
 * identity, Flat clamps to one period, Linear is not supported.  Linear needs the table path
 * (BSK_EVAL_AUTO / BSK_EVAL_PP). */
bsk_status bsk_spline_eval_extrap(const bsk_spline *s, const void *d_x, int64_t n, int32_t nout,
                                  const int32_t *derivs, void *const *d_outs, int32_t algo,
                                  int32_t extrap, void *stream);
/* Diagnostics of the tables BSK_EVAL_AUTO / BSK_EVAL_PP use for this spline (built here if they are not yet).
 * Every table coefficient is formed in double-double arithmetic and rounded once; each uniform cell and each
 * knot-interval record is CERTIFIED on the host to evaluate D^d S, d = 0 .. k-1, within a quarter of the
 * tolerance (1e-13 Float64 / 1e-5 Float32 of max |D^d S|); what cannot be certified is evaluated by the
 * de Boor recurrence in the reference's operation order (Splines/spline.jl:174-209).
 *   info[0] knot intervals          info[1] interval records routed to de Boor
 *   info[2] uniform cells (0: none) info[3] cells holding one knot (side entries)
 *   info[4] cells routed to the interval table             info[5] 1 = cell table stays in global memory
 *   info[6] shared-memory bytes of the staged cell table   info[7] 1 = BSK_EVAL_AUTO launches the cell kernel
 *   geom[0..3] (optional) = cell origin, domain end, cell width, 1 / cell width as the kernel uses them */
bsk_status bsk_spline_table_info(const bsk_spline *s, int64_t *info, double *geom);
/* Same through HOST buffers (H2D, kernel, D2H, synchronise). */
bsk_status bsk_spline_eval_host(const bsk_spline *s, const void *h_x, int64_t n, int32_t nout,
                                const int32_t *derivs, void *const *h_outs, int32_t algo);

/* Vector-valued splines: M splines sharing one basis, coefficients eltype SVector{M,T}
 * (Splines/spline.jl:180-184; test/periodic.jl:8-23, test/interpolation.jl:57).  d_coefs is
 * [ncoef][M] with M fastest — the layout of Vector{SVector{M,T}} and of bsk_banded_solve's result —
 * d_out is [n][M].  Plain and periodic bases; zero outside the knot domain of a plain basis. */
bsk_status bsk_spline_eval_multi(const bsk_basis *b, const void *d_coefs, int64_t ncoef, int64_t M,
                                 const void *d_x, int64_t n, void *d_out, void *stream);

/* ---- collocation --------------------------------------------------------------------
 * collocation_matrix!(C, B, x, Derivative(deriv); clip_threshold) —
 * Collocation/Collocation.jl:209-251 into CollocationMatrix band storage
 * data[u+1+i-j, j], (2k-3) x length(B) column-major (Collocation/matrix.jl:105-117,
 * Collocation.jl:165-198).  d_x: nx F64 points (nx must equal length(B): the band store is square, like the
 * reference's CollocationMatrix; other shapes: bsk_collocation_rows).
 * *h_n_outside (optional) receives the number of non-clipped entries that fell outside
 * the bands (the reference emits one @warn each, Collocation.jl:239-244).  Synchronises
 * only if h_n_outside != NULL. */
bsk_status bsk_collocation_matrix(const bsk_basis *b, const double *d_x, int64_t nx,
                                  int32_t deriv, double clip_threshold, double *d_band,
                                  int64_t *h_n_outside, void *stream);
/* collocation_matrix! for ANY number of points nx (rectangular matrices: least-squares fitting with more data
 * than basis functions) — the Matrix / SparseMatrixCSC targets of Collocation/Collocation.jl:82-106,162-193,
 * 209-251.  Row form with exactly k slots per point: d_col[nx][k] = 1-based column index (0: clipped entry, or
 * point outside the domain of a non-periodic basis), d_val[nx][k] = D^deriv b_col(x_i), slots in the reference's
 * order (column jlast + 1 - dj, dj = 1..k; periodic columns wrapped, BSplines/periodic.jl:245-253) — the rows of
 * a CSR / ELL matrix, from which the shim builds SparseMatrixCSC.  d_dense (optional): nx x length(B)
 * column-major (a Julia Matrix), zero-filled here.  Any of the three outputs may be NULL.  Plain, periodic and
 * recombined Float64 bases. */
bsk_status bsk_collocation_rows(const bsk_basis *b, const double *d_x, int64_t nx, int32_t deriv,
                                double clip_threshold, int64_t *d_col, double *d_val, double *d_dense,
                                void *stream);
/* Periodic cubic: CyclicTridiagonalMatrix diagonals a, b, c (Collocation/cyclic.jl:21-33). */
bsk_status bsk_collocation_cyclic(const bsk_basis *b, const double *d_x, int64_t nx,
                                  int32_t deriv, double clip_threshold, double *d_a, double *d_b,
                                  double *d_c, int64_t *h_n_outside, void *stream);

/* ---- banded LU + multi-RHS solve ----------------------------------------------------
 * CollocationMatrix / lu! / ldiv!  — Collocation/matrix.jl:40-117,132-201,218-271.
 * d_band: (l+u+1) x n column-major F64 with l == u; copied into the handle. */
bsk_status bsk_banded_create(const double *d_band, int64_t n, int32_t l, int32_t u, int32_t device,
                             bsk_banded **out);
bsk_status bsk_banded_create_host(const double *h_band, int64_t n, int32_t l, int32_t u,
                                  int32_t device, bsk_banded **out);
/* lu!(C): BANFAC without pivoting, in place in the handle.  On a zero pivot returns
 * BSK_ERR_ZERO_PIVOT and *info_pivot = 1-based row (ZeroPivotException(i)). */
bsk_status bsk_banded_lu(bsk_banded *h, int64_t *info_pivot);
/* The same with a mode (BSK_LU_*): BSK_LU_FAST factors partitions of the matrix in parallel — the recurrence
 * forgets its start state, every partition starts 128 columns early and the state it hands to the next one is
 * checked against what that partition computed itself (1e-14 relative).  lu! of C3 (n = 4096) in ~0.1 ms
 * instead of 1.1 ms, n = 10^5 in ~0.15 ms instead of 30 ms; ZeroPivotException(i) is still reported. */
bsk_status bsk_banded_lu_mode(bsk_banded *h, int32_t mode, int64_t *info_pivot);
bsk_status bsk_banded_data(const bsk_banded *h, double *h_out); /* (l+u+1) x n copy */
/* Rows of halo validated at factor time for BSK_SOLVE_WINDOWED (0 = windowed unavailable). */
bsk_status bsk_banded_halo(const bsk_banded *h, int32_t *halo);
/* ldiv!(F, Y): BANSLV on nrhs right-hand sides stored interleaved, d_rhs[(i-1)*nrhs + c]
 * (== Vector{SVector{nrhs,Float64}}; test/interpolation.jl:57), solved IN PLACE.
 * algo: BSK_SOLVE_*. */
bsk_status bsk_banded_solve(const bsk_banded *h, double *d_rhs, int64_t nrhs, int32_t algo,
                            void *stream);
bsk_status bsk_banded_solve_host(const bsk_banded *h, double *h_rhs, int64_t nrhs, int32_t algo);
/* mul!(Y, C, X) for nrhs interleaved vectors — Collocation/matrix.jl:91-94 (BandedMatrices gbmv in the
 * reference).  Only before bsk_banded_lu (which overwrites the band data, like lu!).  X and Y must not alias. */
bsk_status bsk_banded_matvec(const bsk_banded *h, const double *d_x, double *d_y, int64_t nrhs, void *stream);
/* du = F \ (Lm * u): `mul!(y, Lm, u); ldiv!(du, F, y)` of a collocation time step
 * (examples/heat.jl:314-328; Collocation/matrix.jl:91-94 + :218-271) in ONE pass over HBM (16 B per
 * element instead of 32).  lu: factorised; Lm: NOT factorised, same size; d_u is only read, d_out must
 * not alias it.  Shapes the fused kernel does not cover run bsk_banded_matvec + bsk_banded_solve. */
bsk_status bsk_banded_matvec_solve(const bsk_banded *lu, const bsk_banded *Lm, const double *d_u,
                                   double *d_out, int64_t nrhs, void *stream);
bsk_status bsk_banded_destroy(bsk_banded *h);

/* ---- cyclic tridiagonal (periodic cubic interpolation) -------------------------------
 * CyclicTridiagonalMatrix + ldiv! — Collocation/cyclic.jl:21-33,112-197.  The reference
 * re-eliminates the matrix for every right-hand side (SplineInterpolations.jl:147-149);
 * here it is factored once at create. */
bsk_status bsk_cyclic_create(const double *h_a, const double *h_b, const double *h_c, int64_t n,
                             int32_t device, bsk_cyclic **out);
bsk_status bsk_cyclic_create_device(const double *d_a, const double *d_b, const double *d_c,
                                    int64_t n, int32_t device, bsk_cyclic **out);
/* d_rhs interleaved [n][nrhs], solved in place. */
bsk_status bsk_cyclic_solve(const bsk_cyclic *h, double *d_rhs, int64_t nrhs, void *stream);
bsk_status bsk_cyclic_destroy(bsk_cyclic *h);

/* ---- cyclic banded (periodic interpolation, any order but k = 4) ---------------------
 * The reference builds a SparseMatrixCSC and factorises it with UMFPACK
 * (Collocation/Collocation.jl:141-142, SplineInterpolations.jl:109).  Here: banded LU of the
 * non-wrapped part, once, + Woodbury correction of the two bw x bw corners.
 * h_cband: (2 bw + 1) x n column-major, A(i, j) with cyclic offset d = i - j (mod n) in
 * [shift - bw, shift + bw] at h_cband[(bw + d - shift) + (j - 1) (2 bw + 1)] (1-based i, j); `shift` =
 * offset of the band centre (0 for even orders; odd orders have b_j peaking at x_{j+1}: shift = 1).
 * Needs n >= 4 bw + 1. */
bsk_status bsk_cyclic_banded_create(const double *h_cband, int64_t n, int32_t bw, int32_t shift,
                                    int32_t device, bsk_cyclic_banded **out);
/* In place on nrhs interleaved right-hand sides d_rhs[(i-1)*nrhs + c]. */
bsk_status bsk_cyclic_banded_solve(const bsk_cyclic_banded *h, double *d_rhs, int64_t nrhs, void *stream);
bsk_status bsk_cyclic_banded_destroy(bsk_cyclic_banded *h);

/* ---- the interpolation path in Float32 ------------------------------------------------
 * The reference code is generic in the element type: interpolate(x::Vector{Float32}, y, ...) builds a
 * CollocationMatrix{Float32} (Collocation/Collocation.jl:165-198) and runs lu! / ldiv!
 * (Collocation/matrix.jl:132-201,218-271) in Float32; periodic cubic bases use
 * CyclicTridiagonalMatrix{Float32} (Collocation/cyclic.jl:112-197).  Same layouts, argument meaning
 * and status codes as the Float64 entry points above; the basis must have been created with
 * BSK_F32 knots.  Assembly, factors and BSK_SOLVE_TWOPASS solutions keep the reference's operation order in
 * Float32 (bit-identical to a Float32 CPU restatement, see tests/); the cyclic solve is factored once. */
bsk_status bsk_collocation_matrix_f32(const bsk_basis *b, const float *d_x, int64_t nx, int32_t deriv,
                                      float clip_threshold, float *d_band, int64_t *h_n_outside,
                                      void *stream);
bsk_status bsk_collocation_cyclic_f32(const bsk_basis *b, const float *d_x, int64_t nx, int32_t deriv,
                                      float clip_threshold, float *d_a, float *d_b, float *d_c,
                                      int64_t *h_n_outside, void *stream);
bsk_status bsk_banded_f32_create(const float *d_band, int64_t n, int32_t l, int32_t u, int32_t device,
                                 bsk_banded_f32 **out);
bsk_status bsk_banded_f32_create_host(const float *h_band, int64_t n, int32_t l, int32_t u,
                                      int32_t device, bsk_banded_f32 **out);
bsk_status bsk_banded_f32_lu(bsk_banded_f32 *h, int64_t *info_pivot);
bsk_status bsk_banded_f32_data(const bsk_banded_f32 *h, float *h_out);
/* Rows of halo validated at factor time for the single-pass solve (0 = unavailable). */
bsk_status bsk_banded_f32_halo(const bsk_banded_f32 *h, int32_t *halo);
/* algo: BSK_SOLVE_TWOPASS = BANSLV order, bit-identical to the Float32 restatement of ldiv!;
 * BSK_SOLVE_WINDOWED = single pass over HBM (8 B per element), two right-hand sides per lane with packed
 * Float32 FMAs, within 1e-5 of the BANSLV result; BSK_SOLVE_AUTO = single pass when the matrix has a
 * validated halo and the block has >= 64 columns. */
bsk_status bsk_banded_f32_solve(const bsk_banded_f32 *h, float *d_rhs, int64_t nrhs, int32_t algo, void *stream);
bsk_status bsk_banded_f32_destroy(bsk_banded_f32 *h);
bsk_status bsk_cyclic_f32_create(const float *h_a, const float *h_b, const float *h_c, int64_t n,
                                 int32_t device, bsk_cyclic_f32 **out);
bsk_status bsk_cyclic_f32_solve(const bsk_cyclic_f32 *h, float *d_rhs, int64_t nrhs, void *stream);
bsk_status bsk_cyclic_f32_destroy(bsk_cyclic_f32 *h);

/* ---- Galerkin matrices and projections (plain and recombined Float64 bases) ------------
 * galerkin_matrix!(M, B, (Derivative(deriv1), Derivative(deriv2)))  — src/Galerkin/linear.jl:220-311:
 *   M[i, j] = integral of D^deriv1 b_i * D^deriv2 b_j, Gauss-Legendre with the caller's nodes / weights on
 *   [-1, 1] (h_qx, h_qw, nq; the reference uses nq = k, quadratures.jl:23-33) on every knot segment.
 *   d_band: (2 (k - 1) + 1) x length(B) column-major band store data[(k - 1) + i - j, j], overwritten.
 *   Contributions are added in the reference's order (segment by segment, node by node).
 * galerkin_projection!(f, phi, B, Derivative(deriv))  — src/Galerkin/projection.jl:44-99:
 *   phi[i] = integral of f * D^deriv b_i.  f is the caller's: bsk_galerkin_nodes returns the abscissae
 *   (host, reference order), d_fx holds f there. */
bsk_status bsk_galerkin_nodes(const bsk_basis *b, const double *h_qx, int32_t nq, double *h_x,
                              int64_t capacity, int64_t *npts);
bsk_status bsk_galerkin_matrix(const bsk_basis *b, int32_t deriv1, int32_t deriv2, const double *h_qx,
                               const double *h_qw, int32_t nq, double *d_band, void *stream);
/* galerkin_matrix(f, B, ...) (Galerkin/linear.jl:60-96,305): A[i, j] += y * b_i * b_j * f(x); d_fx = f at the
 * nfx nodes bsk_galerkin_nodes lists for the same quadrature (NULL: f = 1, = bsk_galerkin_matrix). */
bsk_status bsk_galerkin_matrix_weighted(const bsk_basis *b, int32_t deriv1, int32_t deriv2, const double *h_qx,
                                        const double *h_qw, int32_t nq, const double *d_fx, int64_t nfx,
                                        double *d_band, void *stream);
/* galerkin_matrix((B1, B2), (Derivative(m), Derivative(n))) for two bases that share their parent B-spline basis,
 * e.g. (R, B) or two recombinations (Galerkin/linear.jl:134-199,220-311; test/airy.jl:66-69): N1 x N2 with
 * l = (k-1) + dl sub- and u = (k-1) - dl super-diagonals, dl = cl(B2) - cl(B1), band store data[u + i - j, j] of
 * (l + u + 1) x N2 doubles.  d_band == NULL only returns *h_l, *h_u.  d_fx: optional weight at the nodes (or NULL). */
bsk_status bsk_galerkin_matrix_pair(const bsk_basis *b1, const bsk_basis *b2, int32_t deriv1, int32_t deriv2,
                                    const double *h_qx, const double *h_qw, int32_t nq, const double *d_fx, int64_t nfx,
                                    double *d_band, int32_t *h_l, int32_t *h_u, void *stream);
bsk_status bsk_galerkin_projection(const bsk_basis *b, int32_t deriv, const double *h_qx,
                                   const double *h_qw, int32_t nq, const double *d_fx, int64_t nfx,
                                   double *d_phi, void *stream);

/* ---- timing helper (CUDA events on `stream`) ----------------------------------------- */
bsk_status bsk_event_create(int32_t device, void **event);
bsk_status bsk_event_record(void *event, void *stream);
bsk_status bsk_event_elapsed_ms(void *start, void *stop, float *ms); /* synchronises stop */
bsk_status bsk_event_destroy(void *event);

#ifdef __cplusplus
}
#endif
#endif /* BSK_H */
