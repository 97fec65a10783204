// bsk_solve_win.cu — K6b: single-pass multi-RHS banded solve (windowed backward sweep).
//
// Replaces ldiv!(x, F::CollocationLU, y) (src/Collocation/matrix.jl:218-271) for batches of
// right-hand sides stored interleaved (RHS index fastest == Vector{SVector{M,Float64}},
// test/interpolation.jl:57).
//
// One warp per CTA, one right-hand side (column) per lane.  Three TMA streams (1-D bulk async
// copies, cp.async.bulk, completion on mbarriers) feed the warp from HBM / L2 without touching
// registers or L1:
//   A  the column group's rows, 8 rows (8 x 256 B) per chunk, PG chunks ahead, landing in the
//      warp's ring of (M + H + 8 PG) rows in shared memory;
//   B  the per-row records {L(r,r-bw..r-1), 1/U(r,r)} of the same 8 rows (same mbarrier as A);
//   C  the per-row records {U(r,r+bw..r+1)/U(r,r)} of the same rows, kept beside the ring for
//      as long as the rows are (the backward sweeps read them 1 + H/M times).
// The forward sweep (BANSLV's recurrence, FMA-contracted) runs in place in the ring; the
// backward recurrence is restarted every M rows from H rows further down with a zero state.
// The influence of that wrong start state on the rows that are kept decays like |U^{-1}| away
// from the diagonal and was bounded below 1e-17 (relative to the solution's magnitude) at
// factor time for this H (bsk_banded_lu), i.e. far below one ulp.
// HBM traffic: every element is read once and written once (16 B per element); the two-pass
// BANSLV-order kernels (bsk_solve.cu) move 32 B per element.
//
// Everything in the inner loops is organised in chunks of 8 rows that never straddle a ring
// wrap, so ring accesses use immediate offsets, and the records are read from shared memory
// with conflict-free broadcast loads.
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include <cuda.h>            // CUtensorMap (types only; the encoder is fetched at run time)

#include "handles.cuh"
#include "tma.cuh"

using namespace bsk;

namespace {

constexpr int G = 8;   // rows per chunk

template <int BW> struct RecW {
    static constexpr int LW = bsk::win_lw(BW);                  // {l_BW .. l_1, dinv} padded to even
    static constexpr int UW = bsk::win_uw(BW);                  // {u'_BW .. u'_1} padded to even
};

// Column-oriented ("right-looking", exactly BANSLV's own loop order, matrix.jl:246-266) records:
//   lrec[i] = { L(i+1,i), ..., L(i+BW,i), 1/U(i,i), pad }      multipliers applied once z_i is final
//   urec[r] = { U(r-1,r)/U(r-1,r-1), ..., U(r-BW,r)/U(r-BW,r-BW), pad }   (targets pre-scaled)
// rows r > n are identity rows (zero multipliers, unit pivot).
template <int BW>
__global__ void k_win_repack(const double *__restrict__ w, int n, int n_pad, double *__restrict__ lrec,
                             double *__restrict__ urec) {
    constexpr int LW = RecW<BW>::LW, UW = RecW<BW>::UW;
    const int nroww = 2 * BW + 1, middle = BW + 1;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x + 1; r <= n_pad; r += gridDim.x * blockDim.x) {
        double *lr = lrec + (size_t)(r - 1) * LW;
        double *ur = urec + (size_t)(r - 1) * UW;
        for (int j = 0; j < LW; ++j) lr[j] = 0.0;
        for (int j = 0; j < UW; ++j) ur[j] = 0.0;
        if (r > n) { lr[BW] = 1.0; continue; }
        lr[BW] = 1.0 / w[(middle - 1) + (size_t)(r - 1) * nroww];
        for (int j = 1; j <= BW; ++j) {
            if (r + j <= n) lr[j - 1] = w[(middle + j - 1) + (size_t)(r - 1) * nroww];
            if (r - j >= 1)
                ur[j - 1] = w[(middle - j - 1) + (size_t)(r - 1) * nroww] / w[(middle - 1) + (size_t)(r - j - 1) * nroww];
        }
    }
}

// (A TMA tile store of the results from the ring was tried and measured slower than direct
// coalesced STG on B200: 1.28 ms vs 1.22 ms on C3.)

template <int W>
__device__ __forceinline__ void load_rec_s(const double *p, double *out) {      // shared memory, 16-B aligned
#pragma unroll
    for (int i = 0; i < W / 2; ++i) {
        const double2 v = *((const double2 *)p + i);
        out[2 * i] = v.x;
        out[2 * i + 1] = v.y;
    }
}

// `ld` = elements between consecutive rows of rhs.
template <int BW, int PG>
__global__ void __launch_bounds__(32)
k_banded_windowed(const double *__restrict__ lrec, const double *__restrict__ urec,
                  double *__restrict__ rhs, int64_t ngroups, int64_t ld, int n, int n_pad_all, int M, int H,
                  const __grid_constant__ CUtensorMap tmap, int use_tmap, int nseg, int SEG, int HF,
                  const double *__restrict__ scratch, const __grid_constant__ CUtensorMap tmap_s) {
    constexpr int TB = 32;
    constexpr int LW = RecW<BW>::LW, UW = RecW<BW>::UW;
    extern __shared__ __align__(128) double smem_d[];
    const int RW = M + H + PG * G;
    // shared-memory layout (doubles): ring[RW][32] | cur[RW][UW] | cfw[PG][8*LW] | mbar[PG]
    // cur = backward records of the rows currently in the ring (same slot index as the data)
    double *const ring = smem_d;
    double *const cur = ring + (size_t)RW * TB;
    double *const cfw = cur + (size_t)RW * UW;
    const int lane = threadIdx.x;
    double *const my = ring + lane;                          // my[slot * TB]
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    const unsigned cur_s = (unsigned)__cvta_generic_to_shared(cur);
    const unsigned cfw_s = (unsigned)__cvta_generic_to_shared(cfw);
    const unsigned barA_s = cfw_s + (unsigned)(PG * G * LW * 8);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < PG; ++i) mbar_init(barA_s + 8 * i, 1);
        fence_proxy_async();
    }
    __syncwarp();
    unsigned issA = 0, conA = 0;                             // chunk counters (mbarrier = counter % PG)

    // Work item = (column group g, row segment seg).  A segment finalises the rows [k0, k1); its
    // forward sweep starts HF rows earlier from a zero state (the forward recurrence forgets its
    // start state like |L^{-1}| decays; HF validated at factor time like H) and runs H rows past
    // k1 for the last backward restart.  With nseg == 1 this is the whole column, exactly.
    // The solve is in place, so the halo rows of a segment (which belong to its neighbours and
    // may already hold their results) are read from `scratch`, a copy of the rows
    // [kb - HF, kb + H) around every segment boundary kb taken before the kernel started.
    for (int64_t w = blockIdx.x; w < ngroups * nseg; w += gridDim.x) {
        const int64_t g = w / nseg;
        const int seg = (int)(w - g * nseg);
        const int k0 = seg * SEG, k1 = min(n_pad_all, k0 + SEG);
        const int f0 = max(0, k0 - HF);                      // first row of the forward sweep
        const int n_pad = min(n_pad_all, k1 + H) - f0;       // local number of rows (multiple of 8)
        const int nloc = min(n, f0 + n_pad) - f0;            // local rows that exist in HBM
        const int kl0 = k0 - f0, kl1 = k1 - f0;              // kept rows, local
        const int nblk = (n_pad + M - 1) / M;
        double *const p = rhs + (int64_t)f0 * ld + g * TB + lane;
        const double *const grow = rhs + (int64_t)f0 * ld + g * TB;   // local row 0 of this column group
        const double *const lrec_l = lrec + (size_t)f0 * LW;
        const double *const urec_l = urec + (size_t)f0 * UW;

        // ---- producer: 8 data rows + their forward and backward records, one mbarrier phase --------
        int pf_row = 0, pf_slot = 0;
        auto issue_fwd = [&]() {
            const unsigned i = issA % PG;
            const unsigned bar = barA_s + 8 * i;
            const int nvalid = min(G, nloc - pf_row);        // rows that exist in HBM
            if (!use_tmap && nvalid < G && pf_row < n_pad) { // identity padding rows (TMA zero-fills them)
#pragma unroll
                for (int q = 0; q < G; ++q)
                    if (q >= nvalid) my[(pf_slot + q) * TB] = 0.0;
            }
            __syncwarp();                                    // all lanes are done with these slots
            if (elect_one()) {
                if (pf_row < n_pad) {
                    // result tiles stored from these slots (a ring lap ago) must have been read out
                    fence_proxy_async();                     // generic accesses before async writes
                    const unsigned dst = ring_s + (unsigned)(pf_slot * TB * 8);
                    // halo chunk of a segment -> scratch copy (chunks never straddle k0 / k1)
                    const int ra = f0 + pf_row;
                    const bool halo = (nseg > 1) && (ra < k0 || ra >= k1);
                    const int srow = halo ? ((ra < k0 ? seg - 1 : seg) * (HF + H) + ra - ((ra < k0 ? k0 : k1) - HF)) : 0;
                    if (use_tmap) {
                        mbar_expect_tx(bar, (unsigned)(G * TB * 8 + G * (LW + UW) * 8));
                        if (halo) tma_load_2d(dst, &tmap_s, (int)(g * TB), srow, bar);
                        else tma_load_2d(dst, &tmap, (int)(g * TB), ra, bar);
                    } else {
                        mbar_expect_tx(bar, (unsigned)(nvalid * TB * 8 + G * (LW + UW) * 8));
                        const double *src = halo ? scratch + (int64_t)srow * ld + g * TB : grow + (int64_t)pf_row * ld;
                        if (nvalid == G) {
#pragma unroll
                            for (int q = 0; q < G; ++q) bulk_g2s(dst + q * TB * 8, src + (int64_t)q * ld, TB * 8, bar);
                        } else {
                            for (int q = 0; q < nvalid; ++q) bulk_g2s(dst + q * TB * 8, src + (int64_t)q * ld, TB * 8, bar);
                        }
                    }
                    bulk_g2s(cfw_s + i * (G * LW * 8), lrec_l + (size_t)pf_row * LW, G * LW * 8, bar);
                    bulk_g2s(cur_s + (unsigned)(pf_slot * UW * 8), urec_l + (size_t)pf_row * UW, G * UW * 8, bar);
                } else {
                    mbar_arrive(bar);
                }
            }
            ++issA;
            pf_row += G;
            pf_slot += G;
            if (pf_slot == RW) pf_slot = 0;
        };
#pragma unroll 1
        for (int gg = 0; gg < PG; ++gg) issue_fwd();

        int f_slot = 0;                                      // ring slot of the next forward row
        double tf[BW > 0 ? BW : 1];                          // rows r0 .. r0+BW-1, partially eliminated
        {
            mbar_wait(barA_s + 8 * (conA % PG), (conA / PG) & 1);      // chunk 0
            ++conA;
#pragma unroll
            for (int j = 0; j < BW; ++j) tf[j] = my[j * TB];
        }
        int fpos = 0;                                        // rows forwarded so far
        for (int bb = 0; bb < nblk; ++bb) {
            // ---- forward sweep up to the restart row of block bb, (bb + 1) M + H (right-looking: z_i final ->
            //      update rows i+1..i+BW); H need not be a multiple of M (see k_banded_tmem) ----
            {
                const int r1 = min(n_pad, (bb + 1) * M + H);
#pragma unroll 1
                for (int r0 = fpos; r0 < r1; r0 += G) {
                    const unsigned i = (conA - 1) % PG;                    // chunk of rows r0 .. r0+7
                    mbar_wait(barA_s + 8 * (conA % PG), (conA / PG) & 1);  // next chunk (look-ahead rows)
                    ++conA;
                    double *const rs = my + f_slot * TB;
                    const double *const lr = cfw + i * (G * LW);
                    int n_slot = f_slot + G;
                    if (n_slot == RW) n_slot = 0;
                    double t[G + BW];
#pragma unroll
                    for (int j = 0; j < BW; ++j) t[j] = tf[j];
#pragma unroll
                    for (int q = BW; q < G; ++q) t[q] = rs[q * TB];
                    if (r0 + G < n_pad) {
#pragma unroll
                        for (int j = 0; j < BW; ++j) t[G + j] = my[(n_slot + j) * TB];
                    } else {
#pragma unroll
                        for (int j = 0; j < BW; ++j) t[G + j] = 0.0;
                    }
                    double rec[G][LW];
#pragma unroll
                    for (int q = 0; q < G; ++q) load_rec_s<LW>(lr + q * LW, rec[q]);
#pragma unroll
                    for (int q = 0; q < G; ++q) {
                        const double z = t[q];
#pragma unroll
                        for (int j = 1; j <= BW; ++j) t[q + j] = fma(-z, rec[q][j - 1], t[q + j]);
                        rs[q * TB] = z * rec[q][BW];         // stored pre-scaled by 1/U(r,r)
                    }
#pragma unroll
                    for (int j = 0; j < BW; ++j) tf[j] = t[G + j];
                    f_slot = n_slot;
                    issue_fwd();
                }
                fpos = max(fpos, r1);
            }
            // ---- backward sweep that finalises block bb ------------------------------------------
            if (bb * M >= kl0 && bb * M < kl1) {
                const int keep_lo = bb * M;
                const int keep_hi = min(n_pad, (bb + 1) * M);
                const int r_top = min(n_pad, (bb + 1) * M + H);   // exclusive, multiple of 8
                double tb[BW > 0 ? BW : 1];                  // rows rc-1 .. rc-BW, partially eliminated
                int slot = r_top % RW;                       // once per block
                slot = (slot == 0 ? RW : slot) - G;
                double *dst = p + (int64_t)(keep_hi - 1) * ld;
#pragma unroll 1
                for (int rc = r_top - G; rc >= keep_lo; rc -= G) {
                    const double *const rs = my + slot * TB;
                    const double *const ur = cur + slot * UW;
                    const int p_slot = (slot == 0 ? RW : slot) - G;
                    double t[G];
                    double la[BW > 0 ? BW : 1];
                    if (rc == r_top - G) {                   // start of the sweep: zero state above
#pragma unroll
                        for (int q = 0; q < G; ++q) t[q] = rs[q * TB];
                    } else {
#pragma unroll
                        for (int j = 0; j < BW; ++j) t[G - 1 - j] = tb[j];
#pragma unroll
                        for (int q = 0; q < G - BW; ++q) t[q] = rs[q * TB];
                    }
                    if (rc > keep_lo) {
#pragma unroll
                        for (int j = 0; j < BW; ++j) la[j] = my[(p_slot + G - 1 - j) * TB];
                    } else {
#pragma unroll
                        for (int j = 0; j < BW; ++j) la[j] = 0.0;
                    }
                    double rec[G][UW];
#pragma unroll
                    for (int q = 0; q < G; ++q) load_rec_s<UW>(ur + q * UW, rec[q]);
#pragma unroll
                    for (int q = G - 1; q >= 0; --q) {
                        const double xq = t[q];
#pragma unroll
                        for (int j = 1; j <= BW; ++j) {
                            if (q - j >= 0) t[q - j] = fma(-xq, rec[q][j - 1], t[q - j]);
                            else la[j - q - 1] = fma(-xq, rec[q][j - 1], la[j - q - 1]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < BW; ++j) tb[j] = la[j];
                    if (rc < keep_hi) {                      // kept rows (warp-uniform branch): store
#pragma unroll
                            for (int q = G - 1; q >= 0; --q) {
                                if (rc + q < nloc) *dst = t[q];
                                dst -= ld;
                            }
                    }
                    slot = p_slot;
                }
            }
        }
        // drain the chunks issued past the end of this column group
        while (conA < issA) { mbar_wait(barA_s + 8 * (conA % PG), (conA / PG) & 1); ++conA; }
    }
}

// ---- TMEM-resident ring --------------------------------------------------------------------------
// Same algorithm, but the per-lane ring of z' values lives in Tensor Memory instead of shared
// memory: tcgen05.st / tcgen05.ld with the 32x32b shape give every lane of a warp private access to
// "its" TMEM lane, 512 x 32-bit columns per SM-quarter, i.e. lane-private on-chip storage that does
// not compete with shared memory.  A CTA of 4 warps allocates 256 columns (128 doubles per thread =
// 128 ring rows >= M + H), so two CTAs = 8 warps are resident per SM instead of ~4.4, and a whole
// 8-row chunk moves with ONE tcgen05.ld/st (x16) instead of 8 LDS/STS.  Shared memory only holds
// the TMA landing buffers (PG chunks of RHS rows and forward records) and the backward records.
__device__ __forceinline__ void tmem_st8(unsigned taddr, const double *v) {
    unsigned r[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { r[2 * i] = (unsigned)__double2loint(v[i]); r[2 * i + 1] = (unsigned)__double2hiint(v[i]); }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, double *v) {
    unsigned r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// Arithmetic on the 8-byte lane elements of the ring.  ArF64: one double per lane.  ArF32x2: TWO
// Float32 right-hand sides per lane, packed in the 64 bits of a `double` register and updated with the
// packed fma.rn.f32x2 / mul.rn.f32x2 of sm_100 — data movement (TMA tiles of 32 x 8-byte elements, TMEM
// x16 transfers, 8-byte stores) is bit-for-bit the Float64 kernel's, each instruction now serves two
// columns.  The packed FMA has no operand negation, so the Float32 records hold the multipliers NEGATED
// (and duplicated in both halves).
struct ArF64 {
    static __device__ __forceinline__ double fnma(double z, double r, double t) { return fma(-z, r, t); }   // t - z r
    static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
};
struct ArF32x2 {
    static __device__ __forceinline__ double fnma(double z, double nr, double t) {                          // t + z (-r)
        unsigned long long o;
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(o) : "l"(__double_as_longlong(z)), "l"(__double_as_longlong(nr)),
                                               "l"(__double_as_longlong(t)));
        return __longlong_as_double((long long)o);
    }
    static __device__ __forceinline__ double mul(double a, double b) {
        unsigned long long o;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(o) : "l"(__double_as_longlong(a)), "l"(__double_as_longlong(b)));
        return __longlong_as_double((long long)o);
    }
};

// TMEM columns per CTA = 512 / OCC (OCC resident CTAs of 4 warps per SM); ring rows per lane = half of that
// (one 8-byte element = two 32-bit columns).
template <int OCC> struct TmemCfg {
    static constexpr int kCols = 512 / OCC;
    static constexpr int kRows = kCols / 2;
};

// FUSED = 1: the right-hand side is produced on the fly as  y = Lm * u  (Lm: a second banded matrix
// with at most BW sub/super-diagonals, row records `mrec`; u read through `tmap`, never modified)
// and the solution goes to `rhs` (a different buffer): the `mul!` + `ldiv!` pair of a collocation
// time step (examples/heat.jl:314-328) in one pass over HBM.  Needs 2 BW <= 8 so that the u rows
// one 8-row chunk of y depends on are exactly the chunk being eliminated and the next one.
template <int BW, int PG, int FUSED, typename AR, int OCC>
__global__ void __launch_bounds__(128, OCC)
k_banded_tmem(const double *__restrict__ lrec, const double *__restrict__ urec,
              double *__restrict__ rhs, int64_t ngroups, int64_t ld, int n, int n_pad_all, int M, int H,
              const __grid_constant__ CUtensorMap tmap, int nseg, int SEG, int HF,
              const double *__restrict__ scratch, const __grid_constant__ CUtensorMap tmap_s,
              const double *__restrict__ mrec, const double *__restrict__ uin) {
    constexpr int TB = 32;
    constexpr int LW = RecW<BW>::LW, UW = RecW<BW>::UW;
    constexpr int MW = 2 * BW + 2;                            // {Lm(r,r-BW) .. Lm(r,r+BW), pad}
    constexpr int MWS = FUSED ? PG * G * MW : 0;
    constexpr int kTmemCols = TmemCfg<OCC>::kCols;
    constexpr int RWT = TmemCfg<OCC>::kRows;                  // TMEM ring rows
    constexpr int RC = RWT + PG * G;                          // rows of backward records kept in smem
    // per-warp shared memory (doubles): stage[PG][8][32] | cur[RC][UW] | cfw[PG][8*LW] | cmv[PG][8*MW] | mbar[PG]
    constexpr int WARP_D = ((PG * G * TB + RC * UW + PG * G * LW + MWS + PG + 15) / 16) * 16;   // 128-byte multiple
    extern __shared__ __align__(128) double smem_d[];
    __shared__ unsigned s_tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tbase = s_tmem_base + ((unsigned)(warp * 32) << 16);   // this warp's 32 TMEM lanes

    double *const stage = smem_d + (size_t)warp * WARP_D;
    double *const cur = stage + PG * G * TB;
    double *const cfw = cur + RC * UW;
    double *const cmv = cfw + PG * G * LW;
    const unsigned stage_s = (unsigned)__cvta_generic_to_shared(stage);
    const unsigned cur_s = (unsigned)__cvta_generic_to_shared(cur);
    const unsigned cfw_s = (unsigned)__cvta_generic_to_shared(cfw);
    const unsigned cmv_s = (unsigned)__cvta_generic_to_shared(cmv);
    const unsigned barA_s = cmv_s + (unsigned)(MWS * 8);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < PG; ++i) mbar_init(barA_s + 8 * i, 1);
        fence_proxy_async();
    }
    __syncwarp();
    unsigned issA = 0, conA = 0;

    const int64_t nwork = ngroups * nseg;
    for (int64_t w = (int64_t)blockIdx.x * 4 + warp; w < nwork; w += (int64_t)gridDim.x * 4) {
        const int64_t g = w / nseg;
        const int seg = (int)(w - g * nseg);
        const int k0 = seg * SEG, k1 = min(n_pad_all, k0 + SEG);
        const int f0 = max(0, k0 - HF);
        const int n_pad = min(n_pad_all, k1 + H) - f0;
        const int nloc = min(n, f0 + n_pad) - f0;
        const int kl0 = k0 - f0, kl1 = k1 - f0;
        const int nblk = (n_pad + M - 1) / M;
        double *const p = rhs + (int64_t)f0 * ld + g * TB + lane;
        const double *const lrec_l = lrec + (size_t)f0 * LW;
        const double *const urec_l = urec + (size_t)f0 * UW;
        const double *const mrec_l = FUSED ? mrec + (size_t)f0 * MW : nullptr;

        int pf_row = 0, pf_cslot = 0;                         // next row to prefetch, its slot in `cur`
        auto issue_fwd = [&]() {
            const unsigned i = issA % PG;
            const unsigned bar = barA_s + 8 * i;
            __syncwarp();                                    // all lanes are done with this staging slot
            if (elect_one()) {
                if (pf_row < n_pad) {
                    fence_proxy_async();
                    const int ra = f0 + pf_row;
                    const bool halo = !FUSED && (nseg > 1) && (ra < k0 || ra >= k1);
                    const int srow = halo ? ((ra < k0 ? seg - 1 : seg) * (HF + H) + ra - ((ra < k0 ? k0 : k1) - HF)) : 0;
                    mbar_expect_tx(bar, (unsigned)(G * TB * 8 + G * (LW + UW) * 8 + (FUSED ? G * MW * 8 : 0)));
                    if (halo) tma_load_2d(stage_s + i * (G * TB * 8), &tmap_s, (int)(g * TB), srow, bar);
                    else tma_load_2d(stage_s + i * (G * TB * 8), &tmap, (int)(g * TB), ra, bar);
                    bulk_g2s(cfw_s + i * (G * LW * 8), lrec_l + (size_t)pf_row * LW, G * LW * 8, bar);
                    bulk_g2s(cur_s + (unsigned)(pf_cslot * UW * 8), urec_l + (size_t)pf_row * UW, G * UW * 8, bar);
                    // records of the y rows this chunk brings into the window: rows pf_row + BW ...
                    if (FUSED) bulk_g2s(cmv_s + i * (G * MW * 8), mrec_l + (size_t)(pf_row + BW) * MW, G * MW * 8, bar);
                } else if (FUSED && pf_row < n_pad + G) {
                    // one chunk of u past the window: the last y rows of the window depend on it
                    fence_proxy_async();
                    mbar_expect_tx(bar, (unsigned)(G * TB * 8));
                    tma_load_2d(stage_s + i * (G * TB * 8), &tmap, (int)(g * TB), f0 + pf_row, bar);
                } else {
                    mbar_arrive(bar);
                }
            }
            ++issA;
            pf_row += G;
            pf_cslot += G;
            if (pf_cslot == RC) pf_cslot = 0;
        };
#pragma unroll 1
        for (int gg = 0; gg < PG; ++gg) issue_fwd();

        int f_slot = 0;                                      // TMEM ring slot of the next forward row
        double tf[BW > 0 ? BW : 1];
        {
            mbar_wait(barA_s + 8 * (conA % PG), (conA / PG) & 1);
            const double *const ys = stage + (conA % PG) * (G * TB) + lane;
            ++conA;
            if (FUSED) {
                // y rows f0 .. f0+BW-1 also need the u rows just above the window: read them directly
                const double *const ug = uin + g * TB + lane;
#pragma unroll
                for (int j = 0; j < BW; ++j) {
                    double acc = 0.0;
#pragma unroll
                    for (int m = 0; m <= 2 * BW; ++m) {
                        const int row = f0 + j + m - BW;
                        if (row >= 0 && row < n) acc = fma(mrec_l[(size_t)j * MW + m], ug[(int64_t)row * ld], acc);
                    }
                    tf[j] = acc;
                }
            } else {
#pragma unroll
                for (int j = 0; j < BW; ++j) tf[j] = ys[j * TB];
            }
        }
        // The forward sweep runs in 8-row chunks; block bb (M rows) is finalised by a backward sweep as soon as the
        // forward sweep has passed its restart row (bb + 1) M + H.  H need not be a multiple of M: the ring holds
        // M + H rows, the backward sweeps touch every row (M + H) / M times — with M = ring - H that is 1.8 (Float64,
        // H = 56) or 1.45 (Float32, H = 40) instead of the 2.0 of M = H.
        int fpos = 0;                                        // rows forwarded so far
        for (int bb = 0; bb < nblk; ++bb) {
            {
                const int r1 = min(n_pad, (bb + 1) * M + H);
#pragma unroll 1
                for (int r0 = fpos; r0 < r1; r0 += G) {
                    const unsigned i = (conA - 1) % PG;
                    const unsigned inext = conA % PG;
                    mbar_wait(barA_s + 8 * inext, (conA / PG) & 1);
                    ++conA;
                    const double *const ys = stage + i * (G * TB) + lane;
                    const double *const yn = stage + inext * (G * TB) + lane;
                    const double *const lr = cfw + i * (G * LW);
                    double t[G + BW];
#pragma unroll
                    for (int j = 0; j < BW; ++j) t[j] = tf[j];
                    if (FUSED) {
                        // y(r0+BW+q) = sum_m Lm(r, r-BW+m) u(r0+q+m), u rows r0 .. r0+15 = this chunk + next
                        double uu[2 * G];
#pragma unroll
                        for (int q = 0; q < G; ++q) uu[q] = ys[q * TB];
#pragma unroll
                        for (int q = 0; q < G; ++q) uu[G + q] = yn[q * TB];
                        const double *const mr = cmv + i * (G * MW);
#pragma unroll
                        for (int q = 0; q < G; ++q) {
                            double mrow[MW];
                            load_rec_s<MW>(mr + q * MW, mrow);
                            double acc = 0.0;
#pragma unroll
                            for (int m = 0; m <= 2 * BW; ++m) acc = fma(mrow[m], uu[q + m], acc);
                            t[BW + q] = acc;
                        }
                    } else {
#pragma unroll
                        for (int q = BW; q < G; ++q) t[q] = ys[q * TB];
                        if (r0 + G < n_pad) {
#pragma unroll
                            for (int j = 0; j < BW; ++j) t[G + j] = yn[j * TB];
                        } else {
#pragma unroll
                            for (int j = 0; j < BW; ++j) t[G + j] = 0.0;
                        }
                    }
                    double rec[G][LW];
#pragma unroll
                    for (int q = 0; q < G; ++q) load_rec_s<LW>(lr + q * LW, rec[q]);
                    double zs[G];
#pragma unroll
                    for (int q = 0; q < G; ++q) {
                        const double z = t[q];
#pragma unroll
                        for (int j = 1; j <= BW; ++j) t[q + j] = AR::fnma(z, rec[q][j - 1], t[q + j]);
                        zs[q] = AR::mul(z, rec[q][BW]);
                    }
                    tmem_st8(tbase + (unsigned)(f_slot * 2), zs);
#pragma unroll
                    for (int j = 0; j < BW; ++j) tf[j] = t[G + j];
                    f_slot += G;
                    if (f_slot == RWT) f_slot = 0;
                    issue_fwd();
                }
                fpos = max(fpos, r1);
            }
            if (bb * M >= kl0 && bb * M < kl1) {
                const int keep_lo = bb * M;
                const int keep_hi = min(n_pad, (bb + 1) * M);
                const int r_top = min(n_pad, (bb + 1) * M + H);
                double tb[BW > 0 ? BW : 1];
                int slot = r_top % RWT;
                slot = (slot == 0 ? RWT : slot) - G;
                int cslot = r_top % RC;
                cslot = (cslot == 0 ? RC : cslot) - G;
                double *dst = p + (int64_t)(keep_hi - 1) * ld;
                tmem_wait_st();                              // forward results are in TMEM
#pragma unroll 1
                for (int rc = r_top - G; rc >= keep_lo; rc -= G) {
                    const double *const ur = cur + cslot * UW;
                    const int p_slot = (slot == 0 ? RWT : slot) - G;
                    double t[G], tl[G];
                    double la[BW > 0 ? BW : 1];
                    tmem_ld8(tbase + (unsigned)(slot * 2), t);
                    if (rc != r_top - G) {
#pragma unroll
                        for (int j = 0; j < BW; ++j) t[G - 1 - j] = tb[j];
                    }
                    if (rc > keep_lo) {
                        tmem_ld8(tbase + (unsigned)(p_slot * 2), tl);
#pragma unroll
                        for (int j = 0; j < BW; ++j) la[j] = tl[G - 1 - j];
                    } else {
#pragma unroll
                        for (int j = 0; j < BW; ++j) la[j] = 0.0;
                    }
                    double rec[G][UW];
#pragma unroll
                    for (int q = 0; q < G; ++q) load_rec_s<UW>(ur + q * UW, rec[q]);
#pragma unroll
                    for (int q = G - 1; q >= 0; --q) {
                        const double xq = t[q];
#pragma unroll
                        for (int j = 1; j <= BW; ++j) {
                            if (q - j >= 0) t[q - j] = AR::fnma(xq, rec[q][j - 1], t[q - j]);
                            else la[j - q - 1] = AR::fnma(xq, rec[q][j - 1], la[j - q - 1]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < BW; ++j) tb[j] = la[j];
                    if (rc < keep_hi) {
#pragma unroll
                        for (int q = G - 1; q >= 0; --q) {
                            if (rc + q < nloc) *dst = t[q];
                            dst -= ld;
                        }
                    }
                    slot = p_slot;
                    cslot = (cslot == 0 ? RC : cslot) - G;
                }
            }
        }
        while (conA < issA) { mbar_wait(barA_s + 8 * (conA % PG), (conA / PG) & 1); ++conA; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(s_tmem_base), "n"(kTmemCols) : "memory");
    }
}

template <int BW>
bsk_status prepare_t(bsk_banded *h, const double *w) {
    constexpr int LW = RecW<BW>::LW, UW = RecW<BW>::UW;
    const int n = (int)h->n;
    const int n_pad = ((n + G - 1) / G) * G;
    h->n_pad = n_pad;
    if (!h->d_lrec) BSK_CUDA(bsk::dev_malloc((void **)&h->d_lrec, sizeof(double) * (size_t)n_pad * LW));
    if (!h->d_urec) BSK_CUDA(bsk::dev_malloc((void **)&h->d_urec, sizeof(double) * (size_t)n_pad * UW));
    int grid = std::max(1, std::min((n_pad + 255) / 256, 148 * 4));
    k_win_repack<BW><<<grid, 256>>>(w, n, n_pad, h->d_lrec, h->d_urec);
    BSK_CUDA(cudaGetLastError());
    return BSK_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// RHS block as a 2-D tensor: inner dimension = columns (contiguous), outer = rows (ld elements apart)
bool make_rhs_tensor_map(CUtensorMap *tm, double *d_rhs, int64_t ncols, int64_t ld, int64_t nrows) {
    EncodeTiledFn enc = get_encoder();
    if (!enc || getenv("BSK_NO_TENSORMAP")) return false;
    if ((((uintptr_t)d_rhs) & 127) != 0 || ld % 2 != 0) return false;
    cuuint64_t dims[2] = {(cuuint64_t)ncols, (cuuint64_t)nrows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {32, (cuuint32_t)G};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)d_rhs, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// What the solve kernels need to know about a factorised matrix (Float64 handle or Float32 handle).
struct WinMat {
    int device;
    int n, n_pad;
    int halo, halo_f;
    const double *lrec, *urec;       // 8-byte records (Float32: two packed, negated floats per entry)
};

template <int BW, typename AR, int OCC>
bsk_status launch_tmem(const WinMat &h, double *d_rhs, int64_t ngroups, int64_t ld, cudaStream_t st, int M, int H,
                       const CUtensorMap &tmap, int nseg, int SEG, int HF, const double *scratch,
                       const CUtensorMap &tmap_s, const double *mrec, const double *d_u) {
    constexpr int PG = 4;
    constexpr int LW = RecW<BW>::LW, UW = RecW<BW>::UW;
    constexpr int RC = TmemCfg<OCC>::kRows + PG * G;
    const int64_t nwork = ngroups * nseg;
    const int64_t wslots = (int64_t)sm_count(h.device) * 4 * OCC;
    const int64_t wr = (nwork + wslots - 1) / wslots;
    const int64_t nwarps = (nwork + wr - 1) / wr;
    const int gridt = (int)std::max<int64_t>(1, (nwarps + 3) / 4);
    if (mrec != nullptr) {
        if constexpr (2 * BW <= G && OCC == 2 && std::is_same<AR, ArF64>::value) {
            constexpr int MW = 2 * BW + 2;
            const size_t smem_f = (size_t)4 * (((PG * G * 32 + RC * UW + PG * G * LW + PG * G * MW + PG + 15) / 16) * 16) * sizeof(double);
            auto kf = k_banded_tmem<BW, PG, 1, AR, OCC>;
            BSK_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
            kf<<<gridt, 128, smem_f, st>>>(h.lrec, h.urec, d_rhs, ngroups, ld, h.n, h.n_pad, M, H, tmap, nseg, SEG, HF,
                                           nullptr, tmap_s, mrec, d_u);
        } else {
            set_error("fused mat-vec + solve unavailable for this shape");
            return BSK_ERR_UNSUPPORTED;
        }
    } else {
        const size_t smem_t = (size_t)4 * (((PG * G * 32 + RC * UW + PG * G * LW + PG + 15) / 16) * 16) * sizeof(double);
        auto kt = k_banded_tmem<BW, PG, 0, AR, OCC>;
        BSK_CUDA(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
        kt<<<gridt, 128, smem_t, st>>>(h.lrec, h.urec, d_rhs, ngroups, ld, h.n, h.n_pad, M, H, tmap, nseg, SEG, HF,
                                       scratch, tmap_s, nullptr, nullptr);
    }
    return BSK_OK;
}

// mrec / d_u non-null: fused  d_rhs = (LU)^{-1} (Lm * d_u)  (d_u is only read).
// `d_rhs`, `ld`, `ngroups` are in 8-byte lane elements: doubles, or PAIRS of Float32 columns (AR = ArF32x2).
template <int BW, typename AR>
bsk_status solve_t(const WinMat &hm, double *d_rhs, int64_t ngroups, int64_t ld, cudaStream_t st,
                   const double *mrec = nullptr, const double *d_u = nullptr) {
    const WinMat *h = &hm;
    const bool fused = mrec != nullptr;
    constexpr bool kF64 = std::is_same<AR, ArF64>::value;
    constexpr int PG = 4;                                     // 4 chunks x 8 rows in flight
    constexpr int LW = RecW<BW>::LW, UW = RecW<BW>::UW;
    // halo validated at factor time (multiple of 8); the kernel restarts every M rows and needs
    // the effective halo to be a multiple of M
    int M = h->halo <= 64 ? h->halo : ((h->halo + 1) / 2 + 7) / 8 * 8;
    int per_sm_cap = 16;
    if (const char *env = getenv("BSK_WINDOW_M")) {           // tuning knobs
        int m = atoi(env);
        if (m >= 8 && m % 8 == 0) M = m;
    }
    if (const char *env = getenv("BSK_WINDOW_CTAS")) per_sm_cap = std::max(1, atoi(env));
    int H = ((h->halo + M - 1) / M) * M;
    // TMEM ring: 4 resident CTAs per SM (16 warps, 64 ring rows per lane) when the window fits, else 2 (128 rows)
    int occ = (M + H <= TmemCfg<4>::kRows && !fused) ? 4 : 2;
    if (const char *env = getenv("BSK_TMEM_OCC")) { if (atoi(env) == 2) occ = 2; }
    bool tmem_ok = M + H <= TmemCfg<2>::kRows && getenv("BSK_NO_TMEM") == nullptr && get_encoder() != nullptr;
    // The TMEM kernel takes any halo that is a multiple of 8 (k_banded_tmem: the forward position is decoupled from
    // the backward blocks): H = the validated halo itself, M = what is left of the ring.  The backward sweeps then
    // touch every row (M + H) / M times instead of twice.
    if (get_encoder() != nullptr && getenv("BSK_NO_TMEM") == nullptr && getenv("BSK_WINDOW_M") == nullptr &&
        getenv("BSK_WINDOW_EQUAL") == nullptr && h->halo % G == 0 && h->halo + G <= TmemCfg<2>::kRows) {
        H = h->halo;
        // (two CTAs per SM with the long blocks of a 128-row ring beat four CTAs with a 64-row ring also for the
        //  short halos of tridiagonal systems: C4's cyclic solve 0.527 vs 0.550 ms; BSK_TMEM_OCC=4 keeps the other)
        const char *eo = getenv("BSK_TMEM_OCC");
        occ = (!fused && eo && atoi(eo) == 4 && 2 * H <= TmemCfg<4>::kRows) ? 4 : 2;
        M = (occ == 4 ? TmemCfg<4>::kRows : TmemCfg<2>::kRows) - H;
        tmem_ok = true;
    }
    const size_t smem = ((size_t)(M + H + PG * G) * (32 + UW) + (size_t)PG * G * LW + PG) * sizeof(double);
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(per_sm_cap, (226 * 1024) / (smem + 1024)));
    const int64_t slots = (int64_t)sm_count(h->device) * per_sm;
    // Row segments: only when there are too few column groups to fill the machine and the forward
    // recurrence decays too (halo_f > 0).  Redundant work per segment: (HF + H) rows.
    int nseg = 1, SEG = h->n_pad, HF = 0;
    const int64_t fill = tmem_ok ? (int64_t)sm_count(h->device) * 4 * occ * 4 / 5     // >= 80 % of the warp slots busy
                                 : (slots * 3 + 1) / 2;
    if (h->halo_f > 0 && ngroups < fill && getenv("BSK_NO_SEGMENTS") == nullptr) {
        HF = ((h->halo_f + M - 1) / M) * M;
        int want = (int)((fill + ngroups - 1) / ngroups);
        int most = std::max(1, h->n_pad / (6 * (HF + H)));          // keep redundancy below ~1/6
        nseg = std::max(1, std::min(want, most));
        if (const char *env = getenv("BSK_WINDOW_NSEG")) nseg = std::max(1, atoi(env));
        SEG = ((h->n_pad + nseg - 1) / nseg + M - 1) / M * M;
        nseg = (h->n_pad + SEG - 1) / SEG;
        if (nseg == 1) { SEG = h->n_pad; HF = 0; }
    }
    const int64_t nwork = ngroups * nseg;
    const int64_t rounds = (nwork + slots - 1) / slots;
    const int grid = (int)std::max<int64_t>(1, (nwork + rounds - 1) / rounds);   // balanced rounds
    CUtensorMap tmap, tmap_s;
    memset(&tmap, 0, sizeof(tmap));
    memset(&tmap_s, 0, sizeof(tmap_s));
    int use_tmap = make_rhs_tensor_map(&tmap, fused ? const_cast<double *>(d_u) : d_rhs, ngroups * 32, ld, h->n) ? 1 : 0;
    double *scratch = nullptr;
    if ((fused || !kF64) && !(use_tmap && tmem_ok && (!fused || 2 * BW <= G))) {
        set_error(fused ? "fused mat-vec + solve unavailable for this shape"
                        : "single-pass Float32 solve unavailable for this shape");
        return BSK_ERR_UNSUPPORTED;
    }
    if (nseg > 1 && !fused) {
        // copy of the rows around every segment boundary (contiguous in memory: full rows)
        const int HS = HF + H;
        // freed scratch blocks stay cached in the stream-ordered pool (default: returned to the OS at
        // every synchronisation, which makes the next cudaMallocAsync cost ~1 ms)
        ensure_pool(h->device);
        BSK_CUDA(cudaMallocAsync((void **)&scratch, sizeof(double) * (size_t)(nseg - 1) * HS * ld, st));
        cudaError_t ce = cudaSuccess;
        for (int b = 0; b + 1 < nseg && ce == cudaSuccess; ++b) {
            const int64_t kb = (int64_t)(b + 1) * SEG;
            const int64_t r0 = kb - HF, r1 = std::min<int64_t>(h->n, kb + H);
            if (r1 > r0)
                ce = cudaMemcpyAsync(scratch + (size_t)b * HS * ld, d_rhs + r0 * ld, sizeof(double) * (size_t)(r1 - r0) * ld,
                                     cudaMemcpyDeviceToDevice, st);
        }
        if (ce != cudaSuccess) {                      // do not leak the scratch block on a failed copy
            cudaFreeAsync(scratch, st);
            BSK_CUDA(ce);
        }
        if (use_tmap && !make_rhs_tensor_map(&tmap_s, scratch, ngroups * 32, ld, (int64_t)(nseg - 1) * HS)) use_tmap = 0;
    }
    bsk_status rc = BSK_OK;
    if (use_tmap && tmem_ok) {
        if (occ == 4) rc = launch_tmem<BW, AR, 4>(*h, d_rhs, ngroups, ld, st, M, H, tmap, nseg, SEG, HF, scratch, tmap_s, mrec, d_u);
        else rc = launch_tmem<BW, AR, 2>(*h, d_rhs, ngroups, ld, st, M, H, tmap, nseg, SEG, HF, scratch, tmap_s, mrec, d_u);
    } else if constexpr (kF64) {
        auto kern = k_banded_windowed<BW, PG>;
        if (smem > 48 * 1024)
            rc = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess ? BSK_OK : BSK_ERR_CUDA;
        if (rc == BSK_OK)
            kern<<<grid, 32, smem, st>>>(h->lrec, h->urec, d_rhs, ngroups, ld, h->n, h->n_pad, M, H, tmap, use_tmap,
                                         nseg, SEG, HF, scratch, tmap_s);
    } else {
        set_error("single-pass Float32 solve needs the TMEM kernel");
        rc = BSK_ERR_UNSUPPORTED;
    }
    cudaError_t le = cudaGetLastError();
    if (scratch) cudaFreeAsync(scratch, st);
    if (rc != BSK_OK) return rc;
    BSK_CUDA(le);
    return BSK_OK;
}

inline WinMat winmat(const bsk_banded *h) {
    WinMat m;
    m.device = h->device; m.n = (int)h->n; m.n_pad = h->n_pad; m.halo = h->halo; m.halo_f = h->halo_f;
    m.lrec = h->d_lrec; m.urec = h->d_urec;
    return m;
}
inline WinMat winmat(const bsk_banded_f32 *h) {
    WinMat m;
    m.device = h->device; m.n = (int)h->n; m.n_pad = h->n_pad; m.halo = h->halo; m.halo_f = h->halo_f;
    m.lrec = h->d_lrec; m.urec = h->d_urec;
    return m;
}

// Float32 factors -> records of the single-pass kernel: every entry is a pair of identical floats, the
// multipliers NEGATED (see ArF32x2); same column-oriented layout as k_win_repack.
template <int BW>
__global__ void k_win_repack_f32(const float *__restrict__ w, int n, int n_pad, float2 *__restrict__ lrec,
                                 float2 *__restrict__ urec) {
    constexpr int LW = RecW<BW>::LW, UW = RecW<BW>::UW;
    const int nroww = 2 * BW + 1, middle = BW + 1;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x + 1; r <= n_pad; r += gridDim.x * blockDim.x) {
        float2 *lr = lrec + (size_t)(r - 1) * LW;
        float2 *ur = urec + (size_t)(r - 1) * UW;
        for (int j = 0; j < LW; ++j) lr[j] = make_float2(0.0f, 0.0f);
        for (int j = 0; j < UW; ++j) ur[j] = make_float2(0.0f, 0.0f);
        if (r > n) { lr[BW] = make_float2(1.0f, 1.0f); continue; }
        const float dinv = 1.0f / w[(middle - 1) + (size_t)(r - 1) * nroww];
        lr[BW] = make_float2(dinv, dinv);
        for (int j = 1; j <= BW; ++j) {
            if (r + j <= n) {
                const float v = -w[(middle + j - 1) + (size_t)(r - 1) * nroww];
                lr[j - 1] = make_float2(v, v);
            }
            if (r - j >= 1) {
                const float v = -(w[(middle - j - 1) + (size_t)(r - 1) * nroww] / w[(middle - 1) + (size_t)(r - j - 1) * nroww]);
                ur[j - 1] = make_float2(v, v);
            }
        }
    }
}

template <int BW>
bsk_status prepare_f32_t(bsk_banded_f32 *h) {
    constexpr int LW = RecW<BW>::LW, UW = RecW<BW>::UW;
    const int n = (int)h->n;
    const int n_pad = ((n + G - 1) / G) * G;
    h->n_pad = n_pad;
    BSK_CUDA(bsk::dev_malloc((void **)&h->d_lrec, sizeof(double) * (size_t)n_pad * LW));
    BSK_CUDA(bsk::dev_malloc((void **)&h->d_urec, sizeof(double) * (size_t)n_pad * UW));
    int grid = std::max(1, std::min((n_pad + 255) / 256, 148 * 4));
    k_win_repack_f32<BW><<<grid, 256>>>(h->d_w, n, n_pad, (float2 *)h->d_lrec, (float2 *)h->d_urec);
    BSK_CUDA(cudaGetLastError());
    return BSK_OK;
}

}  // namespace

namespace bsk {

// Row records of a second banded matrix Lm for the fused kernel: (n_pad + G) rows of
// { Lm(r, r-bw) .. Lm(r, r+bw), pad }, zero outside Lm's own band, outside the matrix and for r > n.
// Band store (Collocation/Collocation.jl:165-198): element (i, j) at w[(u + i - j) + (j-1) (l+u+1)].
bsk_status windowed_matrec(const bsk_banded *Lm, int bw, int n_pad, double **out) {
    const int n = (int)Lm->n, lm = Lm->l, um = Lm->u, MW = 2 * bw + 2;
    const size_t rows = (size_t)n_pad + G;
    std::vector<double> w((size_t)(lm + um + 1) * n), rec(rows * MW, 0.0);
    BSK_CUDA(cudaSetDevice(Lm->device));
    BSK_CUDA(cudaMemcpy(w.data(), Lm->d_w, w.size() * 8, cudaMemcpyDeviceToHost));
    for (int i = 1; i <= n; ++i)
        for (int m = 0; m <= 2 * bw; ++m) {
            const int j = i - bw + m;
            if (j < 1 || j > n || i - j > lm || j - i > um) continue;
            rec[(size_t)(i - 1) * MW + m] = w[(size_t)(um + i - j) + (size_t)(j - 1) * (lm + um + 1)];
        }
    double *d = nullptr;
    BSK_CUDA(bsk::dev_malloc((void **)&d, rec.size() * 8));
    BSK_CUDA(cudaMemcpy(d, rec.data(), rec.size() * 8, cudaMemcpyHostToDevice));
    *out = d;
    return BSK_OK;
}

bsk_status windowed_solve_fused(const bsk_banded *h, const double *mrec, const double *d_u, double *d_out,
                                int64_t ngroups, int64_t ld, cudaStream_t st) {
    switch (h->l) {
        case 0: return solve_t<0, ArF64>(winmat(h), d_out, ngroups, ld, st, mrec, d_u);
        case 1: return solve_t<1, ArF64>(winmat(h), d_out, ngroups, ld, st, mrec, d_u);
        case 2: return solve_t<2, ArF64>(winmat(h), d_out, ngroups, ld, st, mrec, d_u);
        case 3: return solve_t<3, ArF64>(winmat(h), d_out, ngroups, ld, st, mrec, d_u);
        case 4: return solve_t<4, ArF64>(winmat(h), d_out, ngroups, ld, st, mrec, d_u);
    }
    set_error("fused mat-vec + solve needs at most 4 sub-diagonals");
    return BSK_ERR_UNSUPPORTED;
}

// `w`: the factors to repack (h->d_w, or the second band buffer of the partitioned factorisation)
bsk_status windowed_prepare(bsk_banded *h, const double *w) {
    switch (h->l) {
        case 0: return prepare_t<0>(h, w);
        case 1: return prepare_t<1>(h, w);
        case 2: return prepare_t<2>(h, w);
        case 3: return prepare_t<3>(h, w);
        case 4: return prepare_t<4>(h, w);
        case 5: return prepare_t<5>(h, w);
        case 6: return prepare_t<6>(h, w);
    }
    return BSK_ERR_UNSUPPORTED;
}

// Solves the first 32 * ngroups columns of an interleaved RHS block whose rows are `ld`
// elements apart.  Requires 16-byte aligned rows (d_rhs 16-byte aligned, ld even).
bsk_status windowed_solve(const bsk_banded *h, double *d_rhs, int64_t ngroups, int64_t ld, cudaStream_t st) {
    switch (h->l) {
        case 0: return solve_t<0, ArF64>(winmat(h), d_rhs, ngroups, ld, st);
        case 1: return solve_t<1, ArF64>(winmat(h), d_rhs, ngroups, ld, st);
        case 2: return solve_t<2, ArF64>(winmat(h), d_rhs, ngroups, ld, st);
        case 3: return solve_t<3, ArF64>(winmat(h), d_rhs, ngroups, ld, st);
        case 4: return solve_t<4, ArF64>(winmat(h), d_rhs, ngroups, ld, st);
        case 5: return solve_t<5, ArF64>(winmat(h), d_rhs, ngroups, ld, st);
        case 6: return solve_t<6, ArF64>(winmat(h), d_rhs, ngroups, ld, st);
    }
    return BSK_ERR_UNSUPPORTED;
}

// ---- Float32: two right-hand sides per lane -------------------------------------------------------
bsk_status windowed_prepare_f32(bsk_banded_f32 *h) {
    switch (h->l) {
        case 0: return prepare_f32_t<0>(h);
        case 1: return prepare_f32_t<1>(h);
        case 2: return prepare_f32_t<2>(h);
        case 3: return prepare_f32_t<3>(h);
        case 4: return prepare_f32_t<4>(h);
        case 5: return prepare_f32_t<5>(h);
        case 6: return prepare_f32_t<6>(h);
    }
    return BSK_ERR_UNSUPPORTED;
}

// Solves the first 64 * ngroups columns of an interleaved Float32 RHS block whose rows are `ld_pairs`
// column PAIRS apart (nrhs even, block 128-byte aligned).
bsk_status windowed_solve_f32(const bsk_banded_f32 *h, float *d_rhs, int64_t ngroups, int64_t ld_pairs, cudaStream_t st) {
    double *p = (double *)d_rhs;
    switch (h->l) {
        case 0: return solve_t<0, ArF32x2>(winmat(h), p, ngroups, ld_pairs, st);
        case 1: return solve_t<1, ArF32x2>(winmat(h), p, ngroups, ld_pairs, st);
        case 2: return solve_t<2, ArF32x2>(winmat(h), p, ngroups, ld_pairs, st);
        case 3: return solve_t<3, ArF32x2>(winmat(h), p, ngroups, ld_pairs, st);
        case 4: return solve_t<4, ArF32x2>(winmat(h), p, ngroups, ld_pairs, st);
        case 5: return solve_t<5, ArF32x2>(winmat(h), p, ngroups, ld_pairs, st);
        case 6: return solve_t<6, ArF32x2>(winmat(h), p, ngroups, ld_pairs, st);
    }
    return BSK_ERR_UNSUPPORTED;
}

}  // namespace bsk
