// bsk_probe.cu — measurement aid for the global-table evaluation kernels (no reference counterpart).
//
// k_probe_gather has the memory behaviour of k_spline_cell<4, *, double, 1, 0, 0> (bsk_eval_cell.cu) and
// nothing else: n doubles streamed in, for each one record of 32 bytes gathered from a table that lives in
// L2 (index = (int)(x * nrec), one 256-bit load = one sector), a cubic Horner, n doubles streamed out.
// It states what the memory system gives that access pattern — the bound of the C4 / C5 evaluations —
// and lets the variants of the point stream (LDG.128 with software prefetch, TMA rings of several shapes,
// cache hints) be compared in seconds instead of one five-minute build of the product kernel each.
#include <algorithm>

#include "handles.cuh"
#include "tma.cuh"

using namespace bsk;

namespace {

constexpr int kPT = 1024;

template <int HINT>
__device__ __forceinline__ void gather256(const void *p, int4 &lo, int4 &hi) {
    if (HINT == 1)
        asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                     : "l"(p));
    else
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                     : "l"(p));
}

// IT, DEPTH: per-warp TMA ring of DEPTH tiles of IT x 32 sixteen-byte vectors (DEPTH = 0: LDG.128 + prefetch).
// UNR: vectors of one lane whose gathers are in flight together.
template <int IT, int DEPTH, int HINT, int UNR>
__global__ void __launch_bounds__(kPT, 1)
k_probe_gather(const double *__restrict__ x, double *__restrict__ out, int64_t nvec, const double *__restrict__ table,
               int nrec) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int4 *x4 = (const int4 *)x;
    const double scale = (double)nrec;
    const int nrec1 = nrec - 1;
    auto gather = [&](const int4 raw, int4 (&lo)[2], int4 (&hi)[2], double (&u)[2]) {
        double xv[2];
        memcpy(xv, &raw, 16);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double s = xv[q] * scale;
            int c = (int)s;
            c = c < 0 ? 0 : (c > nrec1 ? nrec1 : c);
            u[q] = s - (double)c;
            gather256<HINT>(table + 4 * (size_t)c, lo[q], hi[q]);
        }
    };
    auto finish = [&](const int4 (&lo)[2], const int4 (&hi)[2], const double (&u)[2], const int64_t v) {
        double r[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            double c[4];
            memcpy(&c[0], &lo[q], 16);
            memcpy(&c[2], &hi[q], 16);
            r[q] = fma(fma(fma(c[3], u[q], c[2]), u[q], c[1]), u[q], c[0]);
        }
        int4 o;
        memcpy(&o, r, 16);
        ((int4 *)out)[v] = o;
    };
    if constexpr (DEPTH > 0) {
        constexpr int TILEV = 32 * IT;
        constexpr int NW = kPT / 32;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        unsigned char *const ring = smem + (size_t)warp * DEPTH * (TILEV * 16);
        const unsigned bars = (unsigned)__cvta_generic_to_shared(smem + (size_t)NW * DEPTH * (TILEV * 16)) + (unsigned)warp * DEPTH * 8u;
        const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
        const int64_t ntile = nvec / TILEV;
        const int64_t GW = (int64_t)gridDim.x * NW;
        const int64_t gw = (int64_t)blockIdx.x * NW + warp;
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) mbar_init(bars + 8u * d, 1);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
            fence_proxy_async();
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) {
                const int64_t Tn = gw + d * GW;
                if (Tn < ntile) {
                    mbar_expect_tx(bars + 8u * d, TILEV * 16);
                    bulk_g2s(ring_s + (unsigned)d * (TILEV * 16), x4 + Tn * TILEV, TILEV * 16, bars + 8u * d);
                }
            }
        }
        __syncwarp();
        int slot = 0;
        unsigned parity = 0;
        for (int64_t Tc = gw; Tc < ntile; Tc += GW) {
            mbar_wait(bars + 8u * slot, parity);
            const int4 *tile = (const int4 *)(ring + (size_t)slot * (TILEV * 16));
            int4 raws[IT];
#pragma unroll
            for (int it = 0; it < IT; ++it) raws[it] = tile[it * 32 + lane];
            __syncwarp();
            const int64_t Tn = Tc + DEPTH * GW;
            if (lane == 0 && Tn < ntile) {
                // The LDS above has been ISSUED, not performed: behind a queue of divergent gathers it can sit in
                // the load/store pipe for longer than a DRAM round trip, and the bulk copy (async proxy) is not
                // ordered behind it — without this fence the refill overwrote half-read tiles (measured: 7 % of
                // the points wrong with 512-byte tiles).  The issuing lane took part in the warp-wide LDS, so
                // fence.proxy.async here waits for that instruction before the copy may start.
                fence_proxy_async();
                mbar_expect_tx(bars + 8u * slot, TILEV * 16);
                bulk_g2s(ring_s + (unsigned)slot * (TILEV * 16), x4 + Tn * TILEV, TILEV * 16, bars + 8u * slot);
            }
#pragma unroll
            for (int it0 = 0; it0 < IT; it0 += UNR) {
                int4 lo[UNR][2], hi[UNR][2];
                double u[UNR][2];
#pragma unroll
                for (int j = 0; j < UNR; ++j) gather(raws[it0 + j], lo[j], hi[j], u[j]);
#pragma unroll
                for (int j = 0; j < UNR; ++j) finish(lo[j], hi[j], u[j], Tc * TILEV + (it0 + j) * 32 + lane);
            }
            if (++slot == DEPTH) { slot = 0; parity ^= 1u; }
        }
        const int64_t v = ntile * TILEV + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (v < nvec) {
            int4 lo[2], hi[2];
            double u[2];
            gather(x4[v], lo, hi, u);
            finish(lo, hi, u, v);
        }
    } else {
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        int4 nxt[UNR];
#pragma unroll
        for (int j = 0; j < UNR; ++j)
            if (v + j * stride < nvec) nxt[j] = x4[v + j * stride];
        for (; v < nvec; v += UNR * stride) {
            int4 raw[UNR];
#pragma unroll
            for (int j = 0; j < UNR; ++j) raw[j] = nxt[j];
#pragma unroll
            for (int j = 0; j < UNR; ++j)
                if (v + (UNR + j) * stride < nvec) nxt[j] = x4[v + (UNR + j) * stride];
            int4 lo[UNR][2], hi[UNR][2];
            double u[UNR][2];
#pragma unroll
            for (int j = 0; j < UNR; ++j)
                if (v + j * stride < nvec) gather(raw[j], lo[j], hi[j], u[j]);
#pragma unroll
            for (int j = 0; j < UNR; ++j)
                if (v + j * stride < nvec) finish(lo[j], hi[j], u[j], v + j * stride);
        }
    }
}

template <int IT, int DEPTH, int HINT, int UNR>
cudaError_t launch(const double *x, double *out, int64_t nvec, const double *table, int nrec, int sms, int carve,
                   cudaStream_t st) {
    auto kern = k_probe_gather<IT, DEPTH, HINT, UNR>;
    const size_t smem = DEPTH > 0 ? (size_t)(kPT / 32) * DEPTH * (32 * IT * 16 + 8) : 0;
    if (smem > 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (carve >= 0) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    kern<<<sms, kPT, smem, st>>>(x, out, nvec, table, nrec);
    return cudaGetLastError();
}

}  // namespace

extern "C" bsk_status bsk_probe_gather(const double *d_x, double *d_out, int64_t n, const double *d_table, int64_t nrec,
                                       int32_t variant, void *stream) {
    BSK_REQUIRE(d_x && d_out && d_table, BSK_ERR_ARGUMENT, "NULL argument");
    BSK_REQUIRE(n >= 0 && n % 2 == 0 && nrec > 0 && nrec < (1ll << 31), BSK_ERR_ARGUMENT, "n must be even, 0 < nrec < 2^31");
    BSK_REQUIRE((((uintptr_t)d_x | (uintptr_t)d_out) & 15) == 0 && ((uintptr_t)d_table & 31) == 0, BSK_ERR_ARGUMENT,
                "points / results must be 16-byte aligned, the table 32-byte aligned");
    if (n == 0) return BSK_OK;
    int dev = 0;
    BSK_CUDA(cudaGetDevice(&dev));
    const int sms = sm_count(dev);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    switch (variant) {
        case 0: e = launch<1, 0, 0, 1>(d_x, d_out, n / 2, d_table, (int)nrec, sms, -1, st); break;   // LDG.128 points + prefetch, 256-bit gathers
        case 1: e = launch<2, 2, 0, 1>(d_x, d_out, n / 2, d_table, (int)nrec, sms, -1, st); break;   // points by per-warp TMA rings (64 KB)
        case 2: e = launch<1, 0, 1, 1>(d_x, d_out, n / 2, d_table, (int)nrec, sms, -1, st); break;   // as 0, gathers L1::no_allocate (the product kernel)
        case 3: e = launch<2, 2, 1, 1>(d_x, d_out, n / 2, d_table, (int)nrec, sms, -1, st); break;   // as 1, gathers L1::no_allocate
        case 4: e = launch<1, 0, 0, 2>(d_x, d_out, n / 2, d_table, (int)nrec, sms, -1, st); break;   // as 0, four gathers per lane in flight
        default: set_error("unknown variant %d", variant); return BSK_ERR_ARGUMENT;
    }
    if (e != cudaSuccess) { set_error("probe launch failed: %s", cudaGetErrorString(e)); return BSK_ERR_CUDA; }
    return BSK_OK;
}
