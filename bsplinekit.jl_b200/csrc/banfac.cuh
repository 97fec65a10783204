// banfac.cuh — K5: banded LU without pivoting (BANFAC, src/Collocation/matrix.jl:132-201), shared by the
// Float64 (bsk_solve.cu) and Float32 (bsk_solve_f32.cu) paths.  Same statements as the reference in the
// element type T, non-contracting arithmetic: bit-identical factors.
#pragma once
#include "handles.cuh"

namespace bsk {

// ---- K5: BANFAC (matrix.jl:132-201), a strictly serial recurrence: one thread -----------
template <typename T>
__global__ void k_banded_lu(T *__restrict__ w, int nbandl, int nbandu, int nrow, long long *info) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int nroww = nbandl + nbandu + 1;
    const int middle = nbandu + 1;
#define W(r, c) w[((r) - 1) + (size_t)((c) - 1) * nroww]
    *info = 0;
    if (nrow == 1) { if (W(middle, 1) == (T)0) *info = 1; return; }
    if (nbandl == 0) {
        for (int i = 1; i <= nrow; ++i)
            if (W(middle, i) == (T)0) { *info = i; return; }
        return;
    }
    for (int i = 1; i <= nrow - 1; ++i) {
        const T pivot = W(middle, i);
        if (pivot == (T)0) { *info = i; return; }
        const T ipiv = xdiv((T)1, pivot);
        const int jmax = min(nbandl, nrow - i);
        for (int j = 1; j <= jmax; ++j) W(middle + j, i) = xmul(W(middle + j, i), ipiv);
        const int kmax = min(nbandu, nrow - i);
        for (int kk = 1; kk <= kmax; ++kk) {
            const T factor = W(middle - kk, i + kk);
            for (int j = 1; j <= jmax; ++j)
                W(middle - kk + j, i + kk) = xsub(W(middle - kk + j, i + kk), xmul(factor, W(middle + j, i)));
        }
    }
    if (W(middle, nrow) == (T)0) *info = nrow;
#undef W
}

// Same recurrence, same operations in the same order (bit-identical factors), but the band streams
// through shared memory: all threads load / store chunks of columns with coalesced accesses while
// warp 0 eliminates from a window of kLuChunk + nbandu columns (one band row per lane), so the serial
// chain runs at register / shuffle latency instead of DRAM latency.
constexpr int kLuChunk = 256;

template <int BW, typename T>
__global__ void __launch_bounds__(128)
k_banded_lu_stream(T *__restrict__ w, int nrow, long long *info) {
    extern __shared__ __align__(16) unsigned char banfac_smem[];
    T *const sw = (T *)banfac_smem;                         // (kLuChunk + nbandu) columns x nroww
    __shared__ long long s_info;
    constexpr int nbandl = BW, nbandu = BW;
    constexpr int nroww = nbandl + nbandu + 1;
    constexpr int middle = nbandu + 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) s_info = 0;
    // initial window: columns 1 .. min(nrow, kLuChunk + nbandu)
    {
        const int ncol = min(nrow, kLuChunk + nbandu);
        for (int e = tid; e < ncol * nroww; e += nt) sw[e] = w[e];
    }
    __syncthreads();
#define SW(r, idx) sw[((r) - 1) + (size_t)(idx) * nroww]
    for (int c0 = 0; c0 < nrow; c0 += kLuChunk) {
        const int ndo = min(kLuChunk, nrow - c0);           // columns eliminated in this round
        if (tid < 32) {
            // Warp 0 eliminates.  Lane r (< nroww) owns band row r of a register window of BW + 1 columns:
            // column idx and the BW columns it updates.  The bw x bw rank-one update of a column step is
            // spread over the lanes (pivot, multipliers and factors travel by shuffle), so the serial
            // chain per column is  shuffle -> 1/pivot -> multiply -> shuffle -> multiply, subtract
            // instead of ~bw^2 dependent operations in one thread.  Same operations on the same values
            // as BANFAC: bit-identical factors.
            const int r = tid;
            const unsigned full = 0xffffffffu;
            T win[BW + 1];
#pragma unroll
            for (int c = 0; c <= BW; ++c) win[c] = (r < nroww && c0 + c < nrow) ? sw[r + (size_t)c * nroww] : (T)0;
            bool stop = false;
            int idx = 0;
            for (; idx < ndo; ++idx) {
                const int i = c0 + idx + 1;                 // 1-based column
                if (i <= nrow - 1) {
                    const T pivot = __shfl_sync(full, win[0], middle - 1);
                    if (pivot == (T)0) { if (r == 0) s_info = i; stop = true; break; }
                    const T ipiv = xdiv((T)1, pivot);
                    const int jmax = min(nbandl, nrow - i);
                    if (r >= middle && r - (middle - 1) <= jmax) win[0] = xmul(win[0], ipiv);
#pragma unroll
                    for (int kk = 1; kk <= BW; ++kk) {
                        // column i + kk: W(middle - kk + j, i + kk) -= W(middle - kk, i + kk) * W(middle + j, i)
                        const T factor = __shfl_sync(full, win[kk], middle - kk - 1);
                        const T lj = __shfl_down_sync(full, win[0], kk);
                        const int j = r - (middle - kk - 1);
                        if (kk <= jmax && j >= 1 && j <= jmax) win[kk] = xsub(win[kk], xmul(factor, lj));
                    }
                }
                // retire column idx, shift the window, bring in column idx + BW + 1
                if (r < nroww) sw[r + (size_t)idx * nroww] = win[0];
#pragma unroll
                for (int c = 0; c < BW; ++c) win[c] = win[c + 1];
                const bool have = (idx + BW + 1 < kLuChunk + BW) && (c0 + idx + BW + 1 < nrow);
                win[BW] = (have && r < nroww) ? sw[r + (size_t)(idx + BW + 1) * nroww] : (T)0;
            }
            if (!stop) {
                // the BW columns after the last eliminated one carry partial updates: back to the window
#pragma unroll
                for (int c = 0; c < BW; ++c)
                    if (r < nroww && c0 + ndo + c < nrow && ndo + c < kLuChunk + BW)
                        sw[r + (size_t)(ndo + c) * nroww] = win[c];
                __syncwarp();
                if (r == 0 && c0 + ndo >= nrow && SW(middle, nrow - 1 - c0) == (T)0) s_info = nrow;
            }
        }
        __syncthreads();
        // write the finished columns back
        for (int e = tid; e < ndo * nroww; e += nt) w[(size_t)c0 * nroww + e] = sw[e];
        if (s_info != 0) break;
        __syncthreads();
        // carry the nbandu partially updated columns over, then load the next ones
        const int carry = min(nbandu, nrow - (c0 + ndo));
        if (carry <= 0) break;
        for (int e = tid; e < carry * nroww; e += nt) {
            const T v = sw[(size_t)kLuChunk * nroww + e];
            sw[e] = v;                                      // kLuChunk >= nbandu: source and target do not overlap
        }
        __syncthreads();                                    // the loads below overwrite the carried columns' old slots
        const int first = c0 + kLuChunk + nbandu;           // 0-based global column of window slot `nbandu`
        const int nnew = min(kLuChunk, nrow - first);
        for (int e = tid; e < nnew * nroww; e += nt) sw[(size_t)nbandu * nroww + e] = w[(size_t)first * nroww + e];
        __syncthreads();
    }
#undef SW
    if (tid == 0) *info = s_info;
}


// ---- partitioned BANFAC ---------------------------------------------------------------------------
// The recurrence is serial in the column index, but it forgets its start state: eliminating from column
// s - H as if the matrix began there reproduces the factors of the columns >= s up to an error that decays
// with H like the Schur-complement recurrence contracts (diagonally dominant collocation matrices: about
// one bit per column).  Partition p = columns [pL, (p+1)L) is eliminated by one CTA from the ORIGINAL matrix
// starting H columns early, and continues V = BW + 1 columns into the next partition; those V columns carry
// the complete elimination state at the boundary and are compared afterwards with what the next partition
// (which reached them after only H warm-up columns) produced (k_banfac_verify).  Agreement to `tol` relative
// at every boundary certifies the whole factorisation (partition 0 is exact, every later start state is then
// within tol of the true one, and its error keeps decaying); any disagreement, non-finite value or zero
// pivot sends the caller back to the serial kernel, which is also what reports ZeroPivotException(i).
// Out of place (w_in is read by neighbouring partitions while w_out is written).  Factors agree with
// BANFAC's to ~1e-16 relative, not bit for bit: the serial kernels stay the bit-exact option.
template <int BW, typename T>
__global__ void __launch_bounds__(128)
k_banfac_part(const T *__restrict__ w_in, T *__restrict__ w_out, T *__restrict__ ver, int nrow, int L, int H,
              int *__restrict__ fail) {
    extern __shared__ __align__(16) unsigned char banfac_smem[];
    T *const sw = (T *)banfac_smem;
    constexpr int nroww = 2 * BW + 1;
    constexpr int middle = BW + 1;
    constexpr int V = BW + 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int p = blockIdx.x;
    const int s0 = p * L, e0 = min(nrow, s0 + L);
    const int start = max(0, s0 - H);
    const int stop = min(nrow, e0 + V);                      // == nrow for the last partition
    const int ncol = stop - start;
    for (int e = tid; e < ncol * nroww; e += nt) sw[e] = w_in[(size_t)start * nroww + e];
    __syncthreads();
    if (tid < 32) {
        const int r = tid;
        const unsigned full = 0xffffffffu;
        T win[BW + 1];
#pragma unroll
        for (int c = 0; c <= BW; ++c) win[c] = (r < nroww && c < ncol) ? sw[r + (size_t)c * nroww] : (T)0;
        bool bad = false;
        for (int idx = 0; idx < ncol; ++idx) {
            const int i = start + idx + 1;                  // 1-based global column
            const T pivot = __shfl_sync(full, win[0], middle - 1);
            if (!(pivot != (T)0) || !(fabs((double)pivot) < 1e300)) { bad = true; break; }   // zero / NaN / Inf pivot
            if (i <= nrow - 1) {
                const T ipiv = xdiv((T)1, pivot);
                const int jmax = min(BW, nrow - i);
                if (r >= middle && r - (middle - 1) <= jmax) win[0] = xmul(win[0], ipiv);
#pragma unroll
                for (int kk = 1; kk <= BW; ++kk) {
                    const T factor = __shfl_sync(full, win[kk], middle - kk - 1);
                    const T lj = __shfl_down_sync(full, win[0], kk);
                    const int j = r - (middle - kk - 1);
                    if (kk <= jmax && j >= 1 && j <= jmax) win[kk] = xsub(win[kk], xmul(factor, lj));
                }
            }
            if (r < nroww) sw[r + (size_t)idx * nroww] = win[0];
#pragma unroll
            for (int c = 0; c < BW; ++c) win[c] = win[c + 1];
            win[BW] = (idx + BW + 1 < ncol && r < nroww) ? sw[r + (size_t)(idx + BW + 1) * nroww] : (T)0;
        }
        if (bad && r == 0) atomicExch(fail, 1);
    }
    __syncthreads();
    const int own0 = s0 - start, nown = e0 - s0;
    for (int e = tid; e < nown * nroww; e += nt) w_out[(size_t)s0 * nroww + e] = sw[(size_t)own0 * nroww + e];
    const int nver = stop - e0;                              // columns continued into the next partition
    if (e0 < nrow)
        for (int e = tid; e < nver * nroww; e += nt) ver[(size_t)p * V * nroww + e] = sw[(size_t)(own0 + nown) * nroww + e];
}

// Boundary check of the partitioned factorisation: the V columns partition p carried past its end against
// the same columns as the next partition produced them.
template <int BW, typename T>
__global__ void __launch_bounds__(64)
k_banfac_verify(const T *__restrict__ w_out, const T *__restrict__ ver, int nrow, int L, int nparts, double tol,
                int *__restrict__ fail) {
    constexpr int nroww = 2 * BW + 1;
    constexpr int V = BW + 1;
    const int p = blockIdx.x;
    if (p + 1 >= nparts) return;
    const int e0 = min(nrow, (p + 1) * L);
    const int nver = min(nrow, e0 + V) - e0;
    bool bad = false;
    for (int c = threadIdx.x; c < nver; c += blockDim.x) {
        double cmax = 0.0;
        for (int r = 0; r < nroww; ++r) cmax = fmax(cmax, fabs((double)w_out[(size_t)(e0 + c) * nroww + r]));
        for (int r = 0; r < nroww; ++r) {
            const double a = (double)w_out[(size_t)(e0 + c) * nroww + r], b = (double)ver[(size_t)p * V * nroww + c * nroww + r];
            if (!(fabs(a - b) <= tol * cmax)) bad = true;    // also catches NaN
        }
    }
    if (bad) atomicExch(fail, 1);
}

// Partitioned factorisation of w_in into w_out (both (2 BW + 1) x n); *d_fail (device int, zeroed here) != 0
// afterwards means "not certified: use the serial kernel".  `d_ver` needs nparts * (BW + 1) * (2 BW + 1) elements.
template <typename T>
inline cudaError_t launch_banfac_part(const T *w_in, T *w_out, T *d_ver, int bw, int n, int L, int H, double tol,
                                      int *d_fail, cudaStream_t st) {
    const int nparts = (n + L - 1) / L;
    cudaError_t e = cudaMemsetAsync(d_fail, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    const size_t smem = sizeof(T) * (size_t)(L + H + bw + 1) * (2 * bw + 1);
#define BSK_BANFAC_PART(BWV)                                                                                              \
    case BWV: {                                                                                                           \
        auto kern = k_banfac_part<BWV, T>;                                                                                \
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        if (e == cudaSuccess) {                                                                                           \
            kern<<<nparts, 128, smem, st>>>(w_in, w_out, d_ver, n, L, H, d_fail);                                         \
            k_banfac_verify<BWV, T><<<nparts, 64, 0, st>>>(w_out, d_ver, n, L, nparts, tol, d_fail);                      \
        }                                                                                                                 \
    } break;
    switch (bw) {
        BSK_BANFAC_PART(1) BSK_BANFAC_PART(2) BSK_BANFAC_PART(3) BSK_BANFAC_PART(4) BSK_BANFAC_PART(5) BSK_BANFAC_PART(6)
        default: return cudaErrorInvalidValue;
    }
#undef BSK_BANFAC_PART
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}
inline int banfac_part_count(int n, int L) { return (n + L - 1) / L; }

// lu!(C) on the device: picks the streamed kernel (any n, l = u = 1..6) or the plain serial one.
template <typename T>
inline cudaError_t launch_banfac(T *d_w, int l, int u, int n, long long *d_info, bool serial) {
    if (n == 1 || l == 0 || serial) {
        k_banded_lu<T><<<1, 1>>>(d_w, l, u, n, d_info);
    } else {
        const size_t smem = sizeof(T) * (size_t)(kLuChunk + u) * (l + u + 1);
        switch (l) {
            case 1: k_banded_lu_stream<1, T><<<1, 128, smem>>>(d_w, n, d_info); break;
            case 2: k_banded_lu_stream<2, T><<<1, 128, smem>>>(d_w, n, d_info); break;
            case 3: k_banded_lu_stream<3, T><<<1, 128, smem>>>(d_w, n, d_info); break;
            case 4: k_banded_lu_stream<4, T><<<1, 128, smem>>>(d_w, n, d_info); break;
            case 5: k_banded_lu_stream<5, T><<<1, 128, smem>>>(d_w, n, d_info); break;
            default: k_banded_lu_stream<6, T><<<1, 128, smem>>>(d_w, n, d_info); break;
        }
    }
    return cudaGetLastError();
}

}  // namespace bsk
