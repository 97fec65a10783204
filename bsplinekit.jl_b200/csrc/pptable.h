// pptable.h — HOST-side construction of the local-polynomial evaluation tables of a spline
// (per-knot-interval records for K3b, uniform-cell records for K3c; bsk_eval.cu uploads them,
// pp.cuh / bsk_eval_cell.cu read them).  Plain C++: no CUDA calls, so that the construction and its
// accuracy certificate are unit-tested on the CPU (tests/c_driver/pptable_emul.cpp).
//
// What the tables replace: the knot search + de Boor triangle of (S::Spline)(x)
// (src/Splines/spline.jl:144-209) and of (Derivative(n) * S)(x) (spline.jl:272-330).
//
// Accuracy.  Every coefficient is formed in double-double arithmetic (dd.h) from the polynomial
// piece itself (a polynomial-valued de Boor recurrence, i.e. the reference's recurrence carried out
// on polynomials in u = x - z) and rounded ONCE to the element type, so a table evaluation differs
// from the exact spline only by the rounding of its own Horner recurrence.  That rounding is bounded
// a priori for every cell and every derivative order d = 0 .. k-1,
//     E_d = u_T * sum_m (2 (m - d) + 3) |c_m| m!/(m-d)! U^(m-d)        (U = largest |x - centre|),
// plus the terms of the knot-cell encodings, and a cell (or interval record) is CERTIFIED only if
// E_d <= kCertBudget * tol_T * scale_d for all d, with tol_T the north-star bar (1e-13 / 1e-5) and
// scale_d the magnitude of D^d S over the domain.  Cells that are not certified are routed to the
// interval table, interval records that are not certified carry a flag and the kernels evaluate
// those points with the bit-faithful de Boor recurrence instead.
#pragma once
#include <atomic>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "core.cuh"
#include "dd.h"

namespace bsk {

constexpr double kCertBudget = 0.25;     // share of the tolerance granted to the table's own rounding

template <typename T> struct TableTol;
template <> struct TableTol<double> { static constexpr double tol = 1e-13, u = 1.1102230246251565e-16, vmax = 1.0e307; };
template <> struct TableTol<float> { static constexpr double tol = 1e-5, u = 5.9604644775390625e-8, vmax = 3.0e38; };

struct TableInput {
    int kind;                            // 0 plain, 1 periodic
    int k;
    int64_t N;                           // number of basis functions of the evaluation basis
    const std::vector<double> *knots;    // plain: N + k knots; periodic: N + 1 breakpoints
    const std::vector<double> *coefs;    // coefficients in the evaluation basis
    bool allow_global_cell;              // tables too large for shared memory may stay in global memory
    int global_cells_x2;                 // cells per interval (x2) of a global-memory table
    int64_t global_max_rec = 400000;     // most knot intervals a global-memory cell table is built for
    int64_t global_cap_bytes = 32ll << 20;   // its size cap (the cell count shrinks to fit, down to 3 per interval)
};

template <typename T>
struct EvalTables {
    // ---- per-interval records { t_lo, mid | flag, c_0 .. c_{k-1}, pad }: S(x) = sum c_m (x - mid)^m ----
    int rec = 0;
    int64_t nrec = 0;
    std::vector<T> recs;
    std::vector<uint32_t> lut;
    int lut_ncell = 0;
    double lut_scale = 0;
    int64_t n_uncertified = 0;           // records flagged for the de Boor route
    // largest derivative order d for which every term c_m m!/(m-d)! of every record / cell stays inside the range of
    // T (Float32: on very short intervals the high coefficients ~ h^-m do not); calls that ask for more take the
    // de Boor route, which forms D^d S in the reference's own way
    int range_safe_d = BSK_MAXK;
    // ---- uniform-cell table --------------------------------------------------------------------
    bool cell_valid = false, cell_global = false, cell_jump = false;
    std::vector<T> cell_rec, cell_rc;    // device layout
    int ncell = 0, nside = 0;
    double x0 = 0, b = 0, scale = 0, cw = 0;
    size_t smem = 0;
    int64_t n_cell_slow = 0;             // cells routed to the interval table
    int64_t n_cell_flagged = 0;          // cells with one knot (side entry)
    double scale_d[BSK_MAXK] = {0};      // magnitude of D^d S used by the certificate
};

// Derivative coefficients, _derivative (Splines/spline.jl:293-330; periodic: Splines/periodic.jl:76-117)
// computed in T (the reference differentiates in the coefficient type).  t = stored knots
// (plain: full vector; periodic: breakpoints incl. endpoint).  Output: plain ncoef - nd, periodic ncoef.
template <typename T>
inline void derivative_coefs(int kind, int k, const std::vector<double> &tk, const std::vector<double> &c,
                      int nd, std::vector<double> &out) {
    const int64_t n = (int64_t)c.size();
    std::vector<T> du(n);
    if (kind == 0) {
        for (int64_t i = 0; i < n; ++i) du[i] = (T)c[i];
        for (int m = 1; m <= nd; ++m)
            for (int64_t i = n; i >= 1; --i) {
                T dt = (T)tk[i + k - m - 1] - (T)tk[i - 1];
                if (dt == (T)0 || i == 1) du[i - 1] = (T)0;
                else du[i - 1] = (T)(k - m) * (du[i - 1] - du[i - 2]) / dt;
            }
        out.resize(n - nd);
        for (int64_t i = 0; i < n - nd; ++i) out[i] = (double)du[i + nd];
    } else {
        // knots(B') : same breakpoints, order k - nd (BSplines/periodic.jl:255-260)
        std::vector<T> tt(tk.size());
        for (size_t i = 0; i < tk.size(); ++i) tt[i] = (T)tk[i];
        KnotView<T> kd;
        kd.t = tt.data(); kd.nt = (int)tt.size(); kd.N = (int)tt.size() - 1; kd.k = k - nd; kd.kind = 1;
        kd.ilast = 0; kd.a = tt.front(); kd.b = tt.back(); kd.period = tt.back() - tt.front();
        kd.lut = nullptr; kd.ncell = 0; kd.lut_scale = 0;
        const int64_t N = n;
        const int delta = k / 2 - (k - nd) / 2;
        for (int64_t i = 1; i <= N; ++i) {
            int64_t j = i - delta;
            while (j < 1) j += N;
            while (j > N) j -= N;
            du[j - 1] = (T)c[i - 1];
        }
        for (int m = 1; m <= nd; ++m) {
            T du0 = du[N - 1];
            for (int64_t i = N; i >= 1; --i) {
                T dt = knot_at<1>(kd, kd.t, (int)(i + k - m)) - knot_at<1>(kd, kd.t, (int)i);
                T prev = (i == 1) ? du0 : du[i - 2];
                du[i - 1] = (T)(k - m) * (du[i - 1] - prev) / dt;
            }
        }
        out.resize(N);
        for (int64_t i = 0; i < N; ++i) out[i] = (double)du[i];
    }
}

inline void table_parallel_for(int64_t n, const std::function<void(int64_t, int64_t, int)> &body) {
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::min<int64_t>(std::max(1u, std::min(hw, 16u)), std::max<int64_t>(1, n / 2048));
    if (nt <= 1) { body(0, n, 0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t]() { body(n * t / nt, n * (t + 1) / nt, t); });
    for (auto &x : th) x.join();
}
inline int table_parallel_slots(int64_t n) {
    unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min<int64_t>(std::max(1u, std::min(hw, 16u)), std::max<int64_t>(1, n / 2048));
}

// Taylor coefficients about z of the polynomial piece of knot interval n (1-based, t_n < t_{n+1}):
// de Boor's recurrence (spline.jl:174-209) with every d_j a polynomial in u = x - z, in dd.
template <int KIND, typename T>
inline void piece_taylor_dd(const KnotView<T> &kv, const T *t, const double *c, int ncoef, int k, int n, double z,
                            dd *a) {
    dd P[BSK_MAXK][BSK_MAXK];
    for (int j = 1; j <= k; ++j) P[j - 1][0] = dd(coef_at<KIND>(c, ncoef, 1, j + n - k));
    for (int r = 2; r <= k; ++r) {
        for (int j = k; j >= r; --j) {
            const double ti = (double)knot_at<KIND>(kv, t, j - k + n);
            const double tj = (double)knot_at<KIND>(kv, t, j - r + n + 1);
            const dd den = dd(tj) - dd(ti);
            const dd a0 = (dd(z) - dd(ti)) / den;
            const dd a1 = dd(1.0) / den;
            dd D[BSK_MAXK];
            for (int m = 0; m <= r - 2; ++m) D[m] = P[j - 1][m] - P[j - 2][m];
            for (int m = r - 1; m >= 0; --m) {
                dd v = m <= r - 2 ? P[j - 2][m] + a0 * D[m] : dd(0.0);
                if (m >= 1) v = v + a1 * D[m - 1];
                P[j - 1][m] = v;
            }
        }
    }
    for (int m = 0; m < k; ++m) a[m] = P[k - 1][m];
}

// Re-centre a polynomial: coefficients about z -> about z + delta (repeated synthetic division).
inline void taylor_shift_dd(dd *a, int k, const dd &delta) {
    for (int i = 0; i < k - 1; ++i)
        for (int j = k - 2; j >= i; --j) a[j] = a[j] + delta * a[j + 1];
}

inline double falling_d(int m, int d) {
    double ff = 1.0;
    for (int q = 0; q < d; ++q) ff *= (double)(m - q);
    return ff;
}

// Rounding bound of the Horner evaluation of D^d (sum c_m u^m), |u| <= U, in units of the unit
// roundoff: coefficient rounding, the product with m!/(m-d)!, (m - d + 1) fused steps, and the
// relative rounding of u itself propagated through the term.
inline double horner_bound(const double *absc, int k, int d, double U) {
    double s = 0.0, up = 1.0;
    for (int m = d; m < k; ++m) {
        s += (2.0 * (m - d) + 3.0) * absc[m] * falling_d(m, d) * up;
        up *= U;
    }
    return s;
}

template <typename T>
struct TableBuilder {
    typedef TableTol<T> Tol;
    const TableInput in;
    const int k, kind;
    std::vector<T> tT;                         // knots in the element type (periodic wraps are formed in T)
    KnotView<T> kv;
    std::vector<double> plo, phi;              // non-empty polynomial pieces: ends
    std::vector<int> pid;                      // knot-interval index n of the piece (1-based)
    std::vector<dd> pa;                        // nrec x k Taylor coefficients about pmid (dd)
    std::vector<double> pmid;
    double scale_d[BSK_MAXK];

    explicit TableBuilder(const TableInput &i) : in(i), k(i.k), kind(i.kind) {
        const auto &t = *in.knots;
        tT.resize(t.size());
        for (size_t q = 0; q < t.size(); ++q) tT[q] = (T)t[q];
        kv.t = tT.data(); kv.nt = (int)tT.size(); kv.N = (int)in.N; kv.k = k; kv.kind = kind; kv.ilast = 0;
        kv.a = tT.front(); kv.b = tT.back(); kv.period = (T)(tT.back() - tT.front());
        kv.lut = nullptr; kv.ncell = 0; kv.lut_scale = 0;
        if (kind == 0) {
            for (int64_t q = k; q <= in.N; ++q)
                if (t[q - 1] < t[q]) { plo.push_back(t[q - 1]); phi.push_back(t[q]); pid.push_back((int)q); }
        } else {
            for (int64_t s = 1; s <= in.N; ++s)
                if (t[s - 1] < t[s]) { plo.push_back(t[s - 1]); phi.push_back(t[s]); pid.push_back((int)s + k / 2); }
        }
    }

    void taylor(int r, double z, dd *a) const {
        const auto &c = *in.coefs;
        if (kind == 0) piece_taylor_dd<0, T>(kv, tT.data(), c.data(), (int)c.size(), k, pid[r], z, a);
        else piece_taylor_dd<1, T>(kv, tT.data(), c.data(), (int)c.size(), k, pid[r], z, a);
    }

    // piece coefficients about an arbitrary centre (from the stored expansion about the piece midpoint)
    void piece_about(int r, double centre, dd *a) const {
        for (int m = 0; m < k; ++m) a[m] = pa[(size_t)r * k + m];
        taylor_shift_dd(a, k, dd(centre) - dd(pmid[r]));
    }

    void prepare_pieces() {
        const int64_t nrec = (int64_t)plo.size();
        pa.resize((size_t)nrec * k);
        pmid.resize(nrec);
        table_parallel_for(nrec, [&](int64_t r0, int64_t r1, int) {
            for (int64_t r = r0; r < r1; ++r) {
                pmid[r] = 0.5 * (plo[r] + phi[r]);
                taylor((int)r, pmid[r], &pa[(size_t)r * k]);
            }
        });
        // magnitude of D^d S: largest value at the ends and the midpoint of every piece that is not
        // negligibly narrow (clusters of nearly equal knots would otherwise set the scale with
        // derivative values that no ordinary evaluation point ever sees)
        std::vector<double> w(nrec);
        for (int64_t r = 0; r < nrec; ++r) w[r] = phi[r] - plo[r];
        std::vector<double> ws = w;
        std::nth_element(ws.begin(), ws.begin() + nrec / 2, ws.end());
        const double wmin = 1e-3 * ws[nrec / 2];
        for (int d = 0; d < k; ++d) scale_d[d] = 0.0;
        for (int64_t r = 0; r < nrec; ++r) {
            if (w[r] < wmin) continue;
            for (int d = 0; d < k; ++d) {
                for (int e = -1; e <= 1; ++e) {
                    const double uu = 0.5 * e * w[r];
                    double acc = 0.0;
                    for (int m = k - 1; m >= d; --m) acc = acc * uu + pa[(size_t)r * k + m].hi * falling_d(m, d);
                    scale_d[d] = std::max(scale_d[d], std::fabs(acc));
                }
            }
        }
    }

    // certificate of a Horner evaluation with T-rounded coefficients `c` over |u| <= U; `extra[d]` =
    // additional absolute error (already in units of the unit roundoff) of the cell encoding
    bool certified(const double *absc, double U, const double *extra) const {
        for (int d = 0; d < k; ++d) {
            const double e = Tol::u * (horner_bound(absc, k, d, U) + (extra ? extra[d] : 0.0));
            if (!(e <= kCertBudget * Tol::tol * scale_d[d])) return false;
        }
        note_range(absc);
        return true;
    }

    // Range of the element type: the kernels form c_m * m!/(m-d)! before the Horner step scales it down by u^(m-d);
    // on very short intervals the high coefficients (~ h^-m) leave Float32's range although the sum itself is
    // harmless (found by the randomised tests: order 7, geometric knots, D^1 S came out as inf at the boundary).
    // Recorded table-wide as the largest derivative order that is safe everywhere.
    mutable std::atomic<int> range_safe{BSK_MAXK};
    void note_range(const double *absc) const {
        for (int d = 0; d < k; ++d)
            for (int m = d; m < k; ++m)
                if (!(absc[m] * falling_d(m, d) <= Tol::vmax)) {
                    int cur = range_safe.load();
                    while (d - 1 < cur && !range_safe.compare_exchange_weak(cur, d - 1)) {}
                    return;
                }
    }

    static T set_lsb(T v, int f) {
        if (sizeof(T) == 8) {
            uint64_t u; double d = (double)v; memcpy(&u, &d, 8);
            u = (u & ~1ull) | (uint64_t)f; memcpy(&d, &u, 8); return (T)d;
        }
        uint32_t u; float g = (float)v; memcpy(&u, &g, 4);
        u = (u & ~1u) | (uint32_t)f; memcpy(&g, &u, 4); return (T)g;
    }
    static T index_bits(uint32_t idx) {       // flagged cells: last slot = (index << 1) | 1 as raw bits
        const uint32_t raw = (idx << 1) | 1u;
        T out;
        if (sizeof(T) == 8) { uint64_t u = raw; double d; memcpy(&d, &u, 8); out = (T)d; }
        else { float g; memcpy(&g, &raw, 4); out = (T)g; }
        return out;
    }
    static void put_index_bits(T *rec, uint32_t idx) {      // bits 1..3 of c_0 .. c_3 (Float64 only)
        for (int m = 0; m < 4; ++m) {
            uint64_t u; double d = (double)rec[m]; memcpy(&u, &d, 8);
            u = (u & ~0xEull) | ((uint64_t)((idx >> (3 * m)) & 7u) << 1);
            memcpy(&d, &u, 8); rec[m] = (T)d;
        }
    }

    // ---- per-interval records -----------------------------------------------------------------
    void build_records(EvalTables<T> &out) {
        const int VPA = 16 / (int)sizeof(T);
        const int REC = ((k + 2 + VPA - 1) / VPA) * VPA;
        const int64_t nrec = (int64_t)plo.size();
        out.rec = REC;
        out.nrec = nrec;
        out.recs.assign((size_t)nrec * REC, (T)0);
        std::vector<int64_t> bad(table_parallel_slots(nrec), 0);
        table_parallel_for(nrec, [&](int64_t r0, int64_t r1, int slot) {
            for (int64_t r = r0; r < r1; ++r) {
                const T tlo = (T)plo[r], thi = (T)phi[r];
                T mid = set_lsb((T)0.5 * (tlo + thi), 0);
                dd a[BSK_MAXK];
                piece_about((int)r, (double)mid, a);
                double absc[BSK_MAXK];
                for (int m = 0; m < k; ++m) absc[m] = std::fabs(a[m].hi);
                const double U = std::max((double)mid - plo[r], phi[r] - (double)mid);
                if (!certified(absc, U, nullptr)) {
                    mid = set_lsb(mid, 1);
                    piece_about((int)r, (double)mid, a);
                    ++bad[slot];
                }
                T *R = &out.recs[(size_t)r * REC];
                R[0] = tlo; R[1] = mid;
                for (int m = 0; m < k; ++m) R[2 + m] = (T)a[m].value();
            }
        });
        out.n_uncertified = 0;
        for (int64_t b : bad) out.n_uncertified += b;
        // LUT over the record left ends (same bracket construction as the knot LUT, bsk_api.cu)
        std::vector<double> xi(nrec + 1);
        for (int64_t r = 0; r < nrec; ++r) xi[r] = (double)(T)plo[r];
        xi[nrec] = (*in.knots)[in.knots->size() - 1];
        const int n = (int)xi.size();
        const double range = xi[n - 1] - xi[0];
        int nc = 64;
        while (nc < 4 * n && nc < 65536) nc <<= 1;
        double sc = (double)nc / range;
        if (sizeof(T) == 4) sc = (double)(float)sc;
        const double margin = sizeof(T) == 4 ? 0.05 : 1e-6;
        out.lut.resize(nc);
        for (int c = 0; c < nc; ++c) {
            const double e_lo = xi[0] + ((double)c - margin) / sc;
            const double e_hi = xi[0] + ((double)c + 1.0 + margin) / sc;
            int lo = (int)(std::lower_bound(xi.begin(), xi.end(), e_lo) - xi.begin());
            int hi = (int)(std::upper_bound(xi.begin(), xi.end(), e_hi) - xi.begin());
            if (c == 0) lo = 0;
            if (c == nc - 1) hi = n;
            int w = hi - lo;
            if (w > 254) w = 255;
            out.lut[c] = (uint32_t)lo | ((uint32_t)w << 24);
        }
        out.lut_ncell = nc;
        out.lut_scale = sc;
    }

    // ---- uniform-cell table (bsk_eval_cell.cu) ---------------------------------------------------
    void build_cells(EvalTables<T> &out) {
        out.cell_valid = false;
        const int VPA = 16 / (int)sizeof(T);
        const int RECV = ((k + VPA - 1) / VPA) * VPA;
        const int NCH = RECV / VPA;
        const int XI = (k == 1) ? 1 : 0;
        const size_t budget = 214 * 1024;
        const int64_t nrec = (int64_t)plo.size();
        const int nbp = (int)nrec - 1;
        const double a = plo[0], bb = phi[nrec - 1];
        const uint32_t kMulti = 0x7FFFFFFFu, kJumpMulti = 4095u;
        const bool jump_ok = cell_jump_scheme((int)sizeof(T), k);
        const size_t side_bytes = jump_ok ? 16 : (size_t)2 * RECV * sizeof(T);
        int64_t nc = (int64_t)((budget - std::min<size_t>(budget, (size_t)(nbp + 8) * side_bytes)) / (RECV * sizeof(T)));
        nc = std::min<int64_t>(nc, 16 * nrec);
        nc = std::min<int64_t>(nc, 60000);
        const bool in_global = nc < 3 * nrec && nrec <= in.global_max_rec && in.allow_global_cell;
        const bool jump = jump_ok && !in_global;
        const bool jpad = jump && cell_jump_pad(k);
        if (in_global) {
            const int m2 = std::max(6, in.global_cells_x2);
            nc = (nrec * m2) / 2;
            const int64_t cap = in.global_cap_bytes / (int64_t)(cell_global_stride((int)sizeof(T), RECV) * sizeof(T));
            nc = std::max<int64_t>(std::min(nc, cap), 3 * nrec);
        }
        for (int attempt = 0; attempt < 6 && nc >= 3 * nrec && nc >= 8; ++attempt, nc = nc * 9 / 10) {
            const T x0 = (T)a;
            const T cwT = (T)((bb - a) / (double)nc);
            const T scT = (T)((double)nc / (bb - a));
            const double cw = (double)cwT;
            // The device forms the cell index as trunc(fl(fl(x - x0) * scale)) in T: it can differ from the
            // host's (x - a) / cw by 4 u_T nc cells at the far end of the domain (two roundings on the device,
            // the roundings of cw and scale themselves).  Pieces are assigned to cells with that margin.
            const double margin_cells = (sizeof(T) == 4 ? 0.01 : 1e-6) + 6.0 * (double)nc * Tol::u;
            const double margin = margin_cells * cw;
            const double U = (0.5 + margin_cells) * cw;
            std::vector<T> crec((size_t)nc * RECV, (T)0);
            struct Side { int64_t cell; T knot; T l[BSK_MAXK], r[BSK_MAXK]; T jump; };
            const int nslot = table_parallel_slots(nc);
            std::vector<std::vector<Side>> sides(nslot);
            std::vector<int64_t> nslow(nslot, 0);
            table_parallel_for(nc, [&](int64_t c0, int64_t c1, int slot) {
                std::vector<Side> &mine = sides[slot];
                for (int64_t c = c0; c < c1; ++c) {
                    const double e0 = a + (double)c * cw, e1 = a + (double)(c + 1) * cw;
                    const T centerT = std::fma((T)c + (T)0.5, cwT, x0);     // exactly the device's formula
                    const double centre = (double)centerT;
                    int pl = (int)(std::lower_bound(plo.begin() + 1, plo.end(), e0 - margin) - (plo.begin() + 1));
                    int nb = 0;
                    while (pl + 1 + nb < (int)nrec && plo[pl + 1 + nb] <= e1 + margin) ++nb;
                    if (c == nc - 1) nb = (int)nrec - 1 - pl;        // everything up to the right end
                    const double Uc = (c == nc - 1) ? std::max(U, bb - centre) : U;
                    T *R = &crec[(size_t)c * RECV];
                    dd cf[BSK_MAXK];
                    double absl[BSK_MAXK];
                    piece_about(pl, centre, cf);
                    for (int m = 0; m < k; ++m) { R[m] = (T)cf[m].value(); absl[m] = std::fabs(cf[m].hi); }
                    auto mark_slow = [&]() {
                        if (jpad) R[RECV - 1] = index_bits(kMulti);
                        else if (jump) { put_index_bits(R, kJumpMulti); R[k - 1] = set_lsb(R[k - 1], 1); }
                        else R[k - 1] = index_bits(kMulti);
                        ++nslow[slot];
                    };
                    if (nb == 0) {
                        if (certified(absl, Uc, nullptr)) { if (!jpad) R[k - 1] = set_lsb(R[k - 1], 0); }
                        else mark_slow();
                        continue;
                    }
                    if (nb >= 2) { mark_slow(); continue; }
                    dd cr[BSK_MAXK];
                    double absr[BSK_MAXK];
                    piece_about(pl + 1, centre, cr);
                    for (int m = 0; m < k; ++m) absr[m] = std::fabs(cr[m].hi);
                    Side s;
                    s.cell = c;
                    s.knot = (T)plo[pl + 1];
                    if (jump) {
                        // simple knot <=> the two pieces are consecutive knot intervals; the first and the
                        // last cell stay unflagged-or-slow (boundary pieces feed the extrapolations)
                        const bool simple = pid[pl + 1] == pid[pl] + 1 && c != 0 && c != nc - 1;
                        bool ok = simple;
                        if (simple) {
                            // right of the knot the kernel forms P_L + J dx^(k-1) with P_L continued beyond its
                            // own interval: both terms enter the bound with their absolute values, and the
                            // index bits perturb c_0 .. c_3 of an even-order record by up to 8 ulp each
                            const double jmp = std::fabs((cr[k - 1] - cf[k - 1]).hi);
                            const double dxm = std::max(0.0, (e1 + margin) - plo[pl + 1]);
                            double extra[BSK_MAXK];
                            for (int d = 0; d < k; ++d) {
                                extra[d] = (double)(k - d + 2) * jmp * falling_d(k - 1, d) * std::pow(dxm, k - 1 - d);
                                if (!jpad) {
                                    double up = 1.0;
                                    for (int m = d; m < std::min(k, 4); ++m) { extra[d] += 16.0 * absl[m] * falling_d(m, d) * up; up *= Uc; }
                                }
                            }
                            ok = certified(absl, Uc, extra) && jmp * (double)k <= Tol::vmax;
                        }
                        if (!ok) { mark_slow(); continue; }
                        s.jump = (T)(cr[k - 1] - cf[k - 1]).value();
                        if (!jpad) R[k - 1] = set_lsb(R[k - 1], 1);
                        mine.push_back(s);
                    } else {
                        if (!certified(absl, Uc, nullptr) || !certified(absr, Uc, nullptr)) { mark_slow(); continue; }
                        for (int m = 0; m < k; ++m) { s.l[m] = (T)cf[m].value(); s.r[m] = (T)cr[m].value(); }
                        R[XI] = s.knot;
                        mine.push_back(s);
                    }
                }
            });
            // side indices in cell order
            int64_t nmulti = 0;
            for (int64_t v : nslow) nmulti += v;
            std::vector<T> side_l, side_r;
            int64_t nflag = 0;
            for (auto &vec : sides) {
                for (const Side &s : vec) {
                    T *R = &crec[(size_t)s.cell * RECV];
                    if (jump) {
                        const uint32_t j = (uint32_t)(side_l.size() / 2);
                        if (!jpad && j >= kJumpMulti) {          // out of index bits: the cell takes the interval table
                            put_index_bits(R, kJumpMulti);
                            ++nmulti;
                            continue;
                        }
                        side_l.push_back(s.knot);
                        side_l.push_back(s.jump);
                        if (jpad) R[RECV - 1] = index_bits(j);
                        else put_index_bits(R, j);
                    } else {
                        const uint32_t j = (uint32_t)(side_l.size() / RECV);
                        R[k - 1] = index_bits(j);
                        const size_t base = side_l.size();
                        side_l.resize(base + RECV, (T)0);
                        side_r.resize(base + RECV, (T)0);
                        for (int m = 0; m < k; ++m) { side_l[base + m] = s.l[m]; side_r[base + m] = s.r[m]; }
                    }
                    ++nflag;
                }
            }
            if (jump && side_l.empty()) side_l.resize(2, (T)0);
            if (side_l.empty()) { side_l.resize(RECV, (T)0); side_r.resize(RECV, (T)0); }
            const int nside = jump ? (int)(side_l.size() / 2) : (int)(side_l.size() / RECV);
            const size_t smem = jump ? ((size_t)nc * RECV + 2 * (size_t)nside) * sizeof(T) + 16
                                     : ((size_t)nc + 2 * (size_t)nside) * RECV * sizeof(T) + 16;
            if (!in_global && smem > 226 * 1024) continue;      // retry with fewer cells
            if (nmulti * 50 > nc) break;                        // clustered knots: not worth it
            // device layout: chunk-major for shared memory, record-major for global memory (CellView)
            const int RS = in_global ? cell_global_stride((int)sizeof(T), RECV) : RECV;
            out.cell_rec.assign((size_t)nc * RS, (T)0);
            out.cell_rc.assign(jump ? side_l.size() : 2 * (size_t)nside * RS, (T)0);
            if (jump) out.cell_rc = side_l;                     // {knot, jump} pairs, 16 bytes each
            for (int64_t c = 0; c < nc; ++c)
                for (int w = 0; w < NCH; ++w)
                    for (int e = 0; e < VPA; ++e)
                        out.cell_rec[in_global ? (size_t)c * RS + w * VPA + e : ((size_t)w * nc + c) * VPA + e] =
                            crec[c * RECV + w * VPA + e];
            for (int side = 0; side < 2 && !jump; ++side) {
                const std::vector<T> &src = side ? side_r : side_l;
                for (int64_t j = 0; j < nside; ++j)
                    for (int w = 0; w < NCH; ++w)
                        for (int e = 0; e < VPA; ++e)
                            out.cell_rc[in_global ? ((size_t)side * nside + j) * RS + w * VPA + e
                                                  : (((size_t)side * NCH + w) * nside + j) * VPA + e] = src[j * RECV + w * VPA + e];
            }
            out.ncell = (int)nc;
            out.nside = nside;
            out.x0 = (double)x0;
            out.b = (double)(T)bb;
            out.scale = (double)scT;
            out.cw = (double)cwT;
            out.smem = in_global ? 0 : smem;
            out.cell_global = in_global;
            out.cell_jump = jump;
            out.n_cell_slow = nmulti;
            out.n_cell_flagged = nflag;
            out.cell_valid = true;
            break;
        }
    }
};

// Returns false (with a message) when the spline has no non-empty knot interval.
template <typename T>
inline bool build_eval_tables(const TableInput &in, EvalTables<T> &out, std::string *err) {
    TableBuilder<T> tb(in);
    if (tb.plo.empty()) {
        if (err) *err = "spline has no non-empty knot interval";
        return false;
    }
    tb.prepare_pieces();
    for (int d = 0; d < BSK_MAXK; ++d) out.scale_d[d] = d < in.k ? tb.scale_d[d] : 0.0;
    tb.build_records(out);
    tb.build_cells(out);
    out.range_safe_d = tb.range_safe.load();
    return true;
}

}  // namespace bsk
