// pptable_emul.cpp — TEST INFRASTRUCTURE: CPU emulation of the table-evaluation kernels
// (k_spline_cell in bsk_eval_cell.cu, k_spline_pp in bsk_eval.cu) on the tables that the product's
// host-side builder (csrc/pptable.h) produces.  The arithmetic below restates what the kernels do
// per point — cell index, record gather, knot flags, Horner with fused multiply-adds, jump term,
// interval-table fallback, de Boor route for records that are not certified — so that the accuracy
// certificate of the builder can be checked against the oracle without a GPU.  Built by
// tests/test_pptable_cpu.py with g++ -O2 -ffp-contract=off; never linked into the product.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../bsplinekit.jl_b200/csrc/pptable.h"

using namespace bsk;

namespace {

template <typename T> bool lsb(T v);
template <> bool lsb<double>(double v) { uint64_t u; memcpy(&u, &v, 8); return u & 1; }
template <> bool lsb<float>(float v) { uint32_t u; memcpy(&u, &v, 4); return u & 1; }
template <typename T> unsigned idx_bits(T v);
template <> unsigned idx_bits<double>(double v) { uint64_t u; memcpy(&u, &v, 8); return ((unsigned)(u & 0xFFFFFFFFu)) >> 1; }
template <> unsigned idx_bits<float>(float v) { uint32_t u; memcpy(&u, &v, 4); return u >> 1; }
unsigned lo3(double v) { uint64_t u; memcpy(&u, &v, 8); return (unsigned)((u >> 1) & 7u); }

template <typename T> T fallingT(int m, int d) {
    T ff = (T)1;
    for (int q = 0; q < d; ++q) ff *= (T)(m - q);
    return ff;
}

struct EmulBase {
    virtual ~EmulBase() {}
    virtual void info(int64_t *iv, double *scale) const = 0;
    virtual void eval(const double *x, int64_t n, int deriv, int path, double *out, int8_t *route) const = 0;
};

template <typename T>
struct Emul : EmulBase {
    int kind, k;
    int64_t N;
    std::vector<double> knots, coefs;
    std::vector<T> tT;
    EvalTables<T> tb;
    std::vector<std::vector<T>> dcoef;      // coefficients of Derivative(d) * S in T
    KnotView<T> kv;
    T xa, xb;

    Emul(int kind_, int k_, const double *kn, int64_t nk, int64_t N_, const double *cf, int64_t nc, int allow_global)
        : kind(kind_), k(k_), N(N_), knots(kn, kn + nk), coefs(cf, cf + nc) {
        TableInput in;
        in.kind = kind; in.k = k; in.N = N; in.knots = &knots; in.coefs = &coefs;
        in.allow_global_cell = allow_global != 0; in.global_cells_x2 = 64;
        std::string err;
        build_eval_tables<T>(in, tb, &err);
        tT.assign(knots.begin(), knots.end());
        dcoef.resize(k);
        for (int d = 0; d < k; ++d) {
            std::vector<double> dc;
            if (d == 0) dc = coefs;
            else derivative_coefs<T>(kind, k, knots, coefs, d, dc);
            dcoef[d].assign(dc.begin(), dc.end());
        }
        kv.t = tT.data(); kv.nt = (int)tT.size(); kv.N = (int)N; kv.k = k; kv.kind = kind;
        kv.a = tT.front(); kv.b = tT.back(); kv.period = (T)(tT.back() - tT.front());
        kv.lut = nullptr; kv.ncell = 0; kv.lut_scale = 0;
        int il = kv.nt;
        do { il -= 1; } while (il > 1 && tT[il - 1] == tT.back());
        kv.ilast = il;
        xa = kind == 0 ? tT[k - 1] : tT.front();
        xb = kind == 0 ? tT[N] : tT.back();
    }

    void info(int64_t *iv, double *scale) const override {
        iv[0] = tb.nrec; iv[1] = tb.n_uncertified; iv[2] = tb.cell_valid; iv[3] = tb.cell_global; iv[4] = tb.cell_jump;
        iv[5] = tb.ncell; iv[6] = tb.nside; iv[7] = tb.n_cell_slow; iv[8] = tb.n_cell_flagged; iv[9] = (int64_t)tb.smem;
        for (int d = 0; d < BSK_MAXK; ++d) scale[d] = tb.scale_d[d];
    }

    // deboor_exact of pp.cuh
    T exact(int d, T xs, T xq) const {
        KnotView<T> kd = kv;
        kd.k = k - d;
        int ncoef = (int)dcoef[d].size();
        if (kind == 0) { kd.t += d; kd.nt -= 2 * d; kd.N -= d; kd.ilast -= d; }
        int z;
        if (kind == 0) {
            const int i = find_interval<0>(kd, kd.t, xs, z);
            return deboor_generic<0, T>(kd, kd.t, dcoef[d].data(), ncoef, 1, kd.k, i, xq);
        }
        const int i = find_interval<1>(kd, kd.t, xs, z);
        return deboor_generic<1, T>(kd, kd.t, dcoef[d].data(), ncoef, 1, kd.k, i, xq);
    }

    T horner(const T *cm, T u, int d) const {
        T acc = (T)0;
        for (int m = k - 1; m >= d; --m) acc = std::fma(acc, u, d == 0 ? cm[m] : cm[m] * fallingT<T>(m, d));
        return acc;
    }

    // pp_lookup of pp.cuh
    bool pp_lookup(T xs, T *cm, T *u) const {
        const int REC = tb.rec;
        const T *rec = tb.recs.data();
        int c = (int)((xs - rec[0]) * (T)tb.lut_scale);
        c = c < 0 ? 0 : (c >= tb.lut_ncell ? tb.lut_ncell - 1 : c);
        const uint32_t e = tb.lut[c];
        int lo = (int)(e & 0xFFFFFFu);
        const int w = (int)(e >> 24);
        int hi = (w != 255) ? lo + w : (int)tb.nrec;
        hi = hi > (int)tb.nrec ? (int)tb.nrec : hi;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (xs < rec[(size_t)mid * REC]) hi = mid; else lo = mid + 1;
        }
        int r = lo - 1;
        r = r < 0 ? 0 : (r >= (int)tb.nrec ? (int)tb.nrec - 1 : r);
        const T *R = rec + (size_t)r * REC;
        for (int m = 0; m < k; ++m) cm[m] = R[2 + m];
        *u = xs - R[1];
        return lsb<T>(R[1]);
    }

    void eval(const double *x, int64_t n, int deriv, int path, double *out, int8_t *route) const override {
        const int VPA = 16 / (int)sizeof(T);
        const int RECV = ((k + VPA - 1) / VPA) * VPA;
        const bool use_cell = path == 0 && tb.cell_valid && tb.n_uncertified == 0;
        const bool STAGE = !tb.cell_global;
        const int RS = STAGE ? RECV : cell_global_stride((int)sizeof(T), RECV);
        const bool JUMP = STAGE && cell_jump_scheme((int)sizeof(T), k);
        const bool JPAD = JUMP && cell_jump_pad(k);
        const int XI = (k == 1) ? 1 : 0;
        const T x0 = (T)tb.x0, scale = (T)tb.scale, cw = (T)tb.cw;
        const int nc = tb.ncell, nside = tb.nside;
        auto cell_at = [&](const std::vector<T> &tab, int count, size_t base_rec, int c, int slot) -> T {
            // shared: chunk-major [(w * count + c) * VPA + e]; global: record-major
            if (STAGE) return tab[base_rec + ((size_t)(slot / VPA) * count + c) * VPA + slot % VPA];
            return tab[base_rec + (size_t)c * RS + slot];
        };
        for (int64_t p = 0; p < n; ++p) {
            const T xq = (T)x[p];
            int8_t rt = 0;
            if (!(xq == xq)) { out[p] = (double)xq; if (route) route[p] = -1; continue; }
            bool outside = kind == 1 ? !(xq >= xa && xq < xb) : !(xq >= xa && xq <= xb);
            if (outside) {
                if (kind == 0) { out[p] = 0.0; if (route) route[p] = -2; continue; }
                out[p] = (double)exact(deriv, xq, xq);          // periodic: the reference's own arithmetic
                if (route) route[p] = 4;
                continue;
            }
            const T xs = xq;
            T cm[BSK_MAXK];
            T u, jump = (T)0, dxk = (T)0;
            bool slow = !use_cell;
            if (use_cell) {
                int c = (int)((xs - x0) * scale);
                c = c < 0 ? 0 : (c > nc - 1 ? nc - 1 : c);
                T rec[BSK_MAXK + 4];
                for (int m = 0; m < RECV; ++m) rec[m] = cell_at(tb.cell_rec, nc, 0, c, m);
                const T last = rec[JPAD ? RECV - 1 : k - 1];
                if (JUMP) {
                    if (lsb<T>(last)) {
                        unsigned j;
                        if (JPAD) j = idx_bits<T>(last);
                        else j = lo3((double)rec[0]) | (lo3((double)rec[1]) << 3) | (lo3((double)rec[2]) << 6) | (lo3((double)rec[3]) << 9);
                        if (j == (JPAD ? 0x7FFFFFFFu : 4095u)) slow = true;
                        else {
                            rt = 1;
                            const T kn = tb.cell_rc[2 * (size_t)j], J = tb.cell_rc[2 * (size_t)j + 1];
                            if (!(xs < kn)) { jump = J; dxk = xs - kn; }
                        }
                    }
                } else if (lsb<T>(last)) {
                    const unsigned j = idx_bits<T>(last);
                    if (j == 0x7FFFFFFFu) slow = true;
                    else {
                        rt = 1;
                        const int right = (xs < rec[XI]) ? 0 : 1;
                        const size_t base = STAGE ? (size_t)right * (RECV / VPA) * nside * VPA : (size_t)right * nside * RS;
                        for (int m = 0; m < RECV; ++m) rec[m] = cell_at(tb.cell_rc, nside, base, (int)j, m);
                    }
                }
                for (int m = 0; m < k; ++m) cm[m] = rec[m];
                u = xs - std::fma((T)c + (T)0.5, cw, x0);
            }
            if (slow) {
                rt = 2;
                if (pp_lookup(xs, cm, &u)) {
                    out[p] = (double)exact(deriv, xs, xs);
                    if (route) route[p] = 3;
                    continue;
                }
                jump = (T)0; dxk = (T)0;
            }
            T val = horner(cm, u, deriv);
            if (use_cell && JUMP) {
                T pw = jump;
                for (int q = 0; q < k - 1; ++q)
                    if (q < k - 1 - deriv) pw = pw * dxk;
                if (deriv <= k - 1) val = std::fma(pw, fallingT<T>(k - 1, deriv), val);
            }
            out[p] = (double)val;
            if (route) route[p] = rt;
        }
    }
};

}  // namespace

extern "C" {

void *ppe_create(int kind, int k, int is_f32, const double *knots, int64_t nk, int64_t N, const double *coefs,
                 int64_t nc, int allow_global) {
    if (is_f32) return (EmulBase *)new Emul<float>(kind, k, knots, nk, N, coefs, nc, allow_global);
    return (EmulBase *)new Emul<double>(kind, k, knots, nk, N, coefs, nc, allow_global);
}
void ppe_info(void *h, int64_t *iv, double *scale) { ((EmulBase *)h)->info(iv, scale); }
void ppe_eval(void *h, const double *x, int64_t n, int deriv, int path, double *out, int8_t *route) {
    ((EmulBase *)h)->eval(x, n, deriv, path, out, route);
}
void ppe_destroy(void *h) { delete (EmulBase *)h; }

}  // extern "C"
