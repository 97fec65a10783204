"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY — see bsk_oracle.c).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package (bsplinekit.jl_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KIND_PLAIN, KIND_PERIODIC, KIND_RECOMBINED = 0, 1, 2


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libbsk_oracle.so")
        src_m = max(os.path.getmtime(os.path.join(_HERE, f))
                    for f in ("bsk_oracle.c", "bsk_oracle_tmpl.h", "bsk_oracle_band_tmpl.h"))
        if not os.path.exists(path) or os.path.getmtime(path) < src_m:
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_basis_function.restype = C.c_double
        _LIB.orc_collocation_matrix.restype = C.c_int64
        _LIB.orc_collocation_cyclic.restype = C.c_int64
        _LIB.orc_banded_lu.restype = C.c_int64
        _LIB.orc_collocation_matrix_f32.restype = C.c_int64
        _LIB.orc_collocation_cyclic_f32.restype = C.c_int64
        _LIB.orc_banded_lu_f32.restype = C.c_int64
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _np(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


_SFX = {("f64", "f64"): "dd", ("f64", "f32"): "df", ("f32", "f32"): "ff"}
_DT = {"f64": np.float64, "f32": np.float32}


class Recomb:
    """Corner description of a RecombineMatrix (matrices.jl:70-110)."""

    def __init__(self, cl, cr, nl, nr, ul, lr):
        self.cl, self.cr, self.nl, self.nr = cl, cr, nl, nr
        self.ul = np.asfortranarray(np.asarray(ul, dtype=np.float64).reshape(-1, nl) if nl else np.zeros((cl, 0)))
        self.lr = np.asfortranarray(np.asarray(lr, dtype=np.float64).reshape(-1, nr) if nr else np.zeros((cr, 0)))


def find_knot_interval(kind, k, t, x, kt="f64"):
    t = _np(t, _DT[kt]); x = _np(x, _DT[kt])
    idx = np.empty(x.size, np.int64); zone = np.empty(x.size, np.int8)
    getattr(lib(), "orc_find_knot_interval_" + _SFX[(kt, kt)])(
        C.c_int(kind), C.c_int(k), _p(t), C.c_int64(t.size), _p(x), C.c_int64(x.size), _p(idx), _p(zone))
    return idx, zone


def evaluate_all(kind, k, t, x, nderiv=0, kt="f64", vt="f64", ileft=None, recomb=None):
    t = _np(t, _DT[kt]); x = _np(x, _DT[kt])
    n = x.size
    idx = np.empty(n, np.int64); bs = np.empty((n, k), _DT[vt])
    il = _np(ileft, np.int64) if ileft is not None else None
    sfx = _SFX[(kt, vt)]
    if recomb is None:
        getattr(lib(), "orc_evaluate_all_" + sfx)(
            C.c_int(kind), C.c_int(k), _p(t), C.c_int64(t.size), _p(x), C.c_int64(n),
            C.c_int(nderiv), _p(il), _p(idx), _p(bs))
    else:
        r = recomb
        getattr(lib(), "orc_evaluate_all_recombined_" + sfx)(
            C.c_int(k), _p(t), C.c_int64(t.size), C.c_int(r.cl), C.c_int(r.cr), C.c_int(r.nl),
            C.c_int(r.nr), C.c_int(r.ul.shape[0]), _p(r.ul), C.c_int(r.lr.shape[0]), _p(r.lr),
            _p(x), C.c_int64(n), C.c_int(nderiv), _p(il), _p(idx), _p(bs))
    return idx, bs


def spline_eval(kind, k, t, coefs, x, kt="f64"):
    t = _np(t, _DT[kt]); x = _np(x, _DT[kt]); c = _np(coefs, _DT[kt])
    out = np.empty(x.size, _DT[kt])
    getattr(lib(), "orc_spline_eval_" + _SFX[(kt, kt)])(
        C.c_int(kind), C.c_int(k), _p(t), C.c_int64(t.size), _p(c), C.c_int64(c.size),
        _p(x), C.c_int64(x.size), _p(out))
    return out


def spline_derivative(kind, k, t, coefs, nd, kt="f64"):
    """Returns (new_k, new_knots, new_coefs) of Derivative(nd) * S."""
    t = _np(t, _DT[kt]); c = _np(coefs, _DT[kt])
    out = np.empty(c.size - nd if kind == KIND_PLAIN else c.size, _DT[kt])
    rc = getattr(lib(), "orc_spline_derivative_" + _SFX[(kt, kt)])(
        C.c_int(kind), C.c_int(k), _p(t), C.c_int64(t.size), _p(c), C.c_int64(c.size),
        C.c_int(nd), _p(out))
    if rc != 0:
        raise ValueError("cannot differentiate order %d spline %d times" % (k, nd))
    tn = t[nd:t.size - nd].copy() if kind == KIND_PLAIN else t.copy()
    return k - nd, tn, out


def basis_function(k, t, i, x, nd=0):
    t = _np(t, np.float64)
    return lib().orc_basis_function(C.c_int(k), _p(t), C.c_int64(t.size), C.c_int64(i),
                                    C.c_double(x), C.c_int(nd))


def collocation_matrix(kind, k, t, x, nderiv=0, clip=np.finfo(np.float64).eps, recomb=None):
    """Band storage (2k-3, ncols) column-major (returned as Fortran-ordered array)."""
    t = _np(t, np.float64); x = _np(x, np.float64)
    r = recomb if recomb is not None else Recomb(0, 0, 0, 0, np.zeros((0, 0)), np.zeros((0, 0)))
    ncols = t.size - k - r.cl - r.cr
    band = np.zeros((2 * k - 3, ncols), np.float64, order="F")
    nout = lib().orc_collocation_matrix(
        C.c_int(kind), C.c_int(k), _p(t), C.c_int64(t.size), C.c_int(r.cl), C.c_int(r.cr),
        C.c_int(r.nl), C.c_int(r.nr), C.c_int(r.ul.shape[0]), _p(r.ul), C.c_int(r.lr.shape[0]),
        _p(r.lr), _p(x), C.c_int64(x.size), C.c_int(nderiv), C.c_double(clip), _p(band),
        C.c_int64(ncols))
    return band, int(nout)


def collocation_cyclic(k, data, x, nderiv=0, clip=np.finfo(np.float64).eps):
    data = _np(data, np.float64); x = _np(x, np.float64)
    n = data.size - 1
    a = np.empty(n); b = np.empty(n); c = np.empty(n)
    nout = lib().orc_collocation_cyclic(C.c_int(k), _p(data), C.c_int64(data.size), _p(x),
                                        C.c_int64(x.size), C.c_int(nderiv), C.c_double(clip),
                                        _p(a), _p(b), _p(c))
    return a, b, c, int(nout)


def collocation_matrix_f32(kind, k, t, x, nderiv=0, clip=np.finfo(np.float32).eps, recomb=None):
    """Float32 basis and points: CollocationMatrix{Float32} (Collocation.jl:165-198)."""
    t = _np(t, np.float32); x = _np(x, np.float32)
    r = recomb if recomb is not None else Recomb(0, 0, 0, 0, np.zeros((0, 0)), np.zeros((0, 0)))
    ncols = t.size - k - r.cl - r.cr
    band = np.zeros((2 * k - 3, ncols), np.float32, order="F")
    nout = lib().orc_collocation_matrix_f32(
        C.c_int(kind), C.c_int(k), _p(t), C.c_int64(t.size), C.c_int(r.cl), C.c_int(r.cr),
        C.c_int(r.nl), C.c_int(r.nr), C.c_int(r.ul.shape[0]), _p(r.ul), C.c_int(r.lr.shape[0]),
        _p(r.lr), _p(x), C.c_int64(x.size), C.c_int(nderiv), C.c_float(clip), _p(band),
        C.c_int64(ncols))
    return band, int(nout)


def collocation_cyclic_f32(k, data, x, nderiv=0, clip=np.finfo(np.float32).eps):
    data = _np(data, np.float32); x = _np(x, np.float32)
    n = data.size - 1
    a = np.empty(n, np.float32); b = np.empty(n, np.float32); c = np.empty(n, np.float32)
    nout = lib().orc_collocation_cyclic_f32(C.c_int(k), _p(data), C.c_int64(data.size), _p(x),
                                            C.c_int64(x.size), C.c_int(nderiv), C.c_float(clip),
                                            _p(a), _p(b), _p(c))
    return a, b, c, int(nout)


def banded_lu(band):
    """In place on a Fortran-ordered (l+u+1, n) array, l == u. Returns info (0 = ok)."""
    assert band.flags.f_contiguous and band.dtype in (np.float64, np.float32)
    h = (band.shape[0] - 1) // 2
    if band.dtype == np.float32:
        return int(lib().orc_banded_lu_f32(_p(band), C.c_int(h), C.c_int(h), C.c_int64(band.shape[1])))
    return int(lib().orc_banded_lu(_p(band), C.c_int(h), C.c_int(h), C.c_int64(band.shape[1])))


def banded_solve(band_lu, rhs):
    """rhs: C-ordered (n, nrhs) array (rhs index fastest), solved in place."""
    assert band_lu.flags.f_contiguous and rhs.flags.c_contiguous and rhs.dtype == band_lu.dtype
    h = (band_lu.shape[0] - 1) // 2
    n = band_lu.shape[1]
    r2 = rhs.reshape(n, -1)
    if rhs.dtype == np.float32:
        lib().orc_banded_solve_f32(_p(band_lu), C.c_int(h), C.c_int(h), C.c_int64(n), _p(r2),
                                   C.c_int64(r2.shape[1]))
        return rhs
    assert rhs.dtype == np.float64
    lib().orc_banded_solve(_p(band_lu), C.c_int(h), C.c_int(h), C.c_int64(n), _p(r2),
                           C.c_int64(r2.shape[1]))
    return rhs


def banded_matvec(band, x):
    h = (band.shape[0] - 1) // 2
    n = band.shape[1]
    x2 = np.ascontiguousarray(x, dtype=np.float64).reshape(n, -1)
    y = np.empty_like(x2)
    lib().orc_banded_matvec(_p(band), C.c_int(h), C.c_int(h), C.c_int64(n), _p(x2), _p(y),
                            C.c_int64(x2.shape[1]))
    return y.reshape(np.shape(x))


def cyclic_solve_f32(a, b, c, d):
    a = _np(a, np.float32); b = _np(b, np.float32); c = _np(c, np.float32)
    n = a.size
    d2 = np.array(d, dtype=np.float32, order="C").reshape(n, -1)
    x = np.empty_like(d2)
    lib().orc_cyclic_solve_f32(_p(a), _p(b), _p(c), C.c_int64(n), _p(d2), _p(x), C.c_int64(d2.shape[1]))
    return x.reshape(np.shape(d))


def cyclic_solve(a, b, c, d):
    a = _np(a, np.float64); b = _np(b, np.float64); c = _np(c, np.float64)
    n = a.size
    d2 = np.array(d, dtype=np.float64, order="C").reshape(n, -1)
    x = np.empty_like(d2)
    lib().orc_cyclic_solve(_p(a), _p(b), _p(c), C.c_int64(n), _p(d2), _p(x), C.c_int64(d2.shape[1]))
    return x.reshape(np.shape(d))


def make_knots(x, k):
    x = _np(x, np.float64)
    t = np.empty(x.size + k)
    lib().orc_make_knots(_p(x), C.c_int64(x.size), C.c_int(k), _p(t))
    return t


def make_knots_periodic(x, k, L):
    x = _np(x, np.float64)
    t = np.empty(x.size + 1)
    lib().orc_make_knots_periodic(_p(x), C.c_int64(x.size), C.c_int(k), C.c_double(L), _p(t))
    return t


def greville(kind, k, t, cl=0, N=None):
    t = _np(t, np.float64)
    if N is None:
        N = t.size - k
    out = np.empty(N)
    lib().orc_greville(C.c_int(kind), C.c_int(k), _p(t), C.c_int64(t.size), C.c_int(cl),
                       C.c_int64(N), _p(out))
    return out


def robin_corners(k, t, alpha):
    """Corner blocks for a single BC op = sum_m alpha[m] D^m on both sides."""
    t = _np(t, np.float64); alpha = _np(alpha, np.float64)
    q = alpha.size - 1
    ul = np.zeros((q + 1, q), order="F"); lr = np.zeros((q + 1, q), order="F")
    lib().orc_robin_corners(C.c_int(k), _p(t), C.c_int64(t.size), _p(alpha), C.c_int(q), _p(ul), _p(lr))
    return Recomb(1, 1, q, q, ul, lr)


def augment_knots(breaks, k):
    """augment_knots! (knots.jl:68-84): pad both ends to multiplicity k."""
    b = np.asarray(breaks)
    return np.concatenate([np.repeat(b[:1], k - 1), b, np.repeat(b[-1:], k - 1)])


def natural_corners(k, t):
    """Natural boundary conditions: corner blocks of the recombination matrix
    (src/Recombinations/natural.jl:9-134).  For each side: the derivatives D^2..D^h (h = k/2) of the
    h+1 boundary B-splines at the boundary, two (h+1)x(h+1) systems (rows: sum = h, normalised
    derivative rows = 0, last row picks b_{h+1} resp. b_1), near-zeros removed."""
    t = _np(t, np.float64)
    assert k % 2 == 0
    h = k // 2
    N = t.size - k
    blocks = []
    for side in ("left", "right"):
        x = t[k - 1] if side == "left" else t[N]
        js = list(range(1, h + 2)) if side == "left" else list(range(N - h, N + 1))
        D = np.zeros((h - 1, h + 1))
        for i in range(1, h):                                # natural.jl:66-97: evaluate_all(B, x, Derivative(i+1))
            idx, bs = evaluate_all(0, k, t, [x], nderiv=i + 1)
            jlast = int(idx[0])
            for n, j in enumerate(js):
                dj = jlast + 1 - j
                if 1 <= dj <= k:
                    D[i - 1, n] = bs[0, dj - 1]
        rhs = np.zeros(h + 1)
        rhs[0] = h
        cols = []
        for which in (1, 2):                                 # natural.jl:103-134
            M = np.zeros((h + 1, h + 1))
            M[0, :] = 1.0
            for i in range(2, h + 1):
                M[i - 1, :] = D[i - 2] / np.max(np.abs(D[i - 2]))
            M[h, h if which == 1 else 0] = 1.0
            v = np.linalg.solve(M, rhs)
            v[np.abs(v) < 10 * np.finfo(np.float64).eps * np.max(np.abs(v))] = 0.0      # natural.jl:44-48
            cols.append(v)
        blocks.append(np.stack(cols, axis=1))
    return Recomb(h - 1, h - 1, 2, 2, blocks[0], blocks[1])


def band_to_dense(band):
    """Band store data[u + i - j, j] (Collocation.jl:165-198) -> dense matrix."""
    nb, n = band.shape
    u = (nb - 1) // 2
    A = np.zeros((n, n))
    for j in range(n):
        for i in range(max(0, j - u), min(n, j + u + 1)):
            A[i, j] = band[u + i - j, j]
    return A


def fit_smoothing(xs, ys, lam, weights=None):
    """fit(BSplineOrder(4), xs, ys, λ, Natural(); weights) restated with dense numpy algebra
    (src/SplineInterpolations/smoothing.jl:43-124): returns the coefficients in the natural
    recombined basis.  ys: (N,) or (N, M).  Small N only (dense N x N matrices)."""
    xs = _np(xs, np.float64)
    N = xs.size
    t = augment_knots(xs, 4)                                 # BSplineBasis(order, copy(xs)), smoothing.jl:57
    rc = natural_corners(4, t)                               # RecombinedBSplineBasis(B, Natural()), :58
    A = band_to_dense(collocation_matrix(KIND_RECOMBINED, 4, t, xs, 0, recomb=rc)[0])   # :61-64
    D = band_to_dense(collocation_matrix(KIND_RECOMBINED, 4, t, xs, 2, recomb=rc)[0])
    dl = np.zeros((N, N))                                    # :66-74
    for i in range(1, N - 1):
        dl[i, i] = 2 * (xs[i + 1] - xs[i - 1])
        dl[i, i + 1] = xs[i + 1] - xs[i]
    dl = dl + np.triu(dl, 1).T                               # Hermitian(Δ_upper)
    H = dl @ D                                               # :77
    Bm = H.T @ D                                             # :83
    W = np.diag(np.asarray(weights, dtype=np.float64)) if weights is not None else np.eye(N)
    M = (Bm + Bm.T) * (lam / 6) + 2 * (A.T @ W @ A)          # :84-92
    Lc = np.linalg.cholesky((M + M.T) / 2)                   # cholesky!(Hermitian(M)), :93
    z = 2 * (A.T @ (W @ np.asarray(ys, dtype=np.float64)))   # :113-114
    return np.linalg.solve(Lc.T, np.linalg.solve(Lc, z)), t, rc


def periodic_knots(data, k):
    """Materialised PeriodicKnots t[1..N+k] (src/BSplines/periodic.jl:62,102-115): t[i] = xi[i - k/2] +- L."""
    data = _np(data, np.float64)
    N = data.size - 1
    L = data[-1] - data[0]
    out = np.empty(N + k)
    for i in range(1, N + k + 1):
        x = 0.0
        j = i - k // 2
        while j < 1:
            x -= L
            j += N
        while j > N + 1:
            x += L
            j -= N
        out[i - 1] = x + data[j - 1]
    return out


def collocation_points_periodic(k, data):
    """collocation_points(::PeriodicBSplineBasis): even k -> the knots (SameAsKnots), odd k -> Greville
    averages of the periodic knots clamped to the period (src/Collocation/points.jl:56-104,164-225)."""
    data = _np(data, np.float64)
    N = data.size - 1
    if k % 2 == 0:
        return data[:-1].copy()
    t = periodic_knots(data, k)
    x = np.empty(N)
    tsum = comp = 0.0
    for i in range(N):
        if i == 0:
            tsum = 0.0
            for j in range(2, k + 1):
                tsum += t[i + j - 1]
        else:
            prev = tsum
            y = t[i + k - 1] - comp
            tsum = prev + y
            comp = (tsum - prev) - y
            prev = tsum
            y = -t[i] - comp
            tsum = prev + y
            comp = (tsum - prev) - y
        x[i] = min(max(tsum / (k - 1), data[0]), data[-1])
    return x


def periodic_collocation_dense(k, data, x, nderiv=0, clip=np.finfo(np.float64).eps):
    """collocation_matrix!(C, B::PeriodicBSplineBasis, x, Derivative(n)) as a dense N x N matrix
    (src/Collocation/Collocation.jl:209-251; column wrap = basis_to_array_index,
    src/BSplines/periodic.jl:245-253).  The reference stores it sparse and factorises with UMFPACK
    (Collocation.jl:141-142); a dense pivoted solve of this matrix is the oracle for those orders."""
    data = _np(data, np.float64); x = _np(x, np.float64)
    N = data.size - 1
    idx, bs = evaluate_all(KIND_PERIODIC, k, data, x, nderiv=nderiv)
    A = np.zeros((x.size, N))
    for i in range(x.size):
        for d in range(1, k + 1):
            v = bs[i, d - 1]
            if abs(v) <= clip:
                continue
            A[i, (int(idx[i]) - d) % N] = v                  # j = jlast + 1 - d (1-based), wrapped
    return A


def spline_extrapolate(method, k, t, coefs, x):
    """extrapolate(S, method)(x) for a plain basis, method in {"flat", "linear", "smooth"}
    (src/SplineExtrapolations/SplineExtrapolations.jl:22-71).  Linear: the slope is the value of
    Derivative(1) * S at the boundary (the reference differentiates S with ForwardDiff, the same
    number up to rounding)."""
    t = _np(t, np.float64); c = _np(coefs, np.float64); x = _np(x, np.float64)
    N = c.size
    a, b = t[k - 1], t[N]
    if method == "flat":
        return spline_eval(0, k, t, c, np.clip(x, a, b))
    if method == "smooth":
        out = np.empty_like(x)
        lib().orc_spline_eval_smooth_dd(C.c_int(k), _p(t), C.c_int64(t.size), _p(c), C.c_int64(N), _p(x),
                                        C.c_int64(x.size), _p(out))
        return out
    assert method == "linear"
    k1, t1, c1 = spline_derivative(0, k, t, c, 1)
    va, vb = spline_eval(0, k, t, c, [a, b])
    sa, sb = spline_eval(0, k1, t1, c1, [a, b])
    out = spline_eval(0, k, t, c, np.clip(x, a, b))
    out = np.where(x < a, va + sa * (x - a), out)
    return np.where(x > b, vb + sb * (x - b), out)


def galerkin_matrix(k, t, d1=0, d2=0, recomb=None, f=None, nq=None):
    """galerkin_matrix!(M, B, (Derivative(d1), Derivative(d2))) as a dense matrix, the reference's own
    loop (src/Galerkin/linear.jl:268-311): knot segments in order, k Gauss-Legendre nodes each
    (quadratures.jl:23-65), evaluate_all with ileft, A[i, j] += y * bi * bj."""
    t = _np(t, np.float64)
    kind = KIND_RECOMBINED if recomb is not None else KIND_PLAIN
    off = recomb.cl if recomb is not None else 0
    Np = t.size - k
    N = Np - (recomb.cl + recomb.cr if recomb is not None else 0)
    xa, xb = t[k - 1], t[Np]
    qx, qw = np.polynomial.legendre.leggauss(k if nq is None else nq)
    A = np.zeros((N, N))
    for n in range(1, t.size):
        tn, tn1 = t[n - 1], t[n]
        if tn1 == tn or tn1 <= xa or tn >= xb:
            continue
        alpha, beta = (tn1 - tn) / 2, (tn + tn1) / 2
        xs = alpha * qx + beta
        il = np.full(qx.size, n - off, dtype=np.int64)
        _, b1 = evaluate_all(kind, k, t, xs, nderiv=d1, ileft=il, recomb=recomb)
        b2 = b1 if d2 == d1 else evaluate_all(kind, k, t, xs, nderiv=d2, ileft=il, recomb=recomb)[1]
        ilast = n - off
        for q in range(qx.size):
            y = alpha * qw[q]
            fq = None if f is None else float(f(xs[q]))
            for dj in range(1, k + 1):
                j = ilast + 1 - dj
                if not 1 <= j <= N:
                    continue
                for di in range(1, k + 1):
                    i = ilast + 1 - di
                    if not 1 <= i <= N:
                        continue
                    term = y * b1[q, di - 1] * b2[q, dj - 1]              # y * bi * bj * f(x), left to right (:305)
                    A[i - 1, j - 1] += term if fq is None else term * fq
    return A


def galerkin_projection(f, k, t, nderiv=0, recomb=None):
    """galerkin_projection!(f, phi, B, Derivative(n)) (src/Galerkin/projection.jl:60-99)."""
    t = _np(t, np.float64)
    kind = KIND_RECOMBINED if recomb is not None else KIND_PLAIN
    off = recomb.cl if recomb is not None else 0
    Np = t.size - k
    N = Np - (recomb.cl + recomb.cr if recomb is not None else 0)
    qx, qw = np.polynomial.legendre.leggauss(k)
    phi = np.zeros(N)
    for n in range(1, t.size):
        tn, tn1 = t[n - 1], t[n]
        if tn1 == tn:
            continue
        alpha, beta = (tn1 - tn) / 2, (tn + tn1) / 2
        xs = alpha * qx + beta
        il = np.full(k, n - off, dtype=np.int64)
        _, bs = evaluate_all(kind, k, t, xs, nderiv=nderiv, ileft=il, recomb=recomb)
        ilast = n - off
        for q in range(k):
            y = alpha * qw[q] * f(xs[q])
            for di in range(1, k + 1):
                i = ilast + 1 - di
                if 1 <= i <= N:
                    phi[i - 1] += y * bs[q, di - 1]
    return phi


def galerkin_matrix_pair(k, t, rc1, rc2, d1=0, d2=0):
    """galerkin_matrix!(M, (B1, B2), (Derivative(d1), Derivative(d2))) as a dense N1 x N2 matrix: the same loop
    (src/Galerkin/linear.jl:268-311) with two bases derived from one parent; rc1 / rc2 = Recomb or None (the parent)."""
    t = _np(t, np.float64)
    Np = t.size - k
    kind1 = KIND_RECOMBINED if rc1 is not None else KIND_PLAIN
    kind2 = KIND_RECOMBINED if rc2 is not None else KIND_PLAIN
    off1 = rc1.cl if rc1 is not None else 0
    off2 = rc2.cl if rc2 is not None else 0
    N1 = Np - ((rc1.cl + rc1.cr) if rc1 is not None else 0)
    N2 = Np - ((rc2.cl + rc2.cr) if rc2 is not None else 0)
    xa, xb = t[k - 1], t[Np]
    qx, qw = np.polynomial.legendre.leggauss(k)
    A = np.zeros((N1, N2))
    for n in range(1, t.size):
        tn, tn1 = t[n - 1], t[n]
        if tn1 == tn or tn1 <= xa or tn >= xb:
            continue
        alpha, beta = (tn1 - tn) / 2, (tn + tn1) / 2
        xs = alpha * qx + beta
        _, b1 = evaluate_all(kind1, k, t, xs, nderiv=d1, ileft=np.full(k, n - off1, dtype=np.int64), recomb=rc1)
        _, b2 = evaluate_all(kind2, k, t, xs, nderiv=d2, ileft=np.full(k, n - off2, dtype=np.int64), recomb=rc2)
        for q in range(k):
            y = alpha * qw[q]
            for dj in range(1, k + 1):
                j = n - off2 + 1 - dj
                if not 1 <= j <= N2:
                    continue
                for di in range(1, k + 1):
                    i = n - off1 + 1 - di
                    if not 1 <= i <= N1:
                        continue
                    A[i - 1, j - 1] += y * b1[q, di - 1] * b2[q, dj - 1]
    return A


def galerkin_matrix_periodic(k, data, d1=0, d2=0):
    """galerkin_matrix!(M, B::PeriodicBSplineBasis, (Derivative(d1), Derivative(d2))) as a dense matrix: the same
    loop (src/Galerkin/linear.jl:268-311) over the PeriodicKnots ts[1 .. N+k]; only the segments inside the main
    period pass the boundary test (:277-279); array indices wrap (basis_to_array_index, BSplines/periodic.jl:245-253)."""
    data = _np(data, np.float64)
    N = data.size - 1
    ts = periodic_knots(data, k)
    xa, xb = data[0], data[-1]
    qx, qw = np.polynomial.legendre.leggauss(k)
    A = np.zeros((N, N))
    for n in range(1, ts.size):
        tn, tn1 = ts[n - 1], ts[n]
        if tn1 == tn or tn1 <= xa or tn >= xb:
            continue
        alpha, beta = (tn1 - tn) / 2, (tn + tn1) / 2
        xs = alpha * qx + beta
        il = np.full(k, n, dtype=np.int64)
        _, b1 = evaluate_all(KIND_PERIODIC, k, data, xs, nderiv=d1, ileft=il)
        b2 = b1 if d2 == d1 else evaluate_all(KIND_PERIODIC, k, data, xs, nderiv=d2, ileft=il)[1]
        for q in range(k):
            y = alpha * qw[q]
            for dj in range(1, k + 1):
                j = (n + 1 - dj - 1) % N
                for di in range(1, k + 1):
                    i = (n + 1 - di - 1) % N
                    A[i, j] += y * b1[q, di - 1] * b2[q, dj - 1]
    return A


def galerkin_projection_periodic(f, k, data, nderiv=0):
    """galerkin_projection!(f, phi::Vector, B::PeriodicBSplineBasis) (src/Galerkin/projection.jl:60-99): ALL segments
    of ts[1 .. N+k] (no boundary test in this loop), indices outside 1..N skipped (:90-92) — every b_i, i = 1..N, is
    integrated over its own unwrapped support, f being evaluated at the unwrapped nodes."""
    data = _np(data, np.float64)
    N = data.size - 1
    ts = periodic_knots(data, k)
    qx, qw = np.polynomial.legendre.leggauss(k)
    phi = np.zeros(N)
    for n in range(1, ts.size):
        tn, tn1 = ts[n - 1], ts[n]
        if tn1 == tn:
            continue
        alpha, beta = (tn1 - tn) / 2, (tn + tn1) / 2
        xs = alpha * qx + beta
        il = np.full(k, n, dtype=np.int64)
        _, bs = evaluate_all(KIND_PERIODIC, k, data, xs, nderiv=nderiv, ileft=il)
        for q in range(k):
            y = alpha * qw[q] * f(xs[q])
            for di in range(1, k + 1):
                i = n + 1 - di
                if 1 <= i <= N:
                    phi[i - 1] += y * bs[q, di - 1]
    return phi


def fit_smoothing_periodic(xs, ys, lam, L, weights=None):
    """fit(BSplineOrder(4), xs, ys, λ, Periodic(L); weights) with dense numpy algebra
    (src/SplineInterpolations/smoothing.jl:126-213).  Returns (coefficients, breakpoints)."""
    xs = _np(xs, np.float64)
    N = xs.size
    data = make_knots_periodic(xs, 4, L)
    A = periodic_collocation_dense(4, data, xs, 0)
    D = periodic_collocation_dense(4, data, xs, 2)
    dl = np.zeros((N, N))                                    # :149-167 (ts[i + h] == xs[i])
    for i in range(N):
        xm = xs[i - 1] - (L if i == 0 else 0.0)
        xp = xs[(i + 1) % N] + (L if i == N - 1 else 0.0)
        dl[i, (i - 1) % N] += xs[i] - xm
        dl[i, i] += 2 * (xp - xm)
        dl[i, (i + 1) % N] += xp - xs[i]
    H = dl @ D
    Bm = H.T @ D
    W = np.diag(np.asarray(weights, dtype=np.float64)) if weights is not None else np.eye(N)
    M = (Bm + Bm.T) * (lam / 6) + 2 * (A.T @ W @ A)
    z = 2 * (A.T @ (W @ np.asarray(ys, dtype=np.float64)))
    return np.linalg.solve(M, z), data
