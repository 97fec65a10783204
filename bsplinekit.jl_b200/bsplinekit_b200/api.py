"""Host-side mirror of the BSplineKit.jl interface for the two hot paths.

Julia is not available in this image, so this module is the executable stand-in for the
Julia shim (julia/BSplineKitB200.jl): same names, argument meaning and error behaviour
as the reference, every numerical body delegated to libbsplinekit_b200.so through the
C ABI (include/bsk.h).  Only constructor / type logic lives here (knot augmentation,
make_knots, Greville points, recombination corner blocks) — exactly the part that stays
Julia in the reference-side integration.  torch is used for device memory and streams.

Reference citations are relative to the BSplineKit.jl source tree (v0.19.2).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (ArgumentError, DimensionMismatch, ZeroPivotException, check, lib,
                   BSK_F32, BSK_F64, BASIS_PLAIN, BASIS_PERIODIC, BASIS_RECOMBINED,
                   EVAL_AUTO, EVAL_DEBOOR, EVAL_PP, SOLVE_AUTO, SOLVE_TWOPASS, SOLVE_WINDOWED)

_NP = {BSK_F32: np.float32, BSK_F64: np.float64}
_TORCH = {BSK_F32: torch.float32, BSK_F64: torch.float64}


def _dtype_code(dt):
    dt = np.dtype(dt)
    if dt == np.float32:
        return BSK_F32
    if dt == np.float64:
        return BSK_F64
    raise ArgumentError("unsupported element type %s (Float32 / Float64 only)" % dt)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _device():
    if not torch.cuda.is_available():
        raise _lib.CudaError("no CUDA device available (there is no CPU fallback)")
    return torch.cuda.current_device()


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _to_device(x, dtype_code):
    """Accept a CUDA tensor (used in place), a CPU tensor or anything numpy understands."""
    tdt = _TORCH[dtype_code]
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            if x.dtype != tdt:
                raise ArgumentError("device vector has dtype %s, expected %s" % (x.dtype, tdt))
            return x.contiguous(), True
        return x.to(tdt).contiguous().cuda(), False
    arr = np.ascontiguousarray(np.asarray(x, dtype=_NP[dtype_code]))
    return torch.from_numpy(arr).cuda(), False


# ---- operator / order vocabulary (DifferentialOps.jl:88-160, BSplines.jl:106-108) ----------

class BSplineOrder:
    def __init__(self, k):
        self.k = int(k)

    def __eq__(self, o):
        return isinstance(o, BSplineOrder) and o.k == self.k


class DifferentialOp:
    """sum_m alpha[m] D^m  (Derivative, ScaledDerivative, DifferentialOpSum)."""

    def __init__(self, terms):
        self.terms = dict(terms)           # order -> coefficient, insertion order kept

    def __add__(self, o):
        t = dict(self.terms)
        for m, a in o.terms.items():
            t[m] = t.get(m, 0.0) + a
        return DifferentialOp(t)

    def __sub__(self, o):
        return self + (-1.0) * o

    def __rmul__(self, a):
        if isinstance(a, (int, float)):
            if a == 0:
                raise ArgumentError("scale factor cannot be zero")      # DifferentialOps.jl:117-119
            return DifferentialOp({m: a * c for m, c in self.terms.items()})
        return NotImplemented

    def __neg__(self):
        return (-1.0) * self

    @property
    def max_order(self):
        return max(self.terms) if self.terms else -1

    def __eq__(self, o):                                   # S == S' whatever the order of the terms (DifferentialOps.jl:150-158)
        return isinstance(o, DifferentialOp) and {m: float(a) for m, a in self.terms.items()} == \
            {m: float(a) for m, a in o.terms.items()}

    def __hash__(self):
        return hash(tuple(sorted(self.terms.items())))

    def __repr__(self):
        return " + ".join(("D{%d}" % m) if a == 1 else ("%s * D{%d}" % (a, m)) for m, a in self.terms.items())


class LeftNormal:
    """LeftNormal() . op: odd-order derivatives change sign on the left boundary (DifferentialOps.jl:36-42,128-131)."""

    def dot(self, op):
        return DifferentialOp({m: (-a if m % 2 else a) for m, a in op.terms.items()})


class RightNormal:
    """RightNormal() . op == op (DifferentialOps.jl:44-51,78)."""

    def dot(self, op):
        return DifferentialOp(dict(op.terms))


def dot(direction, op):
    """LinearAlgebra.dot(dir, op) / dir ⋅ op (DifferentialOps.jl:64-78)."""
    if isinstance(op, (LeftNormal, RightNormal)):
        direction, op = op, direction
    return direction.dot(op)


def max_order(*ops):
    """max_order(ops...) (DifferentialOps.jl:56-81): -1 for no operator."""
    return max((o.max_order for o in ops), default=-1)


class DerivativeUnitRange:
    """Derivative(m:n) (DifferentialOps.jl:162-190): the operators D^m .. D^n as a tuple-like range."""

    def __init__(self, m, n):
        self.m, self.n = int(m), int(n)

    def __iter__(self):
        return iter(Derivative(d) for d in range(self.m, self.n + 1))

    def __len__(self):
        return max(0, self.n - self.m + 1)

    def __eq__(self, o):
        return isinstance(o, DerivativeUnitRange) and (o.m, o.n) == (self.m, self.n)

    def __repr__(self):
        return "Derivative(%d:%d)" % (self.m, self.n)


class Derivative(DifferentialOp):
    def __new__(cls, n=1, *rest):
        if isinstance(n, range):                           # Derivative(2:5) (DifferentialOps.jl:170-176)
            return DerivativeUnitRange(n.start, n.stop - 1)
        return super().__new__(cls)

    def __init__(self, n=1):
        super().__init__({int(n): 1.0})
        self.n = int(n)

    def __pow__(self, p):                                  # D2^3 === Derivative(6) (DifferentialOps.jl:95-96)
        return Derivative(self.n * int(p))

    def __mul__(self, S):
        if isinstance(S, (Spline, SplineInterpolation)):
            return S.derivative(self.n)                                 # spline.jl:272-282
        if isinstance(S, (int, float)):
            return S * DifferentialOp(self.terms)
        return NotImplemented

    def __eq__(self, o):
        if isinstance(o, Derivative):
            return o.n == self.n
        return DifferentialOp.__eq__(self, o)              # 1 * D == D (DifferentialOps.jl:124-126)

    def __hash__(self):
        return hash(("D", self.n))

    def __repr__(self):
        return "D{%d}" % self.n


class Periodic:
    """BoundaryConditions.Periodic(L) (BoundaryConditions.jl:50-61)."""

    def __init__(self, L):
        self.L = L


def period(x):
    return x.L if isinstance(x, Periodic) else x.period


class Natural:
    """BoundaryConditions.Natural (BoundaryConditions.jl:36-48): generalised natural BCs,
    derivatives 2 .. k/2 vanish at the boundary (even k only)."""


def _order(k):
    return k.k if isinstance(k, BSplineOrder) else int(k)


# ---- bases ------------------------------------------------------------------------------------

def augment_knots(breaks, k):
    """augment_knots! (BSplines/knots.jl:68-84)."""
    b = np.asarray(breaks)
    return np.concatenate([np.repeat(b[:1], k - 1), b, np.repeat(b[-1:], k - 1)])


class AbstractBSplineBasis:
    _handle = None

    def __del__(self):
        h, self._handle = self._handle, None
        if h is not None:
            try:
                lib.bsk_basis_destroy(h)
            except Exception:
                pass

    def __len__(self):
        return self._len

    def __call__(self, x, *args, **kw):                                 # BSplines.jl:84-85
        return evaluate_all(self, x, *args, **kw)

    @property
    def dtype(self):
        return _NP[self._dt]


def _create_basis(kind, k, knots, recomb=None):
    dt = _dtype_code(knots.dtype)
    out = C.c_void_p()
    desc = None
    keep = None
    if recomb is not None:
        ul = np.asfortranarray(recomb["ul"], dtype=np.float64)
        lr = np.asfortranarray(recomb["lr"], dtype=np.float64)
        keep = (ul, lr)
        desc = _lib.RecombDesc(recomb["cl"], recomb["cr"], recomb["nl"], recomb["nr"],
                               ul.shape[0], lr.shape[0],
                               ul.ctypes.data_as(C.POINTER(C.c_double)),
                               lr.ctypes.data_as(C.POINTER(C.c_double)))
    kn = np.ascontiguousarray(knots)
    check(lib.bsk_basis_create(kind, k, dt, kn.ctypes.data_as(C.c_void_p), kn.size,
                               C.byref(desc) if desc is not None else None, _device(), C.byref(out)))
    del keep
    return out, dt


class BSplineBasis(AbstractBSplineBasis):
    """BSplineBasis(BSplineOrder(k), knots; augment=Val(true))  (BSplines/basis.jl:70-90)."""

    def __init__(self, order, knots, augment=True):
        k = _order(order)
        if k <= 0:
            raise ArgumentError("B-spline order must be k ≥ 1")
        t = np.asarray(knots)
        if not np.issubdtype(t.dtype, np.floating):
            t = t.astype(np.float64)
        self.k = k
        self.t = augment_knots(t, k) if augment else t.copy()
        self._len = self.t.size - k
        self._handle, self._dt = _create_basis(BASIS_PLAIN, k, self.t)

    def knots(self):
        return self.t

    def boundaries(self):                                               # basis.jl:217-222
        return self.t[self.k - 1], self.t[self._len]


class PeriodicBSplineBasis(AbstractBSplineBasis):
    """PeriodicBSplineBasis(BSplineOrder(k), ts) with ts including the endpoint
    (BSplines/periodic.jl:181-205)."""

    def __init__(self, order, breaks):
        k = _order(order)
        if k <= 0:
            raise ArgumentError("B-spline order must be k ≥ 1")
        t = np.asarray(breaks)
        if not np.issubdtype(t.dtype, np.floating):
            t = t.astype(np.float64)
        self.k = k
        self.t = t.copy()
        self._len = t.size - 1
        self.period = t[-1] - t[0]
        self._handle, self._dt = _create_basis(BASIS_PERIODIC, k, self.t)

    def knots(self):
        return self.t

    def boundaries(self):
        return self.t[0], self.t[-1]


def _bfun_value(k, t, i, x):
    """_evaluate(::BSplineOrder{k}, t, i, x, T), i 1-based (BSplines/basis_function.jl:200-256)."""
    T = t.dtype.type
    tend = t[-1]

    def supp(ta, tb):
        return (ta <= x < tb) or (x == tb == tend)
    if k == 1:
        return T(1) if supp(t[i - 1], t[i]) else T(0)
    ta, tb = t[i - 1], t[i + k - 1]
    if not supp(ta, tb):
        return T(0)
    ta1, tb1 = t[i], t[i + k - 2]
    y = T(0)
    if tb1 != ta:
        y = y + _bfun_value(k - 1, t, i, x) * (x - ta) / (tb1 - ta)
    if ta1 != tb:
        y = y + _bfun_value(k - 1, t, i + 1, x) * (tb - x) / (tb - ta1)
    return y


def _bfun_diff(nd, k, t, i, x):
    """evaluate_diff(::Derivative{N}, ...) (BSplines/basis_function.jl:138-170)."""
    if nd == 0:
        return _bfun_value(k, t, i, x)
    T = t.dtype.type
    y = T(0)
    dt = t[i + k - 2] - t[i - 1]
    if dt != 0:
        y = y + _bfun_diff(nd - 1, k - 1, t, i, x) / dt
    dt = t[i + k - 1] - t[i]
    if dt != 0:
        y = y - _bfun_diff(nd - 1, k - 1, t, i + 1, x) / dt
    return y * T(k - 1)


def _eval_op(op, k, t, i, x, left):
    """evaluate(B, i, x, dot(op, LeftNormal())) (DifferentialOps.jl:128-131,160;
    BSplines/basis_function.jl:178-191): each term alpha * D^m b_i, summed left to right;
    odd derivatives change sign on the left boundary."""
    T = t.dtype.type
    s = None
    for m, a in op.terms.items():
        am = -a if (left and m % 2 == 1) else a
        v = T(am) * _bfun_diff(m, k, t, i, x)
        s = v if s is None else s + v
    return s


def _make_submatrix(ops, B, side):
    """_make_submatrix (Recombinations/matrices.jl:219-313, 331-357).
    Returns (ndrop, nconstraints, A) with A the corner block (rows x recombined functions)."""
    T = B.t.dtype.type
    ops = list(ops)
    nc = len(ops)
    if nc == 0:
        return 0, 0, np.zeros((0, 0))                                  # free boundary (:219-223)
    for o in ops:
        if not isinstance(o, DifferentialOp):
            raise ArgumentError("boundary condition must be a differential operator")
    if max(o.max_order for o in ops) >= B.k:
        raise ArgumentError("cannot resolve operators with B-splines of order %d" % B.k)   # :321-329
    last = ops[-1]
    is_d = lambda o, n: isinstance(o, Derivative) and o.n == n
    if nc == 1 and is_d(last, 0):
        return 1, 1, np.zeros((1, 0))                                  # Dirichlet (:227-231)
    if nc == 1 and is_d(last, 1):
        return 0, 1, np.ones((2, 1))                                   # Neumann (:234-238)
    if nc == 1:
        ndrop = 0                                                      # :332
    else:
        for j, o in enumerate(ops[:-1]):
            if not is_d(o, j):
                raise ArgumentError("unsupported combination of boundary conditions")   # :338-343
        m = nc - 1
        if is_d(last, m):
            ndrop = m + 1
        elif last.max_order <= m:
            raise ArgumentError("order of last operator is too low")   # :351-354
        else:
            ndrop = m
    nmax = last.max_order
    if nmax + 1 == ndrop:
        return ndrop, nc, np.zeros((nc, 0))                            # (D0..D(n-1)) (:263-271)
    q = nmax - ndrop
    A = np.zeros((q + nc, q), dtype=np.float64)
    k, t = B.k, B.t
    if side == "left":
        x = t[k - 1]
        idx = [ndrop + d for d in range(1, q + 2)]
        bs = [_eval_op(last, k, t, i, x, True) for i in idx]
        for l in range(1, q + 1):
            b0, b1 = bs[l - 1], bs[l]
            r = T(2) / (b0 - b1)
            A[l + nc - 2, l - 1] = -b1 * r
            A[l + nc - 1, l - 1] = b0 * r
    else:
        N = len(B)
        x = t[N]
        M = N - ndrop
        idx = [M - d + 1 for d in range(1, q + 2)]
        bs = [_eval_op(last, k, t, i, x, False) for i in idx]
        for l in range(1, q + 1):
            b0, b1 = bs[q - l + 1], bs[q - l]
            r = T(2) / (b0 - b1)
            A[l - 1, l - 1] = -b1 * r
            A[l, l - 1] = b0 * r
    return ndrop, nc, A


def _natural_submatrix(B, side):
    """_make_submatrix(::Natural, ...) (Recombinations/natural.jl:9-134).  The derivatives of the
    boundary B-splines come from the device evaluate_all (reference arithmetic); the two small
    (h+1) x (h+1) systems are solved on the host (the reference uses StaticArrays `\\`)."""
    k = B.k
    if k % 2 == 1:
        raise ArgumentError("`Natural` boundary condition only supported for even-order splines (got k = %d)" % k)
    T = B.t.dtype.type
    h = k // 2
    N = len(B)
    x = B.t[k - 1] if side == "left" else B.t[N]
    js = list(range(1, h + 2)) if side == "left" else list(range(N - h, N + 1))
    D = np.zeros((h - 1, h + 1), dtype=B.t.dtype)                     # _natural_eval_derivatives
    for i in range(1, h):
        jlast, bs = evaluate_all(B, float(x) if T is np.float64 else T(x), Derivative(i + 1))
        for n, j in enumerate(js):
            dj = jlast + 1 - j
            if 1 <= dj <= k:
                D[i - 1, n] = bs[dj - 1]
    rhs = np.zeros(h + 1, dtype=B.t.dtype)
    rhs[0] = h
    cols = []
    for which in (1, 2):                                               # _natural_system_matrix
        M = np.zeros((h + 1, h + 1), dtype=B.t.dtype)
        M[0, :] = 1
        for i in range(2, h + 1):
            d = D[i - 2]
            M[i - 1] = d / np.max(np.abs(d))
        if which == 1:
            M[h, h] = 1
        else:
            M[h, 0] = 1
        v = np.linalg.solve(M, rhs)
        eps_ = 10 * np.finfo(B.t.dtype).eps * np.max(np.abs(v))       # _remove_near_zeros
        v[np.abs(v) < eps_] = 0
        cols.append(v)
    return 0, h - 1, np.stack(cols, axis=1).astype(np.float64)


class RecombinedBSplineBasis(AbstractBSplineBasis):
    """RecombinedBSplineBasis(B, ops) / (B, ops_left, ops_right)  (Recombinations/bases.jl:223-245).

    ops: a DifferentialOp, a tuple of them, or () / None for a free boundary."""

    def __init__(self, B, ops_left, ops_right="same"):
        if not isinstance(B, BSplineBasis):
            raise ArgumentError("parent must be a BSplineBasis")
        if ops_right == "same":
            ops_right = ops_left
        norm = lambda o: () if o is None else (o if isinstance(o, Natural) else
                                               ((o,) if isinstance(o, DifferentialOp) else tuple(o)))
        self.parent = B
        self.k = B.k
        self.t = B.t
        self.ops = (norm(ops_left), norm(ops_right))
        sub = lambda ops, side: (_natural_submatrix(B, side) if isinstance(ops, Natural)
                                 else _make_submatrix(ops, B, side))
        ldrop, cl, ul = sub(self.ops[0], "left")
        rdrop, cr, lr = sub(self.ops[1], "right")
        self.cl, self.cr = cl, cr
        self.nl, self.nr = ul.shape[1], lr.shape[1]
        self.ul, self.lr = ul, lr
        self._len = len(B) - cl - cr
        self._handle, self._dt = _create_basis(
            BASIS_RECOMBINED, self.k, self.t,
            dict(cl=cl, cr=cr, nl=self.nl, nr=self.nr, ul=ul, lr=lr))

    def knots(self):
        return self.t

    def boundaries(self):
        return self.parent.boundaries()

    def num_constraints(self):
        return self.cl, self.cr

    def recombination_matrix(self):
        """Dense RecombineMatrix (N x M) for tests (Recombinations/matrices.jl:416-452)."""
        N, M = len(self.parent), len(self)
        A = np.zeros((N, M))
        for j in range(self.nl):
            A[:self.ul.shape[0], j] = self.ul[:, j]
        for j in range(self.nl, M - self.nr):
            A[j + self.cl, j] = 1.0
        h = M - self.nr
        for j in range(self.nr):
            A[h + self.cl:h + self.cl + self.lr.shape[0], h + j] = self.lr[:, j]
        return A


def knots(B):
    return B.knots()


def basis(obj):
    """basis(S) / basis(itp) / basis(b::BasisFunction) (Splines/spline.jl:99, SplineInterpolations.jl:75)."""
    return obj.basis


def coefficients(S):
    """coefficients(S) (Splines/spline.jl:93-97)."""
    return S.coefficients()


def spline(obj):
    """spline(itp) / spline(extrapolation) / spline(approximation) (SplineInterpolations.jl:77)."""
    return obj.spline


def has_parent_basis(B):
    """has_parent_basis(B): true for bases derived from a B-spline basis (BSplines/basis.jl:258-266)."""
    return isinstance(B, RecombinedBSplineBasis)


def constraints(R):
    """constraints(R) -> (left, right) differential operators (Recombinations/bases.jl:302-312); () for plain bases."""
    return R.ops if isinstance(R, RecombinedBSplineBasis) else ((), ())


def recombination_matrix(R):
    """recombination_matrix(R) as a dense array (Recombinations/bases.jl:281-286)."""
    return R.recombination_matrix()


def knots_in_main_period(B):
    """knots_in_main_period(ts::PeriodicKnots): the N knots of one period (BSplines/periodic.jl:68-75)."""
    return B.knots()[:-1]


def basis_to_array_index(B, i):
    """basis_to_array_index(B, axes, i), 1-based: identity for plain / recombined bases, mod1(i, N) for periodic
    ones (BSplines/basis.jl:268-283, periodic.jl:245-253)."""
    if isinstance(B, PeriodicBSplineBasis):
        return (int(i) - 1) % len(B) + 1
    return int(i)


def order(B):
    return B.k


def boundaries(B):
    return B.boundaries()


def find_knot_interval(B, x):
    """find_knot_interval(knots(B), x) broadcast over x (BSplines/evaluate_all.jl:102-119,
    periodic.jl:85-100).  Returns (i, zone): int64 / int8 arrays (CUDA tensors if x is one)."""
    xd, on_dev = _to_device(x, B._dt)
    xd = xd.reshape(-1)
    idx = torch.empty(xd.numel(), dtype=torch.int64, device=xd.device)
    zone = torch.empty(xd.numel(), dtype=torch.int8, device=xd.device)
    check(lib.bsk_find_knot_interval(B._handle, _ptr(xd), xd.numel(), _ptr(idx), _ptr(zone), _stream()))
    if on_dev:
        return idx, zone
    return idx.cpu().numpy(), zone.cpu().numpy()


def evaluate_all(B, x, op=None, T=None, ileft=None):
    """evaluate_all(B, x, [op], [T]; [ileft]) -> (i, bs)  (BSplines/evaluate_all.jl:9-64).

    x may be a scalar (returns (int, tuple) like the reference) or a vector / CUDA tensor
    (returns int64[n] and T[n, k], reversed order per row as in the reference)."""
    if op is not None and not isinstance(op, DifferentialOp):
        op, T = None, op                                                # evaluate_all(B, x, T)
    nd = 0 if op is None else op.n
    scalar = np.isscalar(x)
    xin = [x] if scalar else x
    out_dt = B._dt if T is None else _dtype_code(T)
    xd, on_dev = _to_device(xin, B._dt)
    xd = xd.reshape(-1)
    n = xd.numel()
    idx = torch.empty(n, dtype=torch.int64, device=xd.device)
    bs = torch.empty((n, B.k), dtype=_TORCH[out_dt], device=xd.device)
    il = None
    if ileft is not None:
        il, _ = (ileft, True) if isinstance(ileft, torch.Tensor) and ileft.is_cuda else (
            torch.as_tensor(np.atleast_1d(np.asarray(ileft, dtype=np.int64))).cuda(), False)
        if il.numel() != n:
            raise DimensionMismatch("ileft must have one entry per point")
    check(lib.bsk_evaluate_all(B._handle, _ptr(xd), n, nd, out_dt, _ptr(il) if il is not None else None,
                               _ptr(idx), _ptr(bs), _stream()))
    if scalar:
        return int(idx[0].item()), tuple(bs[0].cpu().numpy().tolist()) if out_dt == BSK_F64 else \
            tuple(np.float32(v) for v in bs[0].cpu().numpy())
    if on_dev:
        return idx, bs
    return idx.cpu().numpy(), bs.cpu().numpy()


# ---- single basis functions (BSplines/basis_function.jl) ----------------------------------------

def support(B, i=None):
    """support(B, i) / support(b::BasisFunction) -> (first, last) knot indices, 1-based: b_i lives on
    [t_i, t_{i+k}) (basis_function.jl:58-72)."""
    if isinstance(B, BasisFunction):
        B, i = B.basis, B.i
    return int(i), int(i) + B.k


def nonzero_in_segment(B, n):
    """nonzero_in_segment(B::BSplineBasis, n) -> (first, last) basis functions that are non-zero on the knot
    segment [t_n, t_{n+1}] (basis_function.jl:74-99)."""
    return max(n - B.k + 1, 1), min(n, len(B))


def isindomain(B, x):
    """isindomain(B, x): a <= x <= b; always true for periodic bases (BSplines/basis.jl:232-237, periodic.jl:81)."""
    if isinstance(B, PeriodicBSplineBasis):
        return True
    a, b = B.boundaries()
    return bool(a <= x <= b)


def evaluate(B, i, x, op=None, T=None):
    """evaluate(B, i, x, [op], [T]) — the i-th basis function at x (scalar, array or CUDA tensor)
    (basis_function.jl:118-197).  The reference recurses on the order (Cox-de Boor); here the value is picked out
    of the k non-zero functions the batched evaluate_all returns for every point (same numbers up to rounding,
    one pass over the points whatever k).  i outside eachindex(B) raises DomainError (:136-138)."""
    N, k = len(B), B.k
    if not 1 <= int(i) <= N:
        raise DomainError("Basis function index must be in 1:%d" % N)
    scalar = np.isscalar(x)
    xin = np.atleast_1d(np.asarray(x)) if not isinstance(x, torch.Tensor) else x
    idx, bs = evaluate_all(B, xin, op, T)
    on_dev = isinstance(bs, torch.Tensor)
    if not on_dev:
        idx, bs = torch.from_numpy(idx), torch.from_numpy(bs)
    j = idx - int(i)                                        # bs[:, j] is b_{idx - j}: reversed order (evaluate_all.jl:9-64)
    if isinstance(B, PeriodicBSplineBasis):
        j = torch.remainder(j, N)
    valid = (j >= 0) & (j < k)
    out = torch.where(valid, bs.gather(1, j.clamp(0, k - 1).reshape(-1, 1)).reshape(-1), torch.zeros((), dtype=bs.dtype, device=bs.device))
    if on_dev:
        return out.reshape(x.shape)
    out = out.numpy()
    return out[0] if scalar else out.reshape(np.shape(x))


class BasisFunction:
    """BasisFunction(B, i, [T]) == B[i]; callable: b(x, [op]) (basis_function.jl:1-47,114)."""

    def __init__(self, B, i, T=None):
        self.basis, self.i, self.T = B, int(i), T

    def __call__(self, x, op=None):
        if isinstance(op, DerivativeUnitRange):             # b(x, Derivative(0:2)) -> tuple (basis_function.jl:165-175)
            return tuple(evaluate(self.basis, self.i, x, d, self.T) for d in op)
        return evaluate(self.basis, self.i, x, op, self.T)

    def knots(self):
        return self.basis.knots()

    @property
    def k(self):
        return self.basis.k


def _basis_getitem(self, i):
    """B[i] (BSplines/BSplines.jl:75-82): BoundsError (IndexError) outside 1:length(B)."""
    if not 1 <= int(i) <= len(self):
        raise IndexError("attempt to access %d-element basis at index [%d]" % (len(self), int(i)))
    return BasisFunction(self, i)


AbstractBSplineBasis.__getitem__ = _basis_getitem
AbstractBSplineBasis.__repr__ = lambda self: "%d-element %s of order %d, domain [%s, %s]" % (
    (len(self), type(self).__name__, self.k) + tuple(self.boundaries()))


# ---- splines ----------------------------------------------------------------------------------

class Spline:
    """Spline(B, coefs) (Splines/spline.jl:41-60); callable (spline.jl:144)."""

    def __init__(self, B, coefs=None, _handle=None, _basis_info=None):
        self.basis = B
        if _handle is not None:
            self._handle = _handle
            self._dt, self.k = _basis_info
            return
        self._dt = B._dt
        self.k = B.k
        c = np.ascontiguousarray(np.asarray(coefs, dtype=_NP[B._dt]))
        if c.ndim != 1 or c.size != len(B):
            raise DimensionMismatch("wrong number of coefficients")     # spline.jl:52-56
        out = C.c_void_p()
        check(lib.bsk_spline_create(B._handle, c.ctypes.data_as(C.c_void_p), c.size, C.byref(out)))
        self._handle = out
        self._coefs = c

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                lib.bsk_spline_destroy(h)
            except Exception:
                pass

    def order(self):
        return self.k

    def _meta(self):
        k, nc, nk = C.c_int32(), C.c_int64(), C.c_int64()
        check(lib.bsk_spline_order(self._handle, C.byref(k), C.byref(nc), C.byref(nk)))
        return k.value, nc.value, nk.value

    def coefficients(self):
        """Coefficients in the evaluation basis (parent basis for recombined splines)."""
        _, nc, _ = self._meta()
        out = np.empty(nc, dtype=_NP[self._dt])
        check(lib.bsk_spline_coefficients(self._handle, out.ctypes.data_as(C.c_void_p), nc))
        return out

    def knots(self):
        _, _, nk = self._meta()
        out = np.empty(nk, dtype=_NP[self._dt])
        check(lib.bsk_spline_knots(self._handle, out.ctypes.data_as(C.c_void_p), nk))
        return out

    def set_coefficients_device(self, d_coefs, stride=1):
        """Re-point at device-resident coefficients (e.g. a column of a batched solve)."""
        n = len(self.basis)
        check(lib.bsk_spline_set_coefs_device(self._handle, _ptr(d_coefs), n, stride, _stream()))

    def __len__(self):                                                  # length(S) == length(coefficients(S))
        return int(self.coefficients().size)

    def __eq__(self, o):                                                # spline.jl:105-107: same basis, same coefficients
        return isinstance(o, Spline) and self.k == o.k and np.array_equal(self.knots(), o.knots()) and \
            np.array_equal(self.coefficients(), o.coefficients())

    __hash__ = object.__hash__

    def derivative(self, n=1):
        """Derivative(n) * S (Splines/spline.jl:272-330)."""
        if n == 0:
            return self
        out = C.c_void_p()
        check(lib.bsk_spline_derivative(self._handle, int(n), C.byref(out)))
        return Spline(self.basis, _handle=out, _basis_info=(self._dt, self.k - n))

    def table_info(self):
        """Diagnostics of the evaluation tables of the default path (bsk_spline_table_info): how many cells /
        interval records there are and how many of them could not be certified to the tolerance."""
        info = (C.c_int64 * 8)()
        geom = (C.c_double * 4)()
        check(lib.bsk_spline_table_info(self._handle, info, geom))
        names = ("intervals", "intervals_deboor", "cells", "knot_cells", "cells_interval_table", "cells_in_global",
                 "smem_bytes", "auto_uses_cells")
        d = dict(zip(names, [int(v) for v in info]))
        d.update(x0=geom[0], b=geom[1], cell_width=geom[2], scale=geom[3])
        return d

    def evaluate(self, x, derivs=(0,), algo=EVAL_AUTO, out=None, extrap=0):
        """Evaluate D^d S for every d in `derivs` at all points, in one pass over x.
        x: scalar, numpy array (host path, pipelined H2D/kernel/D2H) or CUDA tensor.
        extrap: _lib.EXTRAP_* (see SplineExtrapolation)."""
        derivs = [int(d) for d in derivs]
        nout = len(derivs)
        dv = (C.c_int32 * nout)(*derivs)
        scalar = np.isscalar(x)
        if extrap and not (isinstance(x, torch.Tensor) and x.is_cuda):
            xt = torch.as_tensor(np.atleast_1d(np.asarray(x, dtype=_NP[self._dt]))).cuda()
            res = [o.cpu().numpy() for o in self.evaluate(xt, derivs, algo, None, extrap)]
            return [o[0] for o in res] if scalar else res
        if isinstance(x, torch.Tensor) and x.is_cuda:
            if x.dtype != _TORCH[self._dt]:
                raise ArgumentError("device vector has dtype %s, expected %s" % (x.dtype, _TORCH[self._dt]))
            xd = x.contiguous().reshape(-1)
            outs = out if out is not None else [torch.empty_like(xd) for _ in derivs]
            ptrs = (C.c_void_p * nout)(*[o.data_ptr() for o in outs])
            check(lib.bsk_spline_eval_extrap(self._handle, _ptr(xd), xd.numel(), nout, dv, ptrs, algo, int(extrap),
                                             _stream()))
            return [o.reshape(x.shape) for o in outs]
        xh = np.ascontiguousarray(np.atleast_1d(np.asarray(x, dtype=_NP[self._dt])))
        outs = out if out is not None else [np.empty_like(xh) for _ in derivs]
        ptrs = (C.c_void_p * nout)(*[o.ctypes.data for o in outs])
        _device()
        check(lib.bsk_spline_eval_host(self._handle, xh.ctypes.data_as(C.c_void_p), xh.size, nout, dv,
                                       ptrs, algo))
        if scalar:
            return [o[0] for o in outs]
        return outs

    def __call__(self, x, algo=EVAL_AUTO):
        return self.evaluate(x, (0,), algo)[0]


class ComplexSpline:
    """Spline with complex coefficients (eltype ComplexF32 / ComplexF64 in the reference: the de Boor recurrence
    is linear in the coefficients, Splines/spline.jl:180-184): real and imaginary part are two splines sharing
    one basis, evaluated together by the vector-valued kernel (M = 2)."""

    def __init__(self, B, coefs):
        self.basis = B
        self.coefs = coefs                                   # (N,) complex CUDA tensor
        self._batch = SplineBatch(B, torch.view_as_real(coefs).contiguous())

    def coefficients(self):
        return self.coefs.cpu().numpy()

    def __call__(self, x):
        on_dev = isinstance(x, torch.Tensor) and x.is_cuda
        xd = x if on_dev else torch.from_numpy(np.ascontiguousarray(np.atleast_1d(x), dtype=_NP[self.basis._dt])).cuda()
        out = torch.view_as_complex(self._batch(xd.reshape(-1)).contiguous()).reshape(xd.shape)
        if on_dev:
            return out
        out = out.cpu().numpy()
        return out[0] if np.isscalar(x) else out


def _periodic_knot(data, k, i):
    """getindex(::PeriodicKnots, i), i 1-based (BSplines/periodic.jl:102-115): repeated +-L, in the knot type."""
    T = data.dtype.type
    N = data.size - 1
    L = T(data[-1] - data[0])
    x = T(0)
    i -= k // 2
    while i < 1:
        x = T(x - L)
        i += N
    while i > N + 1:
        x = T(x + L)
        i -= N
    return T(x + data[i - 1])


def diff(S, op=None):
    """diff(S, Derivative(n)) == Derivative(n) * S (Splines/spline.jl:272-282); S may be a SplineInterpolation."""
    n = 1 if op is None else op.n
    return S.derivative(n)


def integral(S):
    """integral(S) (Splines/spline.jl:354-374; periodic: Splines/periodic.jl:122-140): an antiderivative of S as a
    spline of order k + 1 on basis_integral(B) (BSplines/basis.jl:113-122, periodic.jl:264-276), zero at the left
    boundary.  O(N) host code like Derivative(n) * S, same operation order as the reference; the result is a
    device spline like any other.  Periodic splines give a non-periodic spline over one period plus k knots."""
    if isinstance(S, SplineInterpolation):
        S = S.spline
    k = S.k
    if k + 1 > 8:
        raise _lib.UnsupportedError("integral of an order-8 spline has order 9 (orders 1..8 are supported)")
    T = _NP[S._dt]
    u = S.coefficients()                     # evaluation-basis coefficients (parent basis for recombined splines)
    t = S.knots()
    periodic = u.size == t.size - 1          # stored knots of a periodic spline: N + 1 breakpoints
    if not periodic:
        beta = np.zeros(u.size + 1, dtype=T)
        np.cumsum(u * (t[k:] - t[:-k]) / T(k), out=beta[1:])       # beta[i+1] = beta[i] + u[i] (t[i+k] - t[i]) / k
        t_int = np.concatenate([t[:1], t, t[-1:]])
        return Spline(BSplineBasis(k + 1, t_int, augment=False), beta)
    N = u.size
    off = k // 2
    t_int = np.array([_periodic_knot(t, k, i) for i in range(1 + off - k, 1 + off + N + k + 1)], dtype=T)
    Np = t_int.size - (k + 1)
    delta = off - k + 1
    beta = np.zeros(Np, dtype=T)
    for i in range(1, Np):
        j = i + delta
        uj = u[(j - 1) % N]
        beta[i] = T(beta[i - 1] + T(T(uj * T(_periodic_knot(t, k, j + k) - _periodic_knot(t, k, j))) / T(k)))
    Bp = BSplineBasis(k + 1, t_int, augment=False)
    Sp = Spline(Bp, beta)
    xleft = t_int[k]                         # boundaries(B')[1] = t[k + 1]
    beta = (beta - T(Sp(np.array([xleft], dtype=T), algo=EVAL_DEBOOR)[0])).astype(T)
    return Spline(Bp, beta)


class SplineBatch:
    """M splines sharing one basis: Spline(B, coefs) with eltype(coefs) <: SVector{M}
    (Splines/spline.jl:180-184).  coefs: (length(B), M) CUDA tensor, M fastest — exactly what a
    batched interpolate! leaves on the device.  Calling it returns an (n, M) tensor."""

    def __init__(self, B, coefs):
        if not (isinstance(coefs, torch.Tensor) and coefs.is_cuda):
            coefs = torch.as_tensor(np.ascontiguousarray(coefs, dtype=_NP[B._dt])).cuda()
        if coefs.dim() != 2 or coefs.shape[0] != len(B) or coefs.dtype != _TORCH[B._dt]:
            raise DimensionMismatch("wrong number of coefficients")
        self.basis, self.coefs = B, coefs.contiguous()
        self._eval_basis = B
        if isinstance(B, RecombinedBSplineBasis):
            # parent_coefficients (Recombinations/splines.jl:32-65): corner blocks + shifted identity
            Mr, P = len(B), B.parent
            par = torch.zeros((len(P), coefs.shape[1]), dtype=coefs.dtype, device=coefs.device)
            ul = torch.as_tensor(B.ul, dtype=coefs.dtype, device=coefs.device)
            lr = torch.as_tensor(B.lr, dtype=coefs.dtype, device=coefs.device)
            par[:ul.shape[0]] = ul @ self.coefs[:B.nl]
            par[B.nl + B.cl:Mr - B.nr + B.cl] = self.coefs[B.nl:Mr - B.nr]
            h = Mr - B.nr + B.cl
            par[h:h + lr.shape[0]] += lr @ self.coefs[Mr - B.nr:]
            self._eval_basis, self._eval_coefs = P, par
        else:
            self._eval_coefs = self.coefs

    def __call__(self, x):
        xd, on_dev = _to_device(x, self.basis._dt)
        xd = xd.reshape(-1)
        M = self.coefs.shape[1]
        out = torch.empty((xd.numel(), M), dtype=self.coefs.dtype, device=xd.device)
        check(lib.bsk_spline_eval_multi(self._eval_basis._handle, _ptr(self._eval_coefs), self._eval_coefs.shape[0], M,
                                        _ptr(xd), xd.numel(), _ptr(out), _stream()))
        return out if on_dev else out.cpu().numpy()


# ---- collocation ----------------------------------------------------------------------------------

def greville_points(k, t, N, cl=0, force_ends=True):
    """Greville sites (t[i+1] + ... + t[i+k-1]) / (k - 1) with the reference's two-step Kahan
    sliding sum (Collocation/points.jl:56-126).  Pure host arithmetic in the knot dtype.
    cl: constraints on the left of a recombined basis (boundary points skipped);
    force_ends: put the first / last point exactly on the boundaries (regular BSplineBasis)."""
    T = t.dtype.type
    a, b = t[k - 1], t[t.size - k]
    x = np.empty(N, dtype=t.dtype)
    tsum = T(0)
    comp = T(0)
    for i in range(N):
        ii = i + cl
        if i == 0:
            tsum = T(0)
            for j in range(2, k + 1):
                tsum = tsum + t[ii + j - 1]
        else:
            prev = tsum
            y = t[ii + k - 1] - comp
            tsum = prev + y
            comp = (tsum - prev) - y
            prev = tsum
            y = -t[ii] - comp
            tsum = prev + y
            comp = (tsum - prev) - y
        xi = T(tsum / T(k - 1))
        if force_ends:                                                  # ensure_points_at_boundaries
            if i == 0:
                xi = a
            elif i == N - 1:
                xi = b
        x[i] = min(max(xi, a), b)
    return x


def _greville_raw(k, t, N):
    """GrevilleSiteIterator without end forcing / clamping (Collocation/points.jl:56-104)."""
    T = t.dtype.type
    x = np.empty(N, dtype=t.dtype)
    tsum = T(0)
    comp = T(0)
    for i in range(N):
        if i == 0:
            tsum = T(0)
            for j in range(2, k + 1):
                tsum = tsum + t[i + j - 1]
        else:
            prev = tsum
            y = t[i + k - 1] - comp
            tsum = prev + y
            comp = (tsum - prev) - y
            prev = tsum
            y = -t[i] - comp
            tsum = prev + y
            comp = (tsum - prev) - y
        x[i] = T(tsum / T(k - 1))
    return x


def collocation_points(B):
    """collocation_points(B) (Collocation/points.jl:46-126,164-225): Greville sites; periodic
    bases of even order -> the knots themselves (SameAsKnots)."""
    if isinstance(B, PeriodicBSplineBasis):
        if B.k % 2 == 0:
            return B.t[:-1].copy()                                      # points.jl:176 (SameAsKnots)
        # AvgKnots on the periodic knot vector t[i] = xi[i - k/2] (+- L), BSplines/periodic.jl:102-115
        N, k = len(B), B.k
        tp = np.array([_periodic_knot(B.t, k, i) for i in range(1, N + k + 1)], dtype=B.t.dtype)
        a, b = B.t[0], B.t[-1]
        # clamp(x, boundaries(B)...), points.jl:100-102
        return np.minimum(np.maximum(_greville_raw(k, tp, N), a), b)
    cl = B.cl if isinstance(B, RecombinedBSplineBasis) else 0
    return greville_points(B.k, B.t, len(B), cl, isinstance(B, BSplineBasis))


class CollocationMatrix:
    """CollocationMatrix band store (2k-3) x n, column-major, on the device
    (Collocation/matrix.jl:40-117), with lu! / ldiv! (matrix.jl:132-271)."""

    def __init__(self, band, _handle=None):
        self._handle = _handle
        self.factored = False
        self.f32 = False           # CollocationMatrix{Float32}: the *_f32 entry points (bsk_solve_f32.cu)
        if band is None:
            return
        if isinstance(band, torch.Tensor) and band.is_cuda:
            # torch tensor of shape (n, nbands) row-major == (nbands, n) column-major
            nb, n = band.shape[1], band.shape[0]
            self.f32 = band.dtype == torch.float32
            out = C.c_void_p()
            create = lib.bsk_banded_f32_create if self.f32 else lib.bsk_banded_create
            check(create(_ptr(band.contiguous()), n, (nb - 1) // 2, (nb - 1) // 2, _device(), C.byref(out)))
        else:
            band = np.asarray(band)
            self.f32 = band.dtype == np.float32
            bh = np.asfortranarray(band, dtype=np.float32 if self.f32 else np.float64)
            if bh.shape[0] % 2 != 1:
                raise ArgumentError("odd number of bands expected")
            nb, n = bh.shape
            out = C.c_void_p()
            create = lib.bsk_banded_f32_create_host if self.f32 else lib.bsk_banded_create_host
            check(create(bh.ctypes.data_as(C.c_void_p), n, (nb - 1) // 2, (nb - 1) // 2, _device(), C.byref(out)))
        self._handle = out
        self.n, self.nbands = n, nb

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                (lib.bsk_banded_f32_destroy if self.f32 else lib.bsk_banded_destroy)(h)
            except Exception:
                pass

    @property
    def bandwidths(self):
        h = (self.nbands - 1) // 2
        return h, h

    def data(self):
        """(nbands, n) Fortran-ordered copy of the band store (factors after lu!)."""
        out = np.empty((self.nbands, self.n), dtype=np.float32 if self.f32 else np.float64, order="F")
        check((lib.bsk_banded_f32_data if self.f32 else lib.bsk_banded_data)(self._handle, out.ctypes.data_as(C.c_void_p)))
        return out

    def dense(self):
        w = self.data()
        l, u = self.bandwidths
        A = np.zeros((self.n, self.n), dtype=w.dtype)
        for j in range(self.n):
            for i in range(max(0, j - u), min(self.n, j + l + 1)):
                A[i, j] = w[u + i - j, j]
        return A

    def mul(self, X):
        """C * X for an (n, nrhs) block (RHS index fastest): mul!(Y, C, X) (matrix.jl:91-94).
        Only before lu_() (which factorises in place, like lu!)."""
        if self.f32:
            raise _lib.UnsupportedError("mul! is implemented for Float64 collocation matrices")
        if isinstance(X, torch.Tensor) and X.is_cuda:
            Xd, on_dev = X.contiguous(), True
        else:
            Xd, on_dev = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).cuda(), False
        if Xd.dtype != torch.float64 or Xd.shape[0] != self.n:
            raise DimensionMismatch("vectors `x` and `y` must have length %d" % self.n)
        Y = torch.empty_like(Xd)
        check(lib.bsk_banded_matvec(self._handle, _ptr(Xd), _ptr(Y), Xd.numel() // self.n, _stream()))
        return Y if on_dev else Y.cpu().numpy()

    def lu_(self, mode=_lib.LU_EXACT):
        """lu!(C) (Collocation/matrix.jl:132-201).  mode LU_EXACT: BANFAC's statements in BANFAC's order
        (bit-identical factors); LU_FAST: partitioned elimination with validated overlap (Float64, n >= 1024),
        factors within ~1e-16 relative of BANFAC's."""
        info = C.c_int64(0)
        if self.f32:
            st = lib.bsk_banded_f32_lu(self._handle, C.byref(info))
        else:
            st = lib.bsk_banded_lu_mode(self._handle, int(mode), C.byref(info))
        check(st, info.value)
        self.factored = True
        return self

    @property
    def halo(self):
        h = C.c_int32(0)
        check((lib.bsk_banded_f32_halo if self.f32 else lib.bsk_banded_halo)(self._handle, C.byref(h)))
        return h.value

    def ldiv_(self, Y, algo=SOLVE_AUTO):
        """ldiv!(F, Y) in place; Y is (n, nrhs) with the RHS index fastest (matrix.jl:218-271)."""
        if self.f32:
            on_dev = isinstance(Y, torch.Tensor) and Y.is_cuda
            if on_dev:
                if Y.dtype != torch.float32 or not Y.is_contiguous():
                    raise ArgumentError("Y must be a contiguous float32 CUDA tensor")
                Yd = Y
            else:
                if not (isinstance(Y, np.ndarray) and Y.dtype == np.float32 and Y.flags.c_contiguous):
                    raise ArgumentError("Y must be a C-contiguous float32 array (solved in place)")
                Yd = torch.from_numpy(Y).cuda()
            if Yd.shape[0] != self.n:
                raise DimensionMismatch("vectors `x` and `y` must have length %d" % self.n)
            check(lib.bsk_banded_f32_solve(self._handle, _ptr(Yd), Yd.numel() // self.n if self.n else 0, algo, _stream()))
            if not on_dev:
                Y[...] = Yd.cpu().numpy()
            return Y
        if isinstance(Y, torch.Tensor) and Y.is_cuda:
            if Y.dtype != torch.float64 or not Y.is_contiguous():
                raise ArgumentError("Y must be a contiguous float64 CUDA tensor")
            if Y.shape[0] != self.n:
                raise DimensionMismatch("vectors `x` and `y` must have length %d" % self.n)
            nrhs = Y.numel() // self.n if self.n else 0
            check(lib.bsk_banded_solve(self._handle, _ptr(Y), nrhs, algo, _stream()))
            return Y
        if not (isinstance(Y, np.ndarray) and Y.dtype == np.float64 and Y.flags.c_contiguous):
            raise ArgumentError("Y must be a C-contiguous float64 array (solved in place)")
        if Y.shape[0] != self.n:
            raise DimensionMismatch("vectors `x` and `y` must have length %d" % self.n)
        _device()
        check(lib.bsk_banded_solve_host(self._handle, Y.ctypes.data_as(C.c_void_p), Y.size // self.n, algo))
        return Y

    def ldiv_mul(self, L, U, out=None):
        """out = F \\ (L * U) in one pass: the `mul!(y, L, u); ldiv!(du, F, y)` pair of a collocation
        time step (examples/heat.jl:314-328).  `self` is factorised, `L` is not; U is an (n, nrhs)
        float64 CUDA block (RHS index fastest) and is left untouched."""
        if self.f32 or L.f32:
            raise _lib.UnsupportedError("the fused mat-vec + solve is implemented for Float64 matrices")
        if not (isinstance(U, torch.Tensor) and U.is_cuda and U.dtype == torch.float64 and U.is_contiguous()):
            raise ArgumentError("U must be a contiguous float64 CUDA tensor")
        if U.shape[0] != self.n or L.n != self.n:
            raise DimensionMismatch("vectors `x` and `y` must have length %d" % self.n)
        if out is None:
            out = torch.empty_like(U)
        elif not (out.is_cuda and out.dtype == torch.float64 and out.is_contiguous() and out.shape == U.shape):
            raise ArgumentError("out must be a contiguous float64 CUDA tensor of the shape of U")
        nrhs = U.numel() // self.n if self.n else 0
        check(lib.bsk_banded_matvec_solve(self._handle, L._handle, _ptr(U), _ptr(out), nrhs, _stream()))
        return out


class CyclicTridiagonalMatrix:
    """CyclicTridiagonalMatrix(a, b, c) + ldiv! (Collocation/cyclic.jl:21-33,112-197),
    factored once."""

    def __init__(self, a, b, c):
        self.f32 = (a.dtype == torch.float32) if isinstance(a, torch.Tensor) else (np.asarray(a).dtype == np.float32)
        if self.f32:
            a, b, c = (np.ascontiguousarray(v.cpu().numpy() if isinstance(v, torch.Tensor) else v, dtype=np.float32)
                       for v in (a, b, c))
            if not (a.size == b.size == c.size):
                raise DimensionMismatch("all diagonals must have the same length")
            self.n = a.size
            out = C.c_void_p()
            check(lib.bsk_cyclic_f32_create(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                            c.ctypes.data_as(C.c_void_p), self.n, _device(), C.byref(out)))
            self.a, self.b, self.c = a, b, c
        elif isinstance(a, torch.Tensor) and a.is_cuda:
            self.n = a.numel()
            out = C.c_void_p()
            check(lib.bsk_cyclic_create_device(_ptr(a), _ptr(b), _ptr(c), self.n, _device(), C.byref(out)))
            self.a, self.b, self.c = (v.cpu().numpy() for v in (a, b, c))
        else:
            a, b, c = (np.ascontiguousarray(v, dtype=np.float64) for v in (a, b, c))
            if not (a.size == b.size == c.size):
                raise DimensionMismatch("all diagonals must have the same length")
            self.n = a.size
            out = C.c_void_p()
            check(lib.bsk_cyclic_create(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                        c.ctypes.data_as(C.c_void_p), self.n, _device(), C.byref(out)))
            self.a, self.b, self.c = a, b, c
        self._handle = out

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                (lib.bsk_cyclic_f32_destroy if self.f32 else lib.bsk_cyclic_destroy)(h)
            except Exception:
                pass

    def dense(self):
        n = self.n
        A = np.zeros((n, n))
        for i in range(n):
            A[i, i] = self.b[i]
            A[i, (i - 1) % n] = self.a[i]
            A[i, (i + 1) % n] = self.c[i]
        return A

    def ldiv_(self, Y):
        if self.f32:
            on_dev = isinstance(Y, torch.Tensor) and Y.is_cuda
            Yd = Y if on_dev else torch.from_numpy(np.ascontiguousarray(Y, dtype=np.float32)).cuda()
            if Yd.dtype != torch.float32 or not Yd.is_contiguous() or Yd.shape[0] != self.n:
                raise DimensionMismatch("all vectors must have length %d" % self.n)
            check(lib.bsk_cyclic_f32_solve(self._handle, _ptr(Yd), Yd.numel() // self.n, _stream()))
            if not on_dev:
                Y[...] = Yd.cpu().numpy()
            return Y
        if isinstance(Y, torch.Tensor) and Y.is_cuda:
            if Y.dtype != torch.float64 or not Y.is_contiguous() or Y.shape[0] != self.n:
                raise DimensionMismatch("all vectors must have length %d" % self.n)
            check(lib.bsk_cyclic_solve(self._handle, _ptr(Y), Y.numel() // self.n, _stream()))
            return Y
        Yh = np.asarray(Y, dtype=np.float64)
        if Yh.shape[0] != self.n:
            raise DimensionMismatch("all vectors must have length %d" % self.n)
        d = torch.from_numpy(np.ascontiguousarray(Yh)).cuda()
        check(lib.bsk_cyclic_solve(self._handle, _ptr(d), d.numel() // self.n, _stream()))
        Y[...] = d.cpu().numpy()
        return Y


def collocation_matrix(B, x, deriv=None, clip_threshold=None, return_outside=False):
    """collocation_matrix(B, x, [Derivative(n)]; clip_threshold = eps(T))
    (Collocation/Collocation.jl:111-120,209-251).  Periodic cubic bases give a
    CyclicTridiagonalMatrix (Collocation.jl:138-146)."""
    nd = 0 if deriv is None else deriv.n
    if B._dt == BSK_F32:
        return _collocation_matrix_f32(B, x, nd, clip_threshold, return_outside)
    clip = float(np.finfo(np.float64).eps) if clip_threshold is None else float(clip_threshold)
    xd, _ = _to_device(x, BSK_F64)
    xd = xd.reshape(-1)
    if xd.numel() != len(B):
        raise ArgumentError("wrong dimensions of collocation matrix")   # Collocation.jl:214-216
    nout = C.c_int64(0)
    if isinstance(B, PeriodicBSplineBasis):
        if B.k != 4:
            return _periodic_collocation_general(B, xd, nd, clip, return_outside)
        n = len(B)
        a = torch.empty(n, dtype=torch.float64, device="cuda")
        b = torch.empty_like(a)
        c = torch.empty_like(a)
        check(lib.bsk_collocation_cyclic(B._handle, _ptr(xd), n, nd, clip, _ptr(a), _ptr(b), _ptr(c),
                                         C.byref(nout), _stream()))
        if nout.value:
            raise ArgumentError("DomainError: %d entries outside of the matrix bands" % nout.value)
        M = CyclicTridiagonalMatrix(a, b, c)
        return (M, nout.value) if return_outside else M
    k = B.k
    nb = 2 * k - 3
    band = torch.empty((len(B), nb), dtype=torch.float64, device="cuda")   # == (nb, n) column-major
    check(lib.bsk_collocation_matrix(B._handle, _ptr(xd), xd.numel(), nd, clip, _ptr(band),
                                     C.byref(nout), _stream()))
    if nout.value:
        import warnings
        warnings.warn("Non-zero value outside of matrix bands (%d entries)" % nout.value)   # :239-244
    M = CollocationMatrix(band)
    return (M, nout.value) if return_outside else M


def collocation_rows(B, x, deriv=None, clip_threshold=None, dense=False):
    """collocation_matrix(B, x, deriv, SparseMatrixCSC{T} / Matrix{T}) for ANY number of points
    (Collocation/Collocation.jl:82-106,162-193,209-251): returns (col, val) — (nx, k) int64 / float64 CUDA tensors,
    col 1-based, 0 = no entry — or, with dense=True, the (nx, length(B)) matrix itself (column-major on the
    device, returned as a torch tensor of that shape)."""
    nd = 0 if deriv is None else deriv.n
    clip = float(np.finfo(np.float64).eps) if clip_threshold is None else float(clip_threshold)
    xd, _ = _to_device(x, BSK_F64)
    xd = xd.reshape(-1)
    nx, k = xd.numel(), B.k
    if dense:
        Ct = torch.empty((len(B), nx), dtype=torch.float64, device="cuda")      # == nx x len(B) column-major
        check(lib.bsk_collocation_rows(B._handle, _ptr(xd), nx, nd, clip, None, None, _ptr(Ct), _stream()))
        return Ct.t()
    col = torch.empty((nx, k), dtype=torch.int64, device="cuda")
    val = torch.empty((nx, k), dtype=torch.float64, device="cuda")
    check(lib.bsk_collocation_rows(B._handle, _ptr(xd), nx, nd, clip, _ptr(col), _ptr(val), None, _stream()))
    return col, val


def _collocation_matrix_f32(B, x, nd, clip_threshold, return_outside):
    """CollocationMatrix{Float32} / CyclicTridiagonalMatrix{Float32}: T = eltype of the basis' knots
    (Collocation.jl:111-120: `T = eltype(x)` default, clip_threshold = eps(T))."""
    clip = float(np.finfo(np.float32).eps) if clip_threshold is None else float(clip_threshold)
    xd, _ = _to_device(x, BSK_F32)
    xd = xd.reshape(-1)
    if xd.numel() != len(B):
        raise ArgumentError("wrong dimensions of collocation matrix")   # Collocation.jl:214-216
    nout = C.c_int64(0)
    if isinstance(B, PeriodicBSplineBasis):
        if B.k != 4:
            raise _lib.UnsupportedError("Float32 periodic interpolation is implemented for cubic splines")
        n = len(B)
        a = torch.empty(n, dtype=torch.float32, device="cuda")
        b = torch.empty_like(a)
        c = torch.empty_like(a)
        check(lib.bsk_collocation_cyclic_f32(B._handle, _ptr(xd), n, nd, clip, _ptr(a), _ptr(b), _ptr(c),
                                             C.byref(nout), _stream()))
        if nout.value:
            raise ArgumentError("DomainError: %d entries outside of the matrix bands" % nout.value)
        M = CyclicTridiagonalMatrix(a, b, c)
        return (M, nout.value) if return_outside else M
    nb = 2 * B.k - 3
    band = torch.empty((len(B), nb), dtype=torch.float32, device="cuda")   # == (nb, n) column-major
    check(lib.bsk_collocation_matrix_f32(B._handle, _ptr(xd), xd.numel(), nd, clip, _ptr(band),
                                         C.byref(nout), _stream()))
    if nout.value:
        import warnings
        warnings.warn("Non-zero value outside of matrix bands (%d entries)" % nout.value)   # :239-244
    M = CollocationMatrix(band)
    return (M, nout.value) if return_outside else M


class CyclicBandedMatrix:
    """Collocation matrix of a periodic basis of order k != 4: banded with wrap-around corners.  The
    reference uses SparseMatrixCSC + UMFPACK (Collocation/Collocation.jl:141-142); here a banded LU of
    the non-wrapped part + Woodbury correction of the corners, factored once at construction.
    cband: (2 bw + 1, n) array, A[i, j] with cyclic offset d = i - j in [shift - bw, shift + bw] at
    cband[bw + d - shift, j]; shift = offset of the band centre (odd orders: 1)."""

    def __init__(self, cband, shift=0):
        cb = np.asfortranarray(np.asarray(cband, dtype=np.float64))
        if cb.shape[0] % 2 != 1:
            raise ArgumentError("odd number of bands expected")
        self.cband = cb
        self.n = cb.shape[1]
        self.bw = (cb.shape[0] - 1) // 2
        self.shift = int(shift)
        out = C.c_void_p()
        check(lib.bsk_cyclic_banded_create(cb.ctypes.data_as(C.c_void_p), self.n, self.bw, self.shift, _device(),
                                           C.byref(out)))
        self._handle = out

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                lib.bsk_cyclic_banded_destroy(h)
            except Exception:
                pass

    def dense(self):
        n, bw = self.n, self.bw
        A = np.zeros((n, n))
        for j in range(n):
            for d in range(-bw, bw + 1):
                A[(j + d + self.shift) % n, j] += self.cband[bw + d, j]
        return A

    def ldiv_(self, Y):
        if isinstance(Y, torch.Tensor) and Y.is_cuda:
            if Y.dtype != torch.float64 or not Y.is_contiguous() or Y.shape[0] != self.n:
                raise DimensionMismatch("all vectors must have length %d" % self.n)
            check(lib.bsk_cyclic_banded_solve(self._handle, _ptr(Y), Y.numel() // self.n, _stream()))
            return Y
        Yh = np.asarray(Y, dtype=np.float64)
        if Yh.shape[0] != self.n:
            raise DimensionMismatch("all vectors must have length %d" % self.n)
        d = torch.from_numpy(np.ascontiguousarray(Yh)).cuda()
        check(lib.bsk_cyclic_banded_solve(self._handle, _ptr(d), d.numel() // self.n, _stream()))
        Y[...] = d.cpu().numpy()
        return Y


def _periodic_collocation_general(B, xd, nd, clip, return_outside):
    """collocation_matrix!(C, B::PeriodicBSplineBasis, x, Derivative(nd)) for k != 4
    (Collocation/Collocation.jl:209-251 with basis_to_array_index, BSplines/periodic.jl:245-253): the
    rows come from the device evaluate_all (K2); the scatter into the cyclic band store is host work
    done once per interpolation object."""
    n, k = len(B), B.k
    idx, bs = evaluate_all(B, xd, Derivative(nd))
    idx = idx.cpu().numpy() if isinstance(idx, torch.Tensor) else np.asarray(idx)
    bs = bs.cpu().numpy() if isinstance(bs, torch.Tensor) else np.asarray(bs)
    rows = np.repeat(np.arange(n), k)
    cols = (idx[:, None] - np.arange(k)[None, :] - 1) % n          # j = jlast + 1 - dj, wrapped, 0-based
    vals = bs.reshape(n, k)
    keep = (np.abs(vals) > clip).ravel()                           # Collocation.jl:231-234
    rows, cols, vals = rows[keep], cols.ravel()[keep], vals.ravel()[keep]
    d = (rows - cols + n // 2) % n - n // 2                        # cyclic offset
    # centre the band: odd orders have b_j peaking at x_{j+1} (offsets 0..k-1 for k = 3)
    shift = (int(d.min()) + int(d.max())) // 2 if d.size else 0
    d = d - shift
    bw = max(1, int(np.max(np.abs(d)))) if d.size else 1
    if bw > 6 or n < 4 * bw + 1:
        raise _lib.UnsupportedError("periodic collocation matrix with %d cyclic bands on %d points" % (2 * bw + 1, n))
    cband = np.zeros((2 * bw + 1, n), order="F")
    np.add.at(cband, (bw + d, cols), vals)
    M = CyclicBandedMatrix(cband, shift)
    return (M, 0) if return_outside else M


def lu(C_):
    return C_.lu_()


def ldiv(F, Y, **kw):
    return F.ldiv_(Y, **kw)


# ---- interpolation ----------------------------------------------------------------------------------

def make_knots(x, k, bc=None):
    """make_knots (SplineInterpolations/SplineInterpolations.jl:302-326, 333-371)."""
    x = np.asarray(x)
    T = np.float64 if not np.issubdtype(x.dtype, np.floating) else x.dtype.type
    x = x.astype(T)
    N = x.size
    if bc is None:
        t = np.empty(N + k, dtype=T)
        t[:k] = x[0]
        t[N:] = x[-1]
        kinv = T(1) / T(k - 1)
        for i in range(1, N - k + 1):
            ti = T(0)
            for j in range(1, k):
                ti = ti + x[i + j - 1]
            t[k + i - 1] = ti * kinv
        return t
    L = T(bc.L)
    if not (x[0] + L > x[-1]):
        raise ArgumentError("endpoint x₁ + L must *not* be included in the interpolation points")
    ts = np.empty(N + 1, dtype=T)
    if k % 2 == 0:
        ts[:-1] = x
        ts[-1] = x[0] + L
        return ts
    x0 = x[0]
    xn = x0 + L
    xprev = x0
    for i in range(1, N):
        ts[i - 1] = (xprev + x[i]) / T(2)
        xprev = x[i]
    ts[N - 1] = (xprev + xn) / T(2)
    ts[0] = x0
    ts[N] = xn
    return ts


class SplineInterpolation:
    """SplineInterpolation (SplineInterpolations.jl:56-119): basis + factorised collocation
    matrix, reusable for many data sets on the same x through interpolate_ (= interpolate!)."""

    def __init__(self, B, x, lu_mode=_lib.LU_FAST):
        """lu_mode: LU_FAST (default: partitioned factorisation from 1024 points on, within 1e-13 like the default
        solve) or LU_EXACT (BANFAC order: with algo=SOLVE_TWOPASS the coefficients are then the reference's bits)."""
        N = len(B)
        x = np.asarray(x)
        if x.size != N:
            raise DimensionMismatch("incompatible lengths of B-spline basis and collocation points")
        self.basis = B
        self.x = x
        if isinstance(B, PeriodicBSplineBasis) and B.k == 2:
            # the collocation matrix is the identity (Collocation/Collocation.jl:147-157)
            if not np.array_equal(B.t[:-1], x):
                raise ArgumentError("expected collocation points to match B-spline knots")
            self.C = None
        else:
            Cm = collocation_matrix(B, x)
            self.C = Cm if isinstance(Cm, (CyclicTridiagonalMatrix, CyclicBandedMatrix)) else Cm.lu_(lu_mode)
        self.spline = None
        self.coefs = None

    def interpolate_(self, y, algo=SOLVE_AUTO):
        """interpolate!(itp, y) (SplineInterpolations.jl:140-152).  y: length-N vector, or an
        (N, M) batch with the data-set index fastest (== Vector{SVector{M}}); CUDA tensors are
        solved in place and stay on the device."""
        N = len(self.basis)
        # Complex data (test/interpolation.jl:57: ComplexF32): ldiv! is linear with a real matrix, so a complex
        # right-hand side is two interleaved real ones — exactly the memory layout of a complex array.  The
        # coefficients come back with the data's own complex type; integer data is converted to the basis'
        # float type (T = float(eltype(y)), SplineInterpolations.jl:268-270).
        is_cplx = (isinstance(y, torch.Tensor) and y.is_complex()) or (not isinstance(y, torch.Tensor) and np.iscomplexobj(y))
        if is_cplx:
            ctype = torch.complex128 if self.basis._dt == BSK_F64 else torch.complex64
            yc = y if isinstance(y, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(y))
            if yc.shape[0] != N:
                raise DimensionMismatch("input data has incorrect length")
            src_dtype = yc.dtype
            Yc = yc.cuda().to(ctype).contiguous()
            self.interpolate_(torch.view_as_real(Yc).reshape(N, -1), algo=algo)
            self.coefs = Yc.to(src_dtype)
            self.spline = ComplexSpline(self.basis, Yc) if Yc.dim() == 1 else None
            self.splines = None
            return self
        if isinstance(y, torch.Tensor) and y.is_cuda:
            if y.shape[0] != N:
                raise DimensionMismatch("input data has incorrect length")
            if y.dtype != _TORCH[self.basis._dt]:
                raise ArgumentError("data has dtype %s, the basis is %s" % (y.dtype, _TORCH[self.basis._dt]))
            Y = y
        else:
            yh = np.asarray(y)
            if yh.shape[0] != N:
                raise DimensionMismatch("input data has incorrect length")
            Y = torch.from_numpy(np.ascontiguousarray(yh, dtype=_NP[self.basis._dt])).cuda()
        if self.C is None:
            pass                                                        # identity
        elif isinstance(self.C, (CyclicTridiagonalMatrix, CyclicBandedMatrix)):
            self.C.ldiv_(Y)
        else:
            self.C.ldiv_(Y, algo=algo)
        self.coefs = Y
        if Y.dim() == 1:
            self.spline = Spline(self.basis, Y.cpu().numpy())
        else:
            self.spline = None
        self.splines = SplineBatch(self.basis, Y) if Y.dim() == 2 else None
        return self

    def spline_of(self, column=0):
        """Spline for one data set of a batch (device-resident coefficients, strided)."""
        Y = self.coefs
        if Y.dim() == 1:
            return self.spline
        M = Y.shape[1]
        S = Spline(self.basis, np.zeros(len(self.basis)))
        S.set_coefficients_device(Y.reshape(-1)[column:], stride=M)
        return S

    def __call__(self, x, **kw):
        return self.spline(x, **kw)

    # SplineWrapper forwarders (Splines/Splines.jl: order / knots / coefficients / integral / diff of a wrapped spline)
    @property
    def k(self):
        return self.basis.k

    def knots(self):
        return self.basis.knots()

    def coefficients(self):
        return self.spline.coefficients()

    def derivative(self, n=1):
        return self.spline.derivative(n)

    def __len__(self):
        return len(self.basis)

    def __repr__(self):
        return "SplineInterpolation containing the %d-element Spline of order %d" % (len(self.basis), self.basis.k)


def interpolate(x, y, order, bc=None):
    """interpolate(x, y, BSplineOrder(k), [bc]) (SplineInterpolations.jl:260-297)."""
    k = _order(order)
    x = np.asarray(x)
    xf = x.astype(np.float64) if not np.issubdtype(x.dtype, np.floating) else x
    if bc is None:
        t = make_knots(xf, k, None)
        B = BSplineBasis(k, t, augment=False)
    elif isinstance(bc, Periodic):
        ts = make_knots(xf, k, bc)
        B = PeriodicBSplineBasis(k, ts)
    elif isinstance(bc, Natural):
        # knots = data points (SplineInterpolations.jl:275-285)
        B = RecombinedBSplineBasis(BSplineBasis(k, xf.copy()), Natural())
    else:
        raise ArgumentError("unknown boundary condition")
    itp = SplineInterpolation(B, xf)
    return itp.interpolate_(y)


# ---- extrapolation (SplineExtrapolations) ----------------------------------------------------------

class Flat:
    """Flat() (SplineExtrapolations.jl:22-33): boundary values extended to the left and right."""
    code = _lib.EXTRAP_FLAT


class Linear:
    """Linear() (SplineExtrapolations.jl:35-52): linear continuation with the boundary slope."""
    code = _lib.EXTRAP_LINEAR


class Smooth:
    """Smooth() (SplineExtrapolations.jl:54-71): the boundary polynomial pieces continued, derivatives
    up to order k - 2 stay continuous."""
    code = _lib.EXTRAP_SMOOTH


class SplineExtrapolation:
    """SplineExtrapolation(S, method) (SplineExtrapolations.jl:73-101); the method rides along in the
    evaluation kernel (bsk_spline_eval_extrap), no extra pass over the points."""

    def __init__(self, sp, method):
        if isinstance(sp, (SplineInterpolation, SplineApproximation, SplineExtrapolation)):
            sp = sp.spline                                     # SplineWrapper, SplineExtrapolations.jl:90
        if not isinstance(sp, Spline):
            raise ArgumentError("expected a Spline or a spline wrapper")
        if not isinstance(method, (Flat, Linear, Smooth)):
            raise ArgumentError("unknown extrapolation method")
        self.spline, self.method = sp, method

    def evaluate(self, x, derivs=(0,), algo=EVAL_AUTO, out=None):
        return self.spline.evaluate(x, derivs, algo, out, extrap=self.method.code)

    def __call__(self, x, algo=EVAL_AUTO):
        return self.evaluate(x, (0,), algo)[0]

    def __repr__(self):
        return "SplineExtrapolation containing the %d-element Spline of order %d, method %s" % (
            len(self.spline), self.spline.k, type(self.method).__name__)


def extrapolate(sp, method):
    """extrapolate(S, method) (SplineExtrapolations.jl:103-111)."""
    return SplineExtrapolation(sp, method)


# ---- Galerkin (src/Galerkin) -------------------------------------------------------------------------

def _gauss_legendre(n):
    """Nodes and weights on [-1, 1] (the reference takes them from FastGaussQuadrature, quadratures.jl:35-49)."""
    x, w = np.polynomial.legendre.leggauss(int(n))
    return np.ascontiguousarray(x), np.ascontiguousarray(w)


def gausslegendre(n):
    """Galerkin.gausslegendre(Val(n)) (Galerkin/quadratures.jl:3-18): (nodes, weights) on [-1, 1]."""
    return _gauss_legendre(int(n))


def galerkin_matrix(*args, quadrature=None, as_band=False):
    """galerkin_matrix([f], B, [(Derivative(m), Derivative(n))]; [quadrature]) (Galerkin/linear.jl:60-311): banded
    matrix of the integrals of f(x) D^m b_i D^n b_j, k Gauss-Legendre nodes per knot segment by default (exact for
    B-spline products).  Returns a CollocationMatrix handle (band store with k - 1 sub / super-diagonals), a
    CyclicBandedMatrix for periodic bases, or the band store itself as a CUDA tensor when as_band=True (needed for
    k = 8: 15 bands exceed the solver's 13)."""
    args = list(args)
    f = args.pop(0) if callable(args[0]) and not isinstance(args[0], AbstractBSplineBasis) else None
    B = args.pop(0)
    derivs = args.pop(0) if args else None
    if isinstance(B, (tuple, list)):
        # a pair of bases must share the parent basis (linear.jl:134-156)
        if len(B) != 2 or B[0].k != B[1].k or not np.array_equal(B[0].knots(), B[1].knots()):
            raise ArgumentError("all bases must share the same parent B-spline basis")
        if B[0] is not B[1]:
            return _galerkin_matrix_pair(f, B[0], B[1], derivs, quadrature, as_band)
        B = B[0]
    d1, d2 = (0, 0) if derivs is None else (derivs[0].n, derivs[1].n)
    qx, qw = _gauss_legendre(B.k) if quadrature is None else quadrature
    qx = np.ascontiguousarray(qx, dtype=np.float64); qw = np.ascontiguousarray(qw, dtype=np.float64)
    nb = 2 * (B.k - 1) + 1
    band = torch.empty((len(B), nb), dtype=torch.float64, device="cuda")      # == (nb, n) column-major
    fx, nfx = None, 0
    if f is not None:
        npts = C.c_int64(0)
        check(lib.bsk_galerkin_nodes(B._handle, qx.ctypes.data_as(C.c_void_p), qx.size, None, 0, C.byref(npts)))
        xs = np.empty(npts.value)
        check(lib.bsk_galerkin_nodes(B._handle, qx.ctypes.data_as(C.c_void_p), qx.size, xs.ctypes.data_as(C.c_void_p),
                                     xs.size, C.byref(npts)))
        fx = torch.from_numpy(_apply(f, xs)).cuda()
        nfx = fx.numel()
    check(lib.bsk_galerkin_matrix_weighted(B._handle, d1, d2, qx.ctypes.data_as(C.c_void_p), qw.ctypes.data_as(C.c_void_p),
                                           qx.size, _ptr(fx) if fx is not None else None, nfx, _ptr(band), _stream()))
    if as_band:
        return band
    if isinstance(B, PeriodicBSplineBasis):
        # SparseMatrixCSC in the reference (linear.jl:50-54,144): band + wrap-around corners = the cyclic band store
        return CyclicBandedMatrix(band.cpu().numpy().T, shift=0)
    return CollocationMatrix(band)


class BandedRectMatrix:
    """N1 x N2 banded matrix with l sub- and u super-diagonals, band store data[u + i - j, j] on the device — what
    galerkin_matrix((B1, B2)) returns in the reference (a BandedMatrix, Galerkin/linear.jl:176-199)."""

    def __init__(self, band, shape, l, u):
        self.band, self.shape, self.bandwidths = band, shape, (l, u)       # band: (N2, l + u + 1) tensor == column-major store

    def dense(self):
        N1, N2 = self.shape
        l, u = self.bandwidths
        w = self.band.cpu().numpy().T
        A = np.zeros((N1, N2))
        for j in range(N2):
            for i in range(max(0, j - u), min(N1, j + l + 1)):
                A[i, j] = w[u + i - j, j]
        return A

    def __matmul__(self, x):                               # Mf * fs (test/airy.jl:69); small host product
        return self.dense() @ np.asarray(x)


def _galerkin_matrix_pair(f, B1, B2, derivs, quadrature, as_band):
    d1, d2 = (0, 0) if derivs is None else (derivs[0].n, derivs[1].n)
    qx, qw = _gauss_legendre(B1.k) if quadrature is None else quadrature
    qx = np.ascontiguousarray(qx, dtype=np.float64); qw = np.ascontiguousarray(qw, dtype=np.float64)
    l, u = C.c_int32(0), C.c_int32(0)
    check(lib.bsk_galerkin_matrix_pair(B1._handle, B2._handle, d1, d2, qx.ctypes.data_as(C.c_void_p),
                                       qw.ctypes.data_as(C.c_void_p), qx.size, None, 0, None, C.byref(l), C.byref(u), None))
    band = torch.empty((len(B2), l.value + u.value + 1), dtype=torch.float64, device="cuda")
    fx, nfx = None, 0
    if f is not None:
        npts = C.c_int64(0)
        check(lib.bsk_galerkin_nodes(B1._handle, qx.ctypes.data_as(C.c_void_p), qx.size, None, 0, C.byref(npts)))
        xs = np.empty(npts.value)
        check(lib.bsk_galerkin_nodes(B1._handle, qx.ctypes.data_as(C.c_void_p), qx.size, xs.ctypes.data_as(C.c_void_p),
                                     xs.size, C.byref(npts)))
        fx = torch.from_numpy(_apply(f, xs)).cuda()
        nfx = fx.numel()
    check(lib.bsk_galerkin_matrix_pair(B1._handle, B2._handle, d1, d2, qx.ctypes.data_as(C.c_void_p),
                                       qw.ctypes.data_as(C.c_void_p), qx.size, _ptr(fx) if fx is not None else None, nfx,
                                       _ptr(band), C.byref(l), C.byref(u), _stream()))
    return band if as_band else BandedRectMatrix(band, (len(B1), len(B2)), l.value, u.value)


def galerkin_projection(f, B, deriv=None, out=None):
    """galerkin_projection(f, B, [Derivative(n)]) (Galerkin/projection.jl:1-99): phi_i = <f, D^n b_i>; f is any
    callable, a Spline included (test/galerkin.jl:73-79).  out: galerkin_projection!(f, phi, B) — a vector of the
    wrong length raises DimensionMismatch (projection.jl:52-55)."""
    nd = 0 if deriv is None else deriv.n
    if out is not None and np.asarray(out).shape != (len(B),):
        raise DimensionMismatch("incorrect length of output vector phi")
    qx, qw = _gauss_legendre(B.k)
    npts = C.c_int64(0)
    check(lib.bsk_galerkin_nodes(B._handle, qx.ctypes.data_as(C.c_void_p), qx.size, None, 0, C.byref(npts)))
    xs = np.empty(npts.value)
    check(lib.bsk_galerkin_nodes(B._handle, qx.ctypes.data_as(C.c_void_p), qx.size, xs.ctypes.data_as(C.c_void_p),
                                 xs.size, C.byref(npts)))
    fx = torch.from_numpy(_apply(f, xs)).cuda()
    phi = torch.empty(len(B), dtype=torch.float64, device="cuda")
    check(lib.bsk_galerkin_projection(B._handle, nd, qx.ctypes.data_as(C.c_void_p), qw.ctypes.data_as(C.c_void_p),
                                      qx.size, _ptr(fx), fx.numel(), _ptr(phi), _stream()))
    if out is not None:
        out[...] = phi.cpu().numpy()
        return out
    return phi.cpu().numpy()


# ---- approximation (SplineApproximations) ----------------------------------------------------------

class VariationDiminishing:
    """VariationDiminishing() (SplineApproximations/variation_diminishing.jl:1-44): c_i = f(Greville site i)."""


class ApproxByInterpolation:
    """ApproxByInterpolation(xs) / ApproxByInterpolation(B) (SplineApproximations/by_interpolation.jl:3-38)."""

    def __init__(self, xs_or_basis):
        if isinstance(xs_or_basis, AbstractBSplineBasis):
            self.xs = collocation_points(xs_or_basis)
        else:
            self.xs = np.asarray(xs_or_basis)


class MinimiseL2Error:
    """MinimiseL2Error() (SplineApproximations/minimiseL2.jl:10-77): solve M c = phi with the Galerkin mass
    matrix M and the projection phi_i = <b_i, f>."""


class SplineApproximation:
    """SplineApproximation (SplineApproximations.jl:24-48): the spline plus what is needed to approximate
    another function in the same basis without refactorising (approximate!)."""

    def __init__(self, method, basis, data):
        self.method, self.basis, self.data = method, basis, data
        self.spline = None

    def __call__(self, x, **kw):
        return self.spline(x, **kw)

    def coefficients(self):
        return self.spline.coefficients()


def _apply(f, xs):
    """f.(xs) for a Python callable: vectorised call if f accepts arrays, else point by point."""
    try:
        ys = np.asarray(f(xs), dtype=np.float64)
        if ys.shape == xs.shape:
            return ys
    except Exception:
        pass
    return np.array([f(float(x)) for x in xs], dtype=np.float64)


def approximate_(f, A):
    """approximate!(f, A) (SplineApproximations.jl:121-130)."""
    if isinstance(A.method, MinimiseL2Error):
        phi = galerkin_projection(f, A.basis)                # minimiseL2.jl:68-76
        A.data.ldiv_(phi.reshape(-1, 1))
        A.spline = Spline(A.basis, phi)
        return A
    if isinstance(A.method, VariationDiminishing):
        B = A.basis                                          # variation_diminishing.jl:35-44 (Greville sites)
        A.spline = Spline(B, _apply(f, np.asarray(collocation_points(B), dtype=np.float64)))
        return A
    itp = A.data                                             # by_interpolation.jl:47-71 -> interpolate!
    itp.interpolate_(_apply(f, np.asarray(A.method.xs, dtype=np.float64)))
    A.spline = itp.spline
    return A


def approximate(f, B, method=None):
    """approximate(f, B, [method = ApproxByInterpolation(B)]) (SplineApproximations.jl:50-119)."""
    if method is None:
        method = ApproxByInterpolation(B)
    if isinstance(method, MinimiseL2Error):
        if isinstance(B, PeriodicBSplineBasis):              # test/periodic.jl:75; factored at construction
            return approximate_(f, SplineApproximation(method, B, galerkin_matrix(B)))
        Mfact = galerkin_matrix(B).lu_()                     # cholesky!(M) in the reference; M is SPD
        return approximate_(f, SplineApproximation(method, B, Mfact))
    if isinstance(method, VariationDiminishing):
        return approximate_(f, SplineApproximation(method, B, None))
    if not isinstance(method, ApproxByInterpolation):
        raise ArgumentError("unknown approximation method")
    itp = SplineInterpolation(B, np.asarray(method.xs, dtype=np.float64))
    return approximate_(f, SplineApproximation(method, B, itp))


# ---- cubic smoothing splines (SplineInterpolations/smoothing.jl) ------------------------------------

class DomainError(ValueError):
    """Julia DomainError (smoothing.jl:49)."""


def _band_to_sparse(w, l, u):
    """Band store data[u + i - j, j] (Collocation.jl:165-198) -> scipy CSR."""
    import scipy.sparse as sp
    n = w.shape[1]
    diags, offs = [], []
    for d in range(-l, u + 1):                               # diagonal d holds A[i, i + d]
        js = np.arange(max(0, d), min(n, n + d))
        diags.append(w[u - d, js])
        offs.append(d)
    return sp.diags(diags, offs, shape=(n, n), format="csr")


def _sparse_to_band(M, bw):
    n = M.shape[0]
    w = np.zeros((2 * bw + 1, n), dtype=np.float64, order="F")
    Mc = M.tocoo()
    keep = np.abs(Mc.row - Mc.col) <= bw
    np.add.at(w, (bw + Mc.row[keep] - Mc.col[keep], Mc.col[keep]), Mc.data[keep])
    return w


def _sparse_bandwidth(M):
    Mc = M.tocoo()
    nz = Mc.data != 0
    return int(np.max(np.abs(Mc.row[nz] - Mc.col[nz]))) if nz.any() else 0


def _cyclic_sparse(B, xs, nd):
    """Periodic cubic collocation matrix (cyclic tridiagonal, Collocation/cyclic.jl:21-33) as scipy CSR."""
    import scipy.sparse as sp
    n = len(B)
    xd = torch.from_numpy(xs).cuda()
    a = torch.empty(n, dtype=torch.float64, device="cuda")
    b, c = torch.empty_like(a), torch.empty_like(a)
    nout = C.c_int64(0)
    check(lib.bsk_collocation_cyclic(B._handle, _ptr(xd), n, nd, float(np.finfo(np.float64).eps), _ptr(a), _ptr(b),
                                     _ptr(c), C.byref(nout), _stream()))
    a, b, c = (v.cpu().numpy() for v in (a, b, c))
    i = np.arange(n)
    return sp.csr_matrix((np.concatenate([a, b, c]), (np.concatenate([i, i, i]),
                                                      np.concatenate([(i - 1) % n, i, (i + 1) % n]))), shape=(n, n))


class _PeriodicSmoothing:
    """fit(...; Periodic(L)) factorised (smoothing.jl:126-213): cyclic 7-band normal matrix (K7b)."""

    def __init__(self, B, F, At2W, xs):
        self.basis, self.F, self.At2W, self.xs = B, F, At2W, xs

    def fit_(self, ys):
        n = len(self.xs)
        on_dev = isinstance(ys, torch.Tensor) and ys.is_cuda
        yh = ys.cpu().numpy() if on_dev else np.asarray(ys, dtype=np.float64)
        if yh.shape[0] != n:
            raise DimensionMismatch("x and y vectors must have the same length")
        z = np.ascontiguousarray(self.At2W @ yh.reshape(n, -1))      # 2 A' W y (smoothing.jl:196-199), once per fit
        Z = torch.from_numpy(z).cuda()
        self.F.ldiv_(Z)
        if yh.ndim == 1:
            return Spline(self.basis, Z[:, 0].cpu().numpy())
        return SplineBatch(self.basis, Z)


def _fit_periodic(xs, lam, bc, weights):
    import scipy.sparse as sp
    N = xs.size
    ts = make_knots(xs, 4, bc)                               # smoothing.jl:137-138
    B = PeriodicBSplineBasis(4, ts)
    A = _cyclic_sparse(B, xs, 0)                             # smoothing.jl:141-144
    D = _cyclic_sparse(B, xs, 2)
    L = float(bc.L)
    xm, xp = np.roll(xs, 1).copy(), np.roll(xs, -1).copy()   # x_{i-1}, x_{i+1} with the period (ts[i + h] == xs[i])
    xm[0] -= L
    xp[-1] += L
    i = np.arange(N)
    dl = sp.csr_matrix((np.concatenate([xs - xm, 2 * (xp - xm), xp - xs]),
                        (np.concatenate([i, i, i]), np.concatenate([(i - 1) % N, i, (i + 1) % N]))), shape=(N, N))
    H = dl @ D                                               # smoothing.jl:170
    Bm = H.T @ D
    W = sp.diags(weights) if weights is not None else sp.identity(N)
    M = ((Bm + Bm.T) * (lam / 6) + 2 * (A.T @ W @ A)).tocoo()
    d = (M.row - M.col + N // 2) % N - N // 2                # cyclic offsets
    keep = M.data != 0
    bw = max(1, int(np.max(np.abs(d[keep])))) if keep.any() else 1
    cband = np.zeros((2 * bw + 1, N), order="F")
    np.add.at(cband, (bw + d[keep], M.col[keep]), M.data[keep])
    F = CyclicBandedMatrix(cband, 0)
    return _PeriodicSmoothing(B, F, (2 * (A.T @ W)).tocsr(), xs)


class SmoothingSpline:
    """What fit() factorised: the natural cubic basis, the LU of the 7-band normal matrix and the
    3-band matrix 2 A' W of the right-hand side.  `fit_(Y)` fits further data sets on the same points
    (the reference rebuilds everything per call)."""

    def __init__(self, R, F, L, xs):
        self.basis, self.F, self.L, self.xs = R, F, L, xs

    def fit_(self, ys):
        n = len(self.xs)
        if isinstance(ys, torch.Tensor) and ys.is_cuda:
            if ys.dtype != torch.float64 or ys.shape[0] != n:
                raise DimensionMismatch("x and y vectors must have the same length")
            Y = ys.contiguous().reshape(n, -1)
            cs = self.F.ldiv_mul(self.L, Y)                  # 2 A' W y, then M \ . : one pass over HBM
            return SplineBatch(self.basis, cs) if ys.dim() > 1 else Spline(self.basis, cs[:, 0].cpu().numpy())
        ys = np.asarray(ys, dtype=np.float64)
        if ys.shape[0] != n:
            raise DimensionMismatch("x and y vectors must have the same length")
        Y = torch.from_numpy(np.ascontiguousarray(ys.reshape(n, -1))).cuda()
        cs = self.F.ldiv_mul(self.L, Y)
        if ys.ndim == 1:
            return Spline(self.basis, cs[:, 0].cpu().numpy())
        return SplineBatch(self.basis, cs)


def fit(order, xs, ys=None, lam=0.0, bc=None, weights=None):
    """fit(BSplineOrder(4), xs, ys, λ, [bc = Natural()]; weights) — cubic smoothing spline minimising
    sum_i w_i |y_i - S(x_i)|^2 + λ ∫ S''^2 (SplineInterpolations/smoothing.jl:1-124).

    ys: (N,) array -> Spline; (N, M) array or CUDA tensor (M data sets, data-set index fastest) ->
    SplineBatch.  ys=None returns the factorised SmoothingSpline for repeated fit_(ys) calls."""
    import scipy.sparse as sp
    k = _order(order)
    if k != 4:
        raise ArgumentError("only cubic splines (BSplineOrder(4)) are currently supported")
    if not lam >= 0:
        raise DomainError("the smoothing parameter λ must be non-negative (got %r)" % (lam,))    # smoothing.jl:49
    if bc is None:
        bc = Natural()
    if not isinstance(bc, (Natural, Periodic)):
        raise ArgumentError("the boundary condition must be Natural() or Periodic(L)")
    xs = np.ascontiguousarray(np.asarray(xs, dtype=np.float64))
    N = xs.size
    if ys is not None and np.shape(ys)[0] != N:
        raise DimensionMismatch("x and y vectors must have the same length")               # smoothing.jl:50
    if weights is not None:
        weights = np.asarray(weights, dtype=np.float64)
        if weights.shape != (N,):
            raise DimensionMismatch("the `weights` vector must have the same length as the data")   # :110
    if isinstance(bc, Periodic):
        sm = _fit_periodic(xs, lam, bc, weights)
        return sm if ys is None else sm.fit_(ys)
    # natural cubic basis with knots = data points (smoothing.jl:56-58)
    R = RecombinedBSplineBasis(BSplineBasis(4, xs.copy()), Natural())
    Ab = collocation_matrix(R, xs)                           # smoothing.jl:61-64 (3 bands are populated)
    Db = collocation_matrix(R, xs, Derivative(2))
    l, u = Ab.bandwidths
    A = _band_to_sparse(Ab.data(), l, u)
    D = _band_to_sparse(Db.data(), l, u)
    # grid-step matrix Δ (symmetric, rows 2..N-1 populated; smoothing.jl:66-74)
    dlt = sp.lil_matrix((N, N))
    for i in range(1, N - 1):
        dlt[i, i] = 2 * (xs[i + 1] - xs[i - 1])
        dlt[i, i + 1] = xs[i + 1] - xs[i]
    dlt = dlt.tocsr()
    dlt = dlt + sp.triu(dlt, 1).T                            # Hermitian(Δ_upper)
    H = dlt @ D
    Bm = H.T @ D
    W = sp.diags(weights) if weights is not None else sp.identity(N)
    M = (Bm + Bm.T) * (lam / 6) + 2 * (A.T @ W @ A)          # smoothing.jl:79-92
    bw = max(1, _sparse_bandwidth(M))
    F = CollocationMatrix(_sparse_to_band(M, bw)).lu_()
    L = CollocationMatrix(_sparse_to_band((2 * (A.T @ W)).tocsr(), 1))          # z = 2 A' W y (smoothing.jl:113-114)
    sm = SmoothingSpline(R, F, L, xs)
    return sm if ys is None else sm.fit_(ys)
