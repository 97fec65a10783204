"""GPU parity tests of the Float32 interpolation path (bsk_solve_f32.cu) through the C ABI.

The reference is generic in the element type: Float32 `x` gives a Float32 basis, a
CollocationMatrix{Float32} and Float32 lu! / ldiv! (Collocation/Collocation.jl:165-198,
Collocation/matrix.jl:132-271).  Bar: bit-exact against the Float32 restatement in oracle/ for
assembly, factors and BANSLV solutions; within 1e-5 relative of the Float64 results
(BASELINE.json north_star).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL32 = 1e-5


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


@pytest.fixture(scope="module")
def synth():
    from bsplinekit_b200 import synth as s
    return s


def _system(bk, oracle, synth, n, k, seed=6):
    xdata = synth.breakpoints(seed, n).astype(np.float32)
    t = bk.make_knots(xdata, k)
    assert t.dtype == np.float32
    B = bk.BSplineBasis(k, t, augment=False)
    Cm = bk.collocation_matrix(B, xdata)
    ref, nout = oracle.collocation_matrix_f32(0, k, t, xdata)
    assert nout == 0
    return xdata, t, B, Cm, ref


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6, 8])
def test_collocation_lu_solve_f32_bitexact(bk, oracle, synth, k):
    n, nrhs = 301, 77
    xdata, t, B, Cm, ref = _system(bk, oracle, synth, n, k)
    assert Cm.f32 and Cm.data().dtype == np.float32
    assert np.array_equal(Cm.data(), ref)
    D = Cm.dense()
    assert np.allclose(D.sum(axis=1), 1.0, atol=1e-5)                 # partition of unity
    lu_ref = ref.copy(order="F")
    assert oracle.banded_lu(lu_ref) == 0
    Cm.lu_()
    assert np.array_equal(Cm.data(), lu_ref)
    Y = (2 * synth.u_numpy(7, 0, n * nrhs) - 1).reshape(n, nrhs).astype(np.float32)
    Yref = oracle.banded_solve(lu_ref, Y.copy())
    import torch
    Yd = torch.from_numpy(Y).cuda()
    Cm.ldiv_(Yd)
    assert np.array_equal(Yd.cpu().numpy(), Yref)
    Yh = Y.copy()
    Cm.ldiv_(Yh)                                                      # host array, solved in place
    assert np.array_equal(Yh, Yref)
    # against the Float64 solve of the same data
    xd64 = xdata.astype(np.float64)
    t64 = oracle.make_knots(xd64, k)
    band64, _ = oracle.collocation_matrix(0, k, t64, xd64)
    oracle.banded_lu(band64)
    Y64 = oracle.banded_solve(band64, Y.astype(np.float64))
    cond_slack = 50.0                                                 # Float32 rounding x conditioning of the system
    assert relerr(Yref, Y64) < cond_slack * TOL32


def test_collocation_f32_derivative_and_recombined(bk, oracle, synth):
    k = 6
    br = synth.breakpoints(10, 400).astype(np.float32)
    B = bk.BSplineBasis(k, br)
    R = bk.RecombinedBSplineBasis(B, bk.Derivative(0) + 3 * bk.Derivative(1))
    rc = oracle.Recomb(R.cl, R.cr, R.nl, R.nr, R.ul, R.lr)
    x = bk.collocation_points(R).astype(np.float32)
    for nd in (0, 1, 2):
        ref, nout_ref = oracle.collocation_matrix_f32(2, k, B.knots(), x, nderiv=nd, recomb=rc)
        Cm, nout = bk.collocation_matrix(R, x, bk.Derivative(nd), return_outside=True)
        assert nout == nout_ref
        assert np.array_equal(Cm.data(), ref)
    with pytest.raises(bk.ArgumentError):
        bk.collocation_matrix(R, x[:-1])


def test_interpolate_f32_end_to_end(bk, oracle, synth):
    """interpolate(x::Vector{Float32}, y, BSplineOrder(k)): S.(xs) ≈ ys (test/interpolation.jl:61-66)
    and agreement with the Float64 interpolation of the same data."""
    n = 200
    x = synth.breakpoints(6, n).astype(np.float32)
    y = np.sin(6 * x.astype(np.float64)).astype(np.float32)
    for k in (4, 6):
        itp = bk.interpolate(x, y, bk.BSplineOrder(k))
        S = itp.spline
        assert S.coefficients().dtype == np.float32
        got = S(x)
        assert np.allclose(got, y, atol=2e-5)
        itp64 = bk.interpolate(x.astype(np.float64), y.astype(np.float64), bk.BSplineOrder(k))
        xs = synth.u_numpy(12, 0, 5000)
        v64 = itp64.spline(xs)
        v32 = S(xs.astype(np.float32))
        assert relerr(v32, v64) < 20 * TOL32
    # a batch of data sets stays on the device
    import torch
    M = 130
    Y = torch.from_numpy((2 * synth.u_numpy(7, 0, n * M) - 1).reshape(n, M).astype(np.float32)).cuda()
    Y0 = Y.cpu().numpy().copy()
    t = bk.make_knots(x, 4)
    Bi = bk.BSplineBasis(4, t, augment=False)
    itp = bk.SplineInterpolation(Bi, x)
    itp.interpolate_(Y, algo=bk.SOLVE_TWOPASS)                       # BANSLV order: the restatement's bits
    band, _ = oracle.collocation_matrix_f32(0, 4, t, x)
    oracle.banded_lu(band)
    ref = oracle.banded_solve(band, Y0.copy())
    assert np.array_equal(Y.cpu().numpy(), ref)
    Y2 = torch.from_numpy(Y0).cuda()
    itp.interpolate_(Y2)                                             # default: single pass where available
    assert relerr(Y2.cpu().numpy(), ref) < TOL32
    with pytest.raises(bk.ArgumentError):
        itp.interpolate_(Y.double())
    with pytest.raises(bk.DimensionMismatch):
        itp.interpolate_(Y[:-1])


def test_banded_f32_special_cases(bk):
    """test/collocation.jl:4-53 in Float32: 1x1, zero pivots, wrong dimensions."""
    C1 = bk.CollocationMatrix(np.array([[2.0]], dtype=np.float32))
    C1.lu_()
    y = np.array([[3.0]], dtype=np.float32)
    C1.ldiv_(y)
    assert y[0, 0] == 1.5
    with pytest.raises(bk.ZeroPivotException) as ei:
        bk.CollocationMatrix(np.array([[0.0]], dtype=np.float32)).lu_()
    assert ei.value.info == 1
    band = np.zeros((3, 4), dtype=np.float32, order="F")
    band[1, :] = [1.0, 2.0, 0.0, 4.0]
    with pytest.raises(bk.ZeroPivotException) as ei:
        bk.CollocationMatrix(band).lu_()
    assert ei.value.info == 3
    band[1, 2] = 3.0
    Cm = bk.CollocationMatrix(band).lu_()
    with pytest.raises(bk.DimensionMismatch):
        Cm.ldiv_(np.zeros((5, 2), dtype=np.float32))
    with pytest.raises(bk.ArgumentError):
        Cm.ldiv_(np.zeros((4, 2)))                                    # Float64 data on a Float32 matrix
    y = np.ones((4, 3), dtype=np.float32)
    Cm.ldiv_(y)
    assert np.allclose(y[:, 0], [1.0, 0.5, 1 / 3, 0.25], rtol=1e-6)
    # long matrix: the factorisation streams several chunks of columns through shared memory
    n, bw = 2000, 2
    rng = np.random.default_rng(5)
    big = (0.1 * rng.standard_normal((2 * bw + 1, n))).astype(np.float32)
    big[bw] = 2.0 + rng.random(n).astype(np.float32)
    big = np.asfortranarray(big)
    Cb = bk.CollocationMatrix(big.copy(order="F")).lu_()
    from oracle import oracle as O
    ref = big.copy(order="F")
    assert O.banded_lu(ref) == 0
    assert np.array_equal(Cb.data(), ref)


def test_periodic_cubic_f32(bk, oracle, synth):
    n, M = 500, 90
    br = synth.breakpoints(8, n + 1).astype(np.float32)
    B = bk.PeriodicBSplineBasis(4, br)
    x = bk.collocation_points(B).astype(np.float32)
    a, b, c, nout = oracle.collocation_cyclic_f32(4, br, x)
    Mx = bk.collocation_matrix(B, x)
    assert nout == 0 and Mx.f32
    assert np.array_equal(Mx.a, a) and np.array_equal(Mx.b, b) and np.array_equal(Mx.c, c)
    Y = (2 * synth.u_numpy(7, 0, n * M) - 1).reshape(n, M).astype(np.float32)
    ref32 = oracle.cyclic_solve_f32(a, b, c, Y.copy())
    ref64 = np.linalg.solve(Mx.dense().astype(np.float64), Y.astype(np.float64))
    import torch
    Yd = torch.from_numpy(Y.copy()).cuda()
    Mx.ldiv_(Yd)
    got = Yd.cpu().numpy()
    assert relerr(got, ref64) < TOL32
    assert relerr(got, ref32) < TOL32
    # interpolate(x, y, BSplineOrder(4), Periodic(L)) in Float32: S.(xs) ≈ ys (test/periodic.jl:176-180)
    xs = br[:-1]
    ys = np.cos(2 * np.pi * xs.astype(np.float64)).astype(np.float32)
    itp = bk.interpolate(xs, ys, bk.BSplineOrder(4), bk.Periodic(np.float32(br[-1] - br[0])))
    assert np.allclose(itp.spline(xs), ys, atol=2e-5)


def test_interpolate_f32_natural_and_derivative_rows(bk, synth):
    """Natural() interpolation in Float32 (SplineInterpolations.jl:275-285, test/natural.jl): the
    interpolant reproduces the data and its second derivative vanishes at both ends; collocation rows of
    Derivative(1) in Float32 agree with the Float64 rows."""
    n = 64
    x = synth.breakpoints(6, n).astype(np.float32)
    y = np.cos(5 * x.astype(np.float64)).astype(np.float32)
    itp = bk.interpolate(x, y, bk.BSplineOrder(4), bk.Natural())
    S = itp.spline
    assert np.allclose(S(x), y, atol=3e-5)
    d2 = (bk.Derivative(2) * S)(np.array([x[0], x[-1]], dtype=np.float32))
    scale = float(np.max(np.abs((bk.Derivative(2) * S)(x))))
    assert np.all(np.abs(d2) < 2e-3 * scale)
    itp64 = bk.interpolate(x.astype(np.float64), y.astype(np.float64), bk.BSplineOrder(4), bk.Natural())
    xs = synth.u_numpy(13, 0, 2000)
    assert relerr(S(xs.astype(np.float32)), itp64.spline(xs)) < 50 * TOL32
    B32 = bk.BSplineBasis(5, x)
    B64 = bk.BSplineBasis(5, x.astype(np.float64))
    xc = bk.collocation_points(B32).astype(np.float32)
    C32 = bk.collocation_matrix(B32, xc, bk.Derivative(1)).dense()
    C64 = bk.collocation_matrix(B64, xc.astype(np.float64), bk.Derivative(1)).dense()
    assert relerr(C32, C64) < 20 * TOL32


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6, 8])
def test_single_pass_solve_f32(bk, oracle, synth, k):
    """ldiv! on Float32 data in ONE pass over HBM (k_banded_tmem with two right-hand sides per lane, packed
    Float32 FMAs, ring in Tensor Memory): within 1e-5 of the BANSLV-order Float32 restatement of
    src/Collocation/matrix.jl:218-271 for every bandwidth, with and without row segments, ragged column counts
    (the columns past the last group of 64 take the BANSLV kernels), and against the Float64 solve."""
    import torch
    for n, nrhs in ((1000, 64), (4096, 4096 + 68), (4096, 132), (257, 192)):
        xdata, t, B, Cm, ref = _system(bk, oracle, synth, n, k, seed=6 + k)
        lu_ref = ref.copy(order="F")
        assert oracle.banded_lu(lu_ref) == 0
        Cm.lu_()
        assert np.array_equal(Cm.data(), lu_ref)
        Y = (2 * synth.u_numpy(7, 0, n * nrhs) - 1).reshape(n, nrhs).astype(np.float32)
        cols = np.r_[0:40, nrhs - 70:nrhs]
        Yref = oracle.banded_solve(lu_ref, np.ascontiguousarray(Y[:, cols]))
        if Cm.halo == 0:
            with pytest.raises(bk.UnsupportedError):
                Cm.ldiv_(torch.from_numpy(Y).cuda(), algo=bk.SOLVE_WINDOWED)
            continue
        Yd = torch.from_numpy(Y).cuda()
        Cm.ldiv_(Yd, algo=bk.SOLVE_WINDOWED)
        got = Yd[:, torch.from_numpy(cols).cuda()].cpu().numpy()
        assert relerr(got, Yref) < TOL32, (n, nrhs)
        # per column too
        assert np.max(np.max(np.abs(got - Yref), axis=0) / np.max(np.abs(Yref), axis=0)) < TOL32
        Ya = torch.from_numpy(Y).cuda()
        Cm.ldiv_(Ya)                                                  # AUTO == single pass here
        assert torch.equal(Ya, Yd)
        Yt = torch.from_numpy(Y).cuda()
        Cm.ldiv_(Yt, algo=bk.SOLVE_TWOPASS)
        assert np.array_equal(Yt[:, torch.from_numpy(cols).cuda()].cpu().numpy(), Yref)
        # the Float64 solve of the same (Float32) matrix and data
        C64 = bk.CollocationMatrix(ref.astype(np.float64)).lu_()
        Z = torch.from_numpy(Y[:, :64].astype(np.float64)).cuda()
        C64.ldiv_(Z)
        assert relerr(Yd[:, :64].cpu().numpy(), Z.cpu().numpy()) < 20 * TOL32
    # column counts that are not a multiple of 4 (rows not 16-byte multiples apart): WINDOWED refuses, AUTO falls
    # back to BANSLV order
    Yo = torch.from_numpy(Y[:, :77].copy()).cuda()
    with pytest.raises(bk.UnsupportedError):
        Cm.ldiv_(Yo, algo=bk.SOLVE_WINDOWED)
    Cm.ldiv_(Yo)
    assert np.array_equal(Yo.cpu().numpy(), oracle.banded_solve(lu_ref, np.ascontiguousarray(Y[:, :77])))
