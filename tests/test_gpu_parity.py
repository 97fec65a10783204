"""GPU parity tests: every CUDA path, called through the C ABI (via the ctypes host mirror),
against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): knot-interval indices bit-exact; Float64 within 1e-13
relative; Float32 within 1e-5 relative.  The bit-faithful paths (de Boor, evaluate_all,
collocation, LU, two-pass solve) are additionally required to be BIT-EXACT here.
Relative error = max|gpu - ref| / max|ref| over the output array.
"""
import numpy as np
import pytest

from conftest import dec_range

pytestmark = pytest.mark.gpu

TOL64 = 1e-13
TOL32 = 1e-5


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def _points(synth, seed, n, a, b, extra=()):
    x = a + (b - a) * synth.u_numpy(seed, 0, n)
    return np.concatenate([x, np.asarray(extra, dtype=np.float64)])


@pytest.fixture(scope="module")
def synth():
    from bsplinekit_b200 import synth as s
    return s


# ------------------------------------------------------------------------------------------------
# K1: knot search
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [2, 4, 5, 8])
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_find_knot_interval_plain(bk, oracle, synth, k, dt):
    npdt = np.float64 if dt == "f64" else np.float32
    br = synth.breakpoints(3, 200).astype(npdt)
    B = bk.BSplineBasis(k, br)
    t = B.knots()
    x = _points(synth, 11, 200000, -0.1, 1.1, extra=list(t) + [t[0] - 1, t[-1] + 1, np.nan]).astype(npdt)
    i_ref, z_ref = oracle.find_knot_interval(0, k, t, x, kt=dt)
    i, z = bk.find_knot_interval(B, x)
    assert np.array_equal(i, i_ref)
    assert np.array_equal(z, z_ref)
    # reference's own assertions (test/bsplines.jl:18-30)
    a, b = t[0], t[-1]
    N = len(B)
    ii, zz = bk.find_knot_interval(B, np.array([a - 1, a, b, b + 1], dtype=npdt))
    assert list(zip(ii.tolist(), zz.tolist())) == [(1, -1), (k, 0), (N, 0), (N, 1)]


@pytest.mark.parametrize("k", [3, 4, 6])
def test_find_knot_interval_periodic(bk, oracle, synth, k):
    br = 2.0 * synth.breakpoints(8, 101) - 1.0
    B = bk.PeriodicBSplineBasis(k, br)
    x = _points(synth, 12, 100000, -5.0, 7.0, extra=list(br) + [br[0] - 2.0, br[-1] + 2.0])
    i_ref, z_ref = oracle.find_knot_interval(1, k, br, x)
    i, z = bk.find_knot_interval(B, x)
    assert np.array_equal(i, i_ref)
    assert np.all(z == 0)


@pytest.mark.parametrize("dt", ["f64", "f32"])
@pytest.mark.parametrize("xoff,ioff", [(0, 0), (1, 0), (1, 1), (0, 1), (3, 1), (2, 2)])
def test_find_knot_interval_any_alignment(bk, oracle, synth, dt, xoff, ioff):
    """The vector body of K1 starts at the first 16-byte boundary of x; the C ABI only promises natural alignment
    of x / idx / zone: every combination of offsets (and lengths shorter than one vector) gives the same integers."""
    import ctypes
    import torch
    from bsplinekit_b200 import _lib as L
    npdt = np.float64 if dt == "f64" else np.float32
    tdt = torch.float64 if dt == "f64" else torch.float32
    br = synth.breakpoints(3, 300).astype(npdt)
    B = bk.BSplineBasis(4, br)
    t = B.knots()
    for n in (0, 1, 2, 3, 5, 1000, 70001):
        x = _points(synth, 31 + n, n + 8, -0.1, 1.1)[:n].astype(npdt) if n else np.zeros(0, npdt)
        xd = torch.zeros(n + 8, dtype=tdt, device="cuda")
        xd[xoff:xoff + n] = torch.from_numpy(x).cuda()
        idx = torch.full((n + 8,), -7, dtype=torch.int64, device="cuda")
        zone = torch.full((n + 8,), 5, dtype=torch.int8, device="cuda")
        es = xd.element_size()
        L.check(L.lib.bsk_find_knot_interval(B._handle, ctypes.c_void_p(xd.data_ptr() + es * xoff), n,
                                             ctypes.c_void_p(idx.data_ptr() + 8 * ioff),
                                             ctypes.c_void_p(zone.data_ptr() + ioff), None))
        i_ref, z_ref = oracle.find_knot_interval(0, 4, t, x, kt=dt) if n else (np.zeros(0, np.int64), np.zeros(0, np.int8))
        ih, zh = idx.cpu().numpy(), zone.cpu().numpy()
        assert np.array_equal(ih[ioff:ioff + n], i_ref) and np.array_equal(zh[ioff:ioff + n], z_ref)
        # nothing written outside [ioff, ioff + n)
        assert np.all(ih[:ioff] == -7) and np.all(ih[ioff + n:] == -7)
        assert np.all(zh[:ioff] == 5) and np.all(zh[ioff + n:] == 5)


def test_find_knot_interval_large_no_smem(bk, oracle, synth):
    """10^5 knots: the knot vector does not fit in shared memory (global-memory path)."""
    br = synth.breakpoints(10, 99996)
    B = bk.BSplineBasis(8, br)
    x = _points(synth, 13, 300000, 0.0, 1.0, extra=list(br[::97]))
    i_ref, z_ref = oracle.find_knot_interval(0, 8, B.knots(), x)
    i, z = bk.find_knot_interval(B, x)
    assert np.array_equal(i, i_ref) and np.array_equal(z, z_ref)


def test_find_knot_interval_on_a_uniform_range(bk, oracle):
    """test/knots.jl:4-14 restated: on the augmented knots of the range -1:0.1:1 (k = 4) the interval search
    agrees with searchsortedlast on the collected knot vector for x in -1.05:0.05:1.05 — the reference's O(1)
    shortcut for ranges (src/BSplines/knots.jl:32-41) and our cell LUT must give the same integers."""
    k = 4
    breaks = dec_range(-1, 1, 21)
    B = bk.BSplineBasis(k, breaks)
    ys = B.knots()
    xs = dec_range(-1.05, 1.05, 43)
    idx, zone = bk.find_knot_interval(B, xs)
    ssl = np.searchsorted(ys, xs, side="right")                 # Base.searchsortedlast
    inside = (xs >= ys[0]) & (xs < ys[-1])
    assert np.array_equal(idx[inside], ssl[inside])
    assert np.all(ssl[xs < ys[0]] == 0) and np.all(idx[xs < ys[0]] == 1) and np.all(zone[xs < ys[0]] == -1)
    assert np.all(ssl[xs >= ys[-1]] == len(ys)) and np.all(idx[xs >= ys[-1]] == len(B))
    i_ref, z_ref = oracle.find_knot_interval(0, k, ys, xs)
    assert np.array_equal(idx, i_ref) and np.array_equal(zone, z_ref)
    for T in (np.float32,):                                     # same integers on a Float32 range
        Bf = bk.BSplineBasis(k, breaks.astype(T))
        i32, z32 = bk.find_knot_interval(Bf, xs.astype(T))
        r32, rz32 = oracle.find_knot_interval(0, k, Bf.knots(), xs.astype(T), kt="f32")
        assert np.array_equal(i32, r32) and np.array_equal(z32, rz32)


@pytest.mark.parametrize("k", [3, 4])
def test_periodic_far_from_the_main_period(bk, oracle, synth, k):
    """Points ~10^3 periods away: the reference wraps x by REPEATED +-L (src/BSplines/periodic.jl:85-100) and
    shifts the knots the same way (:102-115); up to BSK_FAR_PERIODS = 4096 periods the kernels take exactly
    those steps, so indices, basis values and spline values are the oracle's bit for bit.  Farther out (where
    the reference needs > 4096 sequential steps per point) the shift is one multiplication: indices still
    agree, values to rounding."""
    import torch
    L = 0.7                                                     # not a power of two: repeated sums round
    br = L * synth.breakpoints(8, 41)
    B = bk.PeriodicBSplineBasis(k, br)
    c = 2 * synth.u_numpy(31, 0, len(B)) - 1
    S = bk.Spline(B, c)
    u = synth.u_numpy(32, 0, 4000)
    x = np.concatenate([1.0e3 * L + L * u, -1.0e3 * L + L * u, 3500 * L + L * u[:500], -3500 * L + L * u[:500]])
    i_ref, _ = oracle.find_knot_interval(1, k, br, x)
    i, z = bk.find_knot_interval(B, x)
    assert np.array_equal(i, i_ref) and np.all(z == 0)
    i2_ref, bs_ref = oracle.evaluate_all(1, k, br, x)
    i2, bs = bk.evaluate_all(B, x)
    assert np.array_equal(i2, i2_ref) and np.array_equal(bs, bs_ref)
    ref = oracle.spline_eval(1, k, br, c, x)
    xd = torch.from_numpy(x).cuda()
    assert np.array_equal(S(xd, algo=bk.EVAL_DEBOOR).cpu().numpy(), ref)
    assert np.array_equal(S(xd).cpu().numpy(), ref)            # table path: outside the main period = de Boor
    # beyond the faithful range: one multiplication instead of > 4096 steps
    xf = np.concatenate([2.0e5 * L + L * u[:300], -3.0e6 * L + L * u[:300]])
    ifar, _ = bk.find_knot_interval(B, xf)
    m = np.floor((xf - br[0]) / L)
    within = np.abs((xf - br[0]) / L - np.round((xf - br[0]) / L)) > 1e-6      # away from period boundaries
    i_main, _ = oracle.find_knot_interval(1, k, br, xf - m * L)
    assert np.array_equal(ifar[within], (i_main + m.astype(np.int64) * len(B))[within])
    vf = S(torch.from_numpy(xf).cuda()).cpu().numpy()
    vm = oracle.spline_eval(1, k, br, c, xf - m * L)
    assert np.max(np.abs(vf - vm)[within]) < 1e-6              # ~ |x| eps |S'|: conditioning of the wrap itself
    # absurdly far / infinite: no hang, no fault
    S(torch.tensor([1e300, -1e300, np.inf, -np.inf], dtype=torch.float64).cuda())
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------
# K2: evaluate_all
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [2, 3, 4, 5, 6, 8])
def test_evaluate_all_plain_bitexact(bk, oracle, synth, k):
    br = synth.breakpoints(3, 64)
    B = bk.BSplineBasis(k, br)
    t = B.knots()
    x = _points(synth, 14, 20000, -0.05, 1.05, extra=list(t) + [np.nan])
    for nd in range(0, k + 1):
        i_ref, bs_ref = oracle.evaluate_all(0, k, t, x, nderiv=nd)
        i, bs = bk.evaluate_all(B, x, bk.Derivative(nd))
        assert np.array_equal(i, i_ref), (k, nd)
        assert np.array_equal(bs, bs_ref, equal_nan=True), (k, nd)
    # partition of unity inside the domain (test/bsplines.jl:84-90)
    xin = x[(x >= 0) & (x <= 1)]
    _, bs = bk.evaluate_all(B, xin)
    assert np.max(np.abs(bs.sum(axis=1) - 1)) < 1e-14


@pytest.mark.parametrize("k", [4, 6, 8])
def test_evaluate_all_float32_output(bk, oracle, synth, k):
    """T = Float32 with Float64 knots: knot differences in F64, rounded, triangle in F32."""
    br = synth.breakpoints(3, 64)
    B = bk.BSplineBasis(k, br)
    x = _points(synth, 15, 20000, 0.0, 1.0)
    for nd in (0, 1, 2):
        i_ref, bs_ref = oracle.evaluate_all(0, k, B.knots(), x, nderiv=nd, vt="f32")
        i, bs = bk.evaluate_all(B, x, bk.Derivative(nd), np.float32)
        assert bs.dtype == np.float32
        assert np.array_equal(i, i_ref)
        assert np.array_equal(bs, bs_ref)
        _, bs64 = oracle.evaluate_all(0, k, B.knots(), x, nderiv=nd)
        assert relerr(bs, bs64) < TOL32


@pytest.mark.parametrize("k", [4, 8])
def test_evaluate_all_float32_basis(bk, oracle, synth, k):
    br = synth.breakpoints(3, 64).astype(np.float32)
    B = bk.BSplineBasis(k, br)
    x = _points(synth, 16, 20000, 0.0, 1.0).astype(np.float32)
    i_ref, bs_ref = oracle.evaluate_all(0, k, B.knots(), x, kt="f32", vt="f32")
    i, bs = bk.evaluate_all(B, x)
    assert np.array_equal(i, i_ref) and np.array_equal(bs, bs_ref)


def test_evaluate_all_ileft(bk, oracle, synth):
    br = dec_range(-1, 1, 21)
    B = bk.BSplineBasis(4, br)
    i, bs = B(0.42)
    assert i == 18
    i2, bs2 = B(0.44, ileft=i)
    i_ref, bs_ref = oracle.evaluate_all(0, 4, B.knots(), [0.44], ileft=[18])
    assert i2 == 18 and np.array_equal(np.array(bs2), bs_ref[0])


@pytest.mark.parametrize("k", [3, 4, 6])
def test_evaluate_all_periodic_bitexact(bk, oracle, synth, k):
    br = 2.0 * synth.breakpoints(8, 41) - 1.0
    B = bk.PeriodicBSplineBasis(k, br)
    x = _points(synth, 17, 20000, -5.0, 7.0, extra=list(br))
    for nd in range(0, k):
        i_ref, bs_ref = oracle.evaluate_all(1, k, br, x, nderiv=nd)
        i, bs = bk.evaluate_all(B, x, bk.Derivative(nd))
        assert np.array_equal(i, i_ref)
        assert np.array_equal(bs, bs_ref)


def test_docstring_goldens_on_gpu(bk):
    """Class-A goldens of the reference docstrings, through the CUDA path, bit for bit
    (src/BSplines/BSplines.jl:56-79; src/BSplines/periodic.jl:157-161)."""
    B = bk.BSplineBasis(bk.BSplineOrder(4), dec_range(-1, 1, 21))
    assert B(0.42) == (18, (0.0013333333333333268, 0.28266666666666657, 0.6306666666666667, 0.08533333333333339))
    assert B(0.44, ileft=18) == (18, (0.01066666666666666, 0.4146666666666667, 0.5386666666666665, 0.03599999999999999))
    i, bs = B(0.42, np.float32)
    assert i == 18 and [float(v) for v in bs] == [float(np.float32(v)) for v in
                                                   (0.0013333336, 0.28266668, 0.6306667, 0.085333325)]
    assert B(0.42, bk.Derivative(1)) == (18, (0.19999999999999937, 6.4, -3.3999999999999977, -3.200000000000001))
    P = bk.PeriodicBSplineBasis(bk.BSplineOrder(4), dec_range(-1, 1, 11))
    assert P(-0.42) == (5, (0.12150000000000002, 0.6571666666666667, 0.22116666666666668, 0.00016666666666666563))
    assert P(-0.42 + 2) == (15, (0.12150000000000015, 0.6571666666666667, 0.22116666666666657, 0.00016666666666666674))


def _robin(bk, oracle, k, br):
    B = bk.BSplineBasis(k, br)
    R = bk.RecombinedBSplineBasis(B, bk.Derivative(0) + 3 * bk.Derivative(1))
    rc = oracle.robin_corners(k, B.knots(), [1.0, 3.0])
    return B, R, rc


@pytest.mark.parametrize("k", [4, 5, 8])
def test_evaluate_all_recombined_robin(bk, oracle, synth, k):
    br = synth.breakpoints(10, 50)
    B, R, rc = _robin(bk, oracle, k, br)
    assert np.array_equal(R.ul, rc.ul) and np.array_equal(R.lr, rc.lr)
    x = _points(synth, 18, 20000, 0.0, 1.0, extra=[0.0, 1.0] + list(br[:4]) + list(br[-4:]))
    for nd in (0, 1, 2):
        i_ref, bs_ref = oracle.evaluate_all(0, k, B.knots(), x, nderiv=nd, recomb=rc)
        i, bs = bk.evaluate_all(R, x, bk.Derivative(nd))
        assert np.array_equal(i, i_ref)
        assert np.array_equal(bs, bs_ref)
        i32, bs32 = bk.evaluate_all(R, x, bk.Derivative(nd), np.float32)
        _, bs32_ref = oracle.evaluate_all(0, k, B.knots(), x, nderiv=nd, vt="f32", recomb=rc)
        assert np.array_equal(bs32, bs32_ref)
        assert relerr(bs32, bs_ref) < TOL32


@pytest.mark.parametrize("ops", ["dirichlet", "neumann", "mixed01"])
def test_evaluate_all_recombined_simple(bk, oracle, synth, ops):
    k = 5
    br = synth.breakpoints(10, 30)
    B = bk.BSplineBasis(k, br)
    if ops == "dirichlet":
        R = bk.RecombinedBSplineBasis(B, bk.Derivative(0))
        rc = oracle.Recomb(1, 1, 0, 0, np.zeros((1, 0)), np.zeros((1, 0)))
    elif ops == "neumann":
        R = bk.RecombinedBSplineBasis(B, bk.Derivative(1))
        rc = oracle.Recomb(1, 1, 1, 1, np.ones((2, 1)), np.ones((2, 1)))
    else:
        R = bk.RecombinedBSplineBasis(B, (bk.Derivative(0), bk.Derivative(1)))
        rc = oracle.Recomb(2, 2, 0, 0, np.zeros((2, 0)), np.zeros((2, 0)))
    assert len(R) == len(B) - rc.cl - rc.cr
    x = _points(synth, 19, 5000, 0.0, 1.0, extra=[0.0, 1.0])
    i_ref, bs_ref = oracle.evaluate_all(0, k, B.knots(), x, recomb=rc)
    i, bs = bk.evaluate_all(R, x)
    assert np.array_equal(i, i_ref) and np.array_equal(bs, bs_ref)


# ------------------------------------------------------------------------------------------------
# K3: spline evaluation
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [2, 3, 4, 5, 6, 7, 8])
def test_spline_eval_deboor_bitexact(bk, oracle, synth, k):
    br = synth.breakpoints(3, 100)
    B = bk.BSplineBasis(k, br)
    c = 2 * synth.u_numpy(4, 0, len(B)) - 1
    S = bk.Spline(B, c)
    x = _points(synth, 20, 100001, -0.05, 1.05, extra=list(B.knots()) + [np.nan])
    ref = oracle.spline_eval(0, k, B.knots(), c, x)
    got = S(x, algo=bk.EVAL_DEBOOR)
    assert np.array_equal(got, ref, equal_nan=True)
    for nd in range(1, k):
        kd, td, cd = oracle.spline_derivative(0, k, B.knots(), c, nd)
        dS = bk.Derivative(nd) * S
        assert np.array_equal(dS.coefficients(), cd)
        assert np.array_equal(dS.knots(), td)
        refd = oracle.spline_eval(0, kd, td, cd, x)
        assert np.array_equal(dS(x, algo=bk.EVAL_DEBOOR), refd, equal_nan=True)
    with pytest.raises(bk.ArgumentError):
        bk.Derivative(k) * S


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6, 7, 8])
def test_spline_eval_pp_tolerance(bk, oracle, synth, k):
    """Local-polynomial fast path: value and all derivatives in one pass, within 1e-13."""
    br = synth.breakpoints(3, 300)
    B = bk.BSplineBasis(k, br)
    c = 2 * synth.u_numpy(4, 0, len(B)) - 1
    S = bk.Spline(B, c)
    x = _points(synth, 21, 200003, -0.02, 1.02, extra=list(B.knots()) + [np.nan])
    derivs = list(range(0, min(k, 4)))
    outs = S.evaluate(x, derivs, algo=bk.EVAL_PP)
    for nd, got in zip(derivs, outs):
        kd, td, cd = (k, B.knots(), c) if nd == 0 else oracle.spline_derivative(0, k, B.knots(), c, nd)
        ref = oracle.spline_eval(0, kd, td, cd, x)
        # NaN in, NaN out — except for D^(k-1) S, whose evaluation in the reference never touches x (order-1 de Boor)
        assert np.array_equal(got[-1:], ref[-1:], equal_nan=True) and np.isnan(got[-1]) == (nd < k - 1)
        e = relerr(got[:-1], ref[:-1])
        assert e < TOL64, (k, nd, e)
    # outside the domain -> exact zeros (spline.jl:164-168)
    v0 = outs[0][:-1]
    assert np.all(v0[x[:-1] < 0] == 0) and np.all(v0[x[:-1] > 1] == 0)


def test_spline_eval_device_vectors(bk, oracle, synth):
    """Device-resident x and outputs (the DeviceVector path of the Julia shim)."""
    import torch
    k = 4
    br = synth.breakpoints(3, 1000)
    B = bk.BSplineBasis(k, br)
    c = 2 * synth.u_numpy(4, 0, len(B)) - 1
    S = bk.Spline(B, c)
    n = (1 << 22) + 3
    xd = synth.u_torch(5, 0, n)
    v, dv = S.evaluate(xd, (0, 1))
    assert v.is_cuda and v.shape == xd.shape
    xh = synth.u_numpy(5, 0, n)
    assert np.array_equal(xd.cpu().numpy(), xh)
    ref = oracle.spline_eval(0, k, B.knots(), c, xh)
    kd, td, cd = oracle.spline_derivative(0, k, B.knots(), c, 1)
    refd = oracle.spline_eval(0, kd, td, cd, xh)
    assert relerr(v.cpu().numpy(), ref) < TOL64
    assert relerr(dv.cpu().numpy(), refd) < TOL64
    # unaligned device views: scalar head / tail around the vectorised body
    v2, dv2 = S.evaluate(xd[1:], (0, 1))
    assert relerr(v2.cpu().numpy(), ref[1:]) < TOL64
    assert relerr(dv2.cpu().numpy(), refd[1:]) < TOL64
    o1 = torch.empty(n + 1, dtype=torch.float64, device="cuda")[1:]     # output phase != input phase
    S.evaluate(xd, (0,), out=[o1])
    assert relerr(o1.cpu().numpy(), ref) < TOL64
    # faithful path on the same points is bit-exact
    vb = S.evaluate(xd, (0,), algo=bk.EVAL_DEBOOR)[0]
    assert np.array_equal(vb.cpu().numpy(), ref)


@pytest.mark.parametrize("k", [3, 4, 6])
def test_spline_eval_periodic(bk, oracle, synth, k):
    L = 2.0
    br = L * synth.breakpoints(8, 201) - 1.0
    B = bk.PeriodicBSplineBasis(k, br)
    c = 2 * synth.u_numpy(4, 0, len(B)) - 1
    S = bk.Spline(B, c)
    x = _points(synth, 22, 100000, -1.0 - 2 * L, 1.0 + 2 * L, extra=list(br))
    ref = oracle.spline_eval(1, k, br, c, x)
    assert np.array_equal(S(x, algo=bk.EVAL_DEBOOR), ref)
    derivs = list(range(0, k))[:4]
    outs = S.evaluate(x, derivs)       # AUTO -> pp in the main period, de Boor outside
    for nd, got in zip(derivs, outs):
        if nd == 0:
            r = ref
        else:
            kd, td, cd = oracle.spline_derivative(1, k, br, c, nd)
            dS = bk.Derivative(nd) * S
            assert np.array_equal(dS.coefficients(), cd)
            r = oracle.spline_eval(1, kd, td, cd, x)
            assert np.array_equal(dS(x, algo=bk.EVAL_DEBOOR), r)
        assert relerr(got, r) < TOL64, (k, nd)
    # closed curve: S(a) == S(b) (test/periodic.jl:8-23)
    assert S(np.array([-1.0]))[0] == pytest.approx(S(np.array([1.0]))[0], abs=1e-14)


@pytest.mark.parametrize("k", [4, 8])
def test_spline_eval_float32(bk, oracle, synth, k):
    br = synth.breakpoints(10, 500).astype(np.float32)
    B = bk.BSplineBasis(k, br)
    c = (2 * synth.u_numpy(4, 0, len(B)) - 1).astype(np.float32)
    S = bk.Spline(B, c)
    x = synth.u_numpy(11, 0, 100000).astype(np.float32)
    ref = oracle.spline_eval(0, k, B.knots(), c, x, kt="f32")
    got = S(x, algo=bk.EVAL_DEBOOR)
    assert got.dtype == np.float32 and np.array_equal(got, ref)
    ref64 = oracle.spline_eval(0, k, B.knots().astype(np.float64), c.astype(np.float64), x.astype(np.float64))
    assert relerr(S(x), ref64) < TOL32          # fast path vs a Float64 evaluation
    assert relerr(ref, ref64) < TOL32


def test_spline_eval_recombined(bk, oracle, synth):
    """Recombined spline == parent spline with converted coefficients (Recombinations/splines.jl:32-65)."""
    k = 6
    br = synth.breakpoints(10, 80)
    B, R, rc = _robin(bk, oracle, k, br)
    cr_ = 2 * synth.u_numpy(4, 0, len(R)) - 1
    S = bk.Spline(R, cr_)
    A = R.recombination_matrix()
    cp = S.coefficients()
    assert np.allclose(cp, A @ cr_, rtol=0, atol=1e-15)
    x = _points(synth, 23, 50000, 0.0, 1.0, extra=[0.0, 1.0])
    ref = oracle.spline_eval(0, k, B.knots(), cp, x)
    assert np.array_equal(S(x, algo=bk.EVAL_DEBOOR), ref)
    # the Robin condition u + 3 du/dn = 0 holds at both ends (test/recombination.jl:284-305)
    v, dv = S.evaluate(np.array([0.0, 1.0]), (0, 1), algo=bk.EVAL_DEBOOR)
    assert abs(v[0] - 3 * dv[0]) < 1e-9 * max(1, abs(dv[0]))
    assert abs(v[1] + 3 * dv[1]) < 1e-9 * max(1, abs(dv[1]))
    with pytest.raises(bk.DimensionMismatch):
        bk.Spline(R, np.zeros(len(R) + 1))


# ------------------------------------------------------------------------------------------------
# K4: collocation assembly
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [3, 4, 6, 8])
def test_collocation_matrix_plain(bk, oracle, synth, k):
    br = synth.breakpoints(6, 300)
    B = bk.BSplineBasis(k, br)
    x = bk.collocation_points(B)
    assert np.array_equal(x, oracle.greville(0, k, B.knots()))
    for nd in (0, 1, 2):
        if nd >= k:
            continue
        ref, nout_ref = oracle.collocation_matrix(0, k, B.knots(), x, nderiv=nd)
        Cm, nout = bk.collocation_matrix(B, x, bk.Derivative(nd), return_outside=True)
        assert nout == nout_ref
        assert np.array_equal(Cm.data(), ref)
    # rows sum to one; out-of-domain rows are zero (test/collocation.jl:112-121)
    x2 = x.copy(); x2[0] = -1.0; x2[-1] = 2.0
    Cm = bk.collocation_matrix(B, x2)
    D = Cm.dense()
    assert np.all(D[0] == 0) and np.all(D[-1] == 0)
    assert np.allclose(D[1:-1].sum(axis=1), 1.0, atol=1e-14)


def test_collocation_matrix_robin_c5_shape(bk, oracle, synth):
    """Config C5 (scaled down here; full size in test_full_size): Robin D0 + 3 D1, k = 8, Derivative(2)."""
    k = 8
    br = synth.breakpoints(10, 2000)
    B, R, rc = _robin(bk, oracle, k, br)
    x = bk.collocation_points(R)
    assert np.array_equal(x, oracle.greville(2, k, B.knots(), cl=1, N=len(R)))
    for nd in (0, 2):
        ref, nout_ref = oracle.collocation_matrix(2, k, B.knots(), x, nderiv=nd, recomb=rc)
        Cm, nout = bk.collocation_matrix(R, x, bk.Derivative(nd), return_outside=True)
        assert nout == nout_ref == 0
        assert np.array_equal(Cm.data(), ref)
    with pytest.raises(bk.ArgumentError):
        bk.collocation_matrix(R, x[:-1])


def test_collocation_cyclic(bk, oracle, synth):
    br = synth.breakpoints(8, 1001)
    B = bk.PeriodicBSplineBasis(4, br)
    x = bk.collocation_points(B)
    a, b, c, nout = oracle.collocation_cyclic(4, br, x)
    M = bk.collocation_matrix(B, x)
    assert nout == 0
    assert np.array_equal(M.a, a) and np.array_equal(M.b, b) and np.array_equal(M.c, c)
    assert np.allclose(M.dense().sum(axis=1), 1.0, atol=1e-14)       # test/periodic.jl:166-170


# ------------------------------------------------------------------------------------------------
# K5 / K6: banded LU and multi-RHS solve
# ------------------------------------------------------------------------------------------------
def _interp_system(bk, oracle, synth, n, k, seed=6):
    xdata = synth.breakpoints(seed, n)
    t = bk.make_knots(xdata, k)
    assert np.array_equal(t, oracle.make_knots(xdata, k))
    B = bk.BSplineBasis(k, t, augment=False)
    Cm = bk.collocation_matrix(B, xdata)
    ref, _ = oracle.collocation_matrix(0, k, t, xdata)
    assert np.array_equal(Cm.data(), ref)
    return xdata, t, B, Cm, ref


@pytest.mark.parametrize("k", [2, 3, 4, 6, 8])
def test_banded_lu_bitexact(bk, oracle, synth, k):
    xdata, t, B, Cm, ref = _interp_system(bk, oracle, synth, 257, k)
    lu_ref = ref.copy(order="F")
    assert oracle.banded_lu(lu_ref) == 0
    Cm.lu_()
    assert np.array_equal(Cm.data(), lu_ref)


def test_banded_lu_special_cases(bk):
    """test/collocation.jl:4-53: 1x1, zero pivots, wrong dimensions."""
    C1 = bk.CollocationMatrix(np.array([[2.0]]))
    C1.lu_()
    y = np.array([[3.0]])
    C1.ldiv_(y)
    assert y[0, 0] == 1.5
    with pytest.raises(bk.ZeroPivotException) as ei:
        bk.CollocationMatrix(np.array([[0.0]])).lu_()
    assert ei.value.info == 1
    band = np.zeros((3, 4), order="F")
    band[1, :] = [1.0, 2.0, 0.0, 4.0]            # diagonal with a zero at i = 3, no off-diagonals
    with pytest.raises(bk.ZeroPivotException) as ei:
        bk.CollocationMatrix(band).lu_()
    assert ei.value.info == 3
    band[1, 2] = 3.0
    Cm = bk.CollocationMatrix(band).lu_()
    with pytest.raises(bk.DimensionMismatch):
        Cm.ldiv_(np.zeros((5, 2)))
    y = np.ones((4, 3))
    Cm.ldiv_(y)
    assert np.allclose(y[:, 0], [1.0, 0.5, 1 / 3, 0.25])


@pytest.mark.parametrize("k,n,nrhs", [(4, 300, 37), (6, 513, 1030), (8, 200, 5), (2, 64, 70), (3, 100, 33)])
def test_banded_solve_twopass_bitexact(bk, oracle, synth, k, n, nrhs):
    xdata, t, B, Cm, ref = _interp_system(bk, oracle, synth, n, k)
    lu_ref = ref.copy(order="F")
    oracle.banded_lu(lu_ref)
    Cm.lu_()
    Y = (2 * synth.u_numpy(7, 0, n * nrhs) - 1).reshape(n, nrhs)
    Yref = oracle.banded_solve(lu_ref, Y.copy())
    Yg = Y.copy()
    Cm.ldiv_(Yg, algo=bk.SOLVE_TWOPASS)
    assert np.array_equal(Yg, Yref)
    # residual ||C c - y|| (test/splines.jl:170-187)
    res = oracle.banded_matvec(ref, Yg) - Y
    assert np.max(np.abs(res)) < 1e-12


@pytest.mark.parametrize("k,n,nrhs", [(6, 4096, 2048), (4, 2000, 1500), (8, 3000, 1100), (6, 16384, 2048), (4, 30001, 96)])
def test_banded_solve_windowed_tolerance(bk, oracle, synth, k, n, nrhs):
    """Single-pass windowed solve: coefficients within 1e-13 of BANSLV (in practice ~1e-16)."""
    import torch
    xdata, t, B, Cm, ref = _interp_system(bk, oracle, synth, n, k)
    lu_ref = ref.copy(order="F")
    oracle.banded_lu(lu_ref)
    Cm.lu_()
    assert Cm.halo > 0, "decay bound not met for a well-conditioned interpolation matrix"
    Y = (2 * synth.u_numpy(7, 0, n * nrhs) - 1).reshape(n, nrhs)
    Yref = oracle.banded_solve(lu_ref, Y.copy())
    Yd = torch.from_numpy(Y).cuda()
    Cm.ldiv_(Yd, algo=bk.SOLVE_WINDOWED)
    got = Yd.cpu().numpy()
    e = relerr(got, Yref)
    assert e < TOL64, e
    colmax = np.max(np.abs(Yref), axis=0)
    assert np.max(np.max(np.abs(got - Yref), axis=0) / colmax) < TOL64
    Yd2 = torch.from_numpy(Y).cuda()
    Cm.ldiv_(Yd2)                       # AUTO
    assert relerr(Yd2.cpu().numpy(), Yref) < TOL64


def test_interpolate_readme_case_c1(bk, oracle, synth):
    """Config C1: xdata = (0:10).^2, k = 4 (README.md:41-46, test/interpolation.jl:61-66)."""
    xdata = np.arange(11) ** 2
    ydata = synth.u_numpy(1, 0, 11)
    itp = bk.interpolate(xdata, ydata, bk.BSplineOrder(4))
    S = itp.spline
    got = S(xdata.astype(np.float64), algo=bk.EVAL_DEBOOR)
    assert np.allclose(got, ydata, rtol=1e-12, atol=1e-13)           # S.(xs) ≈ ys
    # against the oracle chain
    xf = xdata.astype(np.float64)
    t = oracle.make_knots(xf, 4)
    band, _ = oracle.collocation_matrix(0, 4, t, xf)
    oracle.banded_lu(band)
    c = oracle.banded_solve(band, ydata.copy().reshape(-1, 1)).ravel()
    assert np.array_equal(S.coefficients(), c)
    xs = 100 * synth.u_numpy(2, 0, 1000000)
    ref = oracle.spline_eval(0, 4, t, c, xs)
    assert np.array_equal(S(xs, algo=bk.EVAL_DEBOOR), ref)
    assert relerr(S(xs), ref) < TOL64


def test_interpolate_docstring_goldens(bk):
    """src/SplineInterpolations/SplineInterpolations.jl:201-257 (class B: cospi inputs)."""
    xs = dec_range(-1, 1, 21)
    ys = np.cos(np.pi * xs)
    S = bk.interpolate(xs, ys, bk.BSplineOrder(4)).spline
    assert S(np.array([-1.0]), algo=bk.EVAL_DEBOOR)[0] == -1.0
    assert S(np.array([-0.99]))[0] == pytest.approx(-0.9996420091470221, abs=2e-15)
    assert S(np.array([0.998]))[0] == pytest.approx(-1.0000122303614758, abs=2e-15)
    d1 = (bk.Derivative(1) * S)(np.array([-1.0]))[0]
    d2 = (bk.Derivative(2) * S)(np.array([-1.0]))[0]
    assert d1 == pytest.approx(-0.01663433622896893, abs=5e-13)
    assert d2 == pytest.approx(10.527273287554928, abs=5e-11)
    xp = dec_range(-1, 0.9, 20)
    Sp = bk.interpolate(xp, np.cos(np.pi * xp), bk.BSplineOrder(4), bk.Periodic(2.0)).spline
    assert Sp(np.array([-0.99]))[0] == pytest.approx(-0.9995032595823043, abs=2e-15)
    assert Sp(np.array([0.998]))[0] == pytest.approx(-0.9999801044078943, abs=2e-15)


def test_batched_interpolation_and_eval(bk, oracle, synth):
    """interpolate!(itp, Y) on a batch + evaluating one data set from device-resident coefficients."""
    import torch
    n, k, M = 512, 6, 96
    xdata = synth.breakpoints(6, n)
    B = bk.BSplineBasis(k, bk.make_knots(xdata, k), augment=False)
    itp = bk.SplineInterpolation(B, xdata)
    Y = (2 * synth.u_numpy(7, 0, n * M) - 1).reshape(n, M)
    Yd = torch.from_numpy(Y).cuda()
    itp.interpolate_(Yd, algo=bk.SOLVE_TWOPASS)
    S = itp.spline_of(5)
    got = S(xdata, algo=bk.EVAL_DEBOOR)
    assert np.allclose(got, Y[:, 5], rtol=0, atol=1e-12)
    with pytest.raises(bk.DimensionMismatch):
        itp.interpolate_(np.zeros(n + 1))


# ------------------------------------------------------------------------------------------------
# K7: cyclic tridiagonal
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,nrhs", [(20, 1), (1000, 65), (10000, 300)])
def test_cyclic_solve(bk, oracle, synth, n, nrhs):
    br = synth.breakpoints(8, n + 1)
    B = bk.PeriodicBSplineBasis(4, br)
    x = bk.collocation_points(B)
    M = bk.collocation_matrix(B, x)
    D = (2 * synth.u_numpy(9, 0, n * nrhs) - 1).reshape(n, nrhs)
    ref = oracle.cyclic_solve(M.a, M.b, M.c, D.copy())
    got = D.copy()
    M.ldiv_(got)
    assert relerr(got, ref) < TOL64
    if n <= 1000:
        assert np.max(np.abs(M.dense() @ got - D)) < 1e-12


# ------------------------------------------------------------------------------------------------
# errors / no silent fallback
# ------------------------------------------------------------------------------------------------
def test_error_mapping(bk):
    with pytest.raises(bk.ArgumentError):
        bk.BSplineBasis(0, np.linspace(0, 1, 5))
    B = bk.BSplineBasis(4, np.linspace(0, 1, 10))
    with pytest.raises(bk.DimensionMismatch):
        bk.Spline(B, np.zeros(3))
    with pytest.raises(bk.ArgumentError):
        bk.collocation_matrix(B, np.linspace(0, 1, 5))


def test_native_library_is_loaded(bk):
    """The parity tests above must have run through the in-tree CUDA library."""
    with open("/proc/self/maps") as f:
        assert "libbsplinekit_b200.so" in f.read()


# ------------------------------------------------------------------------------------------------
# Natural boundary conditions (src/Recombinations/natural.jl)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [2, 4, 6, 8])
def test_natural_recombination(bk, synth, k):
    """test/natural.jl:10-54: derivatives 2..k/2 of a natural spline vanish at the boundaries."""
    ts = np.linspace(-1.0, 1.0, 21)
    B = bk.BSplineBasis(k, ts)
    R = bk.RecombinedBSplineBasis(B, bk.Natural())
    assert len(R) == len(B) - 2 * (k // 2 - 1)
    if k == 2:
        assert np.array_equal(R.recombination_matrix(), np.eye(len(B)))
    S = bk.Spline(R, 1 + synth.u_numpy(42, 0, len(R)))
    for x in R.boundaries():
        xs = np.array([x])
        Sx = S(xs, algo=bk.EVAL_DEBOOR)[0]
        assert Sx > 1
        eps_ = 2e-8 * Sx
        assert abs((bk.Derivative(1) * S)(xs, algo=bk.EVAL_DEBOOR)[0]) > eps_
        for n in range(2, k // 2 + 1):
            assert abs((bk.Derivative(n) * S)(xs, algo=bk.EVAL_DEBOOR)[0]) < eps_
    with pytest.raises(bk.ArgumentError):
        bk.RecombinedBSplineBasis(bk.BSplineBasis(5, ts), bk.Natural())


def test_natural_interpolation_docstring(bk):
    """src/SplineInterpolations/SplineInterpolations.jl:218-257 (class B goldens)."""
    xs = dec_range(-1, 1, 21)
    ys = np.cos(np.pi * xs)
    itp = bk.interpolate(xs, ys, bk.BSplineOrder(4), bk.Natural())
    S = itp.spline
    assert S(np.array([-1.0]), algo=bk.EVAL_DEBOOR)[0] == pytest.approx(-1.0, abs=1e-15)
    assert (bk.Derivative(1) * S)(np.array([-1.0]))[0] == pytest.approx(0.2872618670889516, abs=1e-12)
    assert abs((bk.Derivative(2) * S)(np.array([-1.0]))[0]) < 1e-11
    assert S(np.array([-0.99]))[0] == pytest.approx(-0.9971071640321145, abs=1e-14)
    assert S(np.array([0.998]))[0] == pytest.approx(-0.9994253145274461, abs=1e-14)
    assert np.allclose(S(xs), ys, atol=1e-13)                       # S.(xs) ≈ ys


def test_periodic_linear_interpolation_is_identity(bk, oracle, synth):
    """k = 2 periodic: IdentityMatrix collocation (src/Collocation/Collocation.jl:147-157)."""
    xp = 2.0 * synth.breakpoints(8, 33)[:-1] - 1.0
    yp = synth.u_numpy(3, 0, xp.size)
    itp = bk.interpolate(xp, yp, bk.BSplineOrder(2), bk.Periodic(2.0))
    assert np.array_equal(itp.spline.coefficients(), yp)
    assert np.allclose(itp.spline(xp, algo=bk.EVAL_DEBOOR), yp, atol=1e-15)
    tp = oracle.make_knots_periodic(xp, 2, 2.0)
    xs = -3.0 + 6.0 * synth.u_numpy(4, 0, 1000)
    assert np.array_equal(itp.spline(xs, algo=bk.EVAL_DEBOOR), oracle.spline_eval(1, 2, tp, yp, xs))


# ------------------------------------------------------------------------------------------------
# edge cases: empty / ragged inputs, odd sizes, non-decaying factors
# ------------------------------------------------------------------------------------------------
def test_empty_and_tiny_inputs(bk, oracle, synth):
    import torch
    B = bk.BSplineBasis(4, synth.breakpoints(3, 50))
    c = 2 * synth.u_numpy(4, 0, len(B)) - 1
    S = bk.Spline(B, c)
    empty = torch.empty(0, dtype=torch.float64, device="cuda")
    assert S.evaluate(empty, (0, 1))[0].numel() == 0
    i, z = bk.find_knot_interval(B, empty)
    assert i.numel() == 0 and z.numel() == 0
    i, bs = bk.evaluate_all(B, empty)
    assert i.numel() == 0 and tuple(bs.shape) == (0, 4)
    assert S(np.empty(0)).size == 0
    for n in (1, 2, 3, 5, 7, 33):
        x = synth.u_numpy(5, 0, n)
        ref = oracle.spline_eval(0, 4, B.knots(), c, x)
        assert relerr(S(x), ref) < TOL64
        assert np.array_equal(S(x, algo=bk.EVAL_DEBOOR), ref)
        xd = torch.from_numpy(x).cuda()
        assert relerr(S.evaluate(xd, (0,))[0].cpu().numpy(), ref) < TOL64


@pytest.mark.parametrize("k,n,nrhs", [(6, 4099, 2081), (4, 1001, 2050), (6, 777, 2048)])
def test_banded_solve_ragged_sizes(bk, oracle, synth, k, n, nrhs):
    """n not a multiple of the 8-row chunk, RHS count not a multiple of 32 (tail columns take the
    BANSLV-order kernels), unaligned sub-blocks."""
    import torch
    xdata, t, B, Cm, ref = _interp_system(bk, oracle, synth, n, k)
    lu_ref = ref.copy(order="F")
    oracle.banded_lu(lu_ref)
    Cm.lu_()
    Y = (2 * synth.u_numpy(7, 0, n * nrhs) - 1).reshape(n, nrhs)
    Yref = oracle.banded_solve(lu_ref, Y.copy())
    Yd = torch.from_numpy(Y).cuda()
    Cm.ldiv_(Yd)                                            # AUTO
    assert relerr(Yd.cpu().numpy(), Yref) < TOL64
    if nrhs % 2 == 0:
        Yd = torch.from_numpy(Y).cuda()
        Cm.ldiv_(Yd, algo=bk.SOLVE_WINDOWED)
        assert relerr(Yd.cpu().numpy(), Yref) < TOL64
    Yh = Y.copy()
    Cm.ldiv_(Yh)                                            # host path (chunked, pipelined)
    assert relerr(Yh, Yref) < TOL64


def test_banded_solve_without_decay_falls_back(bk, oracle, synth):
    """A factorisation whose backward recurrence does not decay: the windowed solve must refuse,
    AUTO must take the exact BANSLV-order kernels."""
    import torch
    n, nrhs = 600, 2048
    band = np.zeros((3, n), order="F")
    band[1, :] = 1.0                 # diagonal
    band[0, 1:] = 2.0                # super-diagonal: |x_r| grows like 2^r going up
    Cm = bk.CollocationMatrix(band).lu_()
    assert Cm.halo == 0
    lu = band.copy(order="F")
    assert oracle.banded_lu(lu) == 0
    Y = np.zeros((n, nrhs))
    Y[-40:] = (2 * synth.u_numpy(7, 0, 40 * nrhs) - 1).reshape(40, nrhs)
    ref = oracle.banded_solve(lu, Y.copy())
    Yd = torch.from_numpy(Y).cuda()
    Cm.ldiv_(Yd)
    assert np.array_equal(Yd.cpu().numpy()[-60:], ref[-60:])
    with pytest.raises(bk.UnsupportedError):
        Cm.ldiv_(torch.from_numpy(Y).cuda(), algo=bk.SOLVE_WINDOWED)


def test_concurrent_streams(bk, oracle, synth):
    """Kernels are enqueued on the caller's stream: two torch streams, same handles."""
    import torch
    B = bk.BSplineBasis(4, synth.breakpoints(3, 1000))
    c = 2 * synth.u_numpy(4, 0, len(B)) - 1
    S = bk.Spline(B, c)
    xs = [synth.u_torch(5, i << 22, 1 << 22) for i in range(2)]
    S.evaluate(xs[0], (0,))           # build the tables once
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(2)]
    outs = []
    for st, x in zip(streams, xs):
        with torch.cuda.stream(st):
            outs.append(S.evaluate(x, (0, 1)))
    torch.cuda.synchronize()
    for x, (v, dv) in zip(xs, outs):
        ref = oracle.spline_eval(0, 4, B.knots(), c, x[:4096].cpu().numpy())
        assert relerr(v[:4096].cpu().numpy(), ref) < TOL64


# ------------------------------------------------------------------------------------------------
# vector-valued splines: M splines sharing one basis
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,k,M", [("plain", 4, 7), ("plain", 6, 300), ("periodic", 4, 2), ("periodic", 3, 65)])
def test_spline_eval_multi(bk, oracle, synth, kind, k, M):
    import torch
    if kind == "plain":
        B = bk.BSplineBasis(k, synth.breakpoints(3, 120))
        okind, t = 0, B.knots()
        x = _points(synth, 30, 5000, -0.05, 1.05, extra=list(t))
    else:
        br = 2.0 * synth.breakpoints(8, 121) - 1.0
        B = bk.PeriodicBSplineBasis(k, br)
        okind, t = 1, br
        x = _points(synth, 30, 5000, -4.0, 5.0, extra=list(br))
    C = (2 * synth.u_numpy(31, 0, len(B) * M) - 1).reshape(len(B), M)
    SB = bk.SplineBatch(B, C)
    got = SB(x)
    assert got.shape == (x.size, M)
    for m in sorted(set([0, M // 2, M - 1])):
        ref = oracle.spline_eval(okind, k, t, C[:, m].copy(), x)
        assert relerr(got[:, m], ref) < TOL64
    xs = np.sort(x)                                        # sorted points: cached-coefficient path
    got_s = SB(torch.from_numpy(xs).cuda()).cpu().numpy()
    assert relerr(got_s[:, M - 1], oracle.spline_eval(okind, k, t, C[:, M - 1].copy(), xs)) < TOL64
    with pytest.raises(bk.DimensionMismatch):
        bk.SplineBatch(B, C[:-1])


def test_batched_interpolation_then_multi_eval(bk, oracle, synth):
    """interpolate! on a batch, then evaluate all data sets at once from the device coefficients:
    closed parametric curve S(a) == S(b) and I.(xs) ≈ ys  (test/periodic.jl:8-23, 176-180)."""
    import torch
    n, M = 400, 128
    xp = 2.0 * synth.breakpoints(8, n + 1)[:-1] - 1.0
    B = bk.PeriodicBSplineBasis(4, bk.make_knots(xp, 4, bk.Periodic(2.0)))
    itp = bk.SplineInterpolation(B, xp)
    Y = (2 * synth.u_numpy(7, 0, n * M) - 1).reshape(n, M)
    itp.interpolate_(torch.from_numpy(Y).cuda())
    vals = itp.splines(xp)
    assert np.max(np.abs(vals - Y)) < 1e-12
    ends = itp.splines(np.array([-1.0, 1.0]))
    assert np.max(np.abs(ends[0] - ends[1])) < 1e-13


@pytest.mark.parametrize("k,n,nrhs", [(4, 300, 37), (6, 4096, 1024), (8, 500, 3), (2, 64, 5)])
def test_banded_matvec(bk, oracle, synth, k, n, nrhs):
    """mul!(Y, C, X) on batches (src/Collocation/matrix.jl:91-94) and C * (F \\ y) ≈ y (test/splines.jl:170-187)."""
    xdata, t, B, Cm, ref = _interp_system(bk, oracle, synth, n, k)
    X = (2 * synth.u_numpy(17, 0, n * nrhs) - 1).reshape(n, nrhs)
    Y = Cm.mul(X)
    Yref = oracle.banded_matvec(ref, X)
    assert relerr(Y, Yref) < TOL64
    Cm2 = bk.CollocationMatrix(ref.copy(order="F")).lu_()
    Z = Y.copy()
    Cm2.ldiv_(Z, algo=bk.SOLVE_TWOPASS)
    assert np.max(np.abs(Z - X)) < 1e-11
    with pytest.raises(bk.ArgumentError):
        Cm2.mul(X)


@pytest.mark.parametrize("k,n,nrhs,nd", [(3, 1000, 4096, 1), (4, 4096, 2048, 2), (5, 777, 4096, 2),
                                         (6, 4096, 4096, 2), (6, 520, 2048, 0), (8, 600, 2048, 2),
                                         (6, 300, 37, 2), (4, 64, 2048, 1)])
def test_banded_matvec_solve_fused(bk, oracle, synth, k, n, nrhs, nd):
    """du = F \\ (L u): `mul!` then `ldiv!` of a collocation time step (examples/heat.jl:314-328) in one
    pass.  Checked against the oracle's mat-vec followed by its BANSLV (1e-13), on shapes that take the
    fused kernel (with and without row segments) and shapes that fall back (k = 8, 37 columns)."""
    import torch
    xdata, t, B, Cm, ref = _interp_system(bk, oracle, synth, n, k)
    Lm = bk.collocation_matrix(B, xdata, bk.Derivative(nd))
    Lref, _ = oracle.collocation_matrix(0, k, t, xdata, nd)
    assert np.array_equal(Lm.data(), Lref)
    lu_ref = ref.copy(order="F")
    assert oracle.banded_lu(lu_ref) == 0
    Cm.lu_()
    U = (2 * synth.u_numpy(21, 0, n * nrhs) - 1).reshape(n, nrhs)
    want = oracle.banded_matvec(Lref, U)
    oracle.banded_solve(lu_ref, want)
    Ud = torch.from_numpy(U).cuda()
    out = Cm.ldiv_mul(Lm, Ud)
    torch.cuda.synchronize()
    assert torch.equal(Ud.cpu(), torch.from_numpy(U))           # u is only read
    assert relerr(out.cpu().numpy(), want) < TOL64
    if nd == 0:                                                  # F \ (C u) == u
        assert np.max(np.abs(out.cpu().numpy() - U)) < 1e-11
    with pytest.raises(bk.ArgumentError):
        Cm.ldiv_mul(Lm, Ud, out=Ud)
    with pytest.raises(bk.ArgumentError):
        Lm.ldiv_mul(Cm, Ud)                                      # roles swapped: Lm is not factorised


# ---- approximation and smoothing (callers of the interpolation path) ------------

def _goldens():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "docstring_goldens.json")) as f:
        return json.load(f)


def test_approximate_goldens(bk):
    """approximate(f, B, [method]) against the reference's docstring values
    (SplineApproximations.jl:78-119, README.md:50-63)."""
    g = _goldens()["class_C"]["approximate_k3"]
    B = bk.BSplineBasis(g["k"], dec_range(g["breaks"]["a"], g["breaks"]["b"], g["breaks"]["n"]))
    A = bk.approximate(np.sin, B)
    assert np.allclose(A.method.xs, g["interp_points"], atol=1e-15)
    assert np.allclose(A.coefficients(), g["sin_coefs_6digits"], rtol=6e-6, atol=1e-15)
    assert A(0.3) == pytest.approx(g["sin_S(0.3)"], abs=2e-15)
    bk.approximate_(np.exp, A)                                      # approximate!(exp, S_interp)
    assert np.allclose(A.coefficients(), g["exp_coefs_6digits"], rtol=6e-6)
    assert A(0.3) == pytest.approx(g["exp_S(0.3)"], abs=4e-15)
    assert A(0.34) == pytest.approx(g["exp_S_interp(0.34)"], abs=4e-15)
    V = bk.approximate(np.exp, B, bk.VariationDiminishing())
    assert np.allclose(V.coefficients(), g["exp_vd_coefs_6digits"], rtol=6e-6)
    assert V(0.34) == pytest.approx(g["exp_S_vd(0.34)"], abs=4e-15)
    r = _goldens()["class_C"]["readme_approximate_k6"]
    B6 = bk.BSplineBasis(bk.BSplineOrder(6), np.log2(np.arange(1, 17)))
    fa = bk.approximate(lambda x: np.exp(-x) * np.sin(x), B6)
    assert fa(2.3) == pytest.approx(r["fapprox(2.3)"], abs=1e-14)


def test_fit_smoothing_golden_and_oracle(bk, oracle, synth):
    """fit(BSplineOrder(4), xs, ys, λ; weights): docstring coefficients (smoothing.jl:30-42), the oracle's
    dense Cholesky restatement, weights, λ = 0 == natural interpolation, argument errors."""
    import torch
    g = _goldens()["class_C"]["smoothing_fit"]
    xs = dec_range(0.0, 1.0, 101) ** 2
    ys = np.cos(2 * np.pi * xs) + 0.1 * np.sin(200 * np.pi * xs)
    S = bk.fit(bk.BSplineOrder(4), xs, ys, g["lambda"])
    # coefficients() returns the parent-basis coefficients; compare through values and the oracle
    cs_ref, t, rc = oracle.fit_smoothing(xs, ys, g["lambda"])
    assert np.allclose(cs_ref[:10], g["coefs_head_6digits"], rtol=6e-6)
    xq = np.linspace(0, 1, 777)
    idx, bs = oracle.evaluate_all(2, 4, t, xq, recomb=rc)
    ref_vals = np.array([sum(bs[p, d] * cs_ref[idx[p] - d - 1] for d in range(4) if 0 <= idx[p] - d - 1 < cs_ref.size)
                         for p in range(xq.size)])
    # cond(M) = 9e8 on this grid (steps 1e-4 .. 2e-2): LU and Cholesky of the same matrix already differ
    # by 2e-9 in numpy; the tolerance is 100 eps cond(M)
    assert relerr(S(xq), ref_vals) < 2e-8
    # batch of data sets on the device + weights: one factorisation, fused 2 A' W y + solve
    n, M = 300, 2048
    xb = synth.breakpoints(31, n)
    w = 0.5 + synth.u_numpy(32, 0, n)
    Y = (2 * synth.u_numpy(33, 0, n * M) - 1).reshape(n, M)
    sm = bk.fit(bk.BSplineOrder(4), xb, None, 1e-6, weights=w)
    SB = sm.fit_(torch.from_numpy(Y).cuda())
    for m in (0, 1000, M - 1):
        cm, tb, rcb = oracle.fit_smoothing(xb, Y[:, m], 1e-6, weights=w)
        assert relerr(SB.coefs[:, m].cpu().numpy(), cm) < 1e-8      # ill-conditioned normal equations, see above
    S0 = bk.fit(bk.BSplineOrder(4), xs, ys, 0.0)                    # λ = 0: passes through the data
    assert np.max(np.abs(S0(xs) - ys)) < 1e-9
    # vector-valued evaluation in the recombined (natural) basis == column-by-column evaluation
    vals = SB(xq)
    Sm = bk.Spline(sm.basis, SB.coefs[:, 5].cpu().numpy())
    assert relerr(vals[:, 5], Sm(xq)) < TOL64
    d2 = Sm.evaluate(np.array([0.0, 1.0]), (2,))[0]                 # natural: S''(a) = S''(b) = 0
    assert np.max(np.abs(d2)) < 1e-7 * np.max(np.abs(SB.coefs[:, 5].cpu().numpy())) * n * n
    with pytest.raises(bk.DomainError):
        bk.fit(bk.BSplineOrder(4), xs, ys, -1.0)
    with pytest.raises(bk.DimensionMismatch):
        bk.fit(bk.BSplineOrder(4), xs, ys[:-1], 0.1)
    with pytest.raises(bk.DimensionMismatch):
        bk.fit(bk.BSplineOrder(4), xs, ys, 0.1, weights=np.ones(3))
    with pytest.raises(bk.ArgumentError):
        bk.fit(bk.BSplineOrder(6), xs, ys, 0.1)


@pytest.mark.parametrize("k,N", [(3, 64), (5, 300), (6, 1000), (7, 50), (8, 4096)])
def test_periodic_interpolation_general_order(bk, oracle, synth, k, N):
    """interpolate(x, y, BSplineOrder(k), Periodic(L)) for k not in {2, 4}: the reference factorises a
    SparseMatrixCSC with UMFPACK (Collocation.jl:141-142); here banded LU + Woodbury corners.  Checked
    against a dense pivoted solve of the oracle's collocation matrix and I.(xs) ≈ ys (test/periodic.jl:159-180)."""
    import torch
    L = 2.0
    x = -1.0 + L * synth.breakpoints(41, N + 1)[:-1]
    data = bk.make_knots(x, k, bk.Periodic(L))
    assert np.array_equal(data, oracle.make_knots_periodic(x, k, L))
    B = bk.PeriodicBSplineBasis(k, data)
    assert np.array_equal(bk.collocation_points(B), oracle.collocation_points_periodic(k, data))
    Cm = bk.collocation_matrix(B, x)
    A = oracle.periodic_collocation_dense(k, data, x)
    assert np.array_equal(Cm.dense(), A)                             # K2 rows are bit-exact
    D1 = bk.collocation_matrix(B, x, bk.Derivative(1))
    assert np.array_equal(D1.dense(), oracle.periodic_collocation_dense(k, data, x, 1))
    M = 2048 if N <= 1000 else 256
    Y = (2 * synth.u_numpy(43, 0, N * M) - 1).reshape(N, M)
    itp = bk.SplineInterpolation(B, x)
    itp.interpolate_(torch.from_numpy(Y).cuda())
    cs = itp.coefs.cpu().numpy()
    ref = np.linalg.solve(A, Y[:, :64])
    assert relerr(cs[:, :64], ref) < 1e-12
    vals = itp.splines(x)                                            # all data sets at the data points
    assert np.max(np.abs(np.asarray(vals) - Y)) < 1e-11
    S = bk.interpolate(x, Y[:, 3].copy(), bk.BSplineOrder(k), bk.Periodic(L))      # single data set
    assert np.max(np.abs(S(x) - Y[:, 3])) < 1e-11
    xs = np.linspace(-1, 1, 57)
    assert np.max(np.abs(S(xs) - S(xs + L))) < 1e-11
    assert relerr(S(xs), oracle.spline_eval(1, k, data, ref[:, 3].copy(), xs)) < 1e-11


@pytest.mark.parametrize("bw", [1, 2, 3, 4, 5, 6])
def test_windowed_solve_halo_is_sufficient(bk, oracle, bw):
    """The single-pass solve restarts its backward (and, with row segments, forward) recurrence from a
    zero state `halo` rows away; the halo validated at factor time must make that exact to rounding
    for ANY matrix it accepts: short matrices (n < halo), slowly decaying factors (weak diagonal
    dominance), few columns (row segments).  Reference: ldiv! (Collocation/matrix.jl:218-271)."""
    import torch
    rng = np.random.default_rng(100 + bw)
    for n in (4 * bw + 1, 33, 50, 130, 1000, 4096):
        for alpha in (0.3, 0.9, 0.99):
            band = np.zeros((2 * bw + 1, n), order="F")
            band[bw, :] = 1.0
            off = rng.uniform(-1, 1, size=(2 * bw, n)) * alpha / (2 * bw)
            band[:bw, :] = off[:bw]
            band[bw + 1:, :] = off[bw:]
            for j in range(n):                       # entries outside the matrix
                for r in range(2 * bw + 1):
                    i = j + r - bw
                    if i < 0 or i >= n:
                        band[r, j] = 0.0
            lu_ref = band.copy(order="F")
            assert oracle.banded_lu(lu_ref) == 0
            F = bk.CollocationMatrix(band.copy(order="F")).lu_()
            if F.halo == 0:
                continue                             # windowed unavailable: AUTO uses the two-pass kernels
            for M in (2048, 4096 if n <= 1000 else 2048):
                Y = rng.uniform(-1, 1, size=(n, M))
                want = Y[:, :48].copy()
                oracle.banded_solve(lu_ref, want)
                Z = torch.from_numpy(Y).cuda()
                F.ldiv_(Z, algo=bk.SOLVE_WINDOWED)
                got = Z[:, :48].cpu().numpy()
                assert relerr(got, want) < 1e-13, (bw, n, alpha, M, F.halo)
                # all columns agree with the bit-faithful two-pass kernels
                Z2 = torch.from_numpy(Y).cuda()
                F.ldiv_(Z2, algo=bk.SOLVE_TWOPASS)
                assert float((Z - Z2).abs().max() / Z2.abs().max()) < 1e-13, (bw, n, alpha, M, F.halo)


@pytest.mark.parametrize("k", [2, 4, 5, 8])
def test_spline_extrapolation(bk, oracle, synth, k):
    """extrapolate(S, Flat() / Linear() / Smooth()) (src/SplineExtrapolations/SplineExtrapolations.jl:22-71)
    fused into the evaluation kernels; inside the domain nothing changes."""
    import torch
    breaks = synth.breakpoints(51, 40)
    B = bk.BSplineBasis(k, breaks)
    c = 2 * synth.u_numpy(52, 0, len(B)) - 1
    S = bk.Spline(B, c)
    t = B.knots()
    n = 20001
    x = np.concatenate([-0.7 + 2.4 * synth.u_numpy(53, 0, n), [0.0, 1.0, -1e-300, 1 + 1e-15, -3.0, 4.0, np.nan]])
    xd = torch.from_numpy(x).cuda()
    inside = (x >= 0) & (x <= 1)
    plain = S(x)
    for name, meth in (("flat", bk.Flat()), ("linear", bk.Linear()), ("smooth", bk.Smooth())):
        E = bk.extrapolate(S, meth)
        ref = oracle.spline_extrapolate(name, k, t, c, x[:-1])
        for algo in (bk.EVAL_AUTO, bk.EVAL_PP, bk.EVAL_DEBOOR):
            if name == "linear" and algo == bk.EVAL_DEBOOR:
                with pytest.raises(bk.UnsupportedError):
                    E(xd, algo=algo)
                continue
            got = E(xd, algo=algo).cpu().numpy()
            assert np.isnan(got[-1])
            if algo == bk.EVAL_DEBOOR:                          # reference operation order: bit-exact
                assert np.array_equal(got[:-1], ref)
            else:
                assert relerr(got[:-1], ref) < TOL64
                if name in ("flat", "linear"):                  # boundary value / slope by the reference's own
                    out_ = ~inside[:-1]                         # recurrences: bit-exact outside the domain
                    assert np.array_equal(got[:-1][out_], ref[out_])
            assert np.array_equal(got[inside], S(xd, algo=algo).cpu().numpy()[inside])
        assert relerr(E(x[:-1]), ref) < TOL64                   # host arrays
        assert np.array_equal(plain[~inside & ~np.isnan(x)], np.zeros(np.sum(~inside & ~np.isnan(x))))
    # derivative outputs are extrapolated as their own splines
    if k >= 4:
        E = bk.extrapolate(S, bk.Linear())
        v, d1 = E.evaluate(xd, (0, 1))
        k1, t1, c1 = oracle.spline_derivative(0, k, t, c, 1)
        refd1 = oracle.spline_extrapolate("linear", k1, t1, c1, x[:-1])
        assert relerr(d1.cpu().numpy()[:-1], refd1) < TOL64
        assert np.array_equal(d1.cpu().numpy()[:-1][~inside[:-1]], refd1[~inside[:-1]])
        assert relerr(v.cpu().numpy()[:-1], oracle.spline_extrapolate("linear", k, t, c, x[:-1])) < TOL64
    # wrappers: extrapolate(interpolation)
    itp = bk.interpolate(breaks, np.cos(3 * breaks), bk.BSplineOrder(k))
    Ef = bk.extrapolate(itp, bk.Flat())
    assert Ef(-5.0) == itp(0.0) and Ef(7.0) == itp(1.0)
    # periodic: Smooth is the identity, Flat clamps to one period
    Bp = bk.PeriodicBSplineBasis(4, breaks)
    Sp = bk.Spline(Bp, c[:len(Bp)])
    xs = torch.from_numpy(-2 + 5 * synth.u_numpy(54, 0, 999)).cuda()
    assert torch.equal(bk.extrapolate(Sp, bk.Smooth())(xs), Sp(xs))
    assert relerr(bk.extrapolate(Sp, bk.Flat())(xs).cpu().numpy(), Sp(xs.clamp(0.0, 1.0)).cpu().numpy()) < TOL64
    with pytest.raises(bk.UnsupportedError):
        bk.extrapolate(Sp, bk.Linear())(xs)


@pytest.mark.parametrize("k,nbreaks", [(3, 6), (4, 60), (6, 300), (8, 40)])
def test_galerkin_matrix_and_projection(bk, oracle, synth, k, nbreaks):
    """galerkin_matrix / galerkin_projection (src/Galerkin/linear.jl:220-311, projection.jl:44-99): batched
    evaluate_all at all quadrature nodes + one gather thread per entry adding the segments in the reference's
    order — bit-exact against the oracle's sequential loop; plain and recombined bases, derivative pairs."""
    breaks = synth.breakpoints(61, nbreaks) if nbreaks > 6 else dec_range(-1.0, 1.0, 6)
    B = bk.BSplineBasis(k, breaks)
    t = B.knots()
    N = len(B)

    def dense(band):
        w = band.cpu().numpy().T                        # (nb, N) band store
        bw = (w.shape[0] - 1) // 2
        A = np.zeros((N, N))
        for j in range(N):
            for i in range(max(0, j - bw), min(N, j + bw + 1)):
                A[i, j] = w[bw + i - j, j]
        return A

    for d1, d2 in ((0, 0), (1, 1), (0, min(2, k - 1))):
        got = dense(bk.galerkin_matrix(B, (bk.Derivative(d1), bk.Derivative(d2)), as_band=True))
        assert np.array_equal(got, oracle.galerkin_matrix(k, t, d1, d2)), (d1, d2)
    f = lambda x: np.exp(-x) * np.sin(3 * x)
    phi = bk.galerkin_projection(f, B)
    assert np.array_equal(phi, oracle.galerkin_projection(f, k, t))
    if k <= 6:                                            # recombined basis (Dirichlet + Neumann mix)
        R = bk.RecombinedBSplineBasis(B, bk.Derivative(0), bk.Derivative(1))
        rc = oracle.Recomb(R.cl, R.cr, R.nl, R.nr, R.ul, R.lr)
        Rb = bk.galerkin_matrix(R, as_band=True).cpu().numpy().T
        ref = oracle.galerkin_matrix(k, t, 0, 0, recomb=rc)
        bw = k - 1
        for j in range(len(R)):
            for i in range(max(0, j - bw), min(len(R), j + bw + 1)):
                assert Rb[bw + i - j, j] == ref[i, j]
        assert np.array_equal(bk.galerkin_projection(f, R), oracle.galerkin_projection(f, k, t, recomb=rc))


def test_galerkin_reference_assertions(bk, oracle):
    """test/galerkin.jl restated: analytical mass matrix entries for k = 2 on Chebyshev points (:11-27), a pair of
    bases with different parents is refused (:29-32), galerkin_matrix(f, B) with f = x^2 and the quadrature keyword
    (:37-69), projection of the constant spline == column sums of the matrix (:71-86), sums of the recombined
    matrices (:131-146) — and the weighted matrix bit for bit against the oracle's sequential loop."""
    x = np.array([-np.cos(np.pi * n / 42) for n in range(43)])
    B = bk.BSplineBasis(2, x.copy())
    t = B.knots()

    def dense(M, N):
        w = M.cpu().numpy().T
        bw = (w.shape[0] - 1) // 2
        A = np.zeros((N, N))
        for j in range(N):
            for i in range(max(0, j - bw), min(N, j + bw + 1)):
                A[i, j] = w[bw + i - j, j]
        return A

    G = dense(bk.galerkin_matrix(B, as_band=True), len(B))
    i = 8                                                   # 1-based in the reference
    assert G[i - 1, i - 1] == pytest.approx((t[i + 1] - t[i - 1]) / 3, rel=1e-14)
    assert G[i - 2, i - 1] == pytest.approx((t[i] - t[i - 1]) / 6, rel=1e-14)
    with pytest.raises(bk.ArgumentError):
        bk.galerkin_matrix((B, bk.BSplineBasis(3, x.copy())))

    xs = np.array([-np.cos(np.pi * n / 32) for n in range(33)])
    f = lambda v: v * v
    B4 = bk.BSplineBasis(4, xs.copy())
    N4 = len(B4)
    Gb = dense(bk.galerkin_matrix(B4, as_band=True), N4)
    Gf = dense(bk.galerkin_matrix(f, B4, as_band=True), N4)
    assert np.all((Gb == 0) | (Gb > Gf))                      # x in [-1, 1]: the weight only shrinks the integrals
    assert np.array_equal(Gf, oracle.galerkin_matrix(4, B4.knots(), f=f))
    Gi = dense(bk.galerkin_matrix(f, B4, quadrature=bk.gausslegendre(5), as_band=True), N4)
    assert np.max(np.abs(Gf - Gi)) > 1e-8                     # k nodes are not exact for a degree-2k integrand
    assert np.array_equal(Gi, oracle.galerkin_matrix(4, B4.knots(), f=f, nq=5))
    G2 = dense(bk.galerkin_matrix(f, B4, quadrature=bk.gausslegendre(6), as_band=True), N4)
    assert np.max(np.abs(G2 - Gi)) < 1e-16

    for Bp in (B4, bk.RecombinedBSplineBasis(B4, bk.Derivative(1))):
        S1 = bk.Spline(Bp, np.ones(len(Bp)))
        for d in (0, 1):
            M = dense(bk.galerkin_matrix(Bp, (bk.Derivative(0), bk.Derivative(d)), as_band=True), len(Bp))
            phi = bk.galerkin_projection(S1, Bp, bk.Derivative(d))
            assert np.allclose(phi, M.sum(axis=0), rtol=1e-12, atol=1e-14)
    with pytest.raises(bk.DimensionMismatch):
        bk.galerkin_projection(np.sin, B4, out=np.empty(N4 + 3))

    kn = np.array([-np.cos(np.pi * n / 41) for n in range(42)])
    B6 = bk.BSplineBasis(6, kn.copy())
    L = 2.0
    sums = {}
    for name, Bx in (("B", B6), ("R0", bk.RecombinedBSplineBasis(B6, bk.Derivative(0))),
                     ("R1", bk.RecombinedBSplineBasis(B6, bk.Derivative(1))),
                     ("R2", bk.RecombinedBSplineBasis(B6, bk.Derivative(2)))):
        A = dense(bk.galerkin_matrix(Bx, as_band=True), len(Bx))
        assert np.all(np.linalg.eigvalsh(0.5 * (A + A.T)) > 0)         # isposdef
        sums[name] = A.sum()
    assert len(bk.RecombinedBSplineBasis(B6, bk.Derivative(0))) == len(B6) - 2
    assert sums["B"] == pytest.approx(L, rel=1e-13) and sums["R1"] == pytest.approx(L, rel=1e-13)
    assert sums["R0"] < L and abs(sums["R2"] - L) > 1e-6


@pytest.mark.parametrize("k,N", [(2, 12), (3, 9), (4, 40), (5, 33), (6, 200), (8, 60)])
def test_galerkin_periodic(bk, oracle, synth, k, N):
    """Galerkin matrix / projection of a PeriodicBSplineBasis (src/Galerkin/linear.jl:50-54,144,268-311; the
    reference fills a SparseMatrixCSC): cyclic band store, bit-exact against the oracle's sequential loop, the
    reference's own assertions (test/periodic.jl:184-196: sum(G) == period, column sums == (t[j+k] - t[j]) / k)
    and approximate(f, B, MinimiseL2Error()) on a periodic basis (test/periodic.jl:75)."""
    L = 2.0
    data = -1.0 + L * synth.breakpoints(63, N + 1)
    B = bk.PeriodicBSplineBasis(k, data)
    bw = k - 1

    def dense(band):
        w = band.cpu().numpy().T                        # (2 bw + 1, N)
        A = np.zeros((N, N))
        for j in range(N):
            for d in range(-bw, bw + 1):
                A[(j + d) % N, j] = w[bw + d, j]
        return A

    for d1, d2 in ((0, 0), (1, 1), (0, 1)):
        if max(d1, d2) > k - 1:
            continue
        got = dense(bk.galerkin_matrix(B, (bk.Derivative(d1), bk.Derivative(d2)), as_band=True))
        assert np.array_equal(got, oracle.galerkin_matrix_periodic(k, data, d1, d2)), (d1, d2)
    G = dense(bk.galerkin_matrix(B, as_band=True))
    ts = oracle.periodic_knots(data, k)
    assert G.sum() == pytest.approx(L, rel=1e-14)
    assert np.allclose(G.sum(axis=0), (ts[k:k + N] - ts[:N]) / k, rtol=1e-13, atol=0)
    assert np.max(np.abs(G - G.T)) < 1e-15 * np.max(np.abs(G))      # symmetric up to the order of the additions
    f = lambda x: np.cos(np.pi * x) + 0.3 * np.sin(3 * np.pi * x)      # period 2
    assert np.array_equal(bk.galerkin_projection(f, B), oracle.galerkin_projection_periodic(f, k, data))
    if k <= 7 and N >= 4 * bw + 1:
        A = bk.approximate(f, B, bk.MinimiseL2Error())
        cref = np.linalg.solve(oracle.galerkin_matrix_periodic(k, data), oracle.galerkin_projection_periodic(f, k, data))
        assert relerr(A.coefficients(), cref) < 1e-12
        xs = np.linspace(-1.0, 1.0, 200, endpoint=False)
        if N >= 30 and k >= 4:
            assert np.max(np.abs(A(xs) - f(xs))) < 5e-3
        assert relerr(A(xs), oracle.spline_eval(1, k, data, cref, xs)) < 1e-12


def test_galerkin_periodic_too_small(bk, synth):
    B = bk.PeriodicBSplineBasis(4, np.linspace(0.0, 1.0, 6))      # N = 5 < 2 k - 1
    with pytest.raises(bk.UnsupportedError):
        bk.galerkin_matrix(B, as_band=True)


def test_approximate_minimise_l2_golden(bk):
    """approximate(exp, B, MinimiseL2Error()) docstring values (SplineApproximations.jl:104-119)."""
    g = _goldens()["class_C"]["approximate_k3"]
    B = bk.BSplineBasis(g["k"], dec_range(g["breaks"]["a"], g["breaks"]["b"], g["breaks"]["n"]))
    A = bk.approximate(np.exp, B, bk.MinimiseL2Error())
    assert np.allclose(A.coefficients(), g["exp_l2_coefs_6digits"], rtol=6e-6)
    assert A(0.34) == pytest.approx(g["exp_S_opt(0.34)"], abs=5e-15)
    bk.approximate_(np.sin, A)                                  # approximate! reuses the factorisation
    xs = np.linspace(-1, 1, 101)
    assert np.max(np.abs(A(xs) - np.sin(xs))) < 2e-2
    # polynomials of degree k - 1 are reproduced exactly (test/approximation.jl:38-49)
    P = bk.approximate(lambda x: 1 - 2 * x + 3 * x * x, B, bk.MinimiseL2Error())
    assert np.max(np.abs(P(xs) - (1 - 2 * xs + 3 * xs * xs))) < 1e-13


def test_fit_smoothing_periodic(bk, oracle, synth):
    """fit(BSplineOrder(4), xs, ys, λ, Periodic(L); weights) (smoothing.jl:126-213): cyclic 7-band normal
    matrix through the cyclic banded solve; against the oracle's dense restatement, λ = 0 == periodic
    interpolation, periodicity of the result, a batch of data sets."""
    import torch
    N, L = 200, 2.0
    xs = -1.0 + L * synth.breakpoints(71, N + 1)[:-1]
    ys = np.cos(np.pi * xs) + 0.1 * np.sin(40 * np.pi * xs)
    w = 0.5 + synth.u_numpy(72, 0, N)
    for lam, wt in ((1e-4, None), (1e-6, w), (0.0, None)):
        S = bk.fit(bk.BSplineOrder(4), xs, ys, lam, bk.Periodic(L), weights=wt)
        cref, data = oracle.fit_smoothing_periodic(xs, ys, lam, L, weights=wt)
        xq = np.linspace(-1, 1, 301)[:-1]
        ref = oracle.spline_eval(1, 4, data, cref, xq)
        assert relerr(S(xq), ref) < 1e-9, lam                      # normal equations: cond up to ~1e6 here
        assert np.max(np.abs(S(xq) - S(xq + L))) < 1e-12
    assert np.max(np.abs(S(xs) - ys)) < 1e-10                       # λ = 0 interpolates
    Y = np.stack([ys, 2 * ys, np.sin(np.pi * xs)], axis=1)
    SB = bk.fit(bk.BSplineOrder(4), xs, torch.from_numpy(Y).cuda(), 1e-4, bk.Periodic(L))
    c0, _ = oracle.fit_smoothing_periodic(xs, Y[:, 2], 1e-4, L)
    assert relerr(SB.coefs[:, 2].cpu().numpy(), c0) < 1e-9


@pytest.mark.parametrize("k", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_eval_awkward_knot_vectors(bk, oracle, synth, k, dtype):
    """Repeated interior knots (up to multiplicity k: discontinuous splines), clusters of nearly equal
    knots, shifted / scaled domains: the table paths must pick exactly searchsortedlast's interval
    (src/BSplines/evaluate_all.jl:102-119) and stay within tolerance; de Boor stays bit-exact."""
    import torch
    T = np.float64 if dtype == "f64" else np.float32
    tol = TOL64 if dtype == "f64" else TOL32
    rng = np.random.default_rng(1000 + k)
    for case, (a, b) in enumerate(((0.0, 1.0), (1.0e3, 1.0e3 + 2.0), (-3.0e-3, 5.0e-3))):
        base = np.sort(rng.random(40))
        base[0], base[-1] = 0.0, 1.0
        pieces = [base]
        for j in rng.choice(np.arange(1, 39), size=6, replace=False):        # repeated knots
            pieces.append(np.repeat(base[j], rng.integers(1, k)))
        if dtype == "f64":
            for j in rng.choice(np.arange(1, 39), size=4, replace=False):    # clusters
                pieces.append(base[j] + 1e-9 * np.arange(1, 4))
        br = np.sort(np.concatenate(pieces))
        br = (a + (b - a) * br).astype(T)
        br[0], br[-1] = T(a), T(b)
        br = np.sort(br)
        t = oracle.augment_knots(br, k).astype(T)
        B = bk.BSplineBasis(k, t, augment=False)
        c = (2 * rng.random(len(B)) - 1).astype(T)
        S = bk.Spline(B, c)
        x = np.concatenate([a + (b - a) * rng.random(20000), br.astype(np.float64),
                            np.nextafter(br, T(-np.inf)).astype(np.float64), np.nextafter(br, T(np.inf)).astype(np.float64),
                            [a - 1.0, b + 1.0]]).astype(T)
        xd = torch.from_numpy(x).cuda()
        kt = dtype
        ref = oracle.spline_eval(0, k, t, c, x, kt=kt)
        assert np.array_equal(S(xd, algo=bk.EVAL_DEBOOR).cpu().numpy(), ref), (case,)
        scale = np.max(np.abs(ref))
        for algo in (bk.EVAL_AUTO, bk.EVAL_PP):
            got = S(xd, algo=algo).cpu().numpy()
            bad = np.abs(got.astype(np.float64) - ref.astype(np.float64)) > tol * scale
            # a point may only differ beyond rounding if it sits on a discontinuity side that the
            # reference resolves the same way: there must be none
            assert not bad.any(), (case, algo, x[bad][:5], got[bad][:5], ref[bad][:5])
        # every derivative order at the SAME bar (1e-13 / 1e-5): the table builder certifies each cell and each
        # interval record for all d = 0 .. k-1 and routes what it cannot certify (the 1e-9 clusters) to the
        # bit-faithful de Boor recurrence (csrc/pptable.h)
        for d in range(1, k):
            kd, td, cd = oracle.spline_derivative(0, k, t, c, d, kt=kt)
            refd = oracle.spline_eval(0, kd, td, cd, x, kt=kt)
            sc = np.max(np.abs(refd))
            for algo in (bk.EVAL_AUTO, bk.EVAL_PP):
                gotd = S.evaluate(xd, (d,), algo=algo)[0].cpu().numpy()
                badd = np.abs(gotd.astype(np.float64) - refd.astype(np.float64)) > tol * sc
                assert not badd.any(), (case, d, algo, x[badd][:5], gotd[badd][:5], refd[badd][:5])
        if k >= 3:                                   # fused outputs agree with the single-output launches
            both = S.evaluate(xd, (0, 1))
            assert torch.equal(both[1], S.evaluate(xd, (1,))[0])


@pytest.mark.parametrize("k", [2, 4, 5, 8])
def test_knot_search_and_basis_awkward_knot_vectors(bk, oracle, k):
    """find_knot_interval / evaluate_all on knot vectors with repeated and clustered interior knots:
    indices, zones and all k basis values bit-exact (src/BSplines/evaluate_all.jl:102-204)."""
    import torch
    rng = np.random.default_rng(2000 + k)
    base = np.sort(rng.random(60))
    base[0], base[-1] = 0.0, 1.0
    pieces = [base]
    for j in rng.choice(np.arange(1, 59), size=10, replace=False):
        pieces.append(np.repeat(base[j], rng.integers(1, k)))
    for j in rng.choice(np.arange(1, 59), size=6, replace=False):
        pieces.append(base[j] + 1e-12 * np.arange(1, 5))
    br = np.sort(np.concatenate(pieces))
    t = oracle.augment_knots(br, k)
    B = bk.BSplineBasis(k, t, augment=False)
    x = np.concatenate([rng.random(5000), br, np.nextafter(br, -np.inf), np.nextafter(br, np.inf), [-0.5, 1.5, np.nan]])
    xd = torch.from_numpy(x).cuda()
    idx, zone = bk.find_knot_interval(B, xd)
    ridx, rzone = oracle.find_knot_interval(0, k, t, x)
    assert np.array_equal(idx.cpu().numpy(), ridx) and np.array_equal(zone.cpu().numpy(), rzone)
    xin = x[:-3]                                               # inside the domain
    for nd in range(0, min(3, k)):
        i, bs = bk.evaluate_all(B, torch.from_numpy(xin).cuda(), bk.Derivative(nd))
        ri, rbs = oracle.evaluate_all(0, k, t, xin, nderiv=nd)
        assert np.array_equal(i.cpu().numpy(), ri)
        assert np.array_equal(bs.cpu().numpy(), rbs, equal_nan=True), nd


@pytest.mark.parametrize("k", [3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("uneven", [False, True])
def test_eval_cell_jump_encoding(bk, oracle, synth, k, uneven):
    """Shared-memory cell tables of Float64 splines of order >= 3 keep, for a cell holding one simple
    knot xi, the piece left of it plus {xi, J}: right of the knot the kernel evaluates
    P_L + J (x - xi)^(k-1).  Values and every derivative order, at the knots themselves, one ulp either
    side of them and at random points, against the oracle (src/Splines/spline.jl:144-209,272-330);
    very uneven neighbouring intervals must be caught by the accuracy guard of the table builder."""
    import torch
    nb = 1500 if k == 3 else (1000 if k <= 6 else 600)
    br = synth.breakpoints(3, nb)
    if uneven:
        rng = np.random.default_rng(77 + k)
        extra = []
        for j in rng.choice(np.arange(5, nb - 5), size=60, replace=False):   # tiny intervals next to long ones
            extra.append(br[j] + (br[j + 1] - br[j]) * 10.0 ** (-rng.integers(2, 7)))
        br = np.sort(np.concatenate([br, extra]))
    B = bk.BSplineBasis(k, br)
    t = B.knots()
    c = 2 * synth.u_numpy(4, 0, len(B)) - 1
    S = bk.Spline(B, c)
    x = np.concatenate([synth.u_numpy(5, 0, 300000), br, np.nextafter(br, -np.inf), np.nextafter(br, np.inf)])
    xd = torch.from_numpy(x).cuda()
    for derivs in ((0, 1), tuple(range(2, min(k, 6))), tuple(range(6, k))):
        if not derivs:
            continue
        outs = S.evaluate(xd, derivs)
        for d, got in zip(derivs, outs):
            kd, td, cd = (k, t, c) if d == 0 else oracle.spline_derivative(0, k, t, c, d)
            ref = oracle.spline_eval(0, kd, td, cd, x)
            assert relerr(got.cpu().numpy(), ref) < TOL64, (d,)
    # single-output launches (compile-time derivative mask) and the run-time list agree bit for bit
    v0 = S.evaluate(xd, (0,))[0]
    v02 = S.evaluate(xd, (0, 2))[0]
    assert torch.equal(v0, v02)


@pytest.mark.parametrize("k", [5, 6, 7, 8])
def test_eval_cell_global_wide_records(bk, oracle, synth, k):
    """Float64 orders 5..8 on thousands of intervals: cell tables in global memory with 64-byte records that
    lane pairs fetch half by half (full warps) or each lane alone (ragged last warp, unaligned slices).
    Values and derivatives against the oracle, at knots, one ulp either side of them and at random points."""
    import torch
    br = synth.breakpoints(3, 4000)
    B = bk.BSplineBasis(k, br)
    t = B.knots()
    c = 2 * synth.u_numpy(4, 0, len(B)) - 1
    S = bk.Spline(B, c)
    x = np.concatenate([synth.u_numpy(5, 0, 200003), br, np.nextafter(br, -np.inf), np.nextafter(br, np.inf),
                        [-0.5, 1.5, np.nan]])
    xd = torch.from_numpy(x).cuda()
    refs = {}
    for d in (0, 1, 2):
        kd, td, cd = (k, t, c) if d == 0 else oracle.spline_derivative(0, k, t, c, d)
        refs[d] = oracle.spline_eval(0, kd, td, cd, x[:-1])
    for derivs in ((0,), (0, 1), (0, 1, 2)):
        outs = S.evaluate(xd, derivs)
        for d, got in zip(derivs, outs):
            g = got.cpu().numpy()
            assert np.isnan(g[-1])
            assert relerr(g[:-1], refs[d]) < TOL64, (derivs, d)
    # short and unaligned streams: the last warp is ragged, head / tail elements take the scalar path
    for lo, hi in ((0, 37), (1, 1000), (3, 65)):
        got = S.evaluate(xd[lo:hi], (0,))[0].cpu().numpy()
        assert relerr(got, refs[0][lo:hi]) < TOL64


@pytest.mark.parametrize("k", [2, 4])
@pytest.mark.parametrize("nb", [10000, 60000])
def test_float32_large_tables_pick_the_right_piece(bk, oracle, synth, k, nb):
    """Float32 splines on 10^4 .. 10^5 intervals (cell tables of up to 2 x 10^6 cells in global memory): the
    cell index is formed in Float32 on the device and drifts against the host's cell edges by ~4 eps ncell
    cells; the builder's margin (csrc/pptable.h) covers it.  Points at the knots and their nextafter
    neighbours, value and D^(k-1) S (piecewise constant: a wrong piece is an O(1) error); the same inputs run
    on the CPU emulation in tests/test_pptable_cpu.py."""
    import torch
    brf = synth.breakpoints(10, nb).astype(np.float32)
    B = bk.BSplineBasis(k, brf)
    t = B.knots()
    c = (2 * synth.u_numpy(4, 0, len(B)) - 1).astype(np.float32)
    S = bk.Spline(B, c)
    x = np.concatenate([synth.u_numpy(11, 0, 200000).astype(np.float32), brf, np.nextafter(brf, np.float32(-1)),
                        np.nextafter(brf, np.float32(2))])
    xd = torch.from_numpy(x).cuda()
    assert S.table_info()["cells_in_global"] == 1
    top = k - 1 if nb < 50000 else 1
    got = S.evaluate(xd, (0, top))
    for d, g in zip((0, top), got):
        kd, td, cd = (k, t, c) if d == 0 else oracle.spline_derivative(0, k, t, c, d, kt="f32")
        ref = oracle.spline_eval(0, kd, td, cd, x, kt="f32")
        sc = np.max(np.abs(ref))
        bad = np.abs(g.cpu().numpy().astype(np.float64) - ref.astype(np.float64)) > TOL32 * sc
        assert not bad.any(), (d, int(bad.sum()), x[bad][:3])


@pytest.mark.parametrize("k,n", [(3, 1024), (4, 3000), (6, 4096), (8, 100000), (5, 20011)])
def test_partitioned_lu(bk, oracle, synth, k, n):
    """lu! by partitions (BSK_LU_FAST, csrc/banfac.cuh): every partition is eliminated from the original matrix
    starting 128 columns early and hands a checked elimination state to the next one.  Factors within 1e-13 of
    BANFAC's (src/Collocation/matrix.jl:132-201) column by column, the validated halos and the solutions of the
    solve are those of the exact factorisation."""
    import torch
    xdata = synth.breakpoints(6 + k, n)
    t = bk.make_knots(xdata, k)
    B = bk.BSplineBasis(k, t, augment=False)
    band, nout = oracle.collocation_matrix(0, k, t, xdata)
    assert nout == 0
    lu = band.copy(order="F")
    assert oracle.banded_lu(lu) == 0
    Ce = bk.CollocationMatrix(band.copy(order="F")).lu_(bk.LU_EXACT)
    assert np.array_equal(Ce.data(), lu)
    Cf = bk.CollocationMatrix(band.copy(order="F")).lu_(bk.LU_FAST)
    f = Cf.data()
    assert not np.array_equal(f, band)                       # it did factorise
    err = np.max(np.abs(f - lu), axis=0) / np.max(np.abs(lu), axis=0)
    assert err.max() < 1e-13, (err.max(), int(np.argmax(err)))
    assert Cf.halo == Ce.halo
    Y = (2 * synth.u_numpy(7, 0, n * 64) - 1).reshape(n, 64)
    ref = oracle.banded_solve(lu, Y.copy())
    for Cm in (Ce, Cf):
        Z = torch.from_numpy(Y).cuda()
        Cm.ldiv_(Z)
        assert relerr(Z.cpu().numpy(), ref) < TOL64


def test_partitioned_lu_zero_pivot_and_fallback(bk, oracle):
    """A zero pivot deep inside a large matrix: BSK_LU_FAST must still raise ZeroPivotException(i) with BANFAC's
    row (the partition that meets it gives up, the serial kernel reports it).  A matrix whose elimination does
    not forget its start state (no diagonal dominance) must fall back to the exact factors."""
    n, bw = 5000, 2
    rng = np.random.default_rng(5)
    band = np.zeros((2 * bw + 1, n), order="F")
    band[bw, :] = 4.0 + rng.random(n)
    band[:bw, :] = rng.random((bw, n)) - 0.5
    band[bw + 1:, :] = rng.random((bw, n)) - 0.5
    sing = band.copy(order="F")
    sing[:, 3210] = 0.0                                    # column 3211 of the matrix is zero
    ref = sing.copy(order="F")
    info = oracle.banded_lu(ref)
    assert info > 0
    with pytest.raises(bk.ZeroPivotException) as ei:
        bk.CollocationMatrix(sing).lu_(bk.LU_FAST)
    assert ei.value.info == info
    # slowly decaying recurrence: off-diagonals as large as the diagonal
    slow = np.zeros((3, n), order="F")
    slow[1, :] = 2.0
    slow[0, 1:] = -0.999999
    slow[2, :-1] = -0.999999
    ref = slow.copy(order="F")
    assert oracle.banded_lu(ref) == 0
    got = bk.CollocationMatrix(slow.copy(order="F")).lu_(bk.LU_FAST).data()
    assert np.max(np.abs(got - ref) / np.max(np.abs(ref), axis=0)) < 1e-13


@pytest.mark.parametrize("kind", ["plain", "periodic", "robin"])
def test_collocation_rows_rectangular(bk, oracle, synth, kind):
    """collocation_matrix(B, x, deriv, Matrix / SparseMatrixCSC) with MORE points than basis functions
    (src/Collocation/Collocation.jl:82-106,162-193,209-251): row form (k slots per point) and dense form, bit-exact
    against the entries evaluate_all gives the reference loop; points outside the domain leave empty rows
    (Collocation.jl:224), entries <= clip_threshold are dropped (:237)."""
    import torch
    k = 5
    br = synth.breakpoints(21, 30)
    nx = 500
    if kind == "periodic":
        B = bk.PeriodicBSplineBasis(k, br)
        x = -0.7 + 2.4 * synth.u_numpy(22, 0, nx)
        okind, ot, rc = 1, br, None
    else:
        P = bk.BSplineBasis(k, br)
        ot = P.knots()
        okind = 0
        if kind == "robin":
            B = bk.RecombinedBSplineBasis(P, bk.Derivative(0) + 3 * bk.Derivative(1))
            rc = oracle.robin_corners(k, ot, [1.0, 3.0])
        else:
            B, rc = P, None
        x = np.concatenate([-0.1 + 1.2 * synth.u_numpy(22, 0, nx - 2), [0.0, 1.0]])
    N = len(B)
    eps = np.finfo(np.float64).eps
    for nd in (0, 2):
        i_ref, bs_ref = oracle.evaluate_all(okind, k, ot, x, nderiv=nd, recomb=rc)
        ref = np.zeros((nx, N))
        inside = np.ones(nx, dtype=bool) if kind == "periodic" else (x >= 0) & (x <= 1)
        for i in range(nx):
            if not inside[i]:
                continue
            for dj in range(1, k + 1):
                j = i_ref[i] + 1 - dj
                if kind == "periodic":
                    j = (j - 1) % N + 1
                if 1 <= j <= N and abs(bs_ref[i, dj - 1]) > eps:
                    ref[i, j - 1] = bs_ref[i, dj - 1]
        col, val = bk.collocation_rows(B, x, bk.Derivative(nd))
        col, val = col.cpu().numpy(), val.cpu().numpy()
        got = np.zeros((nx, N))
        for dj in range(k):
            m = col[:, dj] > 0
            got[np.nonzero(m)[0], col[m, dj] - 1] = val[m, dj]
        assert np.array_equal(got, ref), (kind, nd)
        assert np.all(col[~inside] == 0) and np.all(val[col == 0] == 0.0)
        dense = bk.collocation_rows(B, x, bk.Derivative(nd), dense=True).cpu().numpy()
        assert dense.shape == (nx, N) and np.array_equal(dense, ref)
    if kind == "plain":                                       # least squares with the rectangular matrix reproduces a spline
        c = 2 * synth.u_numpy(23, 0, N) - 1
        xin = synth.u_numpy(24, 0, 400)
        A = bk.collocation_rows(B, xin, dense=True).cpu().numpy()
        y = oracle.spline_eval(0, k, ot, c, xin)
        assert np.max(np.abs(np.linalg.lstsq(A, y, rcond=None)[0] - c)) < 1e-11


@pytest.mark.parametrize("k", [2, 4, 5, 7])
def test_integral(bk, oracle, synth, k):
    """integral(S) (src/Splines/spline.jl:354-374): zero at the left boundary (test/splines.jl:62-66), the
    derivative of the integral is S again (coefficients to rounding), I(b) - I(a) equals a Gauss-Legendre
    quadrature of S piece by piece; periodic splines (src/Splines/periodic.jl:122-140): differences of the
    integral agree with those of the equivalent non-periodic spline (test/periodic.jl:153-155)."""
    br = synth.breakpoints(3, 40)
    B = bk.BSplineBasis(k, br)
    c = 2 * synth.u_numpy(4, 0, len(B)) - 1
    S = bk.Spline(B, c)
    I = bk.integral(S)
    assert I.k == k + 1 and I.knots().size == B.knots().size + 2
    t = B.knots()
    ref = np.concatenate([[0.0], np.cumsum(c * (t[k:] - t[:-k]) / k)])
    assert np.array_equal(I.coefficients(), ref)
    assert I(np.array([0.0]), algo=bk.EVAL_DEBOOR)[0] == 0.0        # Sint(a) == 0 (test/splines.jl:66), de Boor order
    assert abs(I(np.array([0.0]))[0]) < 1e-15 * np.max(np.abs(ref))    # default (table) path: to rounding
    dI = bk.Derivative(1) * I
    assert np.max(np.abs(dI.coefficients() - c)) < 1e-13
    xs = synth.u_numpy(5, 0, 2000)
    assert relerr(dI(xs), S(xs)) < 1e-13
    # quadrature of S over [0, x1] piece by piece
    qx, qw = np.polynomial.legendre.leggauss(k)
    x1 = 0.77
    edges = np.concatenate([br[br < x1], [x1]])
    tot = 0.0
    for a_, b_ in zip(edges[:-1], edges[1:]):
        xm = 0.5 * (a_ + b_) + 0.5 * (b_ - a_) * qx
        tot += 0.5 * (b_ - a_) * np.dot(qw, oracle.spline_eval(0, k, t, c, xm))
    assert abs(I(np.array([x1]))[0] - tot) < 1e-14
    # periodic
    if k >= 3:
        L = 2.0
        brp = L * synth.breakpoints(8, 31) - 1.0
        Bp = bk.PeriodicBSplineBasis(k, brp)
        cp = 2 * synth.u_numpy(9, 0, len(Bp)) - 1
        Sp = bk.Spline(Bp, cp)
        Ip = bk.integral(Sp)
        assert Ip.k == k + 1
        assert abs(Ip(np.array([brp[0]]))[0]) < 1e-15
        xa, xb = -0.4, 0.55
        edges = np.concatenate([[xa], brp[(brp > xa) & (brp < xb)], [xb]])
        tot = 0.0
        for a_, b_ in zip(edges[:-1], edges[1:]):
            xm = 0.5 * (a_ + b_) + 0.5 * (b_ - a_) * qx
            tot += 0.5 * (b_ - a_) * np.dot(qw, oracle.spline_eval(1, k, brp, cp, xm))
        got = Ip(np.array([xb]))[0] - Ip(np.array([xa]))[0]
        assert abs(got - tot) < 1e-13
    with pytest.raises(bk.UnsupportedError):
        bk.integral(bk.Spline(bk.BSplineBasis(8, br), np.ones(len(br) + 6)))


@pytest.mark.parametrize("dt", ["int32", "complex64", "complex128", "f32x2"])
def test_interpolate_other_element_types(bk, synth, dt):
    """test/interpolation.jl:53-60 — data of type Int32, ComplexF32 (and ComplexF64), SVector{2,Float32} on
    Float64 points: `S.(xs) ≈ ys` with the coefficient type float(eltype(y))."""
    import torch
    n, k = 40, 4
    xs = np.sort(np.random.default_rng(42).standard_normal(n))
    u = lambda s: 2 * synth.u_numpy(s, 0, n) - 1
    if dt == "int32":
        ys = (np.floor(31 * synth.u_numpy(60, 0, n)) - 15).astype(np.int32)
        S = bk.interpolate(xs, ys, bk.BSplineOrder(k)).spline
        assert S.coefficients().dtype == np.float64
        assert np.allclose(S(xs), ys, rtol=1e-12, atol=1e-12)
    elif dt.startswith("complex"):
        ys = (u(61) + 1j * u(62)).astype(np.complex64 if dt == "complex64" else np.complex128)
        itp = bk.interpolate(xs, ys, bk.BSplineOrder(k))
        S = itp.spline
        assert itp.coefs.dtype == (torch.complex64 if dt == "complex64" else torch.complex128)
        got = S(xs)
        assert np.iscomplexobj(got)
        assert np.allclose(got, ys, rtol=1e-6 if dt == "complex64" else 1e-12, atol=1e-6 if dt == "complex64" else 1e-12)
        # linearity: the real part is the interpolant of the real part
        Sr = bk.interpolate(xs, ys.real.astype(np.float64), bk.BSplineOrder(k)).spline
        xq = np.linspace(xs[0], xs[-1], 301)
        assert np.max(np.abs(S(xq).real - Sr(xq))) < 1e-12
        # natural boundary conditions, complex data (test/interpolation.jl:24-47)
        Sn = bk.interpolate(xs, ys, bk.BSplineOrder(k), bk.Natural()).spline
        assert np.allclose(Sn(xs), ys, rtol=1e-6, atol=1e-6)
    else:
        ys = np.stack([u(63), u(64)], axis=1).astype(np.float32)
        itp = bk.interpolate(xs, ys, bk.BSplineOrder(k))
        got = itp.splines(torch.from_numpy(xs).cuda()).cpu().numpy()
        assert np.allclose(got, ys, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# stream ordering of the programmatically launched evaluation kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nbreaks,periodic", [(1000, False), (10000, True)])
def test_back_to_back_launches_see_the_previous_results(bk, oracle, synth, nbreaks, periodic):
    """The cell-table kernels are launched with programmatic stream serialization (their table staging may overlap
    the tail of the previous kernel on the stream); points and results are only touched after griddepcontrol.wait.
    A chain on ONE stream without any host synchronisation — torch kernel writes x, S(x) -> y, torch kernel maps y
    back into the domain, S(y') -> z, several rounds — must give what the same chain gives step by step."""
    import torch
    k = 4
    if periodic:
        br = synth.breakpoints(8, nbreaks + 1)
        B = bk.PeriodicBSplineBasis(k, br)
        kind, tref = 1, br
    else:
        B = bk.BSplineBasis(k, synth.breakpoints(3, nbreaks))
        kind, tref = 0, B.knots()
    c = 2 * synth.u_numpy(4, 0, len(B)) - 1
    S = bk.Spline(B, c)
    n = 1 << 21
    x0 = synth.u_torch(41, 0, n)
    S.evaluate(x0, (0,))                                   # tables built, nothing pending
    torch.cuda.synchronize()
    x = torch.empty_like(x0)
    y = [torch.empty_like(x0)]
    z = [torch.empty_like(x0)]
    for rnd in range(6):
        x.copy_(x0).mul_(0.5 + 0.07 * rnd)                 # producer kernels on the same stream
        S.evaluate(x, (0,), out=y)
        y[0].abs_().mul_(0.25).clamp_(0.0, 0.999)          # consumer / producer in between
        S.evaluate(y[0], (0,), out=z)
        S.evaluate(z[0].abs_().mul_(0.25).clamp_(0.0, 0.999), (0,), out=y)
    torch.cuda.synchronize()
    # the same chain on the host with the oracle, last round only (rounds are independent)
    xh = x0.cpu().numpy()[: 1 << 16] * (0.5 + 0.07 * 5)
    y1 = oracle.spline_eval(kind, k, tref, c, xh)
    y1 = np.clip(np.abs(y1) * 0.25, 0.0, 0.999)
    z1 = oracle.spline_eval(kind, k, tref, c, y1)
    z1 = np.clip(np.abs(z1) * 0.25, 0.0, 0.999)
    y2 = oracle.spline_eval(kind, k, tref, c, z1)
    # each stage amplifies the 1e-15 of the previous one by |S'| ~ 1e3 .. 1e4: compare with that in mind
    got = y[0].cpu().numpy()[: 1 << 16]
    assert np.max(np.abs(got - y2)) < 1e-6 * np.max(np.abs(y2))
    # and exactly reproducible against a fully synchronised replay of the same chain on the device
    xs = (x0 * (0.5 + 0.07 * 5))
    torch.cuda.synchronize()
    a = S.evaluate(xs, (0,))[0]; torch.cuda.synchronize()
    a = a.abs().mul(0.25).clamp(0.0, 0.999); torch.cuda.synchronize()
    b = S.evaluate(a, (0,))[0]; torch.cuda.synchronize()
    b = b.abs().mul(0.25).clamp(0.0, 0.999); torch.cuda.synchronize()
    cfin = S.evaluate(b, (0,))[0]; torch.cuda.synchronize()
    assert torch.equal(cfin, y[0])


@pytest.mark.parametrize("stage", ["1", None])
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_staged_and_unstaged_kernels_agree(bk, oracle, synth, monkeypatch, stage, dt):
    """K1 / K2 / K3a stage knots + LUT in shared memory only when the point stream is long enough to amortise the
    staging (short calls read them through L1 / L2): both variants of every kernel against the oracle, bit for bit,
    on the same small inputs (BSK_ALWAYS_STAGE forces the staged variant)."""
    if stage:
        monkeypatch.setenv("BSK_ALWAYS_STAGE", stage)
    else:
        monkeypatch.delenv("BSK_ALWAYS_STAGE", raising=False)
    npdt = np.float64 if dt == "f64" else np.float32
    for k, periodic in ((4, False), (6, False), (5, True), (4, True)):
        if periodic:
            br = (2.0 * synth.breakpoints(8, 301) - 1.0).astype(npdt)
            B = bk.PeriodicBSplineBasis(k, br)
            kind, tref = 1, br
            x = _points(synth, 51, 3000, -3.0, 3.0, extra=list(br)).astype(npdt)
        else:
            br = synth.breakpoints(3, 300).astype(npdt)
            B = bk.BSplineBasis(k, br)
            kind, tref = 0, B.knots()
            x = _points(synth, 52, 3000, -0.05, 1.05, extra=list(tref)).astype(npdt)
        i_ref, z_ref = oracle.find_knot_interval(kind, k, tref, x, kt=dt)
        i, z = bk.find_knot_interval(B, x)
        assert np.array_equal(i, i_ref) and np.array_equal(z, z_ref)
        for nd in (0, 1, k - 1):
            ir, bsr = oracle.evaluate_all(kind, k, tref, x, nderiv=nd, kt=dt, vt=dt)
            ig, bsg = bk.evaluate_all(B, x, bk.Derivative(nd))
            assert np.array_equal(ig, ir) and np.array_equal(bsg, bsr, equal_nan=True)
        c = (2 * synth.u_numpy(4, 0, len(B)) - 1).astype(npdt)
        S = bk.Spline(B, c)
        xin = x[(x >= tref[0]) & (x <= tref[-1])] if not periodic else x
        got = S.evaluate(xin, (0,), algo=bk.EVAL_DEBOOR)[0]
        assert np.array_equal(got, oracle.spline_eval(kind, k, tref, c, xin, kt=dt), equal_nan=True)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 6])
def test_basis_function_against_cox_de_boor(bk, oracle, synth, k):
    """B[i](x, op) / evaluate(B, i, x, op) (BSplines/basis_function.jl:114-197) is picked out of the batched
    evaluate_all; the oracle's orc_basis_function is the reference's own Cox-de Boor recursion (:146-256) — two
    independent algorithms (the comparison test/bsplines.jl:76-92 makes).  Also support / nonzero_in_segment /
    isindomain / DomainError, device input, recombined and periodic bases."""
    import torch
    br = synth.breakpoints(21, 12)
    B = bk.BSplineBasis(k, br)
    t, N = B.knots(), len(B)
    x = np.concatenate([_points(synth, 22, 400, -0.05, 1.05), br])
    for nd in range(0, min(k, 3)):
        for i in sorted({1, 2, max(1, N // 2), N - 1, N} & set(range(1, N + 1))):
            ref = np.array([oracle.basis_function(k, t, i, float(v), nd) for v in x])
            got = bk.evaluate(B, i, x, bk.Derivative(nd))
            scale = max(1.0, np.max(np.abs(ref)))
            assert np.max(np.abs(got - ref)) < 1e-12 * scale, (nd, i)
            assert np.array_equal(B[i](x, bk.Derivative(nd)), got)
    # scalar and device input
    assert bk.evaluate(B, 1, float(t[0])) == 1.0 and B[N](float(t[-1])) == 1.0
    xd = torch.from_numpy(x).cuda()
    assert np.array_equal(bk.evaluate(B, 2 if N > 1 else 1, xd).cpu().numpy(), bk.evaluate(B, 2 if N > 1 else 1, x))
    with pytest.raises(bk.DomainError):
        bk.evaluate(B, 0, 0.5)
    with pytest.raises(bk.DomainError):
        bk.evaluate(B, N + 1, 0.5)
    assert bk.support(B, 3) == (3, 3 + k) and bk.support(B[3]) == (3, 3 + k)
    assert bk.nonzero_in_segment(B, k) == (1, k) and bk.nonzero_in_segment(B, N + 5)[1] == N
    assert bk.isindomain(B, 0.0) and bk.isindomain(B, 1.0) and not bk.isindomain(B, 1.0000001)
    # partition of unity from single functions
    tot = sum(bk.evaluate(B, i, x) for i in range(1, N + 1))
    inside = (x >= 0) & (x <= 1)
    assert np.allclose(tot[inside], 1.0, rtol=0, atol=1e-14) and np.all(tot[~inside] == 0)
    if k >= 3:
        # recombined basis: phi_j = sum_i A[i, j] b_i
        R = bk.RecombinedBSplineBasis(B, bk.Derivative(1))
        A = R.recombination_matrix()
        xin = x[inside]
        allb = np.array([[oracle.basis_function(k, t, i, float(v), 0) for v in xin] for i in range(1, N + 1)])
        for j in (1, 2, len(R)):
            assert np.max(np.abs(bk.evaluate(R, j, xin) - A[:, j - 1] @ allb)) < 1e-13
        # periodic basis: b_j of the wrapped basis, evaluated at points far from the main period too
        P = bk.PeriodicBSplineBasis(k, br)
        Np = len(P)
        xp = np.concatenate([x, x + 3.0, x - 2.0])
        totp = sum(bk.evaluate(P, j, xp) for j in range(1, Np + 1))
        assert np.allclose(totp, 1.0, rtol=0, atol=1e-13)
        assert bk.isindomain(P, 17.0)


@pytest.mark.parametrize("k", [3, 4, 5, 6])
def test_polynomial_reproduction_and_wrappers(bk, oracle, k):
    """test/splines.jl:13-81 (test_polynomial) restated: a polynomial of degree k - 1 is interpolated exactly,
    its derivative and antiderivative splines follow it on a fine grid; the wrapper forwards (order / basis /
    knots / coefficients / integral / diff of an interpolation) and the two DimensionMismatch cases."""
    x = np.arange(12, dtype=np.float64) ** 1.3 / 6.0
    P = [float((-d) ** d) for d in range(1, k + 1)]          # P(x) = -1 + 4x - 27x^2 + ... (ntuple(d -> (-d)^d))
    dP = [d * P[d] for d in range(1, k)]
    Pint = [0.0] + [P[d] / (d + 1) for d in range(k)]
    ev = lambda c, v: np.polyval(c[::-1], v)
    y = ev(P, x)
    itp = bk.interpolate(x, y, bk.BSplineOrder(k))
    assert repr(itp).startswith("SplineInterpolation containing")
    S = bk.spline(itp)
    assert len(S) == len(x)
    assert np.allclose(itp(x), y, rtol=1e-12, atol=1e-12 * np.max(np.abs(y)))
    xm = (x[1] + x[2]) / 3
    assert itp(np.array([xm]))[0] == S(np.array([xm]))[0]
    assert bk.order(itp) == bk.order(S) and bk.basis(itp) is bk.basis(S)
    assert np.array_equal(bk.knots(itp), bk.knots(S)) and np.array_equal(bk.coefficients(itp), bk.coefficients(S))
    assert bk.integral(itp) == bk.integral(S)
    assert bk.diff(itp, bk.Derivative(1)) == bk.diff(S, bk.Derivative(1)) == bk.Derivative() * itp
    with pytest.raises(bk.DimensionMismatch):
        bk.SplineInterpolation(bk.basis(S), x[:4])            # "incompatible lengths of B-spline basis and collocation points"
    with pytest.raises(bk.DimensionMismatch):
        itp.interpolate_(np.random.rand(len(x) - 1))           # "input data has incorrect length"
    dS = bk.diff(S, bk.Derivative(1))
    assert bk.Derivative() * S == dS
    Sint = bk.integral(S)
    a, b = bk.boundaries(bk.basis(S))
    assert Sint(np.array([a]), algo=bk.EVAL_DEBOOR)[0] == 0.0
    xs = np.linspace(a, b, 9 * len(S) + 42)
    sc = np.max(np.abs(ev(P, xs)))
    assert np.max(np.abs(S(xs) - ev(P, xs))) < 1e-11 * sc
    assert np.max(np.abs(dS(xs) - ev(dP, xs))) < 1e-10 * np.max(np.abs(ev(dP, xs)))
    assert np.max(np.abs(Sint(xs) - (ev(Pint, xs) - ev(Pint, a)))) < 1e-11 * max(1.0, np.max(np.abs(ev(Pint, xs))))


@pytest.mark.parametrize("k", [2, 4, 5])
def test_bsplines_reference_assertions(bk, k):
    """test/splines.jl:84-140 (test_splines) restated: knot multiplicities of the augmented knots, DomainError /
    BoundsError, derivative ranges of a basis function, nonzero_in_segment near the boundaries, values at the
    boundaries, repr."""
    knots_in = np.array([0.0, 0.1, 0.25, 0.3, 0.55, 0.6, 0.8, 1.0])
    B = bk.BSplineBasis(k, knots_in)
    t, N = bk.knots(B), len(B)
    assert np.all(t[:k] == knots_in[0]) and np.all(t[-k:] == knots_in[-1])
    assert np.array_equal(t[k:-k], knots_in[1:-1])
    with pytest.raises(bk.DomainError):
        bk.evaluate(B, 0, 0.2)
    with pytest.raises(bk.DomainError):
        bk.evaluate(B, N + 1, 0.2)
    b3 = B[3]
    assert isinstance(b3, bk.BasisFunction)
    rng = bk.Derivative(range(0, min(k, 3)))
    assert b3(0.2, rng) == tuple(b3(0.2, bk.Derivative(d)) for d in range(0, min(k, 3)))
    assert tuple(bk.Derivative(range(2, 6))) == tuple(bk.Derivative(d) for d in (2, 3, 4, 5))
    for bad in (0, N + 1):
        with pytest.raises(IndexError):
            B[bad]
    rng_of = lambda r: list(range(r[0], r[1] + 1))
    assert rng_of(bk.nonzero_in_segment(B, k + 1)) == list(range(2, k + 2))
    assert rng_of(bk.nonzero_in_segment(B, 1)) == [1]
    assert rng_of(bk.nonzero_in_segment(B, N)) == list(range(N - k + 1, N + 1))
    assert rng_of(bk.nonzero_in_segment(B, N + 1)) == list(range(N - k + 2, N + 1))
    assert rng_of(bk.nonzero_in_segment(B, N + k - 1)) == [N]
    assert rng_of(bk.nonzero_in_segment(B, N + k)) == [] and rng_of(bk.nonzero_in_segment(B, 0)) == []
    assert rng_of(bk.nonzero_in_segment(B, t.size)) == []
    a, b = bk.boundaries(B)
    assert bk.evaluate(B, 1, a) == 1.0 and bk.evaluate(B, N, b) == 1.0
    assert repr(B).startswith("%d-element BSplineBasis of order %d" % (N, k))


def test_extrapolation_reference_assertions(bk, synth):
    """test/extrapolation.jl:5-47 restated.  The reference's `==` identities (itp(a) == ext(a - 1.3), ...) hold bit
    for bit on the de Boor path (the reference's own arithmetic); on the default table path inside and outside
    values come from different, individually certified routes and agree to rounding (asserted at 4 ulp)."""
    xdata = np.round(np.arange(-1.0, 1.0001, 0.2), 10)
    ydata = 2 * synth.u_numpy(33, 0, xdata.size) - 1
    itp = bk.interpolate(xdata, ydata, bk.BSplineOrder(4))
    itp_nat = bk.interpolate(xdata, ydata, bk.BSplineOrder(4), bk.Natural())
    assert repr(bk.extrapolate(itp, bk.Flat())).startswith("SplineExtrapolation containing the")
    for algo, same in ((bk.EVAL_DEBOOR, lambda a, b: a == b),
                       (bk.EVAL_AUTO, lambda a, b: abs(a - b) <= 4 * np.finfo(float).eps * max(abs(a), abs(b), 1.0))):
        one = lambda f, v: float(f(np.array([v]), algo=algo)[0])
        ext = bk.extrapolate(itp, bk.Flat())
        assert same(one(itp, 0.42), one(ext, 0.42))
        assert same(one(itp, -1.0), one(ext, -2.3)) and same(one(itp, 1.0), one(ext, 2.4))
        exs = bk.extrapolate(itp_nat, bk.Smooth())
        for v in (0.42, -1.0, 1.0):
            assert same(one(itp_nat, v), one(exs, v))
        # natural BCs: zero second derivative at the boundary, the smooth continuation is locally linear
        assert one(exs, -1.01) - one(exs, -1.0) == pytest.approx(one(exs, -1.0) - one(exs, -0.99), rel=1e-3)
        assert one(exs, 1.01) - one(exs, 1.0) == pytest.approx(one(exs, 1.0) - one(exs, 0.99), rel=1e-3)
        if algo == bk.EVAL_DEBOOR:
            continue                                   # Linear needs the local-polynomial path (documented)
        exl = bk.extrapolate(itp, bk.Linear())
        for v in (0.42, -1.0, 1.0):
            assert same(one(itp, v), one(exl, v))
        dS = bk.Derivative() * itp
        assert one(exl, -1.1) == pytest.approx(one(itp, -1.0) - 0.1 * one(dS, -1.0), rel=1e-13)
        assert one(exl, 1.1) == pytest.approx(one(itp, 1.0) + 0.1 * one(dS, 1.0), rel=1e-13)


def test_smoothing_fit_minimises_its_objective(bk, synth):
    """test/smoothing.jl:9-200 restated without automatic differentiation: the objective
    sum_i w_i (y_i - S(x_i))^2 + lambda * int S''^2 is quadratic in the coefficients, its gradient
    2 A' W (A c - y) + 2 lambda Q c (A: collocation rows, Q: Galerkin matrix of (D^2, D^2), both assembled by the
    device kernels, independently of `fit`) vanishes at the fitted coefficients — projected on the natural
    recombination for `Natural()`, directly for `Periodic(L)`; curvature decreases and the distance from the data
    grows with lambda; a heavier weight pulls the curve to its point."""
    D2 = bk.Derivative(2)

    def dense_band(band, N):
        w = band.cpu().numpy().T
        bw = (w.shape[0] - 1) // 2
        A = np.zeros((N, N))
        for j in range(N):
            for i in range(max(0, j - bw), min(N, j + bw + 1)):
                A[i, j] = w[bw + i - j, j]
        return A

    def dense_cyclic(band, N):
        w = band.cpu().numpy().T
        bw = (w.shape[0] - 1) // 2
        A = np.zeros((N, N))
        for j in range(N):
            for d in range(-bw, bw + 1):
                A[(j + d) % N, j] = w[bw + d, j]
        return A

    def check(S, xs, ys, lam, weights=None):
        R = S.basis
        w = np.ones_like(xs) if weights is None else weights
        if isinstance(R, bk.PeriodicBSplineBasis):
            Bp, T = R, None
            Q = dense_cyclic(bk.galerkin_matrix(R, (D2, D2), as_band=True), len(R))
        else:
            Bp, T = R.parent, R.recombination_matrix()
            Q = dense_band(bk.galerkin_matrix(Bp, (D2, D2), as_band=True), len(Bp))
        A = bk.collocation_rows(Bp, xs, dense=True).cpu().numpy()
        c = S.coefficients()
        g = 2 * A.T @ (w * (A @ c - ys)) + 2 * lam * (Q @ c)
        # scale of every component: |M| |c| + |b| for M = 2 (A' W A + lambda Q), b = 2 A' W y.  The knots of this
        # data are clustered ((0:0.01:1).^2), Q spans 12 orders of magnitude: a plain norm would compare the
        # boundary rows with the interior ones; the gradient must vanish to rounding of ITS OWN row (backward error)
        M = 2 * (A.T @ (w[:, None] * A) + lam * Q)
        scale = np.abs(M) @ np.abs(c) + np.abs(2 * A.T @ (w * ys))
        if T is not None:
            g, scale = T.T @ g, np.abs(T.T) @ scale
        assert np.max(np.abs(g) / scale) < 1e-11
        # the reference's own measure (test/smoothing.jl:74-77): roundoff-limited by cond(M) ~ 1e10 at lambda = 1e-2,
        # where dense numpy LU / Cholesky of the same system reach 1e-11 and this solve 2e-10
        assert np.sum(g * g) / np.sum(ys * ys) < 1e-8
        return float(c @ Q @ c), float(np.sqrt(np.mean((A @ c - ys) ** 2)))

    xs = np.arange(0.0, 1.0001, 0.01) ** 2
    ys = np.cos(2 * np.pi * xs) + 0.1 * np.sin(200 * np.pi * xs)
    order = bk.BSplineOrder(4)
    S0 = bk.fit(order, xs, ys, 0.0)
    Si = bk.spline(bk.interpolate(xs, ys, order, bk.Natural()))
    assert np.allclose(S0.coefficients(), Si.coefficients(), rtol=1e-9, atol=1e-9)
    curv, dist = [], []
    for lam in (0.0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2):
        c_, d_ = check(bk.fit(order, xs, ys, lam), xs, ys, lam)
        curv.append(c_); dist.append(d_)
    assert curv == sorted(curv, reverse=True) and dist == sorted(dist)
    lam = 1e-3
    wts = np.ones_like(xs)
    S, Sw = bk.fit(order, xs, ys, lam), bk.fit(order, xs, ys, lam, weights=wts)
    check(Sw, xs, ys, lam, wts)
    assert np.array_equal(S.coefficients(), Sw.coefficients())
    wts[2] = 1000.0
    Sw = bk.fit(order, xs, ys, lam, weights=wts)
    cw, dw = check(Sw, xs, ys, lam, wts)
    c1, d1 = check(S, xs, ys, lam)
    one = lambda f, v: float(f(np.array([v]))[0])
    assert abs(one(Sw, xs[2]) - ys[2]) < abs(one(S, xs[2]) - ys[2])
    assert cw > c1 and dw < d1
    # periodic
    N = 100
    xp = np.array([-np.cos(np.pi * n / N) for n in range(N)])
    yp = np.cos(np.pi * xp) + 0.1 * np.sin(200 * np.pi * xp)
    check(bk.fit(order, xp, yp, 1e-4, bk.Periodic(2.0)), xp, yp, 1e-4)
    wp = np.ones_like(xp); wp[2] = 1000.0
    check(bk.fit(order, xp, yp, 1e-4, bk.Periodic(2.0), weights=wp), xp, yp, 1e-4, wp)


@pytest.mark.parametrize("k", [3, 4, 6, 8])
def test_interpolation_reference_assertions(bk, k):
    """test/interpolation.jl:4-52 restated (Float64 data; the other element types are covered above): interpolation
    at 40 sorted normal deviates without and with natural boundary conditions — S.(xs) ≈ ys, the unique knots of the
    natural interpolant are the data locations, derivatives 2 .. k/2 vanish at both boundaries."""
    rng = np.random.default_rng(42)
    xs = np.sort(rng.standard_normal(40))
    ys = rng.standard_normal(40)
    S = bk.spline(bk.interpolate(xs, ys, bk.BSplineOrder(k)))
    assert np.allclose(S(xs), ys, rtol=1e-9, atol=1e-9)
    xi = np.arange(11) ** 2                                     # "Integer xdata" (:60-65)
    yi = rng.random(xi.size)
    assert np.allclose(bk.interpolate(xi, yi, bk.BSplineOrder(4))(xi.astype(float)), yi, rtol=1e-12, atol=1e-12)
    if k % 2:
        return
    Sn = bk.spline(bk.interpolate(xs, ys, bk.BSplineOrder(k), bk.Natural()))
    assert np.allclose(Sn(xs), ys, rtol=1e-8, atol=1e-8)
    ts = bk.knots(Sn)
    tu = ts[k - 1:ts.size - k + 1]
    assert np.unique(tu).size == tu.size and tu.size == xs.size and np.array_equal(tu, xs)
    a, b = bk.boundaries(bk.basis(Sn))
    for n in range(2, k // 2 + 1):
        Sd = bk.Derivative(n) * Sn
        for v in (a, b):
            assert abs(Sd(np.array([v]), algo=bk.EVAL_DEBOOR)[0]) < abs(Sn(np.array([v]), algo=bk.EVAL_DEBOOR)[0]) * 1e-7


def test_concurrent_evaluation_of_one_handle(bk, oracle, synth):
    """Several host threads evaluate the SAME freshly created spline handle at once, each on its own stream: the
    lazy table build is serialised by the handle's lock, every thread gets the same, correct results (the C ABI's
    thread-safety statement in bsk.h)."""
    import threading
    import torch
    k = 4
    B = bk.BSplineBasis(k, synth.breakpoints(3, 700))
    c = 2 * synth.u_numpy(4, 0, len(B)) - 1
    n = 1 << 20
    xs = [synth.u_torch(90 + t, 0, n) for t in range(4)]
    for rnd in range(3):
        S = bk.Spline(B, c)                                   # fresh handle: tables not built yet
        outs = [None] * 4
        errs = []

        def work(t):
            try:
                st = torch.cuda.Stream()
                with torch.cuda.stream(st):
                    v, dv = S.evaluate(xs[t], (0, 1))
                    w = S.evaluate(xs[t], (0,), algo=bk.EVAL_DEBOOR)[0]
                st.synchronize()
                outs[t] = (v, dv, w)
            except Exception as exc:                          # surfaced below
                errs.append(exc)

        th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        assert not errs, errs
        kd, td, cd = oracle.spline_derivative(0, k, B.knots(), c, 1)
        for t in range(4):
            xh = xs[t][:20000].cpu().numpy()
            ref = oracle.spline_eval(0, k, B.knots(), c, xh)
            refd = oracle.spline_eval(0, kd, td, cd, xh)
            v, dv, w = outs[t]
            assert np.array_equal(w[:20000].cpu().numpy(), ref)
            assert relerr(v[:20000].cpu().numpy(), ref) < TOL64 and relerr(dv[:20000].cpu().numpy(), refd) < TOL64


@pytest.mark.parametrize("k", [4, 6, 8])
def test_collocation_boundary_value_problem(bk, k):
    """The collocation work flow the reference's examples are built from (examples/heat.jl, test/airy.jl without the
    Galerkin tensor): u'' - u = f on [-1, 1] with u(-1) = u(1) = 0 in a Dirichlet-recombined basis
    (RecombinedBSplineBasis(B, Derivative(0))): assemble C2 - C0 from collocation_matrix(R, x, Derivative(2)) and
    collocation_matrix(R, x), factor, solve for a BATCH of right-hand sides on the device, evaluate the resulting
    splines and compare with the manufactured solutions u_m(x) = sin(m pi x)."""
    import torch
    N = 160
    breaks = -np.cos(np.pi * np.arange(N + 1) / N)                  # Chebyshev-type, non-uniform
    B = bk.BSplineBasis(k, breaks)
    R = bk.RecombinedBSplineBasis(B, bk.Derivative(0))
    x = bk.collocation_points(R)
    C0 = bk.collocation_matrix(R, x)
    C2 = bk.collocation_matrix(R, x, bk.Derivative(2))
    A = C2.data() - C0.data()                                        # same band layout (Collocation.jl:165-198)
    M = bk.CollocationMatrix(torch.from_numpy(np.ascontiguousarray(A.T)).cuda())
    M.lu_()
    ms = np.arange(1, 9)
    F = np.stack([-(1 + (m * np.pi) ** 2) * np.sin(m * np.pi * x) for m in ms], axis=1)      # u_m'' - u_m
    Fd = torch.from_numpy(np.ascontiguousarray(F)).cuda()
    M.ldiv_(Fd)                                                      # coefficients in the recombined basis, in place
    xs = np.linspace(-1.0, 1.0, 2001)
    for j, m in enumerate(ms):
        S = bk.Spline(R, Fd[:, j].cpu().numpy())
        err = np.max(np.abs(S(xs) - np.sin(m * np.pi * xs)))
        assert err < {4: 1e-3, 6: 2e-5, 8: 1e-6}[k] * m ** (k - 2), (k, m, err)
        assert abs(S(np.array([-1.0]))[0]) < 1e-14 and abs(S(np.array([1.0]))[0]) < 1e-14     # the BCs hold exactly
        d2 = (bk.Derivative(2) * S)(x)
        assert np.max(np.abs(d2 - S(x) - F[:, j])) < 1e-7 * np.max(np.abs(F[:, j]))            # collocation equations


def test_heat_equation_time_stepping_with_the_fused_step(bk):
    """examples/heat.jl:300-330 in small: u_t = nu u_xx on [-1, 1], u(+-1) = 0, collocation in a
    Dirichlet-recombined cubic basis; every right-hand-side evaluation is `du = F \\ (L u)` — the fused
    mat-vec + solve on a batch of 64 initial conditions (modes sin(m pi (x + 1) / 2)) — inside a classical RK4
    loop that never leaves the device.  Each mode must decay like exp(-nu (m pi / 2)^2 t)."""
    import torch
    k, N, nu = 4, 96, 1.0
    B = bk.BSplineBasis(k, np.linspace(-1.0, 1.0, N + 1))
    R = bk.RecombinedBSplineBasis(B, bk.Derivative(0))
    x = bk.collocation_points(R)
    F = bk.collocation_matrix(R, x)
    L = bk.collocation_matrix(R, x, bk.Derivative(2))
    F.lu_()
    M = 64
    modes = 1 + np.arange(M) % 4                                     # m = 1..4, sixteen copies each
    U0 = np.stack([np.sin(m * np.pi * (x + 1) / 2) for m in modes], axis=1)
    C = torch.from_numpy(np.ascontiguousarray(U0)).cuda()
    F.ldiv_(C)                                                       # coefficients of the initial conditions
    dt, steps = 2.0e-5, 400
    k1, k2, k3, k4, tmp = (torch.empty_like(C) for _ in range(5))
    for _ in range(steps):                                           # du/dt = nu F \\ (L u), classical RK4
        F.ldiv_mul(L, C, out=k1)
        torch.add(C, k1, alpha=0.5 * dt * nu, out=tmp); F.ldiv_mul(L, tmp, out=k2)
        torch.add(C, k2, alpha=0.5 * dt * nu, out=tmp); F.ldiv_mul(L, tmp, out=k3)
        torch.add(C, k3, alpha=dt * nu, out=tmp); F.ldiv_mul(L, tmp, out=k4)
        C.add_(k1 + 2 * k2 + 2 * k3 + k4, alpha=dt * nu / 6)
    T = dt * steps
    assert torch.isfinite(C).all()
    xs = np.linspace(-1.0, 1.0, 501)
    Ch = C.cpu().numpy()
    for j in (0, 1, 2, 3, 63):
        m = modes[j]
        S = bk.Spline(R, Ch[:, j])
        exact = np.exp(-nu * (m * np.pi / 2) ** 2 * T) * np.sin(m * np.pi * (xs + 1) / 2)
        assert np.max(np.abs(S(xs) - exact)) < 2e-4 * m ** 2, (m, np.max(np.abs(S(xs) - exact)))
    assert np.array_equal(Ch[:, 0], Ch[:, 4])                        # identical columns stay identical


@pytest.mark.parametrize("k", [3, 4, 6])
def test_galerkin_matrix_pair_of_bases(bk, oracle, synth, k):
    """galerkin_matrix((B1, B2)) for bases sharing their parent (src/Galerkin/linear.jl:134-199; test/airy.jl:66-69
    `Mf = galerkin_matrix((R, B))`): rectangular band with shifted bandwidths, bit for bit the oracle's loop, for
    (R, B), (B, R), two different recombinations and derivative pairs; consistency with the square matrices
    (R' M_BB R == M_RR through the recombination matrix); a foreign basis is refused."""
    br = synth.breakpoints(64, 40)
    B = bk.BSplineBasis(k, br)
    t = B.knots()
    R0 = bk.RecombinedBSplineBasis(B, bk.Derivative(0))
    R1 = bk.RecombinedBSplineBasis(B, bk.Derivative(1))
    Rm = bk.RecombinedBSplineBasis(B, bk.Derivative(0), bk.Derivative(1))          # different BCs left / right
    rc = lambda R: oracle.Recomb(R.cl, R.cr, R.nl, R.nr, R.ul, R.lr) if isinstance(R, bk.RecombinedBSplineBasis) else None
    for B1, B2 in ((R0, B), (B, R0), (R0, R1), (Rm, B), (R1, Rm)):
        for d1, d2 in ((0, 0), (1, 1), (0, 1)):
            M = bk.galerkin_matrix((B1, B2), (bk.Derivative(d1), bk.Derivative(d2)))
            ref = oracle.galerkin_matrix_pair(k, t, rc(B1), rc(B2), d1, d2)
            assert M.shape == ref.shape == (len(B1), len(B2))
            dl = (B2.cl if rc(B2) else 0) - (B1.cl if rc(B1) else 0)
            assert M.bandwidths == (k - 1 + dl, k - 1 - dl)
            assert np.array_equal(M.dense(), ref), (d1, d2)
    # M_RB == T' M_BB with the recombination matrix T of R (N x M): rows of the square parent matrix recombined
    T = R0.recombination_matrix()
    sq = oracle.galerkin_matrix(k, t)
    assert np.max(np.abs(bk.galerkin_matrix((R0, B)).dense() - T.T @ sq)) < 1e-15
    assert np.max(np.abs(bk.galerkin_matrix((R0, B)) @ np.ones(len(B)) - (T.T @ sq).sum(axis=1))) < 1e-14
    with pytest.raises(bk.ArgumentError):
        bk.galerkin_matrix((B, bk.BSplineBasis(k, br + 0.1)))


def test_airy_equation_galerkin(bk):
    """test/airy.jl:12-95 restated: eps u'' - x u = 0 on [-1, 1] with inhomogeneous Dirichlet data (Olver & Townsend,
    SIAM Rev. 2013), Galerkin method in a Dirichlet-recombined basis of order 6 on 100 intervals.  The reference builds
    the multiplication operator from galerkin_tensor((R, R, B)) contracted with the coefficients of c(x) = x; the same
    matrix <phi_i, x phi_j> is galerkin_matrix(x -> x, R) with k + 1 nodes (the integrand has degree 2k - 1).  Every
    matrix is assembled on the device; the banded system is factorised and solved there."""
    import torch
    from scipy.special import airy
    eps, N, k = 1e-4, 100, 6
    breaks = np.linspace(-1.0, 1.0, N + 1)
    B = bk.BSplineBasis(bk.BSplineOrder(k), breaks)
    R = bk.RecombinedBSplineBasis(B, bk.Derivative(0))
    u_exact = lambda x: airy(np.asarray(x) / eps ** (1.0 / 3.0))[0]
    ua, ub = float(u_exact(-1.0)), float(u_exact(1.0))
    u_0 = lambda x: ((1 - x) * ua + (1 + x) * ub) / 2
    xs = bk.collocation_points(B)
    Cfact = bk.collocation_matrix(B, xs).lu_()
    Mc = bk.galerkin_matrix(lambda x: x, R, quadrature=bk.gausslegendre(k + 1), as_band=True)
    Mcd = Mc.cpu().numpy()
    L11 = bk.galerkin_matrix(R, (bk.Derivative(1), bk.Derivative(1)), as_band=True)
    # "the matrix is practically symmetric" (:40): band store data[bw + i - j, j]
    bw = k - 1
    n = len(R)
    for j in range(0, n, 7):
        for d in range(1, bw + 1):
            if j + d < n:
                assert abs(Mcd[j, bw + d] - Mcd[j + d, bw - d]) < 1e-15
    Lfact = bk.CollocationMatrix((eps * L11 + Mc).contiguous()).lu_()
    f = lambda x: x * u_0(x)
    fs = f(xs).reshape(-1, 1).copy()
    Cfact.ldiv_(fs)                                          # coefficients of the right-hand side in B
    Mf = bk.galerkin_matrix((R, B))
    rhs = -(Mf @ fs[:, 0])
    vR = torch.from_numpy(rhs.reshape(-1, 1).copy()).cuda()
    Lfact.ldiv_(vR)
    vB = R.recombination_matrix() @ vR[:, 0].cpu().numpy()
    vsol = bk.Spline(B, vB)
    xq = np.linspace(-1.0, 1.0, 20001)
    err2 = np.trapezoid((vsol(xq) + u_0(xq) - u_exact(xq)) ** 2, xq)
    assert err2 < 1e-6, err2                                 # (:91-92)
