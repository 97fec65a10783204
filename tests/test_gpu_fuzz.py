"""Randomised differential test of the evaluation and interpolation paths against the oracle (fixed seeds, so the
run is reproducible): random orders, knot vectors with uneven spacing / clusters / repeated interior knots, plain
and periodic bases, Float64 and Float32, every derivative order, points on and around every knot and outside the
domain.  The de Boor path must reproduce the oracle bit for bit; the default path must stay inside the north-star
tolerance (1e-13 / 1e-5 of the output's magnitude)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# BSK_FUZZ_OFFSET=<multiple of 1000> shifts every seed: the same tests over fresh random cases (sweeps beyond the
# committed seeds; a failing seed found that way is added to the fixed list, as seed 23 of the first test was).
_OFF = int(os.environ.get("BSK_FUZZ_OFFSET", "0"))


def _seeds(n):
    return list(range(_OFF, _OFF + n))


@pytest.fixture(scope="module")
def synth():
    from bsplinekit_b200 import synth as s
    return s


def _random_breaks(rng, nb, style):
    if style == "jitter":
        i = np.arange(nb, dtype=np.float64)
        x = (i + 0.45 * (2 * rng.random(nb) - 1)) / (nb - 1)
    elif style == "geometric":
        x = np.cumsum(np.exp(rng.uniform(-4.0, 2.0, nb)))
    elif style == "cluster":
        x = np.sort(rng.random(nb))
        j = rng.integers(1, nb - 2)
        x[j + 1] = x[j] + 1e-9 * (1 + rng.random())
        x = np.sort(x)
    else:                                                    # repeated interior knots (multiplicity 2 or 3)
        x = np.sort(rng.random(nb))
        for _ in range(2):
            j = rng.integers(2, nb - 3)
            x[j + 1] = x[j]
        x = np.sort(x)
    x = (x - x[0]) / (x[-1] - x[0])
    a, L = rng.uniform(-3.0, 3.0), np.exp(rng.uniform(-2.0, 3.0))
    x = a + L * x
    x[0], x[-1] = a, a + L
    return x


@pytest.mark.parametrize("seed", _seeds(60))
def test_random_spline_evaluation(bk, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    k = int(rng.integers(2, 9))
    dt = "f32" if seed % 4 == 3 else "f64"
    npdt = np.float64 if dt == "f64" else np.float32
    tol = 1e-13 if dt == "f64" else 1e-5
    periodic = seed % 3 == 2
    style = ("jitter", "geometric", "cluster", "repeat")[seed % 4] if not periodic else ("jitter", "geometric")[seed % 2]
    nb = int(rng.integers(max(2 * k + 2, 8), 400))
    br = _random_breaks(rng, nb, style).astype(npdt)
    if np.any(np.diff(br) < 0):
        br = np.sort(br)
    if periodic:
        if np.any(np.diff(br) <= 0):
            br = np.unique(br)
        B = bk.PeriodicBSplineBasis(k, br)
        kind, tref = 1, br
    else:
        B = bk.BSplineBasis(k, br)
        kind, tref = 0, B.knots()
    c = (2 * rng.random(len(B)) - 1).astype(npdt)
    S = bk.Spline(B, c)
    a, b = float(br[0]), float(br[-1])
    L = b - a
    xs = np.concatenate([a + L * rng.random(4000), br.astype(np.float64),
                         np.nextafter(br, np.inf).astype(np.float64), np.nextafter(br, -np.inf).astype(np.float64),
                         [a - 0.3 * L, b + 0.3 * L, a - 2.5 * L, b + 2.5 * L]]).astype(npdt)
    # bit-exact path, value
    ref0 = oracle.spline_eval(kind, k, tref, c, xs, kt=dt)
    got0 = S.evaluate(xs, (0,), algo=bk.EVAL_DEBOOR)[0]
    assert np.array_equal(got0, ref0, equal_nan=True), (seed, k, dt, periodic, style)
    # default path, every derivative order (at most 4 outputs per call)
    refs = [ref0]
    for d in range(1, k):
        kd, td, cd = oracle.spline_derivative(kind, k, tref, c, d, kt=dt)
        refs.append(oracle.spline_eval(kind, kd, td, cd, xs, kt=dt))
    for lo in range(0, k, 4):
        ds = tuple(range(lo, min(k, lo + 4)))
        outs = S.evaluate(xs, ds)
        for d, o in zip(ds, outs):
            r = refs[d].astype(np.float64)
            # D^(k-1) S is piecewise constant: a point within rounding of a knot may legitimately take either side
            # only if the two paths disagree on the knot interval — they must not (searchsortedlast semantics)
            scale = np.max(np.abs(r)) or 1.0
            assert np.max(np.abs(o.astype(np.float64) - r)) <= tol * scale, (seed, k, dt, periodic, style, d)


@pytest.mark.parametrize("seed", _seeds(20))
def test_random_interpolation(bk, oracle, seed):
    """interpolate(x, Y, BSplineOrder(k)) on random sites, a batch of data sets: against the oracle's banded
    BANFAC / BANSLV on the same collocation matrix, and S(x_i) == y_i."""
    import torch
    rng = np.random.default_rng(2000 + seed)
    k = int(rng.integers(2, 9))
    n = int(rng.integers(max(2 * k, 8), 600))
    x = np.sort(rng.random(n)) * np.exp(rng.uniform(-1, 2))
    x = x + np.arange(n) * 1e-6                              # strictly increasing
    M = int(rng.integers(1, 70))
    Y = 2 * rng.random((n, M)) - 1
    t = bk.make_knots(x, k)
    assert np.array_equal(t, oracle.make_knots(x, k))
    B = bk.BSplineBasis(k, t, augment=False)
    itp = bk.SplineInterpolation(B, x)
    Yd = torch.from_numpy(Y.copy()).cuda()
    itp.interpolate_(Yd)
    band, _ = oracle.collocation_matrix(0, k, t, x)
    oracle.banded_lu(band)
    ref = oracle.banded_solve(band, Y.copy())
    got = Yd.cpu().numpy()
    assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref))
    S = itp.spline_of(0)
    assert np.max(np.abs(S(x) - Y[:, 0])) < 1e-9 * max(1.0, np.max(np.abs(ref[:, 0])))


@pytest.mark.parametrize("seed", _seeds(20))
def test_random_evaluate_all_and_knot_search(bk, oracle, seed):
    """find_knot_interval and evaluate_all (all derivative orders, Float32 output of Float64 bases, recombined
    bases with random Robin / Dirichlet / Neumann constraints) bit for bit against the oracle."""
    rng = np.random.default_rng(3000 + seed)
    k = int(rng.integers(2, 9))
    style = ("jitter", "geometric", "cluster", "repeat")[seed % 4]
    nb = int(rng.integers(max(2 * k + 2, 8), 300))
    br = _random_breaks(rng, nb, style)
    B = bk.BSplineBasis(k, br)
    t = B.knots()
    a, b = float(br[0]), float(br[-1])
    x = np.concatenate([a + (b - a) * rng.random(3000), br, np.nextafter(br, np.inf), np.nextafter(br, -np.inf),
                        [a - 1.0, b + 1.0, np.nan]])
    i_ref, z_ref = oracle.find_knot_interval(0, k, t, x)
    i, z = bk.find_knot_interval(B, x)
    assert np.array_equal(i, i_ref) and np.array_equal(z, z_ref)
    for nd in range(0, k):
        ir, bsr = oracle.evaluate_all(0, k, t, x, nderiv=nd)
        ig, bsg = bk.evaluate_all(B, x, bk.Derivative(nd))
        assert np.array_equal(ig, ir) and np.array_equal(bsg, bsr, equal_nan=True), (seed, k, nd)
    ir, bsr = oracle.evaluate_all(0, k, t, x, nderiv=min(1, k - 1), vt="f32")
    ig, bsg = bk.evaluate_all(B, x, bk.Derivative(min(1, k - 1)), np.float32)
    assert bsg.dtype == np.float32 and np.array_equal(bsg, bsr, equal_nan=True)
    if k >= 3 and style in ("jitter", "geometric"):
        ops = [bk.Derivative(0), bk.Derivative(1), bk.Derivative(0) + float(rng.uniform(0.5, 4.0)) * bk.Derivative(1)]
        op = ops[seed % 3]
        R = bk.RecombinedBSplineBasis(B, op)
        rc = oracle.Recomb(R.cl, R.cr, R.nl, R.nr, R.ul, R.lr)
        xin = x[:-3]
        for nd in (0, 1):
            ir, bsr = oracle.evaluate_all(2, k, t, xin, nderiv=nd, recomb=rc)
            ig, bsg = bk.evaluate_all(R, xin, bk.Derivative(nd))
            assert np.array_equal(ig, ir) and np.array_equal(bsg, bsr, equal_nan=True), (seed, k, nd, "recombined")


@pytest.mark.parametrize("seed", _seeds(12))
def test_random_float32_interpolation_bitexact(bk, oracle, seed):
    """The Float32 chain (assembly, BANFAC, BANSLV) on random sites and orders: bit for bit the Float32 restatement."""
    import torch
    rng = np.random.default_rng(4000 + seed)
    k = int(rng.integers(2, 9))
    n = int(rng.integers(max(2 * k, 8), 500))
    x = (np.sort(rng.random(n)) + np.arange(n) * 1e-3).astype(np.float32)
    t = bk.make_knots(x, k)
    assert t.dtype == np.float32
    B = bk.BSplineBasis(k, t, augment=False)
    Cm = bk.collocation_matrix(B, x)
    ref, nout = oracle.collocation_matrix_f32(0, k, t, x)
    assert np.array_equal(Cm.data(), ref)
    lu_ref = ref.copy(order="F")
    assert oracle.banded_lu(lu_ref) == 0
    Cm.lu_()
    assert np.array_equal(Cm.data(), lu_ref)
    M = int(rng.integers(1, 100))
    Y = (2 * rng.random((n, M)) - 1).astype(np.float32)
    Yd = torch.from_numpy(Y.copy()).cuda()
    Cm.ldiv_(Yd, algo=bk.SOLVE_TWOPASS)
    assert np.array_equal(Yd.cpu().numpy(), oracle.banded_solve(lu_ref, Y.copy()))


@pytest.mark.parametrize("seed", _seeds(12))
def test_random_periodic_interpolation(bk, oracle, seed):
    """Periodic interpolation of every order (cyclic tridiagonal for k = 4, Woodbury-corrected band otherwise) on
    random sites: against a dense pivoted solve of the oracle's periodic collocation matrix; S(x_i) == y_i and
    S(x + L) == S(x)."""
    rng = np.random.default_rng(5000 + seed)
    k = int(rng.integers(2, 8))
    n = int(rng.integers(4 * k + 2, 300))
    L = float(np.exp(rng.uniform(-1, 2)))
    x = -0.3 * L + L * (np.arange(n) + 0.45 * (2 * rng.random(n) - 1)) / n
    y = 2 * rng.random(n) - 1
    itp = bk.interpolate(x, y, bk.BSplineOrder(k), bk.Periodic(L))
    S = itp.spline
    data = oracle.make_knots_periodic(x, k, L)
    xc = x if k != 2 else data[:-1]
    A = oracle.periodic_collocation_dense(k, data, xc)
    cref = np.linalg.solve(A, y)
    cond = np.linalg.cond(A)
    assert np.max(np.abs(S.coefficients() - cref)) <= 1e-13 * cond * np.max(np.abs(cref))
    assert np.max(np.abs(S(x) - y)) < 1e-12 * cond * max(1.0, np.max(np.abs(cref)))
    xq = x[0] + L * rng.random(500)
    assert np.max(np.abs(S(xq + L) - S(xq))) < 1e-11 * max(1.0, np.max(np.abs(cref)))
    assert np.max(np.abs(S(xq, algo=bk.EVAL_DEBOOR) - oracle.spline_eval(1, k, data, S.coefficients(), xq))) == 0


@pytest.mark.parametrize("seed", _seeds(16))
def test_random_vector_valued_and_extrapolated_evaluation(bk, oracle, seed):
    """Vector-valued splines (M splines sharing a basis, the k_spline_multi kernels incl. the TMA ring variant) and
    the three extrapolation methods on random bases / point sets, against the oracle."""
    import torch
    rng = np.random.default_rng(6000 + seed)
    k = int(rng.integers(2, 9))
    periodic = seed % 4 == 3
    nb = int(rng.integers(max(2 * k + 2, 8), 500))
    br = _random_breaks(rng, nb, ("jitter", "geometric")[seed % 2])
    a, b = float(br[0]), float(br[-1])
    L = b - a
    if periodic:
        if np.any(np.diff(br) <= 0):
            br = np.unique(br)
        B = bk.PeriodicBSplineBasis(k, br)
        kind, tref = 1, br
    else:
        B = bk.BSplineBasis(k, br)
        kind, tref = 0, B.knots()
    M = int(rng.choice([1, 2, 3, 4, 7, 32, 130]))
    Cf = 2 * rng.random((len(B), M)) - 1
    n = int(rng.integers(1, 3000))
    x = a - 0.2 * L + 1.4 * L * rng.random(n)
    if seed % 5 == 0:
        x = np.sort(x)
    SB = bk.SplineBatch(B, torch.from_numpy(Cf).cuda())
    got = SB(torch.from_numpy(x).cuda()).cpu().numpy()
    for j in sorted(set([0, M - 1, M // 2])):
        ref = oracle.spline_eval(kind, k, tref, Cf[:, j], x)
        assert np.max(np.abs(got[:, j] - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref))), (seed, k, M, j)
    # the same batch at every breakpoint, one ulp either side, NaN and (plain bases) +-Inf
    xs = np.concatenate([br, np.nextafter(br, np.inf), np.nextafter(br, -np.inf), [np.nan],
                         [] if periodic else [np.inf, -np.inf]])
    gots = SB(torch.from_numpy(xs).cuda()).cpu().numpy()
    for j in sorted(set([0, M - 1])):
        ref = oracle.spline_eval(kind, k, tref, Cf[:, j], xs)
        assert np.array_equal(np.isnan(gots[:, j]), np.isnan(ref)), (seed, k, M, j)
        ok = ~np.isnan(ref)
        assert np.max(np.abs(gots[ok, j] - ref[ok])) <= 1e-13 * max(1.0, np.max(np.abs(ref[ok]))), (seed, k, M, j)
    if not periodic:
        S = bk.Spline(B, Cf[:, 0])
        for name, meth in (("flat", bk.Flat()), ("linear", bk.Linear()), ("smooth", bk.Smooth())):
            ref = oracle.spline_extrapolate(name, k, tref, Cf[:, 0], x)
            out = bk.extrapolate(S, meth)(x)
            tol = 1e-13 if name != "smooth" else 1e-11          # smooth: a polynomial continued up to 0.2 L outside
            assert np.max(np.abs(out - ref)) <= tol * max(1.0, np.max(np.abs(ref))), (seed, k, name)
            if name != "smooth":
                far = np.array([np.inf, -np.inf, np.nan, 1e300, -1e300, b + 1e-300, a - 1e-300])
                with np.errstate(all="ignore"):
                    rf = oracle.spline_extrapolate(name, k, tref, Cf[:, 0], far)
                gf = bk.extrapolate(S, meth)(far)
                fin = np.isfinite(rf)
                assert np.array_equal(gf[~fin], rf[~fin], equal_nan=True), (seed, k, name, gf, rf)
                assert np.all(np.abs(gf[fin] - rf[fin]) <= 1e-13 * np.maximum(1.0, np.abs(rf[fin]))), (seed, k, name, gf, rf)


@pytest.mark.parametrize("seed", _seeds(12))
def test_random_collocation_matrices(bk, oracle, seed):
    """collocation_matrix(B | R, x, Derivative(n)) on random bases, derivative orders and Robin / Dirichlet /
    Neumann recombinations: the band store bit for bit the oracle's; rectangular row form consistent with it."""
    rng = np.random.default_rng(7000 + seed)
    k = int(rng.integers(3, 9))
    nb = int(rng.integers(max(2 * k + 2, 10), 300))
    br = _random_breaks(rng, nb, ("jitter", "geometric")[seed % 2])
    B = bk.BSplineBasis(k, br)
    t = B.knots()
    nd = int(rng.integers(0, min(k - 1, 3) + 1))
    if seed % 3 == 0:
        basis, rc, kindc = B, None, 0
    else:
        ops = [bk.Derivative(0), bk.Derivative(1), bk.Derivative(0) + float(rng.uniform(0.3, 5.0)) * bk.Derivative(1)]
        basis = bk.RecombinedBSplineBasis(B, ops[seed % 3])
        rc, kindc = oracle.Recomb(basis.cl, basis.cr, basis.nl, basis.nr, basis.ul, basis.lr), 2
    x = bk.collocation_points(basis)
    ref, nout_ref = oracle.collocation_matrix(kindc, k, t, x, nderiv=nd, recomb=rc)
    Cm, nout = bk.collocation_matrix(basis, x, bk.Derivative(nd), return_outside=True)
    assert nout == nout_ref and np.array_equal(Cm.data(), ref), (seed, k, nd)
    # rectangular form at other points: rows agree with evaluate_all
    xr = np.sort(float(br[0]) + (float(br[-1]) - float(br[0])) * rng.random(2 * len(basis) + 7))
    D = bk.collocation_rows(basis, xr, bk.Derivative(nd), dense=True).cpu().numpy()
    idx, bs = oracle.evaluate_all(kindc, k, t, xr, nderiv=nd, recomb=rc)
    eps = np.finfo(float).eps
    for r in rng.integers(0, xr.size, 25):
        row = np.zeros(len(basis))
        for d in range(k):
            j = idx[r] - d
            if 1 <= j <= len(basis) and abs(bs[r, d]) > eps:
                row[j - 1] = bs[r, d]
        assert np.array_equal(D[r], row), (seed, r)


@pytest.mark.parametrize("seed", _seeds(16))
def test_random_large_bases_global_tables(bk, oracle, seed):
    """Bases with 2 000 .. 40 000 intervals: the tables stay in global memory (one or two sectors per record, 256-bit
    gathers), plain and periodic, Float64 and Float32, value and derivatives in one pass, points at and next to knots."""
    rng = np.random.default_rng(8000 + seed)
    k = int(rng.integers(2, 9))
    dt = "f32" if seed % 4 == 1 else "f64"
    npdt = np.float64 if dt == "f64" else np.float32
    tol = 1e-13 if dt == "f64" else 1e-5
    periodic = seed % 3 == 1
    nb = int(rng.integers(2000, 40000))
    br = _random_breaks(rng, nb, ("jitter", "geometric")[seed % 2]).astype(npdt)
    br = np.unique(br)
    if periodic:
        B = bk.PeriodicBSplineBasis(k, br)
        kind, tref = 1, br
    else:
        B = bk.BSplineBasis(k, br)
        kind, tref = 0, B.knots()
    c = (2 * rng.random(len(B)) - 1).astype(npdt)
    S = bk.Spline(B, c)
    a, b = float(br[0]), float(br[-1])
    sub = br[rng.integers(0, br.size, 2000)].astype(np.float64)
    xs = np.concatenate([a + (b - a) * rng.random(30000), sub, np.nextafter(sub, np.inf), np.nextafter(sub, -np.inf),
                         [a, b, a - 0.1 * (b - a), b + 0.1 * (b - a)]]).astype(npdt)
    ds = tuple(range(0, min(k, 3)))
    outs = S.evaluate(xs, ds)
    for d, o in zip(ds, outs):
        kd, td, cd = oracle.spline_derivative(kind, k, tref, c, d, kt=dt) if d else (k, tref, c)
        ref = oracle.spline_eval(kind, kd, td, cd, xs, kt=dt).astype(np.float64)
        scale = np.max(np.abs(ref)) or 1.0
        assert np.max(np.abs(o.astype(np.float64) - ref)) <= tol * scale, (seed, k, dt, periodic, d)
    if seed % 2 == 0:
        # jittered breakpoints always get a uniform-cell table; geometric ones may be too uneven for the size caps
        # (Float32, tens of thousands of intervals) and then take the interval-table kernels, checked above as well
        assert S.table_info()["cells"] > 0


@pytest.mark.parametrize("seed", _seeds(40))
def test_random_extreme_scalings(bk, oracle, seed):
    """Domains far from the unit interval (offsets up to 1e6 periods of the knot spacing, lengths 1e-6 .. 1e6),
    tiny bases (k + 1 .. 2 k breakpoints), coefficient magnitudes 1e-20 .. 1e20: value and all derivatives of the
    default path inside the tolerance, de Boor path bit for bit."""
    rng = np.random.default_rng(9000 + seed)
    k = int(rng.integers(1, 9))
    dt = "f32" if seed % 3 == 2 else "f64"
    npdt = np.float64 if dt == "f64" else np.float32
    tol = 1e-13 if dt == "f64" else 1e-5
    nb = int(rng.integers(2, 3 * k + 3)) if seed % 2 == 0 else int(rng.integers(20, 3000))
    L = 10.0 ** rng.uniform(-6, 6) if dt == "f64" else 10.0 ** rng.uniform(-3, 3)
    a = L * (10.0 ** rng.uniform(-1, 3 if dt == "f64" else 1)) * rng.choice([-1.0, 1.0]) if seed % 4 else 0.0
    i = np.arange(nb, dtype=np.float64)
    br = a + L * (i + 0.4 * (2 * rng.random(nb) - 1)) / max(nb - 1, 1)
    br[0], br[-1] = a, a + L
    br = np.unique(br.astype(npdt))
    if br.size < 2:
        return
    B = bk.BSplineBasis(k, br)
    t = B.knots()
    amp = 10.0 ** rng.uniform(-20, 20) if dt == "f64" else 10.0 ** rng.uniform(-8, 8)
    c = (amp * (2 * rng.random(len(B)) - 1)).astype(npdt)
    S = bk.Spline(B, c)
    lo, hi = float(br[0]), float(br[-1])
    xs = np.concatenate([lo + (hi - lo) * rng.random(3000), br.astype(np.float64), np.nextafter(br, np.inf).astype(np.float64),
                         np.nextafter(br, -np.inf).astype(np.float64), [lo - 0.1 * (hi - lo), hi + 0.1 * (hi - lo)]]).astype(npdt)
    ref0 = oracle.spline_eval(0, k, t, c, xs, kt=dt)
    got0 = S.evaluate(xs, (0,), algo=bk.EVAL_DEBOOR)[0]
    assert np.array_equal(got0, ref0, equal_nan=True), (seed, k, dt)
    for lo_d in range(0, k, 4):
        ds = tuple(range(lo_d, min(k, lo_d + 4)))
        outs = S.evaluate(xs, ds)
        for d, o in zip(ds, outs):
            kd, td, cd = oracle.spline_derivative(0, k, t, c, d, kt=dt) if d else (k, t, c)
            r = oracle.spline_eval(0, kd, td, cd, xs, kt=dt).astype(np.float64)
            if not np.all(np.isfinite(r)):
                assert np.array_equal(np.isfinite(o), np.isfinite(r)), (seed, k, dt, d)     # overflow where the reference overflows
                continue
            scale = np.max(np.abs(r)) or 1.0
            assert np.max(np.abs(o.astype(np.float64) - r)) <= tol * scale, (seed, k, dt, d, L, a, amp)


@pytest.mark.parametrize("seed", _seeds(12))
def test_random_natural_interpolation_smoothing_and_fused_step(bk, oracle, seed):
    """Natural interpolation (orders 4, 6, 8), cubic smoothing fit (against the oracle's dense restatement, scaled by
    the conditioning of the normal matrix) and the fused mat-vec + solve against mul! followed by ldiv!, on random
    sites, weights and batch widths."""
    import torch
    rng = np.random.default_rng(10000 + seed)
    n = int(rng.integers(12, 400))
    x = np.sort(rng.random(n)) + np.arange(n) * 2e-3
    y = 2 * rng.random(n) - 1
    k = int(rng.choice([4, 6, 8]))
    if n >= 2 * k:
        Sn = bk.spline(bk.interpolate(x, y, bk.BSplineOrder(k), bk.Natural()))
        assert np.max(np.abs(Sn(x) - y)) < 1e-7 * max(1.0, np.max(np.abs(Sn.coefficients())))
        a, b = bk.boundaries(bk.basis(Sn))
        sc = np.max(np.abs(Sn.coefficients()))
        for d in range(2, k // 2 + 1):
            Sd = bk.Derivative(d) * Sn
            vals = Sd(np.array([a, b]), algo=bk.EVAL_DEBOOR)
            assert np.max(np.abs(vals)) < 1e-6 * max(np.max(np.abs(Sd.coefficients())), sc), (seed, k, d)
    # smoothing
    lam = 10.0 ** rng.uniform(-8, -2)
    w = 0.5 + rng.random(n) if seed % 2 else None
    S = bk.fit(bk.BSplineOrder(4), x, y, lam, weights=w)
    cs_ref, t, rc = oracle.fit_smoothing(x, y, lam, weights=w)
    xq = np.linspace(x[0], x[-1], 333)
    idx, bs = oracle.evaluate_all(2, 4, t, xq, recomb=rc)
    ref = np.array([sum(bs[p, d] * cs_ref[idx[p] - d - 1] for d in range(4) if 0 <= idx[p] - d - 1 < cs_ref.size)
                    for p in range(xq.size)])
    assert np.max(np.abs(S(xq) - ref)) < 1e-6 * max(1.0, np.max(np.abs(ref)))      # cond(M) up to ~1e9 on clustered sites
    # fused step against the two-call form
    kk = int(rng.choice([4, 5, 6]))
    t2 = bk.make_knots(x, kk)
    B2 = bk.BSplineBasis(kk, t2, augment=False)
    F = bk.collocation_matrix(B2, x)
    Lm = bk.collocation_matrix(B2, x, bk.Derivative(int(rng.integers(0, 3))))
    Fl = bk.collocation_matrix(B2, x).lu_()
    M = int(rng.choice([1, 5, 32, 64, 200]))
    U = torch.from_numpy(2 * rng.random((n, M)) - 1).cuda()
    out = Fl.ldiv_mul(Lm, U)
    two = torch.empty_like(U)
    import ctypes
    from bsplinekit_b200 import _lib as Lb
    Lb.check(Lb.lib.bsk_banded_matvec(Lm._handle, ctypes.c_void_p(U.data_ptr()), ctypes.c_void_p(two.data_ptr()), M, None))
    Fl.ldiv_(two)
    assert torch.max(torch.abs(out - two)).item() <= 1e-12 * max(1.0, torch.max(torch.abs(two)).item()), (seed, kk, M)


@pytest.mark.parametrize("seed", _seeds(14))
def test_random_solve_shapes_single_pass_vs_banslv(bk, oracle, seed):
    """The single-pass solve (halo-restarted backward blocks, M = ring - H, row segments when there are few
    columns) against the BANSLV-order two-pass kernels on the same factors: n = 50 .. 20 000 rows, 1 .. 4096
    columns, orders 2 .. 8, Float64 and Float32; and the cyclic tridiagonal solve against the oracle."""
    import torch
    rng = np.random.default_rng(11000 + seed)
    f32 = seed % 4 == 3
    npdt, tdt = (np.float32, torch.float32) if f32 else (np.float64, torch.float64)
    k = int(rng.integers(2, 9))
    n = int(10 ** rng.uniform(1.7, 4.3))
    M = int(rng.choice([1, 2, 31, 32, 33, 64, 96, 500, 1024, 4096]))
    if n * M > 3e7:
        M = max(1, int(3e7 // n))
    i = np.arange(n, dtype=np.float64)
    x = ((i + 0.45 * (2 * rng.random(n) - 1)) / n).astype(npdt)
    x = np.unique(x)
    n = x.size
    t = bk.make_knots(x, k)
    B = bk.BSplineBasis(k, t, augment=False)
    itp = bk.SplineInterpolation(B, x)
    Y = torch.from_numpy((2 * rng.random((n, M)) - 1).astype(npdt)).cuda()
    Y1, Y2 = Y.clone(), Y.clone()
    itp.C.ldiv_(Y1)                                        # AUTO: single pass when a halo was validated
    itp.C.ldiv_(Y2, algo=bk.SOLVE_TWOPASS)
    sc = torch.max(torch.abs(Y2)).item()
    err = torch.max(torch.abs(Y1 - Y2)).item()
    assert err <= (1e-12 if not f32 else 2e-5) * max(1.0, sc), (seed, k, n, M, f32, err, itp.C.halo)
    # residual of the default solve against the unfactorised matrix (dense rows on the device)
    if n <= 3000 and not f32:                              # (the rectangular row form is Float64-only)
        A = bk.collocation_rows(B, x, dense=True).to(torch.float64)
        r = A @ Y1[:, :min(M, 8)].to(torch.float64) - Y[:, :min(M, 8)].to(torch.float64)
        assert torch.max(torch.abs(r)).item() <= (1e-11 if not f32 else 1e-3) * max(1.0, sc)
    if not f32:
        nc = int(rng.integers(5, 3000))
        a_ = 0.2 + rng.random(nc); c_ = 0.2 + rng.random(nc); b_ = 2.5 + rng.random(nc)      # diagonally dominant
        Mc = int(rng.choice([1, 7, 32, 100, 1000]))
        D = 2 * rng.random((nc, Mc)) - 1
        Cy = bk.CyclicTridiagonalMatrix(a_, b_, c_)
        Dd = torch.from_numpy(D.copy()).cuda()
        Cy.ldiv_(Dd)
        ref = oracle.cyclic_solve(a_, b_, c_, D)
        assert np.max(np.abs(Dd.cpu().numpy() - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref))), (seed, nc, Mc)


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_host_path_chunking_and_staging(bk, synth, dt, monkeypatch):
    """bsk_spline_eval_host on pageable arrays (page-locked staging, copy threads, 2^20-point chunks) and with the
    staging disabled, at sizes around every chunking threshold, from views that are not 16-byte aligned, 1 .. 4
    outputs: bit for bit the results of the device-resident call."""
    import torch
    npdt = np.float64 if dt == "f64" else np.float32
    tdt = torch.float64 if dt == "f64" else torch.float32
    k = 5
    B = bk.BSplineBasis(k, synth.breakpoints(3, 300).astype(npdt))
    S = bk.Spline(B, (2 * synth.u_numpy(4, 0, len(B)) - 1).astype(npdt))
    nmax = 3 * (1 << 20) + 77
    big = (1.2 * synth.u_numpy(50, 0, nmax + 3) - 0.1).astype(npdt)
    bigd = torch.from_numpy(big).cuda()
    refs = {}
    for n, off, ds in ((1, 0, (0,)), (2, 1, (0, 1)), ((1 << 14) - 1, 1, (0,)), (1 << 14, 0, (0, 2)), ((1 << 16) + 3, 1, (1,)),
                       ((1 << 20) + 1, 1, (0, 1, 2, 3)), (3 * (1 << 20) + 5, 0, (0, 1)), (nmax, 3, (0,))):
        x = big[off:off + n]
        want = [o.cpu().numpy() for o in S.evaluate(bigd[off:off + n].clone(), ds)]
        for staging in (True, False):
            if staging:
                monkeypatch.delenv("BSK_HOST_NO_STAGING", raising=False)
            else:
                monkeypatch.setenv("BSK_HOST_NO_STAGING", "1")
            outs = [np.full(n + 2, -7.0, dtype=npdt)[1:-1] for _ in ds]     # unaligned views with guard elements
            bases = [o.base for o in outs]
            got = S.evaluate(x, ds, out=outs)
            for g, w, bb in zip(got, want, bases):
                assert np.array_equal(g, w, equal_nan=True), (n, off, ds, staging)
                assert bb[0] == -7.0 and bb[-1] == -7.0                     # nothing written outside


@pytest.mark.parametrize("seed", _seeds(10))
def test_random_galerkin(bk, oracle, seed):
    """Galerkin matrices / projections on random plain, recombined and periodic bases, derivative pairs and weight
    functions: bit for bit the oracle's sequential loops."""
    rng = np.random.default_rng(12000 + seed)
    k = int(rng.integers(2, 8))
    nb = int(rng.integers(2 * k + 2, 120))
    br = _random_breaks(rng, nb, ("jitter", "geometric")[seed % 2])
    d1 = int(rng.integers(0, min(k, 3)))
    d2 = int(rng.integers(0, min(k, 3)))
    f = (lambda v: 1.0 + 0.3 * np.cos(v)) if seed % 2 else None
    fs = (lambda v: float(1.0 + 0.3 * np.cos(v))) if f else None

    def dense(band, N, cyclic):
        w = band.cpu().numpy().T
        bw = (w.shape[0] - 1) // 2
        A = np.zeros((N, N))
        for j in range(N):
            for d in range(-bw, bw + 1):
                i = j + d
                if cyclic:
                    A[i % N, j] = w[bw + d, j]
                elif 0 <= i < N:
                    A[i, j] = w[bw + d, j]
        return A

    mode = seed % 3
    if mode == 0:
        B = bk.BSplineBasis(k, br)
        args = ((f,) if f else ()) + (B, (bk.Derivative(d1), bk.Derivative(d2)))
        got = dense(bk.galerkin_matrix(*args, as_band=True), len(B), False)
        assert np.array_equal(got, oracle.galerkin_matrix(k, B.knots(), d1, d2, f=fs)), (seed, k, d1, d2)
        g = lambda v: np.exp(-v * v)
        assert np.array_equal(bk.galerkin_projection(g, B), oracle.galerkin_projection(lambda v: float(np.exp(-v * v)), k, B.knots()))
    elif mode == 1 and k >= 3:
        B = bk.BSplineBasis(k, br)
        R = bk.RecombinedBSplineBasis(B, bk.Derivative(int(rng.integers(0, 2))))
        rc = oracle.Recomb(R.cl, R.cr, R.nl, R.nr, R.ul, R.lr)
        got = dense(bk.galerkin_matrix(R, (bk.Derivative(d1), bk.Derivative(d2)), as_band=True), len(R), False)
        assert np.array_equal(got, oracle.galerkin_matrix(k, B.knots(), d1, d2, recomb=rc)), (seed, k, d1, d2)
    else:
        brp = np.unique(br)
        if brp.size - 1 < 2 * k - 1:
            return
        P = bk.PeriodicBSplineBasis(k, brp)
        got = dense(bk.galerkin_matrix(P, (bk.Derivative(d1), bk.Derivative(d2)), as_band=True), len(P), True)
        assert np.array_equal(got, oracle.galerkin_matrix_periodic(k, brp, d1, d2)), (seed, k, d1, d2)
        L = float(brp[-1] - brp[0])
        g = lambda v: np.cos(2 * np.pi * v / L)
        assert np.array_equal(bk.galerkin_projection(g, P),
                              oracle.galerkin_projection_periodic(lambda v: float(np.cos(2 * np.pi * v / L)), k, brp))


@pytest.mark.parametrize("seed", _seeds(24))
def test_random_extreme_magnitudes_and_special_points(bk, oracle, seed):
    """Coefficients scaled to the ends of the exponent range (results and derivatives that overflow or fall into
    the subnormals) and NaN / +-Inf / -0.0 / subnormal points: the de Boor path stays bit-identical to the oracle,
    the default path within the tolerance wherever the oracle's result is finite and non-finite wherever it is not."""
    rng = np.random.default_rng(11000 + seed)
    k = int(rng.integers(2, 9))
    f32 = seed % 3 == 2
    dt, npdt, tol = ("f32", np.float32, 1e-5) if f32 else ("f64", np.float64, 1e-13)
    expo = ((-36, -20, 0, 20, 36) if f32 else (-300, -150, 0, 150, 300))[seed % 5]
    nb = int(rng.integers(max(2 * k + 2, 8), 200))
    br = _random_breaks(rng, nb, ("jitter", "geometric")[seed % 2]).astype(npdt)
    br = np.unique(br)
    periodic = seed % 4 == 1
    if periodic:
        B = bk.PeriodicBSplineBasis(k, br)
        kind, tref = 1, br
    else:
        B = bk.BSplineBasis(k, br)
        kind, tref = 0, B.knots()
    with np.errstate(over="ignore", under="ignore"):
        c = ((2 * rng.random(len(B)) - 1) * 10.0 ** expo).astype(npdt)
    S = bk.Spline(B, c)
    a, b = float(br[0]), float(br[-1])
    tiny = np.finfo(npdt).tiny
    # periodic bases: the reference reduces x by repeated +-L (src/BSplines/periodic.jl:85-100), which never ends
    # for +-Inf or for |x| so large that x - L == x, so those points have no reference behaviour to compare with
    far = [] if periodic else [np.inf, -np.inf, float(np.finfo(npdt).max), -float(np.finfo(npdt).max)]
    xs = np.concatenate([a + (b - a) * rng.random(2000), br.astype(np.float64),
                         [np.nan, -0.0, 0.0, tiny / 4, -tiny / 4, a, b], far]).astype(npdt)
    with np.errstate(all="ignore"):
        ref0 = oracle.spline_eval(kind, k, tref, c, xs, kt=dt)
        got0 = S.evaluate(xs, (0,), algo=bk.EVAL_DEBOOR)[0]
        assert np.array_equal(got0, ref0, equal_nan=True), (seed, k, dt, periodic, expo)
        for d in range(0, k):
            if d == 0:
                r = ref0
            else:
                kd, td, cd = oracle.spline_derivative(kind, k, tref, c, d, kt=dt)
                r = oracle.spline_eval(kind, kd, td, cd, xs, kt=dt)
            g = S.evaluate(xs, (d,), algo=bk.EVAL_DEBOOR)[0]
            assert np.array_equal(g, r, equal_nan=True), (seed, k, dt, periodic, expo, d, "de Boor path")
            o = S.evaluate(xs, (d,))[0].astype(np.float64)
            r = r.astype(np.float64)
            # (D^(k-1) S is piecewise constant: the reference touches x only to pick the piece, so a NaN point comes
            #  back as the constant of the piece searchsortedlast assigns to NaN — on both paths)
            nanx = np.isnan(xs)
            assert np.array_equal(o[nanx], r[nanx], equal_nan=True), (seed, k, dt, periodic, expo, d, "NaN point")
            fin = np.isfinite(r)
            assert not np.any(np.isfinite(o[~fin])), (seed, k, dt, periodic, expo, d, "finite where the reference is not")
            if np.any(fin):
                scale = max(np.max(np.abs(r[fin])), float(tiny))
                assert np.all(np.isfinite(o[fin])), (seed, k, dt, periodic, expo, d, "non-finite where the reference is finite")
                assert np.max(np.abs(o[fin] - r[fin])) <= tol * scale, (seed, k, dt, periodic, expo, d)


@pytest.mark.parametrize("seed", _seeds(16))
def test_random_periodic_evaluate_all_and_knot_search(bk, oracle, seed):
    """Periodic bases: find_knot_interval and evaluate_all bit for bit against the oracle for points in the main
    period, on and beside every breakpoint, several periods away on both sides, and NaN (which passes both wrap
    loops of src/BSplines/periodic.jl:85-100 and lands where searchsortedlast puts it)."""
    rng = np.random.default_rng(14000 + seed)
    k = int(rng.integers(2, 9))
    f32 = seed % 4 == 3
    dt, npdt = ("f32", np.float32) if f32 else ("f64", np.float64)
    nb = int(rng.integers(max(2 * k + 2, 8), 300))
    br = np.unique(_random_breaks(rng, nb, ("jitter", "geometric")[seed % 2]).astype(npdt))
    B = bk.PeriodicBSplineBasis(k, br)
    a, b = float(br[0]), float(br[-1])
    L = b - a
    x = np.concatenate([a + L * rng.random(3000), br.astype(np.float64),
                        np.nextafter(br, npdt(np.inf)).astype(np.float64), np.nextafter(br, npdt(-np.inf)).astype(np.float64),
                        a + L * rng.uniform(-6.0, 7.0, 500), [a - 3 * L, b + 3 * L, np.nan]]).astype(npdt)
    i_ref, z_ref = oracle.find_knot_interval(1, k, br, x, kt=dt)
    i, z = bk.find_knot_interval(B, x)
    assert np.array_equal(i, i_ref) and np.array_equal(z, z_ref), (seed, k, dt)
    for nd in range(0, k):
        ir, bsr = oracle.evaluate_all(1, k, br, x, nderiv=nd, kt=dt, vt=dt)
        ig, bsg = bk.evaluate_all(B, x, bk.Derivative(nd))
        assert np.array_equal(ig, ir) and np.array_equal(bsg, bsr, equal_nan=True), (seed, k, dt, nd)


@pytest.mark.parametrize("seed", _seeds(12))
def test_random_solve_is_exactly_homogeneous(bk, seed):
    """Size-independent property of the banded solves: scaling the data by a power of two scales every coefficient
    by exactly that power (no step of LU substitution rounds differently), for every solve algorithm, wide batches
    (the windowed single-pass kernels) and narrow ones, up to data near the ends of the exponent range."""
    import torch
    rng = np.random.default_rng(13000 + seed)
    k = int(rng.integers(2, 9))
    n = int(rng.integers(max(2 * k, 8), 3000))
    M = (7, 64, 257, 1024)[seed % 4]
    f32 = seed % 3 == 2
    tdt = torch.float32 if f32 else torch.float64
    x = (np.sort(rng.random(n)) + np.arange(n) * 1e-3).astype(np.float32 if f32 else np.float64)
    B = bk.BSplineBasis(k, bk.make_knots(x, k), augment=False)
    Cm = bk.collocation_matrix(B, x)
    Cm.lu_()
    Y = torch.from_numpy(2 * rng.random((n, M)) - 1).to(tdt).cuda()
    for e in ((-100, 90) if f32 else (-900, 900)):
        for algo in (bk.SOLVE_AUTO, bk.SOLVE_WINDOWED, bk.SOLVE_TWOPASS):
            A = Y.clone()
            try:
                Cm.ldiv_(A, algo=algo)
            except (bk.UnsupportedError, bk.ArgumentError):
                assert algo == bk.SOLVE_WINDOWED           # no validated halo for this matrix
                continue
            Bs = torch.ldexp(Y, torch.tensor(e, device="cuda")).to(tdt)
            Cm.ldiv_(Bs, algo=algo)
            back = torch.ldexp(Bs, torch.tensor(-e, device="cuda")).to(tdt)
            assert torch.equal(back, A), (seed, k, n, M, e, algo)


@pytest.mark.parametrize("seed", _seeds(8))
def test_random_order_one_splines(bk, oracle, seed):
    """Order 1 (piecewise constants, the lowest order the reference accepts): knot search, evaluate_all and spline
    evaluation are the oracle's bit for bit — no arithmetic touches x, so there is no rounding to tolerate (Float32
    default path: one ulp, see below) — plain and periodic, points on and beside every knot, outside, and NaN."""
    rng = np.random.default_rng(15000 + seed)
    f32 = seed % 2 == 1
    periodic = seed % 4 >= 2
    dt, npdt = ("f32", np.float32) if f32 else ("f64", np.float64)
    nb = int(rng.integers(4, 400))
    br = np.unique(_random_breaks(rng, nb, ("jitter", "geometric")[seed % 2]).astype(npdt))
    if periodic:
        B = bk.PeriodicBSplineBasis(1, br)
        kind, tref = 1, br
    else:
        B = bk.BSplineBasis(1, br)
        kind, tref = 0, B.knots()
    a, b = float(br[0]), float(br[-1])
    L = b - a
    x = np.concatenate([a + L * rng.random(3000), br.astype(np.float64),
                        np.nextafter(br, npdt(np.inf)).astype(np.float64), np.nextafter(br, npdt(-np.inf)).astype(np.float64),
                        [a - 0.5 * L, b + 0.5 * L, a - 2.25 * L, b + 2.25 * L, np.nan]]).astype(npdt)
    i_ref, z_ref = oracle.find_knot_interval(kind, 1, tref, x, kt=dt)
    i, z = bk.find_knot_interval(B, x)
    assert np.array_equal(i, i_ref) and np.array_equal(z, z_ref), (seed, dt, periodic)
    ir, bsr = oracle.evaluate_all(kind, 1, tref, x, nderiv=0, kt=dt, vt=dt)
    ig, bsg = bk.evaluate_all(B, x, bk.Derivative(0))
    assert np.array_equal(ig, ir) and np.array_equal(bsg, bsr, equal_nan=True), (seed, dt, periodic)
    c = (2 * rng.random(len(B)) - 1).astype(npdt)
    S = bk.Spline(B, c)
    ref = oracle.spline_eval(kind, 1, tref, c, x, kt=dt)
    got = S.evaluate(x, (0,), algo=bk.EVAL_DEBOOR)[0]
    assert np.array_equal(got, ref, equal_nan=True), (seed, dt, periodic)
    got = S.evaluate(x, (0,))[0]
    if f32:
        # Float32 cell tables flag knot cells in the last mantissa bit of the record's last slot, which for order 1
        # is the constant itself: one ulp at most
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert np.all(np.abs(got[ok] - ref[ok]) <= np.spacing(np.abs(ref[ok]))), (seed, dt, periodic)
    else:
        assert np.array_equal(got, ref, equal_nan=True), (seed, dt, periodic)


@pytest.mark.parametrize("seed", _seeds(10))
def test_random_solve_columns_are_independent(bk, seed):
    """Right-hand sides never mix: NaN / Inf / huge columns planted in a batch leave every other column's result
    bit-identical to the batch without them — for every solve algorithm, including the Float32 kernel that packs two
    right-hand sides into one lane, and for the cyclic (periodic) solves."""
    import torch
    rng = np.random.default_rng(16000 + seed)
    k = int(rng.integers(2, 9))
    n = int(rng.integers(max(2 * k, 8), 2500))
    M = (5, 64, 130, 1024, 2049)[seed % 5]
    f32 = seed % 3 == 2
    periodic = seed % 4 == 3 and not f32
    tdt = torch.float32 if f32 else torch.float64
    if periodic:
        br = (np.arange(n + 1) + np.concatenate([0.45 * (2 * rng.random(n) - 1), [0.0]])) / n
        br[-1] = br[0] + 1.0
        B = bk.PeriodicBSplineBasis(k, br)
        Cm = bk.collocation_matrix(B, bk.collocation_points(B))
        algos = (None,)
    else:
        x = (np.sort(rng.random(n)) + np.arange(n) * 1e-3).astype(np.float32 if f32 else np.float64)
        B = bk.BSplineBasis(k, bk.make_knots(x, k), augment=False)
        Cm = bk.collocation_matrix(B, x)
        algos = (bk.SOLVE_AUTO, bk.SOLVE_WINDOWED, bk.SOLVE_TWOPASS)
    if not periodic:
        Cm.lu_()                                           # (the cyclic matrices are factored when they are built)
    Y = torch.from_numpy(2 * rng.random((Cm.n, M)) - 1).to(tdt).cuda()
    bad = rng.choice(M, size=max(1, M // 7), replace=False)
    Yb = Y.clone()
    fmax = torch.finfo(tdt).max
    for j, col in enumerate(bad):
        Yb[:, col] = (float("nan"), float("inf"), -float("inf"), fmax)[j % 4]
        if j % 2:
            Yb[int(rng.integers(0, Cm.n)), col] = 1.0
    good = np.setdiff1d(np.arange(M), bad)
    for algo in algos:
        A, Bd = Y.clone(), Yb.clone()
        try:
            if algo is None:
                Cm.ldiv_(A); Cm.ldiv_(Bd)
            else:
                Cm.ldiv_(A, algo=algo); Cm.ldiv_(Bd, algo=algo)
        except (bk.UnsupportedError, bk.ArgumentError):
            assert algo == bk.SOLVE_WINDOWED
            continue
        assert torch.equal(A[:, good], Bd[:, good]), (seed, k, n, M, algo)
        nonfin = [int(c) for j, c in enumerate(bad) if j % 4 != 3]
        assert not torch.isfinite(Bd[:, nonfin]).all(dim=0).any(), (seed, k, n, M, algo, "planted columns stay non-finite")


@pytest.mark.parametrize("seed", _seeds(16))
def test_random_pointer_phases(bk, oracle, seed):
    """Point and result buffers at arbitrary element offsets (every combination of 16-byte phases: the vector kernels
    need a common phase, anything else takes the scalar kernels), 1 .. 4 outputs, stream lengths from one point to a
    few vectors past 2^20: the default path stays inside the tolerance and never writes outside its slices; the
    de Boor path stays bit-identical to the oracle."""
    import torch
    rng = np.random.default_rng(17000 + seed)
    k = int(rng.integers(2, 9))
    f32 = seed % 3 == 2
    dt, npdt, tdt, tol = ("f32", np.float32, torch.float32, 1e-5) if f32 else ("f64", np.float64, torch.float64, 1e-13)
    periodic = seed % 4 == 1
    nb = int(rng.integers(max(2 * k + 2, 8), 3000))
    br = np.unique(_random_breaks(rng, nb, ("jitter", "geometric")[seed % 2]).astype(npdt))
    if periodic:
        B = bk.PeriodicBSplineBasis(k, br); kind, tref = 1, br
    else:
        B = bk.BSplineBasis(k, br); kind, tref = 0, B.knots()
    c = (2 * rng.random(len(B)) - 1).astype(npdt)
    S = bk.Spline(B, c)
    a, b = float(br[0]), float(br[-1])
    nout = int(rng.integers(1, min(k, 4) + 1))
    derivs = tuple(int(d) for d in rng.choice(k, size=nout, replace=False))
    refs_cache = {}
    # the tolerance is norm-wise (north_star): the magnitude of D^d S over the domain, not of the few points of a call
    xsc = (a + (b - a) * rng.random(2000)).astype(npdt)
    mag = {}
    for d in derivs:
        if d == 0:
            mag[d] = float(np.max(np.abs(oracle.spline_eval(kind, k, tref, c, xsc, kt=dt))))
        else:
            kd, td, cd = oracle.spline_derivative(kind, k, tref, c, d, kt=dt)
            mag[d] = float(np.max(np.abs(oracle.spline_eval(kind, kd, td, cd, xsc, kt=dt))))
    for n in (1, 2, 3, int(rng.integers(4, 70)), int(rng.integers(70, 5000)), (1 << 20) + int(rng.integers(0, 9))):
        xo = int(rng.integers(0, 8))
        xbuf = torch.full((n + 16,), float("nan"), dtype=tdt, device="cuda")
        xh = (a - 0.05 * (b - a) + 1.1 * (b - a) * rng.random(n)).astype(npdt)
        xbuf[xo:xo + n] = torch.from_numpy(xh).cuda()
        for algo in (bk.EVAL_AUTO, bk.EVAL_DEBOOR):
            offs = [int(rng.integers(0, 8)) for _ in derivs]
            bufs = [torch.full((n + 16,), -7.0, dtype=tdt, device="cuda") for _ in derivs]
            S.evaluate(xbuf[xo:xo + n], derivs, algo=algo, out=[bf[o:o + n] for bf, o in zip(bufs, offs)])
            for d, bf, o in zip(derivs, bufs, offs):
                g = bf.cpu().numpy()
                assert np.all(g[:o] == -7.0) and np.all(g[o + n:] == -7.0), (seed, n, d, "wrote outside the slice")
                if n > 5000:
                    sel = np.concatenate([np.arange(0, 600), np.arange(n - 600, n)])
                else:
                    sel = np.arange(n)
                key = (n, d)
                if key not in refs_cache:
                    if d == 0:
                        refs_cache[key] = oracle.spline_eval(kind, k, tref, c, xh[sel], kt=dt)
                    else:
                        kd, td, cd = oracle.spline_derivative(kind, k, tref, c, d, kt=dt)
                        refs_cache[key] = oracle.spline_eval(kind, kd, td, cd, xh[sel], kt=dt)
                r = refs_cache[key]
                got = g[o:o + n][sel]
                if algo == bk.EVAL_DEBOOR:
                    assert np.array_equal(got, r), (seed, k, dt, n, d, xo, o)
                else:
                    scale = max(float(np.max(np.abs(r))), mag[d], 1e-30)
                    assert np.max(np.abs(got.astype(np.float64) - r.astype(np.float64))) <= tol * scale, (seed, k, dt, n, d, xo, o)
