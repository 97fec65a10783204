"""CPU tests: the oracle against the reference's golden vectors and restated property tests.

These pin the oracle (oracle/bsk_oracle.c) before it is trusted as the checker of the CUDA
path: class-A docstring goldens bit for bit, class-B to the printed precision, and the
reference's own property tests (test/bsplines.jl, test/splines.jl, test/collocation.jl,
test/periodic.jl, test/recombination.jl) restated on seeded inputs.
"""
import json
import os

import numpy as np
import pytest

from conftest import dec_range

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "docstring_goldens.json")))


def rng_vals(seed, n):
    # the product's generator is not imported here on purpose: numpy only
    return np.random.default_rng(seed).random(n)


def test_class_a_basis(oracle):
    g = GOLD["class_A"]["basis_k4_range_m1_0p1_1"]
    br = dec_range(g["breaks"]["a"], g["breaks"]["b"], g["breaks"]["n"])
    t = oracle.augment_knots(br, g["k"])
    i, bs = oracle.evaluate_all(0, 4, t, [0.42])
    assert [int(i[0]), bs[0].tolist()] == g["B(0.42)"]
    assert bs[0].sum() == g["sum(bs)"]
    i, bs = oracle.evaluate_all(0, 4, t, [0.44], ileft=[18])
    assert [int(i[0]), bs[0].tolist()] == g["B(0.44;ileft=18)"]
    i, bs = oracle.evaluate_all(0, 4, t, [0.42], vt="f32")
    assert int(i[0]) == 18 and bs.dtype == np.float32
    assert bs[0].tolist() == [float(np.float32(v)) for v in g["B(0.42,Float32)"][1]]
    i, bs = oracle.evaluate_all(0, 4, t, [0.42], nderiv=1)
    assert [int(i[0]), bs[0].tolist()] == g["B(0.42,Derivative(1))"]
    # bs[1] - B[i](0.42) == 0.0 and bs[2] - B[i-1](0.42) == -5.55e-17 (BSplines.jl:66-70)
    _, bs = oracle.evaluate_all(0, 4, t, [0.42])
    assert bs[0][0] - oracle.basis_function(4, t, 18, 0.42) == 0.0
    assert bs[0][1] - oracle.basis_function(4, t, 17, 0.42) == -5.551115123125783e-17


def test_class_a_basis_function(oracle):
    g = GOLD["class_A"]["basis_function_B6"]
    t = oracle.augment_knots(dec_range(-1, 1, 21), 4)
    assert oracle.basis_function(4, t, 6, -0.5) == g["B[6](-0.5)"]
    assert oracle.basis_function(4, t, 6, -0.5, 1) == g["B[6](-0.5,Derivative(1))"]


def test_class_a_periodic(oracle):
    g = GOLD["class_A"]["periodic_k4_range_m1_0p2_1"]
    br = dec_range(g["breaks"]["a"], g["breaks"]["b"], g["breaks"]["n"])
    i, bs = oracle.evaluate_all(1, 4, br, [-0.42])
    assert [int(i[0]), bs[0].tolist()] == g["B(-0.42)"]
    i, bs = oracle.evaluate_all(1, 4, br, [-0.42 + 2])
    assert [int(i[0]), bs[0].tolist()] == g["B(-0.42+2)"]


def test_class_a_knot_layouts(oracle):
    g = GOLD["class_A"]["knot_layouts"]
    assert oracle.augment_knots(np.array([0, 0.1, 0.2, 0.6, 1.0]), 3).tolist() == g["BSplineBasis(3, [0, 0.1, 0.2, 0.6, 1])"]


def test_class_b_interpolation(oracle):
    g = GOLD["class_B"]["interpolate_cospi"]
    xs = dec_range(g["xs"]["a"], g["xs"]["b"], g["xs"]["n"])
    ys = np.cos(np.pi * xs)
    k = g["k"]
    t = oracle.make_knots(xs, k)
    band, nout = oracle.collocation_matrix(0, k, t, xs)
    assert nout == 0 and oracle.banded_lu(band) == 0
    c = oracle.banded_solve(band, ys.copy().reshape(-1, 1)).ravel()
    assert np.allclose(c[:5], g["coefs_head_6digits"], rtol=6e-6)
    ev = lambda x: oracle.spline_eval(0, k, t, c, [x])[0]
    assert ev(-1.0) == g["S(-1)"]
    assert ev(-0.99) == pytest.approx(g["S(-0.99)"], abs=2e-15)
    assert ev(0.998) == pytest.approx(g["S(0.998)"], abs=2e-15)
    k1, t1, c1 = oracle.spline_derivative(0, k, t, c, 1)
    k2, t2, c2 = oracle.spline_derivative(0, k, t, c, 2)
    assert oracle.spline_eval(0, k1, t1, c1, [-1.0])[0] == pytest.approx(g["D1S(-1)"], abs=5e-13)
    assert oracle.spline_eval(0, k2, t2, c2, [-1.0])[0] == pytest.approx(g["D2S(-1)"], abs=5e-11)
    p = g["periodic"]
    xp = dec_range(p["xp"]["a"], p["xp"]["b"], p["xp"]["n"])
    tp = oracle.make_knots_periodic(xp, 4, p["L"])
    a, b, cc, nout = oracle.collocation_cyclic(4, tp, xp)
    assert nout == 0
    cp = oracle.cyclic_solve(a, b, cc, np.cos(np.pi * xp))
    assert np.allclose(cp[:5], p["coefs_head_6digits"], rtol=6e-6)
    assert oracle.spline_eval(1, 4, tp, cp, [-0.99])[0] == pytest.approx(p["Sper(-0.99)"], abs=2e-15)
    assert oracle.spline_eval(1, 4, tp, cp, [0.998])[0] == pytest.approx(p["Sper(0.998)"], abs=2e-15)


@pytest.mark.parametrize("k", [2, 4, 5, 6, 8])
def test_find_knot_interval_identities(oracle, k):
    """test/bsplines.jl:18-30."""
    br = np.sort(rng_vals(1, 30)); br[0], br[-1] = 0.0, 1.0
    t = oracle.augment_knots(br, k)
    N = t.size - k
    a, b = t[0], t[-1]
    i, z = oracle.find_knot_interval(0, k, t, [a - 1, a, b, b + 1])
    assert list(zip(i.tolist(), z.tolist())) == [(1, -1), (k, 0), (N, 0), (N, 1)]
    x = (a + b) / 3
    i, z = oracle.find_knot_interval(0, k, t, [x])
    assert (i[0], z[0]) == (np.searchsorted(t, x, side="right"), 0)
    j = t.size // 2
    i, z = oracle.find_knot_interval(0, k, t, [t[j - 1]])
    assert (i[0], z[0]) == (np.searchsorted(t, t[j - 1], side="right"), 0)


@pytest.mark.parametrize("k", [2, 4, 5, 6, 8])
def test_evaluate_all_properties(oracle, k):
    """test/bsplines.jl:49-123: partition of unity, match with the recursive evaluator,
    dot(cs, bs) == (op * S)(x), Float32 variant."""
    br = np.sort(rng_vals(2, 25)); br[0], br[-1] = 0.0, 1.0
    t = oracle.augment_knots(br, k)
    N = t.size - k
    c = rng_vals(3, N) - 0.5
    xs = np.array([t[0], t[k + 2], 0.2 * t[k + 2] + 0.8 * t[k + 3], t[-1]])
    for nd in (0, 1, 2):
        if nd >= k:
            continue
        i, bs = oracle.evaluate_all(0, k, t, xs, nderiv=nd)
        i32, bs32 = oracle.evaluate_all(0, k, t, xs, nderiv=nd, vt="f32")
        assert np.array_equal(i, i32)
        assert np.allclose(bs, bs32, rtol=2e-5, atol=1e-4 * max(1.0, np.abs(bs).max()))
        tot = bs.sum(axis=1)
        if nd == 0:
            assert np.allclose(tot, 1.0, atol=1e-14)
        else:
            assert np.allclose(tot, 0.0, atol=1e-9 * np.abs(bs).max())
        kd, td, cd = (k, t, c) if nd == 0 else oracle.spline_derivative(0, k, t, c, nd)
        sv = oracle.spline_eval(0, kd, td, cd, xs)
        for p, x in enumerate(xs):
            for d in range(k):
                j = i[p] - d
                if 1 <= j <= N:
                    assert bs[p, d] == pytest.approx(oracle.basis_function(k, t, j, x, nd),
                                                     rel=1e-10, abs=1e-10 * max(1, np.abs(bs).max()))
            dot = sum(c[i[p] - d - 1] * bs[p, d] for d in range(k) if 1 <= i[p] - d <= N)
            assert dot == pytest.approx(sv[p], rel=1e-9, abs=1e-9 * max(1, np.abs(bs).max()))


@pytest.mark.parametrize("k", [3, 4, 6])
def test_periodic_matches_regular_far_from_boundaries(oracle, k):
    """test/periodic.jl:127-157: a periodic basis equals a regular one (bit for bit) away from
    the boundaries, with index shift delta = (k - 1) - k ÷ 2."""
    br = np.linspace(-1.0, 1.0, 41)
    t = oracle.augment_knots(br, k)
    xs = np.linspace(-0.3, 0.3, 17)
    i_r, bs_r = oracle.evaluate_all(0, k, t, xs)
    i_p, bs_p = oracle.evaluate_all(1, k, br, xs)
    assert np.array_equal(bs_r, bs_p)
    assert np.all(i_r - i_p == (k - 1) - k // 2)
    # partition of unity and periodic shift ilast' == ilast + nN (test/periodic.jl:77-122)
    x = 2 * rng_vals(4, 50) - 1
    i0, b0 = oracle.evaluate_all(1, k, br, x)
    i1, b1 = oracle.evaluate_all(1, k, br, x + 2 * 3)
    assert np.allclose(b0.sum(axis=1), 1, atol=1e-14)
    assert np.all(i1 == i0 + 3 * 40)
    assert np.allclose(b0, b1, atol=1e-12)


def test_banded_lu_cases(oracle):
    """test/collocation.jl:4-53."""
    w = np.array([[2.0]], order="F")
    assert oracle.banded_lu(w) == 0
    w = np.array([[0.0]], order="F")
    assert oracle.banded_lu(w) == 1
    band = np.zeros((3, 4), order="F"); band[1] = [1, 2, 0, 4]
    assert oracle.banded_lu(band) == 3
    # random diagonally dominant system: L*U == A and A*(F\y) == y
    n, h = 40, 3
    r = np.random.default_rng(5)
    band = np.asfortranarray(r.random((2 * h + 1, n)))
    band[h] += 2 * h + 1
    A = np.zeros((n, n))
    for j in range(n):
        for i in range(max(0, j - h), min(n, j + h + 1)):
            A[i, j] = band[h + i - j, j]
    lu = band.copy(order="F")
    assert oracle.banded_lu(lu) == 0
    L = np.eye(n); U = np.zeros((n, n))
    for j in range(n):
        for i in range(max(0, j - h), min(n, j + h + 1)):
            if i > j: L[i, j] = lu[h + i - j, j]
            else: U[i, j] = lu[h + i - j, j]
    assert np.allclose(L @ U, A, atol=1e-13)
    y = r.random((n, 3))
    x = oracle.banded_solve(lu, y.copy())
    assert np.allclose(A @ x, y, atol=1e-12)
    assert np.allclose(oracle.banded_matvec(band, x), y, atol=1e-12)


@pytest.mark.parametrize("k", [3, 4, 6, 8])
def test_polynomial_reproduction(oracle, k):
    """test/splines.jl:14-78: interpolating a degree k-1 polynomial at the Greville sites
    reproduces P and P' on 9N+42 points."""
    br = np.sort(rng_vals(6, 12)); br[0], br[-1] = -1.0, 1.0
    t = oracle.augment_knots(br, k)
    N = t.size - k
    xc = oracle.greville(0, k, t)
    assert xc[0] == t[0] and xc[-1] == t[-1] and np.all(np.diff(xc) > 0)
    pc = rng_vals(7, k) - 0.5
    P = np.polynomial.Polynomial(pc)
    band, nout = oracle.collocation_matrix(0, k, t, xc)
    assert nout == 0
    dense_rows = oracle.banded_matvec(band, np.ones(N))
    assert np.allclose(dense_rows, 1.0, atol=1e-14)          # rows sum to one
    assert oracle.banded_lu(band) == 0
    c = oracle.banded_solve(band, P(xc).reshape(-1, 1).copy()).ravel()
    xs = np.linspace(-1, 1, 9 * N + 42)
    assert np.allclose(oracle.spline_eval(0, k, t, c, xs), P(xs), atol=1e-12)
    k1, t1, c1 = oracle.spline_derivative(0, k, t, c, 1)
    assert np.allclose(oracle.spline_eval(0, k1, t1, c1, xs), P.deriv()(xs), atol=1e-10)


@pytest.mark.parametrize("k", [4, 5, 8])
def test_robin_recombination_satisfies_bc(oracle, k):
    """test/recombination.jl:150-172,284-305: every recombined function satisfies
    u + 3 du/dn = 0 at both ends (d/dn = -d/dx on the left)."""
    br = np.sort(rng_vals(8, 15)); br[0], br[-1] = 0.0, 1.0
    t = oracle.augment_knots(br, k)
    N = t.size - k
    rc = oracle.robin_corners(k, t, [1.0, 3.0])
    # column sums of the corner blocks are 2 (normalisation r = 2/(b0-b1), matrices.jl:288-309)
    assert rc.ul.sum() == pytest.approx(2.0) and rc.lr.sum() == pytest.approx(2.0)
    for x, sign in ((0.0, -1.0), (1.0, 1.0)):
        _, v = oracle.evaluate_all(0, k, t, [x], recomb=rc)
        _, d = oracle.evaluate_all(0, k, t, [x], nderiv=1, recomb=rc)
        resid = v[0] + 3 * sign * d[0]
        assert np.max(np.abs(resid)) < 1e-9 * max(1.0, np.abs(d).max())
    # collocation rows of the recombined basis are column combinations of the parent's
    M = N - 2
    xcol = oracle.greville(2, k, t, cl=1, N=M)
    assert np.all((xcol > 0) & (xcol < 1))
    band_r, nout = oracle.collocation_matrix(2, k, t, xcol, recomb=rc)
    assert nout == 0 and band_r.shape == (2 * k - 3, M)


def test_cyclic_solve_against_dense(oracle):
    """test/periodic.jl:159-182: cond(C) < 1e3, rows sum to one, C x == y."""
    br = np.sort(rng_vals(9, 31)); br[0], br[-1] = 0.0, 1.0
    x = br[:-1].copy()
    a, b, c, nout = oracle.collocation_cyclic(4, br, x)
    n = a.size
    A = np.zeros((n, n))
    for i in range(n):
        A[i, i] = b[i]; A[i, (i - 1) % n] = a[i]; A[i, (i + 1) % n] = c[i]
    assert nout == 0 and np.allclose(A.sum(axis=1), 1.0, atol=1e-14)
    assert np.linalg.cond(A) < 1e3
    d = rng_vals(10, n * 3).reshape(n, 3)
    sol = oracle.cyclic_solve(a, b, c, d.copy())
    assert np.allclose(A @ sol, d, atol=1e-12)


def test_greville_kahan_and_make_knots(oracle):
    x = np.cumsum(rng_vals(11, 40))
    for k in (3, 4, 6):
        t = oracle.make_knots(x, k)
        assert t.size == x.size + k and np.all(t[:k] == x[0]) and np.all(t[-k:] == x[-1])
        assert np.all(np.diff(t) >= 0)
        kinv = 1.0 / (k - 1)
        assert t[k] == sum(x[1:k].tolist(), 0.0) * kinv or t[k] == pytest.approx(np.sum(x[1:k]) * kinv, rel=1e-15)


def _interp_coefs(oracle, k, t, xs, ys):
    band, nout = oracle.collocation_matrix(0, k, t, xs)
    assert nout == 0 and oracle.banded_lu(band) == 0
    return oracle.banded_solve(band, np.array(ys, dtype=np.float64).reshape(-1, 1)).ravel()


def test_class_c_approximate(oracle):
    """approximate(f, B) = interpolation at collocation_points(B); VariationDiminishing = f at the
    Greville sites (SplineApproximations.jl:78-119, by_interpolation.jl:40-71, variation_diminishing.jl:35-44)."""
    g = GOLD["class_C"]["approximate_k3"]
    k = g["k"]
    t = oracle.augment_knots(dec_range(g["breaks"]["a"], g["breaks"]["b"], g["breaks"]["n"]), k)
    xs = oracle.greville(0, k, t)
    assert np.allclose(xs, g["interp_points"], atol=1e-15)
    cs = _interp_coefs(oracle, k, t, xs, np.sin(xs))
    assert np.allclose(cs, g["sin_coefs_6digits"], rtol=6e-6, atol=1e-15)
    assert oracle.spline_eval(0, k, t, cs, [0.3])[0] == pytest.approx(g["sin_S(0.3)"], abs=1e-15)
    ce = _interp_coefs(oracle, k, t, xs, np.exp(xs))
    assert np.allclose(ce, g["exp_coefs_6digits"], rtol=6e-6)
    assert oracle.spline_eval(0, k, t, ce, [0.3])[0] == pytest.approx(g["exp_S(0.3)"], abs=2e-15)
    assert oracle.spline_eval(0, k, t, ce, [0.34])[0] == pytest.approx(g["exp_S_interp(0.34)"], abs=2e-15)
    cv = np.exp(xs)
    assert np.allclose(cv, g["exp_vd_coefs_6digits"], rtol=6e-6)
    assert oracle.spline_eval(0, k, t, cv, [0.34])[0] == pytest.approx(g["exp_S_vd(0.34)"], abs=2e-15)
    r = GOLD["class_C"]["readme_approximate_k6"]
    t6 = oracle.augment_knots(np.log2(np.arange(1, 17)), 6)
    x6 = oracle.greville(0, 6, t6)
    c6 = _interp_coefs(oracle, 6, t6, x6, np.exp(-x6) * np.sin(x6))
    assert oracle.spline_eval(0, 6, t6, c6, [2.3])[0] == pytest.approx(r["fapprox(2.3)"], abs=1e-14)


def test_class_c_smoothing_fit(oracle):
    """fit(BSplineOrder(4), xs, ys, λ) docstring coefficients (smoothing.jl:30-42)."""
    g = GOLD["class_C"]["smoothing_fit"]
    xs = dec_range(0.0, 1.0, 101) ** 2
    ys = np.cos(2 * np.pi * xs) + 0.1 * np.sin(200 * np.pi * xs)
    cs, t, rc = oracle.fit_smoothing(xs, ys, g["lambda"])
    assert np.allclose(cs[:10], g["coefs_head_6digits"], rtol=6e-6)
    assert np.allclose(cs[-10:], g["coefs_tail_6digits"], rtol=6e-6)
    # λ = 0: the natural interpolating spline (smoothing.jl:7-8)
    c0, _, _ = oracle.fit_smoothing(xs, ys, 0.0)
    idx, bs = oracle.evaluate_all(2, 4, t, xs, recomb=rc)
    vals = np.array([sum(bs[p, d] * c0[idx[p] - d - 1] for d in range(4) if 0 <= idx[p] - d - 1 < c0.size)
                     for p in range(xs.size)])
    assert np.max(np.abs(vals - ys)) < 1e-9


@pytest.mark.parametrize("k", [3, 5, 6, 8])
def test_periodic_interpolation_general_order(oracle, k):
    """Periodic interpolation for k not in {2, 4}: I.(xs) ≈ ys (test/periodic.jl:159-180); collocation
    points: even k = knots, odd k = knot averages inside the period."""
    rng = np.random.default_rng(5)
    N, L = 40, 2.0
    x = np.sort(-1.0 + L * (np.arange(N) + 0.6 * (rng.random(N) - 0.5)) / N)
    data = oracle.make_knots_periodic(x, k, L)
    xc = oracle.collocation_points_periodic(k, data)
    if k % 2 == 0:
        assert np.array_equal(xc, x)                    # SameAsKnots: make_knots put the knots on the data
    else:
        assert xc.size == N and np.all(np.diff(xc) > 0) and xc[0] >= data[0] and xc[-1] <= data[-1]
    A = oracle.periodic_collocation_dense(k, data, x)
    assert np.allclose(A.sum(axis=1), 1.0, atol=1e-14)  # partition of unity
    y = np.cos(np.pi * x) + 0.3 * np.sin(3 * np.pi * x)
    c = np.linalg.solve(A, y)
    vals = oracle.spline_eval(1, k, data, c, x)
    assert np.max(np.abs(vals - y)) < 1e-12
    # periodicity of the interpolant
    xs = np.linspace(-1, 1, 33)
    assert np.max(np.abs(oracle.spline_eval(1, k, data, c, xs) - oracle.spline_eval(1, k, data, c, xs + L))) < 1e-12


def test_class_c_minimise_l2(oracle):
    """approximate(exp, B, MinimiseL2Error()): Galerkin mass matrix + projection
    (SplineApproximations.jl:104-119, minimiseL2.jl:52-77, Galerkin/linear.jl:268-311, projection.jl:60-99)."""
    g = GOLD["class_C"]["approximate_k3"]
    k = g["k"]
    t = oracle.augment_knots(dec_range(g["breaks"]["a"], g["breaks"]["b"], g["breaks"]["n"]), k)
    M = oracle.galerkin_matrix(k, t)
    assert np.max(np.abs(M - M.T)) < 1e-16 and M.sum() == pytest.approx(2.0, abs=1e-14)   # sum_ij <b_i, b_j> = b - a
    c = np.linalg.solve(M, oracle.galerkin_projection(np.exp, k, t))
    assert np.allclose(c, g["exp_l2_coefs_6digits"], rtol=6e-6)
    assert oracle.spline_eval(0, k, t, c, [0.34])[0] == pytest.approx(g["exp_S_opt(0.34)"], abs=5e-15)
    # stiffness matrix of the derivative: rows sum to zero (constants are in the kernel)
    K1 = oracle.galerkin_matrix(k, t, 1, 1)
    assert np.max(np.abs(K1.sum(axis=1))) < 1e-12


# ------------------------------------------------------------------------------------------------
# Float32 restatement of the interpolation path (oracle/bsk_oracle_band_tmpl.h)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [2, 4, 6, 8])
def test_float32_interpolation_chain(oracle, k):
    """CollocationMatrix{Float32} + lu! + ldiv! in Float32: properties the reference tests state for
    any element type (test/collocation.jl:4-53,112-121; test/interpolation.jl:61-66) and agreement with
    the Float64 chain at Float32 accuracy."""
    rng = np.random.default_rng(3 + k)
    n = 120
    x = np.sort(rng.random(n)); x[0], x[-1] = 0.0, 1.0
    x32 = x.astype(np.float32)
    x64 = x32.astype(np.float64)
    t64 = oracle.make_knots(x64, k)
    t32 = t64.astype(np.float32)
    band32, nout = oracle.collocation_matrix_f32(0, k, t32, x32)
    assert nout == 0 and band32.dtype == np.float32
    A32 = oracle.band_to_dense(band32.astype(np.float64))
    assert np.allclose(A32.sum(axis=1), 1.0, atol=2e-6)              # partition of unity
    band64, _ = oracle.collocation_matrix(0, k, t64, x64)
    assert np.max(np.abs(band32 - band64)) < 2e-4                     # knots rounded to Float32
    lu32 = band32.copy(order="F")
    assert oracle.banded_lu(lu32) == 0
    # L * U == A in Float32 accuracy
    h = k - 2
    L = np.eye(n); U = np.zeros((n, n))
    D = oracle.band_to_dense(lu32.astype(np.float64))
    for i in range(n):
        for j in range(max(0, i - h), min(n, i + h + 1)):
            if j < i:
                L[i, j] = D[i, j]
            else:
                U[i, j] = D[i, j]
    assert np.max(np.abs(L @ U - A32)) < 1e-5
    Y = (2 * rng.random((n, 9)) - 1).astype(np.float32)
    C32 = oracle.banded_solve(lu32, Y.copy())
    assert C32.dtype == np.float32
    ref = np.linalg.solve(A32, Y.astype(np.float64))
    assert np.max(np.abs(C32 - ref)) / np.max(np.abs(ref)) < 5e-4     # Float32 rounding x conditioning
    res = np.max(np.abs(A32 @ C32.astype(np.float64) - Y))            # backward error of the solve
    assert res < 1e-6 * k * max(1.0, float(np.max(np.abs(C32))))
    # zero pivot is reported with its row (ZeroPivotException(i))
    z = np.zeros((3, 4), dtype=np.float32, order="F"); z[1] = [1, 2, 0, 4]
    assert oracle.banded_lu(z) == 3


def test_float32_cyclic_chain(oracle):
    n = 64
    rng = np.random.default_rng(11)
    br = np.concatenate([[0.0], np.sort(rng.random(n - 1)), [1.0]]).astype(np.float32)
    a, b, c, nout = oracle.collocation_cyclic_f32(4, br, br[:-1])
    assert nout == 0
    A = np.zeros((n, n))
    for i in range(n):
        A[i, i] = b[i]; A[i, (i - 1) % n] = a[i]; A[i, (i + 1) % n] = c[i]
    assert np.allclose(A.sum(axis=1), 1.0, atol=2e-6)
    Y = (2 * rng.random((n, 5)) - 1).astype(np.float32)
    X = oracle.cyclic_solve_f32(a, b, c, Y.copy())
    ref = np.linalg.solve(A, Y.astype(np.float64))
    assert np.max(np.abs(X - ref)) / np.max(np.abs(ref)) < 1e-4


def test_oracle_galerkin_identities(oracle):
    """Pins the oracle's Galerkin restatement (src/Galerkin/linear.jl:268-311, projection.jl:60-99) on the
    reference's own assertions: analytical mass-matrix entries for k = 2 (test/galerkin.jl:11-27), sum(G) == length
    of the domain and column sums == (t[j+k] - t[j]) / k (test/galerkin.jl:131-146, test/periodic.jl:184-196) for plain
    and periodic bases, projection of the constant 1 == those column sums."""
    import numpy as np
    x = np.array([-np.cos(np.pi * n / 42) for n in range(43)])
    t = oracle.augment_knots(x, 2)
    G = oracle.galerkin_matrix(2, t)
    i = 8
    assert abs(G[i - 1, i - 1] - (t[i + 1] - t[i - 1]) / 3) < 1e-15
    assert abs(G[i - 2, i - 1] - (t[i] - t[i - 1]) / 6) < 1e-15
    for k in (3, 4, 6):
        t = oracle.augment_knots(x, k)
        G = oracle.galerkin_matrix(k, t)
        N = t.size - k
        assert abs(G.sum() - 2.0) < 1e-13
        assert np.allclose(G.sum(axis=0), (t[k:k + N] - t[:N]) / k, rtol=1e-12, atol=1e-15)
        assert np.allclose(oracle.galerkin_projection(lambda v: 1.0, k, t), (t[k:k + N] - t[:N]) / k, rtol=1e-12, atol=1e-15)
    rng = np.random.default_rng(5)
    for k in (2, 3, 4, 5, 6):
        data = np.sort(np.concatenate([[0.0], 3.0 * rng.random(25), [3.0]]))
        N = data.size - 1
        Gp = oracle.galerkin_matrix_periodic(k, data)
        ts = oracle.periodic_knots(data, k)
        assert abs(Gp.sum() - 3.0) < 1e-13
        assert np.allclose(Gp.sum(axis=0), (ts[k:k + N] - ts[:N]) / k, rtol=1e-12, atol=1e-15)
        assert np.max(np.abs(Gp - Gp.T)) < 1e-15
        assert np.allclose(oracle.galerkin_projection_periodic(lambda v: 1.0, k, data), (ts[k:k + N] - ts[:N]) / k,
                           rtol=1e-12, atol=1e-15)


def test_oracle_nan_point_semantics(oracle):
    """What the reference does with a NaN point, read off its code and pinned here because the GPU tests compare with
    it: `searchsortedlast(t, NaN) == length(t)` (isless sorts NaN last), so a plain basis walks back to the last
    non-empty interval with zone 0 (BSplines/evaluate_all.jl:102-119) and a periodic one — NaN passes both wrap
    loops — lands on `k ÷ 2 + N + 1` (BSplines/periodic.jl:85-100); the value is NaN for every order above 1, and the
    coefficient of that interval for order 1 (no arithmetic touches x)."""
    br = np.array([0.0, 0.3, 0.45, 0.8, 1.0])
    x = np.array([np.nan])
    for k in (1, 2, 4):
        t = np.concatenate([[br[0]] * (k - 1), br, [br[-1]] * (k - 1)])
        i, z = oracle.find_knot_interval(0, k, t, x)
        assert i[0] == len(br) - 1 + (k - 1) and z[0] == 0            # the last interval [0.8, 1.0]
        c = np.arange(1.0, len(t) - k + 1)
        v = oracle.spline_eval(0, k, t, c, x)
        assert (v[0] == c[-1]) if k == 1 else np.isnan(v[0])
        N = len(br) - 1
        ip, zp = oracle.find_knot_interval(1, k, br, x)
        assert ip[0] == k // 2 + N + 1 and zp[0] == 0
        cp = np.arange(1.0, N + 1)
        vp = oracle.spline_eval(1, k, br, cp, x)
        assert (vp[0] == cp[0]) if k == 1 else np.isnan(vp[0])         # order 1: PeriodicVector wraps N + 1 to 1
