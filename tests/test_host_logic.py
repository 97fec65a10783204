"""CPU tests of the host-side mirror's constructor logic (no device needed): make_knots,
Greville points, recombination corner blocks, the synthetic generator, operator algebra."""
import numpy as np
import pytest


class StubBasis:
    """Duck-typed BSplineBasis without a device handle (knots only)."""

    def __init__(self, k, t):
        self.k, self.t = k, t

    def __len__(self):
        return self.t.size - self.k


def stub_basis(bk, k, breaks):
    return StubBasis(k, bk.augment_knots(np.asarray(breaks, dtype=np.float64), k))


def test_make_knots_matches_oracle(bk, oracle):
    r = np.random.default_rng(0)
    x = np.cumsum(r.random(50) + 0.1)
    for k in (2, 3, 4, 6, 8):
        assert np.array_equal(bk.make_knots(x, k), oracle.make_knots(x, k))
    xi = np.arange(11) ** 2                       # README case, integer input
    assert np.array_equal(bk.make_knots(xi, 4), oracle.make_knots(xi.astype(float), 4))
    for k in (3, 4, 6):
        assert np.array_equal(bk.make_knots(x, k, bk.Periodic(x[-1] - x[0] + 0.5)),
                              oracle.make_knots_periodic(x, k, x[-1] - x[0] + 0.5))
    with pytest.raises(bk.ArgumentError):
        bk.make_knots(x, 4, bk.Periodic(x[-1] - x[0]))       # endpoint included


def test_greville_points_match_oracle(bk, oracle):
    r = np.random.default_rng(1)
    br = np.sort(r.random(40)); br[0], br[-1] = 0.0, 1.0
    for k in (3, 4, 6, 8):
        t = bk.augment_knots(br, k)
        N = t.size - k
        assert np.array_equal(bk.greville_points(k, t, N, 0, True), oracle.greville(0, k, t))
        assert np.array_equal(bk.greville_points(k, t, N - 2, 1, False), oracle.greville(2, k, t, cl=1, N=N - 2))
    # large case (config C5 size): Kahan sliding sum stays monotone
    from bsplinekit_b200 import synth
    t = bk.augment_knots(synth.breakpoints(10, 99996), 8)
    x = bk.greville_points(8, t, t.size - 8 - 2, 1, False)
    assert x.size == 100000 and np.all(np.diff(x) > 0)
    assert np.array_equal(x, oracle.greville(2, 8, t, cl=1, N=100000))


def test_recombination_corners(bk, oracle):
    from bsplinekit_b200 import api
    r = np.random.default_rng(2)
    br = np.sort(r.random(20)); br[0], br[-1] = 0.0, 1.0
    for k in (4, 5, 8):
        B = stub_basis(bk, k, br)
        op = bk.Derivative(0) + 3 * bk.Derivative(1)
        for side, ref in (("left", "ul"), ("right", "lr")):
            ndrop, nc, A = api._make_submatrix((op,), B, side)
            rc = oracle.robin_corners(k, B.t, [1.0, 3.0])
            assert (ndrop, nc) == (0, 1)
            assert np.array_equal(A, getattr(rc, ref))
        # Robin-like op of higher order: 1.1 D(2) + 4.2 D(3)  (test/recombination.jl:284-305)
        if k >= 5:
            op2 = 1.1 * bk.Derivative(2) + 4.2 * bk.Derivative(3)
            ndrop, nc, A = api._make_submatrix((op2,), B, "left")
            assert A.shape == (4, 3) and np.allclose(A.sum(axis=0), 2.0)
    B = stub_basis(bk, 5, br)
    assert api._make_submatrix((bk.Derivative(0),), B, "left")[:2] == (1, 1)
    assert np.array_equal(api._make_submatrix((bk.Derivative(1),), B, "right")[2], np.ones((2, 1)))
    nd, nc, A = api._make_submatrix((bk.Derivative(0), bk.Derivative(1)), B, "left")
    assert (nd, nc, A.shape) == (2, 2, (2, 0))
    nd, nc, A = api._make_submatrix((bk.Derivative(0), bk.Derivative(2)), B, "left")     # hybrid
    assert (nd, nc, A.shape) == (1, 2, (3, 1))
    with pytest.raises(bk.ArgumentError):
        api._make_submatrix((bk.Derivative(1), bk.Derivative(2)), B, "left")
    with pytest.raises(bk.ArgumentError):
        api._make_submatrix((bk.Derivative(5),), B, "left")


def test_operator_algebra(bk):
    D = bk.Derivative
    op = D(0) + 3 * D(1)
    assert op.terms == {0: 1.0, 1: 3.0} and op.max_order == 1
    assert (-D(2)).terms == {2: -1.0}
    assert (D(1) - 2 * D(3)).terms == {1: 1.0, 3: -2.0}
    with pytest.raises(bk.ArgumentError):
        0 * D(1)
    assert D(2) == D(2) and D(2) != D(1)


def test_synth_numpy_torch_identical():
    import torch
    from bsplinekit_b200 import synth
    a = synth.u_numpy(5, 12345, 4096)
    b = synth.u_torch(5, 12345, 4096, device="cpu").numpy()
    assert np.array_equal(a, b) and 0 <= a.min() and a.max() < 1
    bp = synth.breakpoints(3, 1000)
    assert bp[0] == 0 and bp[-1] == 1 and np.all(np.diff(bp) > 0)
    # slices generated separately concatenate to the full stream (multi-GPU sharding)
    full = synth.u_numpy(9, 0, 1000)
    assert np.array_equal(np.concatenate([synth.u_numpy(9, 0, 400), synth.u_numpy(9, 400, 600)]), full)


def test_differential_ops_reference_assertions(bk):
    """test/diffops.jl:8-50 restated: orders, powers, sums independent of the order of the terms, projections on the
    boundary normals (odd orders flip on the left), derivative ranges."""
    D2, D3 = bk.Derivative(2), bk.Derivative(3)
    assert bk.max_order(D2) == 2 and bk.max_order(D3) == 3
    assert D2 ** 2 == bk.Derivative(4) and D2 ** 3 == bk.Derivative(6)
    assert bk.Derivative() == bk.Derivative(1) and bk.Derivative() ** 3 == bk.Derivative(3)
    S, S2 = D2 + 5 * D3, 5 * D3 + D2
    assert S == S2 and bk.max_order(S) == bk.max_order(S2) == 3
    assert bk.max_order() == -1
    assert bk.dot(bk.LeftNormal(), D2) == D2 == 1 * D2
    assert bk.dot(bk.RightNormal(), D2) == D2
    assert bk.dot(bk.LeftNormal(), D3) == -D3 == -1 * D3
    assert bk.dot(bk.RightNormal(), D3) == D3 == 1 * D3
    assert bk.dot(bk.LeftNormal(), S) == D2 - 5 * D3
    assert bk.dot(bk.RightNormal(), S) == D2 + 5 * D3
    assert bk.dot(D3, bk.LeftNormal()) == -D3                      # dot(op, dir) == dot(dir, op)
    r = bk.Derivative(range(2, 6))                                 # Derivative(2:5)
    assert isinstance(r, bk.DerivativeUnitRange) and repr(r) == "Derivative(2:5)"
    assert r == bk.DerivativeUnitRange(2, 5) and [d.n for d in r] == [2, 3, 4, 5]
    with pytest.raises(bk.ArgumentError):
        0 * D2                                                     # "scale factor cannot be zero"
    assert repr(D2) == "D{2}"


def test_accessor_vocabulary(bk):
    """The small accessors the reference exports next to the hot path: basis_to_array_index (periodic wrap),
    has_parent_basis, constraints, knots_in_main_period — on objects that need no device."""
    class P(bk.PeriodicBSplineBasis):
        def __init__(self, k, t):                                  # no device handle
            self.k, self.t, self._len = k, np.asarray(t, float), len(t) - 1
    p = P(4, np.linspace(0.0, 1.0, 11))
    assert [bk.basis_to_array_index(p, i) for i in (-1, 0, 1, 10, 11, 12)] == [9, 10, 1, 10, 1, 2]
    assert np.array_equal(bk.knots_in_main_period(p), np.linspace(0.0, 1.0, 11)[:-1])
    assert not bk.has_parent_basis(p) and bk.constraints(p) == ((), ())
    assert bk.isindomain(p, 123.0)
