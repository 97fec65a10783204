"""CPU tests of the host-side table builder (bsplinekit.jl_b200/csrc/pptable.h), through a CPU emulation
of the table-evaluation kernels (tests/c_driver/pptable_emul.cpp) — the accuracy certificate of the
default (BSK_EVAL_AUTO) spline-evaluation path, checked against the oracle WITHOUT a GPU:

* every derivative order d = 0 .. k-1 within the north-star bar (1e-13 Float64 / 1e-5 Float32, relative to
  max|ref| over the array) on knot vectors with repeated and clustered knots, shifted / scaled domains
  (the reference semantics: src/Splines/spline.jl:144-209, 272-330; src/BSplines/evaluate_all.jl:102-119);
* a per-element bound |got - ref| <= tol * max(|ref|, 2^-5 ||c||_inf) on the benchmark shapes;
* Float32 tables with 10^4 .. 10^5 knot intervals, where the device's Float32 cell index drifts against
  the host's cell edges: points at the knots and one ulp either side must take the right piece.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_driver", "pptable_emul.cpp")
SO = os.path.join(ROOT, "tests", "c_driver", "libpptable_emul.so")


@pytest.fixture(scope="module")
def emul():
    deps = [SRC] + [os.path.join(ROOT, "bsplinekit.jl_b200", "csrc", f) for f in ("pptable.h", "dd.h", "core.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                               "-o", SO, SRC])
    L = C.CDLL(SO)
    L.ppe_create.restype = C.c_void_p
    L.ppe_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int]
    L.ppe_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ppe_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ppe_destroy.argtypes = [C.c_void_p]

    class Emul:
        """Tables of one spline + emulated evaluation.  path 0 = what BSK_EVAL_AUTO launches, 1 = BSK_EVAL_PP."""

        def __init__(self, kind, k, knots, N, coefs, f32=False):
            self.kn = np.ascontiguousarray(knots, dtype=np.float64)
            self.c = np.ascontiguousarray(coefs, dtype=np.float64)
            self.h = L.ppe_create(kind, k, int(f32), self.kn.ctypes.data, self.kn.size, N, self.c.ctypes.data,
                                  self.c.size, 1)

        def __del__(self):
            L.ppe_destroy(self.h)

        def info(self):
            iv = np.zeros(10, dtype=np.int64)
            sc = np.zeros(8)
            L.ppe_info(self.h, iv.ctypes.data, sc.ctypes.data)
            names = "nrec n_uncert cell_valid cell_global cell_jump ncell nside n_cell_slow n_flagged smem".split()
            return dict(zip(names, iv.tolist())), sc

        def eval(self, x, d=0, path=0):
            x = np.ascontiguousarray(x, dtype=np.float64)
            out = np.empty_like(x)
            rt = np.zeros(x.size, dtype=np.int8)
            L.ppe_eval(self.h, x.ctypes.data, x.size, d, path, out.ctypes.data, rt.ctypes.data)
            return out, rt

    return Emul


@pytest.fixture(scope="module")
def synth():
    from bsplinekit_b200 import synth as s
    return s


def _ref(oracle, kind, k, t, c, x, d, dtype):
    if d == 0:
        return oracle.spline_eval(kind, k, t, c, x, kt=dtype).astype(np.float64)
    kd, td, cd = oracle.spline_derivative(kind, k, t, c, d, kt=dtype)
    return oracle.spline_eval(kind, kd, td, cd, x, kt=dtype).astype(np.float64)


@pytest.mark.parametrize("k", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_awkward_knot_vectors_all_derivatives(emul, oracle, k, dtype):
    """The inputs of tests/test_gpu_parity.py::test_eval_awkward_knot_vectors; every derivative order at the
    north-star tolerance, on both table paths.  Cluster pieces must be routed to the de Boor recurrence."""
    T = np.float64 if dtype == "f64" else np.float32
    tol = 1e-13 if dtype == "f64" else 1e-5
    rng = np.random.default_rng(1000 + k)
    for case, (a, b) in enumerate(((0.0, 1.0), (1.0e3, 1.0e3 + 2.0), (-3.0e-3, 5.0e-3))):
        base = np.sort(rng.random(40))
        base[0], base[-1] = 0.0, 1.0
        pieces = [base]
        for j in rng.choice(np.arange(1, 39), size=6, replace=False):
            pieces.append(np.repeat(base[j], rng.integers(1, k)))
        if dtype == "f64":
            for j in rng.choice(np.arange(1, 39), size=4, replace=False):
                pieces.append(base[j] + 1e-9 * np.arange(1, 4))
        br = np.sort(np.concatenate(pieces))
        br = (a + (b - a) * br).astype(T)
        br[0], br[-1] = T(a), T(b)
        br = np.sort(br)
        t = oracle.augment_knots(br, k).astype(T)
        N = t.size - k
        c = (2 * rng.random(N) - 1).astype(T)
        x = np.concatenate([a + (b - a) * rng.random(20000), br.astype(np.float64),
                            np.nextafter(br, T(-np.inf)).astype(np.float64),
                            np.nextafter(br, T(np.inf)).astype(np.float64), [a - 1.0, b + 1.0]]).astype(T)
        E = emul(0, k, t, N, c, f32=(dtype == "f32"))
        info, _ = E.info()
        if dtype == "f64" and k >= 3:
            assert info["n_uncert"] > 0            # the 1e-9 clusters cannot be certified: de Boor route
        for d in range(k):
            ref = _ref(oracle, 0, k, t, c, x, d, dtype)
            sc = np.max(np.abs(ref))
            for path in (0, 1):
                got, rt = E.eval(x, d, path)
                bad = np.abs(got - ref) > tol * sc
                assert not bad.any(), (case, d, path, x[bad][:3], got[bad][:3], ref[bad][:3], rt[bad][:3])
                if info["n_uncert"] > 0:
                    ex = rt == 3                   # de Boor route: the oracle's own bits
                    assert ex.any() and np.array_equal(got[ex], ref[ex])


@pytest.mark.parametrize("k,nb,derivs", [(4, 1000, (0, 1, 2, 3)), (6, 1000, (0, 1, 2, 5)), (8, 600, (0, 1, 7)),
                                         (5, 1000, (0, 1, 4)), (3, 1500, (0, 1, 2))])
def test_benchmark_shapes_per_element(emul, oracle, synth, k, nb, derivs):
    """C2's shape (and the other orders on the same breakpoints): norm-wise bar on every derivative order, and
    the per-element bound |got - ref| <= 1e-13 max(|ref|, 2^-5 ||c||) incl. the bit-encoded knot cells."""
    br = synth.breakpoints(3, nb)
    t = oracle.augment_knots(br, k)
    N = t.size - k
    c = 2 * synth.u_numpy(4, 0, N) - 1
    x = np.concatenate([synth.u_numpy(5, 0, 300000), br, np.nextafter(br, -1), np.nextafter(br, 2)])
    E = emul(0, k, t, N, c)
    info, scale = E.info()
    assert info["cell_valid"] and not info["cell_global"] and info["cell_jump"] and info["n_uncert"] == 0
    assert info["n_cell_slow"] <= 2
    for d in derivs:
        ref = _ref(oracle, 0, k, t, c, x, d, "f64")
        got, rt = E.eval(x, d, 0)
        sc = np.max(np.abs(ref))
        assert 0.3 * sc <= scale[d] <= 1.5 * sc      # the certificate's scale is the sup-norm of D^d S
        assert np.max(np.abs(got - ref)) <= 1e-13 * sc
        flagged = rt == 1
        assert flagged.sum() > 10000
        floor = np.max(np.abs(ref)) / 32 if d else np.max(np.abs(c)) / 32
        pe = np.abs(got - ref) / np.maximum(np.abs(ref), floor)
        assert pe.max() <= 1e-13, (d, pe.max())
        assert pe[flagged].max() <= 1e-13


def test_c4_periodic_global_table(emul, oracle, synth):
    """C4's shape: periodic cubic on 10^4 breakpoints, cell table in global memory (record-major), values; the
    first derivative goes through the tables of the derivative spline (Splines/periodic.jl:76-117)."""
    br = synth.breakpoints(8, 10001)
    c = 2 * synth.u_numpy(9, 0, 10000) - 1
    x = np.concatenate([synth.u_numpy(9, 0, 300000), br, np.nextafter(br, -1), np.nextafter(br, 2)])
    x = x[(x >= 0) & (x < 1)]
    E = emul(1, 4, br, 10000, c)
    info, _ = E.info()
    assert info["cell_valid"] and info["cell_global"] and info["n_uncert"] == 0 and info["n_cell_slow"] == 0
    ref = _ref(oracle, 1, 4, br, c, x, 0, "f64")
    got, rt = E.eval(x, 0, 0)
    assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref))
    pe = np.abs(got - ref) / np.maximum(np.abs(ref), np.max(np.abs(c)) / 32)
    assert pe.max() <= 1e-13
    kd, td, cd = oracle.spline_derivative(1, 4, br, c, 1)
    Ed = emul(1, kd, td, 10000, cd)
    refd = oracle.spline_eval(1, kd, td, cd, x)
    gotd, _ = Ed.eval(x, 0, 0)
    assert np.max(np.abs(gotd - refd)) <= 2e-14 * np.max(np.abs(refd))


@pytest.mark.parametrize("nb", [10000, 99996])
@pytest.mark.parametrize("k", [2, 4, 8])
def test_float32_large_tables_pick_the_right_piece(emul, oracle, synth, k, nb):
    """Float32 splines on 10^4 / 10^5 intervals (up to 2 x 10^6 cells): the Float32 cell index slips against the
    host's cell edges by up to ~0.3 cell at the far end; the builder's margin must cover it, so that points AT
    the knots and one ulp either side take exactly searchsortedlast's piece.  Order 2 shows a wrong piece in
    the value, the highest derivative shows it at any order (piecewise constant, jumps at every knot)."""
    brf = synth.breakpoints(10, nb).astype(np.float32)
    t = oracle.augment_knots(brf, k).astype(np.float32)
    N = t.size - k
    c = (2 * synth.u_numpy(4, 0, N) - 1).astype(np.float32)
    x = np.concatenate([synth.u_numpy(11, 0, 200000).astype(np.float32), brf, np.nextafter(brf, np.float32(-1)),
                        np.nextafter(brf, np.float32(2))]).astype(np.float64)
    E = emul(0, k, t, N, c, f32=True)
    info, _ = E.info()
    assert info["cell_valid"] and info["cell_global"]
    top = k - 1 if nb < 50000 else min(k - 1, 2)       # the reference's own differencing overflows Float32 beyond
    for d in sorted({0, top}):
        ref = _ref(oracle, 0, k, t, c, x, d, "f32")
        got, rt = E.eval(x, d, 0)
        sc = np.max(np.abs(ref))
        bad = np.abs(got - ref) > 1e-5 * sc
        assert not bad.any(), (d, int(bad.sum()), x[bad][:3], got[bad][:3], ref[bad][:3], rt[bad][:3])


def test_table_builder_is_fast_enough(emul, oracle, synth):
    """C5's Float32 sweep table (k = 8, 10^5 intervals, 10^6 cells) is built in well under a second per core
    group; the builder runs on the host inside the first bsk_spline_eval call."""
    import time
    brf = synth.breakpoints(10, 99996).astype(np.float32)
    t = oracle.augment_knots(brf, 8).astype(np.float32)
    c = (2 * synth.u_numpy(4, 0, t.size - 8) - 1).astype(np.float32)
    t0 = time.perf_counter()
    E = emul(0, 8, t, t.size - 8, c, f32=True)
    dt = time.perf_counter() - t0
    assert E.info()[0]["ncell"] >= 10 ** 6 and dt < 20.0
