"""GPU tests at BASELINE.json's full sizes (configs C2..C5): size-independent properties on the
whole problem + comparison with the oracle on subsets the oracle finishes in seconds."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL64 = 1e-13
TOL32 = 1e-5


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def per_element(got, ref, floor):
    """Worst per-element relative error, |got - ref| / max(|ref|, floor): the norm-wise bar cannot see an error
    next to a zero of S; this one can, down to |S| = floor ("per-element relative check wherever
    abs(ref) is not tiny").  floor = 2^-5 of the coefficient (or output) magnitude."""
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), floor)))


def knot_cell_mask(S, x):
    """Points whose uniform cell holds a knot: the records that carry a flag (and, for even orders, the side
    index in bits 1..3 of four coefficients — csrc/bsk_eval_cell.cu)."""
    ti = S.table_info()
    assert ti["cells"] > 0
    t = np.unique(S.knots())[1:-1]
    has_knot = np.zeros(ti["cells"] + 1, dtype=bool)
    has_knot[np.clip(np.floor((t - ti["x0"]) / ti["cell_width"]).astype(np.int64), 0, ti["cells"])] = True
    pc = np.clip(np.floor((x - ti["x0"]) * ti["scale"]).astype(np.int64), 0, ti["cells"] - 1)
    return has_knot[pc]


@pytest.fixture(scope="module")
def synth():
    from bsplinekit_b200 import synth as s
    return s


def test_c2_full_size_eval(bk, oracle, synth):
    """C2: k = 4, 1000 non-uniform breakpoints, value + 1st derivative at 2^28 points."""
    import torch
    n = 1 << 28
    k = 4
    B = bk.BSplineBasis(k, synth.breakpoints(3, 1000))
    c1 = 2 * synth.u_numpy(4, 0, len(B)) - 1
    c2 = 2 * synth.u_numpy(40, 0, len(B)) - 1
    S1, S2, S12 = bk.Spline(B, c1), bk.Spline(B, c2), bk.Spline(B, 0.5 * c1 - 2.0 * c2)
    xd = synth.u_torch(5, 0, n)
    v1, d1 = S1.evaluate(xd, (0, 1))
    # oracle on strided subsets of the whole range
    idx = torch.arange(0, n, 4099, device="cuda")
    xs = xd[idx].cpu().numpy()
    ref = oracle.spline_eval(0, k, B.knots(), c1, xs)
    kd, td, cd = oracle.spline_derivative(0, k, B.knots(), c1, 1)
    refd = oracle.spline_eval(0, kd, td, cd, xs)
    assert relerr(v1[idx].cpu().numpy(), ref) < TOL64
    assert relerr(d1[idx].cpu().numpy(), refd) < TOL64
    # per element, and separately for the points in knot cells (bit-encoded records)
    ti = S1.table_info()
    assert ti["auto_uses_cells"] == 1 and ti["intervals_deboor"] == 0 and ti["cells_interval_table"] <= 2
    flagged = knot_cell_mask(S1, xs)
    assert 0.1 < flagged.mean() < 0.3
    f0, f1 = np.max(np.abs(c1)) / 32, np.max(np.abs(refd)) / 32
    assert per_element(v1[idx].cpu().numpy(), ref, f0) < TOL64
    assert per_element(d1[idx].cpu().numpy(), refd, f1) < TOL64
    assert per_element(v1[idx].cpu().numpy()[flagged], ref[flagged], f0) < TOL64
    assert per_element(d1[idx].cpu().numpy()[flagged], refd[flagged], f1) < TOL64
    # linearity over ALL points: S(0.5 c1 - 2 c2) == 0.5 S(c1) - 2 S(c2)
    v2, d2 = S2.evaluate(xd, (0, 1))
    v12, d12 = S12.evaluate(xd, (0, 1))
    err_v = (v12 - (0.5 * v1 - 2.0 * v2)).abs().max().item()
    err_d = (d12 - (0.5 * d1 - 2.0 * d2)).abs().max().item()
    assert err_v < 1e-14 * 4 and err_d < 1e-13 * d12.abs().max().item()
    # range: |S| <= max|c| (convex combinations), exact zeros nowhere inside the domain by accident
    assert v1.abs().max().item() <= np.abs(c1).max() * (1 + 1e-14)
    # fused == separate calls
    v_only = S1.evaluate(xd, (0,))[0]
    d_only = S1.evaluate(xd, (1,))[0]
    assert torch.equal(v_only, v1) and torch.equal(d_only, d1)
    # bit-faithful kernel on the full stream agrees with the fast one
    vb = S1.evaluate(xd, (0,), algo=bk.EVAL_DEBOOR)[0]
    assert ((vb - v1).abs().max() / v1.abs().max()).item() < TOL64
    assert np.array_equal(vb[idx].cpu().numpy(), ref)


def test_c3_full_size_batched_solve(bk, oracle, synth):
    """C3: 65 536 data sets sharing 4 096 non-uniform xdata, k = 6: one LU + multi-RHS solve."""
    import torch
    n, k, nrhs = 4096, 6, 65536
    xdata = synth.breakpoints(6, n)
    t = bk.make_knots(xdata, k)
    B = bk.BSplineBasis(k, t, augment=False)
    itp = bk.SplineInterpolation(B, xdata)                       # default: partitioned lu! (within 1e-13)
    itp_exact = bk.SplineInterpolation(B, xdata, lu_mode=bk.LU_EXACT)     # BANFAC order: the reference's bits
    assert itp.C.halo > 0 and itp_exact.C.halo == itp.C.halo
    Y0 = (2 * synth.u_torch(7, 0, n * nrhs) - 1).reshape(n, nrhs)
    band, nout = oracle.collocation_matrix(0, k, t, xdata)
    assert nout == 0
    lu = band.copy(order="F")
    assert oracle.banded_lu(lu) == 0
    cols = np.r_[0:256, 32768:33024, nrhs - 256:nrhs]
    ref = oracle.banded_solve(lu, np.ascontiguousarray(Y0[:, cols].cpu().numpy()))
    bw = k - 2
    bd = torch.from_numpy(np.ascontiguousarray(band)).cuda()       # (2bw+1, n): band[u+i-j, j]
    assert np.array_equal(itp_exact.C.data(), lu)
    f_fast = itp.C.data()
    assert np.max(np.abs(f_fast - lu) / np.max(np.abs(lu), axis=0)) < 1e-13
    for algo, exact in ((bk.SOLVE_WINDOWED, False), (bk.SOLVE_TWOPASS, True), (bk.SOLVE_AUTO, False)):
        Y = Y0.clone()
        (itp_exact if exact else itp).C.ldiv_(Y, algo=algo)
        got = Y[:, cols].cpu().numpy()
        if exact:
            assert np.array_equal(got, ref)
        else:
            assert relerr(got, ref) < TOL64
            assert np.max(np.max(np.abs(got - ref), axis=0) / np.max(np.abs(ref), axis=0)) < TOL64
        # residual of the WHOLE batch: C * c - y, with C applied diagonal by diagonal
        R = -Y0.clone()
        for d in range(-bw, bw + 1):          # entries C[i, i+d] = band[bw - d, i + d]
            lo, hi = max(0, -d), min(n, n - d)
            R[lo:hi] += bd[bw - d, lo + d:hi + d, None] * Y[lo + d:hi + d]
        assert R.abs().max().item() < 5e-13
        del R, Y


def test_c4_periodic_full(bk, oracle, synth):
    """C4: PeriodicBSplineBasis, k = 4, 10^4 breakpoints: cyclic solve + evaluation."""
    import torch
    n = 10000
    br = synth.breakpoints(8, n + 1)
    B = bk.PeriodicBSplineBasis(4, br)
    x = bk.collocation_points(B)
    M = bk.collocation_matrix(B, x)
    a, b, c, nout = oracle.collocation_cyclic(4, br, x)
    assert nout == 0 and np.array_equal(M.a, a) and np.array_equal(M.b, b) and np.array_equal(M.c, c)
    # (i) solve: 1 RHS (latency case) and 16 384 interleaved RHS
    d1 = 2 * synth.u_numpy(9, 0, n) - 1
    got1 = d1.copy().reshape(n, 1)
    M.ldiv_(got1)
    assert relerr(got1.ravel(), oracle.cyclic_solve(a, b, c, d1.copy())) < TOL64
    nrhs = 16384
    D0 = (2 * synth.u_torch(9, 0, n * nrhs) - 1).reshape(n, nrhs)
    D = D0.clone()
    M.ldiv_(D)
    cols = np.r_[0:128, nrhs - 128:nrhs]
    ref = oracle.cyclic_solve(a, b, c, np.ascontiguousarray(D0[:, cols].cpu().numpy()))
    assert relerr(D[:, cols].cpu().numpy(), ref) < TOL64
    ad, bd, cd = (torch.from_numpy(v).cuda() for v in (a, b, c))
    R = bd[:, None] * D + ad[:, None] * torch.roll(D, 1, 0) + cd[:, None] * torch.roll(D, -1, 0) - D0
    assert R.abs().max().item() < 1e-12
    # (ii) evaluation of the interpolant of data set 0 at 2^28 points + wrap subset in [-2L, 3L)
    coefs = D[:, 0].cpu().numpy()
    S = bk.Spline(B, coefs)
    assert relerr(S(x, algo=bk.EVAL_DEBOOR), D0[:, 0].cpu().numpy()) < 1e-12        # I.(xs) ≈ ys
    xd = synth.u_torch(9, 1 << 40, 1 << 28)
    v = S.evaluate(xd, (0,))[0]
    idx = torch.arange(0, 1 << 28, 8191, device="cuda")
    ref = oracle.spline_eval(1, 4, br, coefs, xd[idx].cpu().numpy())
    assert relerr(v[idx].cpu().numpy(), ref) < TOL64
    assert per_element(v[idx].cpu().numpy(), ref, np.max(np.abs(coefs)) / 32) < TOL64
    ti = S.table_info()
    assert ti["cells_in_global"] == 1 and ti["intervals_deboor"] == 0
    # derivative outputs of a periodic spline come from the tables of Derivative(1) * S itself
    # (Splines/periodic.jl:76-117): same bar, incl. the points next to the period boundary
    xb = np.concatenate([xd[idx].cpu().numpy()[:20000], 1e-5 * synth.u_numpy(9, 7, 4000), 1 - 1e-5 * synth.u_numpy(9, 8, 4000)])
    kd, td, cd = oracle.spline_derivative(1, 4, br, coefs, 1)
    refd = oracle.spline_eval(1, kd, td, cd, xb)
    gotv, gotd = S.evaluate(torch.from_numpy(xb).cuda(), (0, 1))
    assert relerr(gotd.cpu().numpy(), refd) < TOL64
    assert relerr(gotv.cpu().numpy(), oracle.spline_eval(1, 4, br, coefs, xb)) < TOL64
    xw = -2.0 + 5.0 * synth.u_numpy(9, 1 << 41, 1 << 20)
    refw = oracle.spline_eval(1, 4, br, coefs, xw)
    assert relerr(S(xw), refw) < TOL64
    assert np.array_equal(S(xw, algo=bk.EVAL_DEBOOR), refw)


def test_c5_recombined_robin_full(bk, oracle, synth):
    """C5: RecombinedBSplineBasis, Robin D(0) + 3 D(1), k = 8, collocation_matrix with Derivative(2)
    at 10^5 points + Float32 sweeps."""
    import torch
    k = 8
    br = synth.breakpoints(10, 99996)
    B = bk.BSplineBasis(k, br)
    R = bk.RecombinedBSplineBasis(B, bk.Derivative(0) + 3 * bk.Derivative(1))
    assert len(B) == 100002 and len(R) == 100000
    rc = oracle.robin_corners(k, B.knots(), [1.0, 3.0])
    assert np.array_equal(R.ul, rc.ul) and np.array_equal(R.lr, rc.lr)
    x = bk.collocation_points(R)
    assert np.array_equal(x, oracle.greville(2, k, B.knots(), cl=1, N=len(R)))
    ref, nout_ref = oracle.collocation_matrix(2, k, B.knots(), x, nderiv=2, recomb=rc)
    Cm, nout = bk.collocation_matrix(R, x, bk.Derivative(2), return_outside=True)
    assert nout == nout_ref == 0
    got = Cm.data()
    assert got.shape == (13, 100000) and np.array_equal(got, ref)
    assert np.count_nonzero(got) <= 8 * 100000
    # evaluate_all(R, x, Derivative(n), Float32), n = 0..2 (the `T` reading of "Float32 sweep")
    for nd in (0, 1, 2):
        i_ref, bs_ref = oracle.evaluate_all(0, k, B.knots(), x, nderiv=nd, vt="f32", recomb=rc)
        i, bs = bk.evaluate_all(R, x, bk.Derivative(nd), np.float32)
        assert np.array_equal(i, i_ref) and np.array_equal(bs, bs_ref)
        _, bs64 = oracle.evaluate_all(0, k, B.knots(), x, nderiv=nd, recomb=rc)
        assert relerr(bs, bs64) < TOL32
    # Float32 basis + Float32 spline evaluated at 2^28 Float32 points
    brf = br.astype(np.float32)
    Bf = bk.BSplineBasis(k, brf)
    cf = (2 * synth.u_numpy(4, 0, len(Bf)) - 1).astype(np.float32)
    Sf = bk.Spline(Bf, cf)
    xd = synth.u_torch(11, 0, 1 << 28, dtype=torch.float32)
    v = Sf.evaluate(xd, (0,))[0]
    idx = torch.arange(0, 1 << 28, 16381, device="cuda")
    xs = xd[idx].cpu().numpy()
    ref64 = oracle.spline_eval(0, k, Bf.knots().astype(np.float64), cf.astype(np.float64), xs.astype(np.float64))
    assert relerr(v[idx].cpu().numpy(), ref64) < TOL32
    ref32 = oracle.spline_eval(0, k, Bf.knots(), cf, xs, kt="f32")
    assert np.array_equal(Sf.evaluate(xd[idx].contiguous(), (0,), algo=bk.EVAL_DEBOOR)[0].cpu().numpy(), ref32)
    assert v.abs().max().item() <= 1.0 + 1e-5
    assert relerr(v[idx].cpu().numpy(), ref32) < TOL32
    # the Float32 cell index slips against the cell edges by ~0.3 cell at the far end of a 10^6-cell table:
    # points AT the knots and one ulp either side must still take searchsortedlast's piece, for the value and
    # for a derivative (the jump of D^2 S at a knot would show a wrong piece at once)
    xk = np.concatenate([brf, np.nextafter(brf, np.float32(-1)), np.nextafter(brf, np.float32(2))])
    gv, g2 = Sf.evaluate(torch.from_numpy(xk).cuda(), (0, 2))
    assert relerr(gv.cpu().numpy(), oracle.spline_eval(0, k, Bf.knots(), cf, xk, kt="f32")) < TOL32
    k2, t2, c2 = oracle.spline_derivative(0, k, Bf.knots(), cf, 2, kt="f32")
    assert relerr(g2.cpu().numpy(), oracle.spline_eval(0, k2, t2, c2, xk, kt="f32")) < TOL32


def test_c3_shape_float32_solve(bk, oracle, synth):
    """C3's shape in Float32 (n = 4096, k = 6, 65 536 data sets): BANSLV-order solve, bit-exact against the
    Float32 restatement on columns sampled over the whole block; identical columns give identical bits
    wherever they sit in the block; small backward error."""
    import torch
    n, k, M = 4096, 6, 65536
    x32 = synth.breakpoints(6, n).astype(np.float32)
    B = bk.BSplineBasis(k, bk.make_knots(x32, k), augment=False)
    itp = bk.SplineInterpolation(B, x32)
    Y = (2 * synth.u_torch(7, 0, n * M, dtype=torch.float32) - 1).reshape(n, M)
    Y[:, M - 3] = Y[:, 5]                      # the same data set in two far-apart columns
    Y[:, 40001] = Y[:, 5]
    cols = torch.arange(0, M, 509, device="cuda")
    Y0 = Y[:, cols].cpu().numpy().copy()
    Ysp = Y.clone()
    itp.interpolate_(Y, algo=bk.SOLVE_TWOPASS)
    band, _ = oracle.collocation_matrix_f32(0, k, B.knots(), x32)
    assert oracle.banded_lu(band) == 0
    ref = oracle.banded_solve(band, Y0.copy())
    assert np.array_equal(Y[:, cols].cpu().numpy(), ref)
    assert torch.equal(Y[:, 5], Y[:, M - 3]) and torch.equal(Y[:, 5], Y[:, 40001])
    # default: single pass (two columns per lane, packed Float32 FMAs): within 1e-5 of the BANSLV bits on the
    # sampled columns and on every column against the two-pass result; identical columns -> identical bits
    assert itp.C.halo > 0
    itp.interpolate_(Ysp)
    assert relerr(Ysp[:, cols].cpu().numpy(), ref) < TOL32
    assert float((Ysp - Y).abs().max() / Y.abs().max()) < TOL32
    assert torch.equal(Ysp[:, 5], Ysp[:, M - 3]) and torch.equal(Ysp[:, 5], Ysp[:, 40001])
    # backward error of the Float32 solve, measured in Float64 against the Float32 matrix itself (with 4096
    # points on [0, 1] the Float32 knots carry ~1e-4 of a knot interval, so the Float64 interpolant of the
    # same data is a different, equally valid, answer; the reference behaves the same way)
    A32, _ = oracle.collocation_matrix_f32(0, k, B.knots(), x32)
    res = oracle.banded_matvec(A32.astype(np.float64), ref[:, :16].astype(np.float64)) - Y0[:, :16]
    assert np.max(np.abs(res)) < 1e-5 * max(1.0, float(np.max(np.abs(ref[:, :16]))))


def test_order6_full_size_eval(bk, oracle, synth):
    """Order 6 at 2^28 points: shared-memory cell table with jump-encoded knot cells (value + derivative)."""
    import torch
    n, k = 1 << 28, 6
    B = bk.BSplineBasis(k, synth.breakpoints(3, 1000))
    c = 2 * synth.u_numpy(4, 0, len(B)) - 1
    S = bk.Spline(B, c)
    xd = synth.u_torch(5, 0, n)
    v, d = S.evaluate(xd, (0, 1))
    idx = torch.arange(0, n, 4099, device="cuda")
    xs = xd[idx].cpu().numpy()
    ref = oracle.spline_eval(0, k, B.knots(), c, xs)
    assert relerr(v[idx].cpu().numpy(), ref) < TOL64
    kd, td, cd = oracle.spline_derivative(0, k, B.knots(), c, 1)
    refd = oracle.spline_eval(0, kd, td, cd, xs)
    assert relerr(d[idx].cpu().numpy(), refd) < TOL64
    flagged = knot_cell_mask(S, xs)
    assert flagged.mean() > 0.15
    assert per_element(v[idx].cpu().numpy(), ref, np.max(np.abs(c)) / 32) < TOL64
    assert per_element(d[idx].cpu().numpy(), refd, np.max(np.abs(refd)) / 32) < TOL64
    assert per_element(v[idx].cpu().numpy()[flagged], ref[flagged], np.max(np.abs(c)) / 32) < TOL64
    assert v.abs().max().item() <= np.abs(c).max() * (1 + 1e-14)
    assert torch.equal(S.evaluate(xd, (0,))[0], v)


def test_more_than_2_31_elements(bk, oracle, synth):
    """Maximum sizes: point streams and right-hand-side batches of more than 2^31 elements (every element index in
    the kernels is 64-bit).  The one big call must equal, bit for bit, the same work split into calls below 2^31
    elements, and windows at the start, across element 2^31 and at the end are checked against the oracle."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 80e9:
        pytest.skip("needs 80 GB of free device memory")
    n = (1 << 31) + 4096 + 3
    br = synth.breakpoints(3, 1000).astype(np.float32)
    B = bk.BSplineBasis(4, br)
    c = (2 * synth.u_numpy(4, 0, len(B)) - 1).astype(np.float32)
    S = bk.Spline(B, c)
    x = synth.u_torch(31, 0, n, dtype=torch.float32)
    out = torch.empty_like(x)
    S.evaluate(x, (0,), out=[out])
    idx, zone = bk.find_knot_interval(B, x)
    windows = [(0, 4096), ((1 << 31) - 2048, (1 << 31) + 2048), (n - 4099, n)]
    for lo, hi in windows:
        xh = x[lo:hi].cpu().numpy()
        assert relerr(out[lo:hi].cpu().numpy(), oracle.spline_eval(0, 4, B.knots(), c, xh, kt="f32")) < TOL32
        ri, rz = oracle.find_knot_interval(0, 4, B.knots(), xh, kt="f32")
        assert np.array_equal(idx[lo:hi].cpu().numpy(), ri) and np.array_equal(zone[lo:hi].cpu().numpy(), rz)
    step = (1 << 29) + 4                                       # 16-byte aligned pieces, the last one ragged
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        assert torch.equal(S.evaluate(x[lo:hi], (0,))[0], out[lo:hi]), lo
        pi, pz = bk.find_knot_interval(B, x[lo:hi])
        assert torch.equal(pi, idx[lo:hi]) and torch.equal(pz, zone[lo:hi]), lo
    del x, out, idx, zone, pi, pz
    torch.cuda.empty_cache()
    # evaluate_all: (2^28 + 1027) points x 8 values, value + first derivative, and the headline call (value +
    # derivative of a Float64 cubic) on 2^31 + 6 points
    n8 = (1 << 28) + 1027
    B8 = bk.BSplineBasis(8, synth.breakpoints(5, 3000))
    x8 = synth.u_torch(33, 0, n8)
    for nd in (0, 1):
        i8, bs8 = bk.evaluate_all(B8, x8, bk.Derivative(nd))
        assert bs8.numel() > (1 << 31)
        for lo in (0, (1 << 28) - 1500, n8 - 3000):
            ri, rbs = oracle.evaluate_all(0, 8, B8.knots(), x8[lo:lo + 3000].cpu().numpy(), nderiv=nd)
            assert np.array_equal(i8[lo:lo + 3000].cpu().numpy(), ri)
            assert np.array_equal(bs8[lo:lo + 3000].cpu().numpy().reshape(-1, 8), rbs)
        for lo in range(0, n8, (1 << 27) + 2):
            hi = min(n8, lo + (1 << 27) + 2)
            pi, pb = bk.evaluate_all(B8, x8[lo:hi], bk.Derivative(nd))
            assert torch.equal(pi, i8[lo:hi]) and torch.equal(pb.reshape(-1, 8), bs8.reshape(-1, 8)[lo:hi]), (nd, lo)
        del i8, bs8, pi, pb
    del x8
    torch.cuda.empty_cache()
    nd_ = (1 << 31) + 6
    B4 = bk.BSplineBasis(4, synth.breakpoints(3, 1000))
    c4 = 2 * synth.u_numpy(4, 0, len(B4)) - 1
    S4 = bk.Spline(B4, c4)
    xd = synth.u_torch(35, 0, nd_)
    o0, o1 = torch.empty_like(xd), torch.empty_like(xd)
    S4.evaluate(xd, (0, 1), out=[o0, o1])
    kd, td, cd = oracle.spline_derivative(0, 4, B4.knots(), c4, 1)
    for lo, hi in ((0, 4096), ((1 << 31) - 2048, (1 << 31) + 6)):
        xh = xd[lo:hi].cpu().numpy()
        assert relerr(o0[lo:hi].cpu().numpy(), oracle.spline_eval(0, 4, B4.knots(), c4, xh)) < TOL64
        assert relerr(o1[lo:hi].cpu().numpy(), oracle.spline_eval(0, kd, td, cd, xh)) < TOL64
    for lo in range(0, nd_, (1 << 30) + 2):
        hi = min(nd_, lo + (1 << 30) + 2)
        p0, p1 = S4.evaluate(xd[lo:hi], (0, 1))
        assert torch.equal(p0, o0[lo:hi]) and torch.equal(p1, o1[lo:hi]), lo
    del xd, o0, o1, p0, p1
    torch.cuda.empty_cache()
    # right-hand sides: 2048 x (2^20 + 64) elements, Float32 (packed single-pass kernel) and Float64 (C3's kernel)
    rows, k, nrhs = 2048, 6, (1 << 20) + 64
    for dt, tdt, tol in ((np.float32, torch.float32, TOL32), (np.float64, torch.float64, TOL64)):
        xs = synth.breakpoints(6, rows).astype(dt)
        Bi = bk.BSplineBasis(k, bk.make_knots(xs, k), augment=False)
        Cm = bk.collocation_matrix(Bi, xs)
        band = Cm.data().copy()
        Cm.lu_()
        Y0 = (2 * synth.u_torch(7, 0, rows * nrhs, dtype=tdt) - 1).reshape(rows, nrhs)
        Y = Y0.clone()
        Cm.ldiv_(Y)
        cols = [0, 1, 12345, (1 << 19) + 1, nrhs - 2, nrhs - 1]
        lu = np.asfortranarray(band.astype(np.float64))
        assert oracle.banded_lu(lu) == 0
        ref = oracle.banded_solve(lu, Y0[:, cols].cpu().numpy().astype(np.float64))
        assert relerr(Y[:, cols].cpu().numpy(), ref) < tol
        blk = 1 << 18
        for c0 in range(0, nrhs, blk):
            part = Y0[:, c0:c0 + blk].contiguous()
            Cm.ldiv_(part)
            assert torch.equal(part, Y[:, c0:c0 + blk]), (dt, c0)
        del Y0, Y, part
        torch.cuda.empty_cache()
