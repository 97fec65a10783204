"""Two devices driven from ONE process, one host thread per device, through the C ABI only (no torch): what
INTEGRATION.md §4 promises a Julia process with `Threads.@spawn` per GPU — handles are per device, every entry
point selects its handle's device, the error text is thread-local.  Skipped on a single-GPU box."""
import ctypes as C
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_on_device(lib, check, dev, k, knots, coefs, x, res, xdata, Yblock):
    try:
        vp = C.c_void_p
        basis, spline, stream = vp(), vp(), vp()
        check(lib.bsk_stream_create(dev, C.byref(stream)))
        check(lib.bsk_basis_create(0, k, 1, knots.ctypes.data_as(vp), knots.size, None, dev, C.byref(basis)))
        check(lib.bsk_spline_create(basis, coefs.ctypes.data_as(vp), coefs.size, C.byref(spline)))
        n = x.size
        dx, dv, dd = vp(), vp(), vp()
        for p in (dx, dv, dd):
            check(lib.bsk_malloc(dev, 8 * n, C.byref(p)))
        check(lib.bsk_memcpy_h2d(dev, dx, x.ctypes.data_as(vp), 8 * n, stream))
        derivs = (C.c_int32 * 2)(0, 1)
        outs = (vp * 2)(dv, dd)
        for _ in range(3):
            check(lib.bsk_spline_eval(spline, dx, n, 2, derivs, outs, 0, stream))
        v, d = np.empty(n), np.empty(n)
        check(lib.bsk_memcpy_d2h(dev, v.ctypes.data_as(vp), dv, 8 * n, stream))
        check(lib.bsk_memcpy_d2h(dev, d.ctypes.data_as(vp), dd, 8 * n, stream))
        # batched interpolation on the same device: assemble, factor, solve this device's block of columns
        nx, M = xdata.size, Yblock.shape[1]
        kk = 6
        import bsplinekit_b200 as bk
        tk = bk.make_knots(xdata, kk)
        ib = vp()
        check(lib.bsk_basis_create(0, kk, 1, tk.ctypes.data_as(vp), tk.size, None, dev, C.byref(ib)))
        dxd, dband, dY = vp(), vp(), vp()
        check(lib.bsk_malloc(dev, 8 * nx, C.byref(dxd)))
        check(lib.bsk_malloc(dev, 8 * nx * (2 * kk - 3), C.byref(dband)))
        check(lib.bsk_malloc(dev, 8 * nx * M, C.byref(dY)))
        check(lib.bsk_memcpy_h2d(dev, dxd, xdata.ctypes.data_as(vp), 8 * nx, stream))
        check(lib.bsk_stream_sync(dev, stream))
        nout = C.c_int64(0)
        check(lib.bsk_collocation_matrix(ib, dxd, nx, 0, float(np.finfo(np.float64).eps), dband, C.byref(nout), stream))
        band = vp()
        check(lib.bsk_banded_create(dband, nx, kk - 2, kk - 2, dev, C.byref(band)))
        info = C.c_int64(0)
        check(lib.bsk_banded_lu_mode(band, 1, C.byref(info)), info.value)
        Yc = np.ascontiguousarray(Yblock)
        check(lib.bsk_memcpy_h2d(dev, dY, Yc.ctypes.data_as(vp), 8 * nx * M, stream))
        check(lib.bsk_banded_solve(band, dY, M, 0, stream))
        sol = np.empty_like(Yc)
        check(lib.bsk_memcpy_d2h(dev, sol.ctypes.data_as(vp), dY, 8 * nx * M, stream))
        check(lib.bsk_stream_sync(dev, stream))
        # a deliberate error on this thread must not leak its text to the other one
        st = lib.bsk_spline_eval(spline, dx, n, 9, derivs, outs, 0, stream)
        assert st == -1
        res[dev] = (v, d, sol, tk)
        for p in (dx, dv, dd, dxd, dband, dY):
            check(lib.bsk_free(dev, p))
        check(lib.bsk_banded_destroy(band))
        check(lib.bsk_spline_destroy(spline))
        check(lib.bsk_basis_destroy(basis))
        check(lib.bsk_basis_destroy(ib))
        check(lib.bsk_stream_destroy(dev, stream))
    except Exception as exc:          # surfaced by the main thread
        res[dev] = exc


def test_two_devices_one_process(bk, oracle):
    from bsplinekit_b200 import _lib, synth
    if bk.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    lib, check = _lib.lib, _lib.check
    k = 4
    B = bk.BSplineBasis(k, synth.breakpoints(3, 1000))          # only for the knot vector (device 0 handle unused)
    knots = B.knots()
    coefs = 2 * synth.u_numpy(4, 0, knots.size - k) - 1
    n = 1 << 22
    xs = [synth.u_numpy(5, d * n, n) for d in range(2)]         # each device evaluates its own slice of the stream
    nx, M = 2048, 256
    xdata = synth.breakpoints(6, nx)
    Y = (2 * synth.u_numpy(7, 0, nx * 2 * M) - 1).reshape(nx, 2 * M)
    res = {}
    th = [threading.Thread(target=_run_on_device, args=(lib, check, d, k, knots, coefs, xs[d], res, xdata,
                                                        Y[:, d * M:(d + 1) * M])) for d in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    kd, td, cd = oracle.spline_derivative(0, k, knots, coefs, 1)
    for d in range(2):
        if isinstance(res[d], Exception):
            raise res[d]
        v, dv, sol, tk = res[d]
        ref = oracle.spline_eval(0, k, knots, coefs, xs[d])
        refd = oracle.spline_eval(0, kd, td, cd, xs[d])
        assert np.max(np.abs(v - ref)) < 1e-13 * np.max(np.abs(ref))
        assert np.max(np.abs(dv - refd)) < 1e-13 * np.max(np.abs(refd))
        band, _ = oracle.collocation_matrix(0, 6, tk, xdata)
        assert oracle.banded_lu(band) == 0
        want = oracle.banded_solve(band, np.ascontiguousarray(Y[:, d * M:(d + 1) * M]))
        assert np.max(np.abs(sol - want)) < 1e-13 * np.max(np.abs(want))
