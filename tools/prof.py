#!/usr/bin/env python
"""Small driver for ncu captures: runs a few launches of one hot kernel (no timing claims).
  python tools/prof.py eval   [--points N] [--algo auto|deboor]
  python tools/prof.py solve  [--nrhs M] [--algo windowed|twopass]
  python tools/prof.py cyclic / evalc4 / evalf32
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "bsplinekit.jl_b200"))

import numpy as np
import torch
import bsplinekit_b200 as bk
from bsplinekit_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("what")
ap.add_argument("--points", type=int, default=1 << 28)
ap.add_argument("--nrhs", type=int, default=65536)
ap.add_argument("--algo", default="auto")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--order", type=int, default=4)
ap.add_argument("--breaks", type=int, default=1000)
args = ap.parse_args()


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if args.what == "eval":
    B = bk.BSplineBasis(args.order, synth.breakpoints(3, args.breaks))
    S = bk.Spline(B, 2 * synth.u_numpy(4, 0, len(B)) - 1)
    xd = synth.u_torch(5, 0, args.points)
    outs = [torch.empty_like(xd), torch.empty_like(xd)]
    algo = {"auto": bk.EVAL_AUTO, "deboor": bk.EVAL_DEBOOR, "pp": bk.EVAL_PP}[args.algo]
    ms = timed(lambda: S.evaluate(xd, (0, 1), algo=algo, out=outs), args.reps)
    print("eval %s: %.3f ms  %.1f Gpt/s  %.1f GB/s" % (args.algo, ms, args.points / ms / 1e6, 24 * args.points / ms / 1e6))
elif args.what == "gather":
    # the access pattern of the global-table kernels alone (bsk_probe_gather), every variant
    import ctypes
    from bsplinekit_b200 import _lib as _bl
    xd = synth.u_torch(9, 0, args.points)
    out = torch.empty_like(xd)
    for nrec in (320000, 1 << 20):
        tab = synth.u_torch(12, 0, 4 * nrec)
        ref = None
        for variant in [int(v) for v in os.environ.get("VARIANTS", "0,1,2,3,4").split(",")]:
            call = lambda: _bl.check(_bl.lib.bsk_probe_gather(ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                                              args.points, ctypes.c_void_p(tab.data_ptr()), nrec, variant, None))
            out.zero_()
            ms = timed(call, args.reps)
            chk = out[::4099].clone()
            if ref is None:
                ref = chk
            print("gather nrec=%d variant %2d: %.3f ms  %.1f Gpt/s  frac of 16 B/pt line %.3f  %s" % (
                nrec, variant, ms, args.points / ms / 1e6, 16 * args.points / ms / 1e6 / 6547.2,
                "ok" if torch.equal(chk, ref) else "MISMATCH"), flush=True)
elif args.what == "evalc4":
    br = synth.breakpoints(8, 10001)
    B = bk.PeriodicBSplineBasis(4, br)
    S = bk.Spline(B, 2 * synth.u_numpy(4, 0, len(B)) - 1)
    xd = synth.u_torch(9, 0, args.points)
    outs = [torch.empty_like(xd)]
    algo = {"auto": bk.EVAL_AUTO, "deboor": bk.EVAL_DEBOOR, "pp": bk.EVAL_PP}[args.algo]
    ms = timed(lambda: S.evaluate(xd, (0,), algo=algo, out=outs), args.reps)
    print("evalc4 %s: %.3f ms  %.1f Gpt/s  %.1f GB/s" % (args.algo, ms, args.points / ms / 1e6, 16 * args.points / ms / 1e6))
elif args.what == "evalf32":
    br = synth.breakpoints(10, 99996).astype(np.float32)
    B = bk.BSplineBasis(8, br)
    S = bk.Spline(B, (2 * synth.u_numpy(4, 0, len(B)) - 1).astype(np.float32))
    xd = synth.u_torch(11, 0, args.points, dtype=torch.float32)
    outs = [torch.empty_like(xd)]
    algo = {"auto": bk.EVAL_AUTO, "deboor": bk.EVAL_DEBOOR, "pp": bk.EVAL_PP}[args.algo]
    ms = timed(lambda: S.evaluate(xd, (0,), algo=algo, out=outs), args.reps)
    print("evalf32 %s: %.3f ms  %.1f Gpt/s  %.1f GB/s" % (args.algo, ms, args.points / ms / 1e6, 8 * args.points / ms / 1e6))
elif args.what == "solve":
    n, k = 4096, 6
    xdata = synth.breakpoints(6, n)
    Bi = bk.BSplineBasis(k, bk.make_knots(xdata, k), augment=False)
    itp = bk.SplineInterpolation(Bi, xdata)
    Y0 = (2 * synth.u_torch(7, 0, n * args.nrhs) - 1).reshape(n, args.nrhs)
    Y = Y0.clone()
    algo = {"auto": bk.SOLVE_AUTO, "windowed": bk.SOLVE_WINDOWED, "twopass": bk.SOLVE_TWOPASS}[args.algo]
    tot = 0.0
    for it in range(args.reps + 1):
        Y.copy_(Y0)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); itp.C.ldiv_(Y, algo=algo); e1.record(); torch.cuda.synchronize()
        if it > 0:
            tot += e0.elapsed_time(e1)
    ms = tot / args.reps
    print("solve %s halo=%d: %.3f ms  %.2f M RHS/s  %.1f GB/s (16 B/elem)" % (
        args.algo, itp.C.halo, ms, args.nrhs / ms / 1e3, 16 * n * args.nrhs / ms / 1e6))
elif args.what == "cyclic":
    n = 10000
    br = synth.breakpoints(8, n + 1)
    B = bk.PeriodicBSplineBasis(4, br)
    M = bk.collocation_matrix(B, bk.collocation_points(B))
    nrhs = 16384
    Y0 = (2 * synth.u_torch(9, 0, n * nrhs) - 1).reshape(n, nrhs)
    Y = Y0.clone()
    tot = 0.0
    for it in range(args.reps + 1):
        Y.copy_(Y0)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); M.ldiv_(Y); e1.record(); torch.cuda.synchronize()
        if it > 0:
            tot += e0.elapsed_time(e1)
    ms = tot / args.reps
    print("cyclic: %.3f ms  %.2f M RHS/s  %.1f GB/s (16 B/elem)" % (ms, nrhs / ms / 1e3, 16 * n * nrhs / ms / 1e6))
if args.what == "evalall":
    import ctypes as C
    from bsplinekit_b200 import _lib
    k = args.order
    B = bk.BSplineBasis(k, synth.breakpoints(3, args.breaks))
    xd = synth.u_torch(5, 0, args.points)
    idx = torch.empty(args.points, dtype=torch.int64, device="cuda")
    bs = torch.empty(args.points * k, dtype=torch.float64, device="cuda")
    call = lambda: _lib.check(_lib.lib.bsk_evaluate_all(B._handle, C.c_void_p(xd.data_ptr()), args.points, 0, 1, None,
                                                        C.c_void_p(idx.data_ptr()), C.c_void_p(bs.data_ptr()), None))
    ms = timed(call, args.reps)
    print("evaluate_all k=%d: %.3f ms  %.1f Gpt/s  %.1f GB/s (%d B/pt)" % (k, ms, args.points / ms / 1e6,
                                                                          (16 + 8 * k) * args.points / ms / 1e6, 16 + 8 * k))
if args.what == "probefma":
    import ctypes as C
    from bsplinekit_b200 import _lib
    scratch = torch.zeros(16, dtype=torch.float64, device="cuda")
    for name, dt_ in (("fp64", 1), ("fp32", 0)):
        cnt = C.c_double(0)
        call = lambda: _lib.check(_lib.lib.bsk_probe_fma(dt_, 1 << 14, C.c_void_p(scratch.data_ptr()), C.byref(cnt), None))
        ms = timed(call, args.reps)
        print("%s FMA probe: %.3f ms, %.2f TFLOP/s" % (name, ms, 2 * cnt.value / ms / 1e9))
if args.what == "solve32w":
    xd32 = synth.breakpoints(6, 4096).astype(np.float32)
    B32 = bk.BSplineBasis(6, bk.make_knots(xd32, 6), augment=False)
    itp32 = bk.SplineInterpolation(B32, xd32)
    Y32 = (2 * synth.u_torch(7, 0, 4096 * args.nrhs, dtype=torch.float32) - 1).reshape(4096, args.nrhs)
    W32 = torch.empty_like(Y32)
    for name, al in (("windowed", bk.SOLVE_WINDOWED), ("twopass", bk.SOLVE_TWOPASS)):
        tot = 0.0
        for it in range(args.reps + 1):
            W32.copy_(Y32)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); itp32.C.ldiv_(W32, algo=al); e1.record()
            torch.cuda.synchronize()
            if it:
                tot += e0.elapsed_time(e1)
        ms = tot / args.reps
        print("solve32 %s halo=%d: %.3f ms  %.1f M RHS/s  %.1f GB/s (8 B/element)" % (name, itp32.C.halo, ms, args.nrhs / ms / 1e3, 8 * 4096 * args.nrhs / ms / 1e6))
if args.what == "knot":
    B = bk.BSplineBasis(args.order, synth.breakpoints(3, args.breaks))
    xd = synth.u_torch(5, 0, args.points)
    ms = timed(lambda: bk.find_knot_interval(B, xd), args.reps)
    print("find_knot_interval: %.3f ms (incl. output allocation)  %.1f Gpt/s" % (ms, args.points / ms / 1e6))
if args.what == "multi":
    n, k, M = 4096, 6, args.nrhs
    xdata = synth.breakpoints(6, n)
    Bi = bk.BSplineBasis(k, bk.make_knots(xdata, k), augment=False)
    Cc = (2 * synth.u_torch(7, 0, n * M) - 1).reshape(n, M)
    SB = bk.SplineBatch(Bi, Cc)
    for label, xs in (("sorted", torch.from_numpy(np.sort(synth.u_numpy(5, 0, 4096))).cuda()),
                      ("random", synth.u_torch(5, 0, 4096))):
        ms = timed(lambda: SB(xs), args.reps)
        print("multi %s: %d splines x %d points: %.3f ms  %.1f GB/s (out + coefs once)" % (
            label, M, xs.numel(), ms, (8.0 * xs.numel() * M + 8.0 * n * M) / ms / 1e6))
if args.what == "fused":
    # du = F \ (L u): C3-shaped collocation time step, fused vs mul! + ldiv!
    n, k = 4096, 6
    xdata = synth.breakpoints(6, n)
    Bi = bk.BSplineBasis(k, bk.make_knots(xdata, k), augment=False)
    F = bk.collocation_matrix(Bi, xdata).lu_()
    L = bk.collocation_matrix(Bi, xdata, bk.Derivative(2))
    U = (2 * synth.u_torch(7, 0, n * args.nrhs) - 1).reshape(n, args.nrhs)
    out = torch.empty_like(U)
    ms = timed(lambda: F.ldiv_mul(L, U, out=out), args.reps)
    print("fused mat-vec + solve: %.3f ms  %.2f M RHS/s  %.1f GB/s (16 B/elem)" % (
        ms, args.nrhs / ms / 1e3, 16 * n * args.nrhs / ms / 1e6))
    os.environ["BSK_NO_FUSED"] = "1"
    out2 = torch.empty_like(U)
    ms2 = timed(lambda: F.ldiv_mul(L, U, out=out2), args.reps)
    print("mul! then ldiv!:       %.3f ms  (max |diff| vs fused %.2e)" % (ms2, (out - out2).abs().max().item()))
if args.what == "colloc":
    import time
    import ctypes as C
    from bsplinekit_b200 import _lib
    B5 = bk.BSplineBasis(8, synth.breakpoints(10, 99996))
    R5 = bk.RecombinedBSplineBasis(B5, bk.Derivative(0) + 3 * bk.Derivative(1))
    x5 = torch.from_numpy(bk.collocation_points(R5)).cuda()
    if os.environ.get("PRE") == "cyclic":        # same process state as bench.py: a row-segmented solve first
        n = 10000
        Bc = bk.PeriodicBSplineBasis(4, synth.breakpoints(8, n + 1))
        Mc = bk.collocation_matrix(Bc, bk.collocation_points(Bc))
        Yc = (2 * synth.u_torch(9, 0, n * 16384) - 1).reshape(n, 16384)
        Mc.ldiv_(Yc); torch.cuda.synchronize()
        del Yc
    if os.environ.get("PRE") == "big":
        big = torch.empty(1 << 31, dtype=torch.float64, device="cuda"); del big
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        band = torch.empty((len(R5), 13), dtype=torch.float64, device="cuda")
        nout = C.c_int64(0)
        _lib.check(_lib.lib.bsk_collocation_matrix(R5._handle, x5.data_ptr(), x5.numel(), 2, 2.2e-16, band.data_ptr(),
                                                   C.byref(nout), None))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        M = bk.CollocationMatrix(band)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        del M
        torch.cuda.synchronize(); t3 = time.perf_counter()
        print("assembly %.2f ms, band handle %.2f ms, destroy %.2f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
if args.what == "factor":
    import time
    for (n, k) in ((4096, 6), (100000, 8), (11, 4)):
        xdata = synth.breakpoints(6, n)
        Bi = bk.BSplineBasis(k, bk.make_knots(xdata, k), augment=False)
        for mode, name in ((bk.LU_EXACT, "exact"), (bk.LU_FAST, "fast")):
            best = [1e9, 1e9, 1e9]
            for it in range(5):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                Cm = bk.collocation_matrix(Bi, xdata)
                torch.cuda.synchronize(); t1 = time.perf_counter()
                Cm.lu_(mode)
                torch.cuda.synchronize(); t2 = time.perf_counter()
                del Cm
                torch.cuda.synchronize(); t3 = time.perf_counter()
                itp = bk.SplineInterpolation(Bi, xdata, lu_mode=mode)
                torch.cuda.synchronize(); t4 = time.perf_counter()
                halo = itp.C.halo
                del itp
                best = [min(best[0], t1 - t0), min(best[1], t2 - t1), min(best[2], t4 - t3)]
            print("n=%d k=%d %s: assemble %.3f ms, lu! (elimination + repack + halo analysis + records) %.3f ms, "
                  "SplineInterpolation(B, x) %.3f ms, halo %d" % (n, k, name, 1e3 * best[0], 1e3 * best[1], 1e3 * best[2], halo))
if args.what == "solve32":
    xd32 = synth.breakpoints(6, 4096).astype(np.float32)
    B32 = bk.BSplineBasis(args.order if args.order != 4 else 6, bk.make_knots(xd32, args.order if args.order != 4 else 6), augment=False)
    itp32 = bk.SplineInterpolation(B32, xd32)
    Y32 = (2 * synth.u_torch(7, 0, 4096 * args.nrhs, dtype=torch.float32) - 1).reshape(4096, args.nrhs)
    W32 = torch.empty_like(Y32)
    tot = 0.0
    for it in range(args.reps + 1):
        W32.copy_(Y32)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); itp32.C.ldiv_(W32); e1.record()
        torch.cuda.synchronize()
        if it:
            tot += e0.elapsed_time(e1)
    ms = tot / args.reps
    print("solve32: %.3f ms  %.1f M RHS/s  %.1f GB/s (8 B/element)" % (ms, args.nrhs / ms / 1e3, 8 * 4096 * args.nrhs / ms / 1e6))
if args.what == "sweep":
    # evaluation rate over (dtype, order, number of breakpoints): which table variant serves which shape
    for dt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
        for k in (2, 3, 4, 6, 8):
            for nb in (100, 1000, 1700, 3000, 30000):
                br = synth.breakpoints(3, nb).astype(dt)
                B = bk.BSplineBasis(k, br)
                S = bk.Spline(B, (2 * synth.u_numpy(4, 0, len(B)) - 1).astype(dt))
                xd = synth.u_torch(5, 0, args.points, dtype=tdt)
                outs = [torch.empty_like(xd)]
                ms = timed(lambda: S.evaluate(xd, (0,), out=outs), args.reps)
                print("%s k=%d breaks=%6d: %.3f ms  %.1f Gpt/s" % (np.dtype(dt).name, k, nb, ms, args.points / ms / 1e6), flush=True)
                del xd, outs
