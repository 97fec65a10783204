#!/usr/bin/env python
"""Small-size pass over every kernel family, meant to run under compute-sanitizer:
  compute-sanitizer --tool memcheck python tools/sanitize_small.py
No timing, no parity claims (the GPU tests do that); prints DONE at the end."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bsplinekit.jl_b200"))
import numpy as np
import torch
import bsplinekit_b200 as bk
from bsplinekit_b200 import synth

# evaluation: cell table in shared memory, in global memory, pp, de Boor, F32, periodic, extrapolation
for k, nb, dt in ((4, 50, np.float64), (4, 20000, np.float64), (8, 3000, np.float32), (2, 7, np.float64)):
    B = bk.BSplineBasis(k, synth.breakpoints(3, nb).astype(dt))
    S = bk.Spline(B, (2 * synth.u_numpy(4, 0, len(B)) - 1).astype(dt))
    x = torch.from_numpy((-0.1 + 1.2 * synth.u_numpy(5, 0, 10007)).astype(dt)).cuda()
    for algo in (bk.EVAL_AUTO, bk.EVAL_PP, bk.EVAL_DEBOOR):
        S.evaluate(x, (0, 1) if k > 2 else (0,), algo=algo)
        S.evaluate(x[1:], (0,), algo=algo)              # unaligned head / tail
    for m in (bk.Flat(), bk.Linear(), bk.Smooth()):
        bk.extrapolate(S, m)(x)
Bp = bk.PeriodicBSplineBasis(4, synth.breakpoints(8, 301))
Sp = bk.Spline(Bp, 2 * synth.u_numpy(4, 0, len(Bp)) - 1)
Sp.evaluate(torch.from_numpy(-2 + 5 * synth.u_numpy(9, 0, 5000)).cuda(), (0, 1, 2))
bk.evaluate_all(Bp, torch.from_numpy(synth.u_numpy(9, 0, 777)).cuda(), bk.Derivative(1))
bk.find_knot_interval(Bp, torch.from_numpy(synth.u_numpy(9, 0, 777)).cuda())

# banded: LU (streamed), halo analysis, two-pass, windowed (TMEM + segments), fused, mat-vec, multi eval
for k, n, M in ((4, 100, 2048), (6, 520, 2048), (6, 1031, 64), (8, 300, 2048), (3, 40, 96)):
    xdata = synth.breakpoints(6, n)
    Bi = bk.BSplineBasis(k, bk.make_knots(xdata, k), augment=False)
    F = bk.collocation_matrix(Bi, xdata)
    L = bk.collocation_matrix(Bi, xdata, bk.Derivative(min(2, k - 1)))
    U = (2 * synth.u_torch(7, 0, n * M) - 1).reshape(n, M)
    L.mul(U)
    F.lu_()
    for algo in (bk.SOLVE_AUTO, bk.SOLVE_TWOPASS) + ((bk.SOLVE_WINDOWED,) if F.halo else ()):
        F.ldiv_(U.clone(), algo=algo)
    C = F.ldiv_mul(L, U)
    bk.SplineBatch(Bi, C)(torch.from_numpy(synth.u_numpy(5, 0, 333)).cuda())

# periodic interpolation: cyclic tridiagonal, cyclic banded (even / odd orders)
for k in (3, 4, 5, 6, 8):
    n = 90
    x = -1.0 + 2.0 * synth.breakpoints(41, n + 1)[:-1]
    itp = bk.SplineInterpolation(bk.PeriodicBSplineBasis(k, bk.make_knots(x, k, bk.Periodic(2.0))), x)
    itp.interpolate_((2 * synth.u_torch(7, 0, n * 2048) - 1).reshape(n, 2048))
    itp.interpolate_(np.cos(np.pi * x))

# recombined bases, collocation with derivative, natural interpolation, smoothing, Galerkin
B5 = bk.BSplineBasis(8, synth.breakpoints(10, 500))
R5 = bk.RecombinedBSplineBasis(B5, bk.Derivative(0) + 3 * bk.Derivative(1))
bk.collocation_matrix(R5, bk.collocation_points(R5), bk.Derivative(2))
xs = synth.breakpoints(31, 120)
bk.interpolate(xs, np.sin(5 * xs), bk.BSplineOrder(4), bk.Natural())(xs)
bk.fit(bk.BSplineOrder(4), xs, (2 * synth.u_torch(33, 0, 120 * 2048) - 1).reshape(120, 2048), 1e-5)
bk.fit(bk.BSplineOrder(4), xs[:-1], np.sin(5 * xs[:-1]), 1e-5, bk.Periodic(1.0))
Bg = bk.BSplineBasis(5, synth.breakpoints(61, 70))
bk.galerkin_matrix(Bg, (bk.Derivative(1), bk.Derivative(0)))
bk.approximate(np.exp, Bg, bk.MinimiseL2Error())
bk.galerkin_matrix(bk.RecombinedBSplineBasis(Bg, bk.Derivative(0)), as_band=True)
# jump-encoded knot cells (odd / even orders), Float32 interpolation path (plain, recombined, periodic cubic)
for k in (3, 5, 6, 7):
    Bj = bk.BSplineBasis(k, synth.breakpoints(3, 900))
    Sj = bk.Spline(Bj, 2 * synth.u_numpy(4, 0, len(Bj)) - 1)
    Sj.evaluate(torch.from_numpy(synth.u_numpy(5, 0, 20011)).cuda(), (0, 1, 2))
x32 = synth.breakpoints(6, 333).astype(np.float32)
for k in (2, 4, 6, 8):
    itp = bk.interpolate(x32, np.sin(6 * x32), bk.BSplineOrder(k))
    itp.interpolate_((2 * synth.u_torch(7, 0, 333 * 130, dtype=torch.float32) - 1).reshape(333, 130))
bk.interpolate(x32, np.sin(6 * x32), bk.BSplineOrder(4), bk.Natural())
bk.interpolate(x32[:-1], np.sin(2 * np.pi * x32[:-1]), bk.BSplineOrder(4), bk.Periodic(np.float32(1.0)))
torch.cuda.synchronize()
print("DONE")
