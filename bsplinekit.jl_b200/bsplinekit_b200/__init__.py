"""bsplinekit_b200 — host-side mirror of the BSplineKit.jl interface over libbsplinekit_b200.so.

See api.py.  Importing this package loads the CUDA shared library and fails loudly if it
has not been built (no CPU fallback).
"""
from ._lib import (ArgumentError, DimensionMismatch, ZeroPivotException, CudaError, UnsupportedError,
                   EVAL_AUTO, EVAL_DEBOOR, EVAL_PP, SOLVE_AUTO, SOLVE_TWOPASS, SOLVE_WINDOWED, LU_EXACT, LU_FAST,
                   EXTRAP_NONE, EXTRAP_FLAT, EXTRAP_LINEAR, EXTRAP_SMOOTH,
                   LIB_PATH, device_count, version)
from .api import (BSplineOrder, Derivative, DifferentialOp, Periodic, Natural, period,
                  BSplineBasis, PeriodicBSplineBasis, RecombinedBSplineBasis, augment_knots,
                  knots, order, boundaries, find_knot_interval, evaluate_all,
                  Spline, SplineBatch, ComplexSpline, integral, collocation_points, greville_points, collocation_matrix, collocation_rows, CollocationMatrix,
                  CyclicTridiagonalMatrix, CyclicBandedMatrix, lu, ldiv, make_knots, SplineInterpolation, interpolate,
                  approximate, approximate_, ApproxByInterpolation, VariationDiminishing, MinimiseL2Error,
                  SplineApproximation, fit, SmoothingSpline, DomainError,
                  Flat, Linear, Smooth, SplineExtrapolation, extrapolate,
                  galerkin_matrix, galerkin_projection, gausslegendre,
                  BasisFunction, evaluate, support, nonzero_in_segment, isindomain,
                  AbstractBSplineBasis, basis, coefficients, spline, has_parent_basis, constraints,
                  recombination_matrix, knots_in_main_period, basis_to_array_index, LeftNormal, RightNormal, dot,
                  max_order, DerivativeUnitRange, diff)
