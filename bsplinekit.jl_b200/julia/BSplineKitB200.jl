# BSplineKitB200.jl — Julia-side shim that binds libbsplinekit_b200.so (include/bsk.h) behind
# BSplineKit's own call signatures, by adding METHODS (multiple dispatch is the reference's
# plugin interface; nothing in BSplineKit is edited).
#
# NOTE: Julia is not installed in the build image, so this file has been reviewed but never
# executed; the executable stand-in with the same structure is the Python mirror
# (bsplinekit_b200/api.py), which the test-suite drives through the same C ABI.
#
# Usage:
#   using BSplineKit, BSplineKitB200
#   xs = DeviceVector(rand(2^28))                 # host -> device once
#   S  = interpolate(x, y, BSplineOrder(4)) |> spline
#   v  = S.(xs)            # or S(xs): device-resident values   (Splines/spline.jl:144)
#   dv = (Derivative(1) * S)(xs)
#   v, dv = evaluate(S, xs, (0, 1))              # fused, one pass over the points
#   itp = interpolate(x, DeviceMatrix(Y), BSplineOrder(6))   # Y is M x N: M data sets interleaved
module BSplineKitB200

using BSplineKit
using BSplineKit: BSplines, Splines, Collocation, SplineInterpolations, Recombinations, SplineExtrapolations
using BSplineKit.BSplines: BSplineBasis, PeriodicBSplineBasis, AbstractBSplineBasis, knots, order
using BSplineKit.Recombinations: RecombinedBSplineBasis, recombination_matrix, num_constraints,
    num_recombined, submatrices
using LinearAlgebra
using SparseArrays: SparseMatrixCSC, sparse

export DeviceVector, DeviceMatrix, device_spline, evaluate,
    device_galerkin_band, device_galerkin_matrix, device_galerkin_projection

const lib = get(ENV, "BSPLINEKIT_B200_LIB", "libbsplinekit_b200")

# ---- status -> exception (include/bsk.h: bsk_status) ---------------------------------------
function last_error()
    buf = Vector{UInt8}(undef, 512)
    n = ccall((:bsk_last_error, lib), Csize_t, (Ptr{UInt8}, Csize_t), buf, 512)
    String(buf[1:min(n, 511)])
end

function check(st::Int32, info = nothing)
    st == 0 && return nothing
    msg = last_error()
    st == -1 && throw(ArgumentError(msg))
    st == -2 && throw(DimensionMismatch(msg))
    st == -4 && throw(LinearAlgebra.ZeroPivotException(something(info, -1)))
    st == -5 && throw(OutOfMemoryError())
    error("libbsplinekit_b200: $msg (status $st)")        # -3 CUDA, -6 unsupported
end

dtype_code(::Type{Float32}) = Int32(0)
dtype_code(::Type{Float64}) = Int32(1)

# ---- device arrays: raw device pointer + owner finaliser (no CUDA.jl) -----------------------
mutable struct DeviceVector{T} <: AbstractVector{T}
    ptr::Ptr{T}
    len::Int
    device::Int32
    function DeviceVector{T}(::UndefInitializer, n::Integer; device = Int32(0)) where {T}
        p = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:bsk_malloc, lib), Int32, (Int32, Csize_t, Ptr{Ptr{Cvoid}}), device, n * sizeof(T), p))
        v = new{T}(convert(Ptr{T}, p[]), n, device)
        finalizer(x -> ccall((:bsk_free, lib), Int32, (Int32, Ptr{Cvoid}), x.device, x.ptr), v)
    end
end
Base.size(v::DeviceVector) = (v.len,)
Base.getindex(::DeviceVector, i...) = error("scalar indexing of a DeviceVector is not supported; use Array(v)")

function DeviceVector(h::Vector{T}; device = Int32(0)) where {T <: Union{Float32, Float64, Int64, Int8}}
    v = DeviceVector{T}(undef, length(h); device)
    GC.@preserve h check(ccall((:bsk_memcpy_h2d, lib), Int32, (Int32, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                               device, v.ptr, pointer(h), sizeof(h), C_NULL))
    check(ccall((:bsk_stream_sync, lib), Int32, (Int32, Ptr{Cvoid}), device, C_NULL))
    v
end

function Base.Array(v::DeviceVector{T}) where {T}
    h = Vector{T}(undef, v.len)
    GC.@preserve h check(ccall((:bsk_memcpy_d2h, lib), Int32, (Int32, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                               v.device, pointer(h), v.ptr, sizeof(h), C_NULL))
    check(ccall((:bsk_stream_sync, lib), Int32, (Int32, Ptr{Cvoid}), v.device, C_NULL))
    h
end

# M x N matrix, column-major: column i holds the M interleaved values of data point i, i.e. the
# memory layout of Vector{SVector{M,T}} (test/interpolation.jl:57).
struct DeviceMatrix{T}
    data::DeviceVector{T}
    M::Int
    N::Int
end
DeviceMatrix(h::Matrix{T}) where {T} = DeviceMatrix{T}(DeviceVector(vec(h)), size(h, 1), size(h, 2))

# ---- bases ----------------------------------------------------------------------------------
mutable struct BasisHandle
    ptr::Ptr{Cvoid}
end
destroy!(h::BasisHandle) = ccall((:bsk_basis_destroy, lib), Int32, (Ptr{Cvoid},), h.ptr)

struct RecombDesc
    cl::Int32; cr::Int32; nl::Int32; nr::Int32; ul_rows::Int32; lr_rows::Int32
    ul::Ptr{Float64}; lr::Ptr{Float64}
end

# One handle per (basis, device): a handle lives on the device it was created for, so a process that
# drives several GPUs (one host thread per device) gets one handle per device from the same basis object.
const _basis_cache = IdDict{Any, BasisHandle}()
const _cache_lock = ReentrantLock()
_key(obj, device) = (obj, Int32(device))

function handle(B::BSplineBasis{k, T}; device = Int32(0)) where {k, T}
    lock(_cache_lock) do; get!(_basis_cache, _key(B, device)) do
        t = collect(T, knots(B))                       # materialises AugmentedKnots (knots.jl:43-48)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve t check(ccall((:bsk_basis_create, lib), Int32,
            (Int32, Int32, Int32, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}),
            0, k, dtype_code(T), pointer(t), length(t), C_NULL, device, out))
        finalizer(destroy!, BasisHandle(out[]))
    end; end
end

function handle(B::PeriodicBSplineBasis{k, T}; device = Int32(0)) where {k, T}
    lock(_cache_lock) do; get!(_basis_cache, _key(B, device)) do
        ξ = collect(T, parent(knots(B)))               # breakpoints incl. endpoint (periodic.jl:37-55)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve ξ check(ccall((:bsk_basis_create, lib), Int32,
            (Int32, Int32, Int32, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}),
            1, k, dtype_code(T), pointer(ξ), length(ξ), C_NULL, device, out))
        finalizer(destroy!, BasisHandle(out[]))
    end; end
end

function handle(R::RecombinedBSplineBasis{k, T}; device = Int32(0)) where {k, T}
    lock(_cache_lock) do; get!(_basis_cache, _key(R, device)) do
        A = recombination_matrix(R)
        ul, lr = map(m -> Matrix{Float64}(m), submatrices(A))      # corner blocks (matrices.jl:70-110)
        cl, cr = num_constraints(A); nl, nr = num_recombined(A)
        t = collect(T, knots(parent(R)))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve t ul lr begin
            desc = Ref(RecombDesc(cl, cr, nl, nr, size(ul, 1), size(lr, 1), pointer(ul), pointer(lr)))
            check(ccall((:bsk_basis_create, lib), Int32,
                (Int32, Int32, Int32, Ptr{Cvoid}, Int64, Ptr{RecombDesc}, Int32, Ptr{Ptr{Cvoid}}),
                2, k, dtype_code(T), pointer(t), length(t), desc, device, out))
        end
        finalizer(destroy!, BasisHandle(out[]))
    end; end
end

# find_knot_interval(ts, xs)  (BSplines/evaluate_all.jl:102-119, periodic.jl:85-100)
function BSplines.find_knot_interval(B::AbstractBSplineBasis, xs::DeviceVector{T}) where {T}
    is = DeviceVector{Int64}(undef, length(xs)); zs = DeviceVector{Int8}(undef, length(xs))
    check(ccall((:bsk_find_knot_interval, lib), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int8}, Ptr{Cvoid}),
        handle(B; device = xs.device).ptr, xs.ptr, length(xs), is.ptr, zs.ptr, C_NULL))
    is, zs
end

# evaluate_all(B, xs, [op], [T]; [ileft])  (BSplines/evaluate_all.jl:53-64, periodic.jl:228,
# Recombinations/bases.jl:386): returns (is::DeviceVector{Int64}, bs::DeviceMatrix{T} k x n)
function BSplines.evaluate_all(B::AbstractBSplineBasis, xs::DeviceVector{Tx},
                               op::Derivative{n} = Derivative(0), ::Type{T} = Tx;
                               ileft::Union{Nothing, DeviceVector{Int64}} = nothing) where {Tx, n, T}
    k = order(B)
    is = DeviceVector{Int64}(undef, length(xs))
    bs = DeviceVector{T}(undef, k * length(xs))
    check(ccall((:bsk_evaluate_all, lib), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}, Ptr{Cvoid}),
        handle(B; device = xs.device).ptr, xs.ptr, length(xs), n, dtype_code(T),
        ileft === nothing ? Ptr{Int64}(C_NULL) : ileft.ptr, is.ptr, bs.ptr, C_NULL))
    is, DeviceMatrix{T}(bs, k, length(xs))
end
(B::AbstractBSplineBasis)(xs::DeviceVector, args...; kws...) = BSplines.evaluate_all(B, xs, args...; kws...)

# ---- splines --------------------------------------------------------------------------------
mutable struct SplineHandle
    ptr::Ptr{Cvoid}
end
const _spline_cache = IdDict{Any, SplineHandle}()

function device_spline(S::Spline{T}; device = Int32(0)) where {T <: Union{Float32, Float64}}
    lock(_cache_lock) do; get!(_spline_cache, _key(S, device)) do
        B = Splines.basis(S)
        c = collect(T, Splines.unwrap_coefficients(S))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve c check(ccall((:bsk_spline_create, lib), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}), handle(B; device).ptr, pointer(c), length(c), out))
        finalizer(h -> ccall((:bsk_spline_destroy, lib), Int32, (Ptr{Cvoid},), h.ptr), SplineHandle(out[]))
    end; end
end

"""
    evaluate(S::Spline, xs::DeviceVector, derivs::NTuple{N,Int}) -> NTuple{N,DeviceVector}

`D^d S` for every `d` in `derivs`, in one pass over the points (fused value + derivatives).
"""
function evaluate(S::Spline{T}, xs::DeviceVector{T}, derivs::NTuple{N, Int} = (0,); algo = 0) where {T, N}
    outs = ntuple(_ -> DeviceVector{T}(undef, length(xs)), Val(N))
    ptrs = Ptr{Cvoid}[o.ptr for o in outs]
    ds = Int32[d for d in derivs]
    GC.@preserve ptrs ds check(ccall((:bsk_spline_eval, lib), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Int32, Ptr{Cvoid}),
        device_spline(S; device = xs.device).ptr, xs.ptr, length(xs), N, ds, ptrs, algo, C_NULL))
    outs
end

# (S::Spline)(x) and S.(xs)  (Splines/spline.jl:69,144): device vectors take the batched path.
(S::Spline)(xs::DeviceVector) = evaluate(S, xs, (0,))[1]
Base.Broadcast.broadcasted(S::Spline, xs::DeviceVector) = S(xs)
# Derivative(n) * S stays the reference's host code (O(N n), spline.jl:272-330); the resulting
# Spline is uploaded on first use by device_spline.

# Host arrays, staged and pipelined inside the library (H2D | kernel | D2H overlapped):
function evaluate(S::Spline{T}, xs::Vector{T}, derivs::NTuple{N, Int} = (0,); algo = 0, device = Int32(0)) where {T, N}
    outs = ntuple(_ -> Vector{T}(undef, length(xs)), Val(N))
    ptrs = Ptr{Cvoid}[pointer(o) for o in outs]
    ds = Int32[d for d in derivs]
    GC.@preserve xs outs ptrs ds check(ccall((:bsk_spline_eval_host, lib), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Int32),
        device_spline(S; device).ptr, pointer(xs), length(xs), N, ds, ptrs, algo))
    outs
end

# Vector-valued splines: coefficients SVector{M,T} held as an M x N DeviceMatrix (M fastest),
# e.g. the result of a batched interpolate!.  Returns an M x n DeviceMatrix (column p = S(xs[p])).
function evaluate(B::AbstractBSplineBasis, C::DeviceMatrix{T}, xs::DeviceVector{T}) where {T}
    C.N == length(B) || throw(DimensionMismatch("wrong number of coefficients"))
    out = DeviceVector{T}(undef, C.M * length(xs))
    check(ccall((:bsk_spline_eval_multi, lib), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
        handle(B; device = xs.device).ptr, C.data.ptr, C.N, C.M, xs.ptr, length(xs), out.ptr, C_NULL))
    DeviceMatrix{T}(out, C.M, length(xs))
end

# ---- collocation + banded LU / solve ------------------------------------------------------------
"""Device-resident CollocationMatrix (band store (2k-3) x n) and its LU; `<: Factorization` so that
`SplineInterpolation` accepts it exactly like `CyclicLUWrapper` (SplineInterpolations.jl:56-88)."""
mutable struct DeviceBandedLU{T} <: Factorization{T}
    ptr::Ptr{Cvoid}
    n::Int
end
Base.size(F::DeviceBandedLU) = (F.n, F.n)

function Collocation.collocation_matrix(B::AbstractBSplineBasis, xs::DeviceVector{Float64},
                                        deriv::Derivative{n} = Derivative(0);
                                        clip_threshold = eps(Float64)) where {n}
    k = order(B)
    N = length(B)
    length(xs) == N || throw(ArgumentError("wrong dimensions of collocation matrix"))
    band = DeviceVector{Float64}(undef, (2k - 3) * N)
    nout = Ref{Int64}(0)
    check(ccall((:bsk_collocation_matrix, lib), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Float64, Ptr{Float64}, Ptr{Int64}, Ptr{Cvoid}),
        handle(B; device = xs.device).ptr, xs.ptr, N, n, clip_threshold, band.ptr, nout, C_NULL))
    nout[] > 0 && @warn "Non-zero value outside of matrix bands ($(nout[]) entries)"    # Collocation.jl:239-244
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:bsk_banded_create, lib), Int32, (Ptr{Float64}, Int64, Int32, Int32, Int32, Ptr{Ptr{Cvoid}}),
                band.ptr, N, k - 2, k - 2, xs.device, out))
    F = DeviceBandedLU{Float64}(out[], N)
    finalizer(f -> ccall((:bsk_banded_destroy, lib), Int32, (Ptr{Cvoid},), f.ptr), F)
end

# fast = false: BANFAC's own order, bit-identical factors; fast = true: partitioned elimination with checked
# partition boundaries (n >= 1024), factors within ~1e-16 relative, ZeroPivotException(i) still thrown.
# Rectangular collocation matrices (any number of points; least-squares fitting): the Matrix / SparseMatrixCSC
# targets of collocation_matrix (Collocation/Collocation.jl:82-106,162-193).
function Collocation.collocation_matrix(B::AbstractBSplineBasis, xs::DeviceVector{Float64}, deriv::Derivative{n},
                                        ::Type{Matrix{Float64}}; clip_threshold = eps(Float64)) where {n}
    nx, N = length(xs), length(B)
    dense = DeviceVector{Float64}(undef, nx * N)
    check(ccall((:bsk_collocation_rows, lib), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Float64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}),
        handle(B; device = xs.device).ptr, xs.ptr, nx, n, clip_threshold, C_NULL, C_NULL, dense.ptr, C_NULL))
    DeviceMatrix{Float64}(dense, nx, N)
end

function Collocation.collocation_matrix(B::AbstractBSplineBasis, xs::DeviceVector{Float64}, deriv::Derivative{n},
                                        ::Type{SparseMatrixCSC{Float64}}; clip_threshold = eps(Float64)) where {n}
    nx, N, k = length(xs), length(B), order(B)
    col = DeviceVector{Int64}(undef, nx * k); val = DeviceVector{Float64}(undef, nx * k)
    check(ccall((:bsk_collocation_rows, lib), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Float64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}),
        handle(B; device = xs.device).ptr, xs.ptr, nx, n, clip_threshold, col.ptr, val.ptr, C_NULL, C_NULL))
    J, V = Array(col), Array(val)
    I = repeat(1:nx; inner = k)
    keep = J .> 0
    sparse(I[keep], J[keep], V[keep], nx, N)
end

function LinearAlgebra.lu!(F::DeviceBandedLU{Float64}; fast::Bool = false)     # Collocation/matrix.jl:132-201
    info = Ref{Int64}(0)
    st = ccall((:bsk_banded_lu_mode, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Int64}), F.ptr, Int32(fast), info)
    check(st, info[])
    F
end

# ldiv!(F, Y): Y is M x N (M data sets interleaved), solved in place   (Collocation/matrix.jl:218-271)
function LinearAlgebra.ldiv!(F::DeviceBandedLU{Float64}, Y::DeviceMatrix{Float64}; algo = 0)
    Y.N == F.n || throw(DimensionMismatch("vectors `x` and `y` must have length $(F.n)"))
    check(ccall((:bsk_banded_solve, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Ptr{Cvoid}),
                F.ptr, Y.data.ptr, Y.M, algo, C_NULL))
    Y
end

# mul!(Y, L, X) on batches (Collocation/matrix.jl:91-94) — L must not be factorised
function LinearAlgebra.mul!(Y::DeviceMatrix{Float64}, L::DeviceBandedLU, X::DeviceMatrix{Float64})
    (X.N == L.n && Y.N == L.n && X.M == Y.M) || throw(DimensionMismatch("vectors `x` and `y` must have length $(L.n)"))
    check(ccall((:bsk_banded_matvec, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Cvoid}),
                L.ptr, X.data.ptr, Y.data.ptr, X.M, C_NULL))
    Y
end

"""
    ldiv_mul!(dU, F, L, U)

`mul!(Y, L, U); ldiv!(dU, F, Y)` — the right-hand side of a collocation time step
(examples/heat.jl:314-328) — in one pass over device memory.  `F` factorised, `L` not; `U` is only read.
"""
function ldiv_mul!(dU::DeviceMatrix{Float64}, F::DeviceBandedLU, L::DeviceBandedLU, U::DeviceMatrix{Float64})
    (U.N == F.n && dU.N == F.n && U.M == dU.M) || throw(DimensionMismatch("vectors `x` and `y` must have length $(F.n)"))
    check(ccall((:bsk_banded_matvec_solve, lib), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Cvoid}),
                F.ptr, L.ptr, U.data.ptr, dU.data.ptr, U.M, C_NULL))
    dU
end

# ---- the same path in Float32 (the reference methods are generic in T) ------------------------------
# CollocationMatrix{Float32} / lu! / ldiv! in Float32 (Collocation/Collocation.jl:165-198,
# Collocation/matrix.jl:132-271): the `_f32` entry points, same layouts and status codes.
function Collocation.collocation_matrix(B::AbstractBSplineBasis, xs::DeviceVector{Float32},
                                        deriv::Derivative{n} = Derivative(0);
                                        clip_threshold = eps(Float32)) where {n}
    k = order(B)
    N = length(B)
    length(xs) == N || throw(ArgumentError("wrong dimensions of collocation matrix"))
    band = DeviceVector{Float32}(undef, (2k - 3) * N)
    nout = Ref{Int64}(0)
    check(ccall((:bsk_collocation_matrix_f32, lib), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Int64, Int32, Float32, Ptr{Float32}, Ptr{Int64}, Ptr{Cvoid}),
        handle(B; device = xs.device).ptr, xs.ptr, N, n, Float32(clip_threshold), band.ptr, nout, C_NULL))
    nout[] > 0 && @warn "Non-zero value outside of matrix bands ($(nout[]) entries)"
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:bsk_banded_f32_create, lib), Int32, (Ptr{Float32}, Int64, Int32, Int32, Int32, Ptr{Ptr{Cvoid}}),
                band.ptr, N, k - 2, k - 2, xs.device, out))
    F = DeviceBandedLU{Float32}(out[], N)
    finalizer(f -> ccall((:bsk_banded_f32_destroy, lib), Int32, (Ptr{Cvoid},), f.ptr), F)
end

function LinearAlgebra.lu!(F::DeviceBandedLU{Float32})
    info = Ref{Int64}(0)
    st = ccall((:bsk_banded_f32_lu, lib), Int32, (Ptr{Cvoid}, Ptr{Int64}), F.ptr, info)
    check(st, info[])
    F
end

function LinearAlgebra.ldiv!(F::DeviceBandedLU{Float32}, Y::DeviceMatrix{Float32})
    Y.N == F.n || throw(DimensionMismatch("vectors `x` and `y` must have length $(F.n)"))
    check(ccall((:bsk_banded_f32_solve, lib), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64, Int32, Ptr{Cvoid}),
                F.ptr, Y.data.ptr, Y.M, Int32(0), C_NULL))
    Y
end

# ---- periodic cubic bases: CyclicTridiagonalMatrix (Collocation/cyclic.jl:21-33, 112-197) ----------------
"""Device factorisation of a `CyclicTridiagonalMatrix` (periodic cubic interpolation; `default_matrix_type`
picks it for k = 4, Collocation/Collocation.jl:133-149).  The reference re-eliminates the matrix inside every
`ldiv!` (cyclic.jl:121-197, re-copied from `C_copy` at SplineInterpolations.jl:147-149); here the tridiagonal
part is factored once and `ldiv!` only streams the right-hand sides."""
mutable struct DeviceCyclicLU{T} <: Factorization{T}
    ptr::Ptr{Cvoid}
    n::Int
end
Base.size(F::DeviceCyclicLU) = (F.n, F.n)

function DeviceCyclicLU(A::Collocation.CyclicTridiagonalMatrix{Float64}; device = Int32(0))
    a, b, c = collect(Float64, A.a), collect(Float64, A.b), collect(Float64, A.c)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve a b c check(ccall((:bsk_cyclic_create, lib), Int32,
                (Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Ptr{Ptr{Cvoid}}),
                a, b, c, length(b), device, out))
    F = DeviceCyclicLU{Float64}(out[], length(b))
    finalizer(f -> ccall((:bsk_cyclic_destroy, lib), Int32, (Ptr{Cvoid},), f.ptr), F)
end

function LinearAlgebra.ldiv!(F::DeviceCyclicLU, Y::DeviceMatrix{Float64})        # cyclic.jl:112-161
    Y.N == F.n || throw(DimensionMismatch("all vectors must have length $(F.n)"))
    check(ccall((:bsk_cyclic_solve, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Cvoid}),
                F.ptr, Y.data.ptr, Y.M, C_NULL))
    Y
end

# ---- periodic bases of order k ∉ {2, 4}: cyclic banded factorisation ---------------------------------
"""Stands where the reference puts `lu(::SparseMatrixCSC)` (UMFPACK) for periodic bases
(Collocation/Collocation.jl:141-142, SplineInterpolations.jl:107-109).  `cband` is the host-assembled
(2bw+1) x N cyclic band store; `shift` the offset of the band centre (1 for odd orders)."""
mutable struct DeviceCyclicBandedLU{T} <: Factorization{T}
    ptr::Ptr{Cvoid}
    n::Int
end
Base.size(F::DeviceCyclicBandedLU) = (F.n, F.n)

function DeviceCyclicBandedLU(cband::Matrix{Float64}, shift::Integer = 0; device = Int32(0))
    nb, N = size(cband)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:bsk_cyclic_banded_create, lib), Int32,
                (Ptr{Float64}, Int64, Int32, Int32, Int32, Ptr{Ptr{Cvoid}}),
                cband, N, (nb - 1) ÷ 2, shift, device, out))
    F = DeviceCyclicBandedLU{Float64}(out[], N)
    finalizer(f -> ccall((:bsk_cyclic_banded_destroy, lib), Int32, (Ptr{Cvoid},), f.ptr), F)
end

function LinearAlgebra.ldiv!(F::DeviceCyclicBandedLU, Y::DeviceMatrix{Float64})
    Y.N == F.n || throw(DimensionMismatch("all vectors must have length $(F.n)"))
    check(ccall((:bsk_cyclic_banded_solve, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Cvoid}),
                F.ptr, Y.data.ptr, Y.M, C_NULL))
    Y
end

# ---- Galerkin matrices and projections (Galerkin/linear.jl:60-311, projection.jl:1-99) -----------------
"""
    device_galerkin_band(B, (Derivative(m), Derivative(n))) -> Matrix{Float64}   # (2k-1) x N band store

All quadrature nodes go through one batched `evaluate_all` on the device; entry `[(k-1) + i - j + 1, j]` is
`A[i, j]` (plain / recombined bases: the `BandedMatrix` of linear.jl:176-199; periodic bases: read cyclically,
`A[mod1(j + d, N), j]` — the entries of the `SparseMatrixCSC` of linear.jl:144).  Bit-identical to the
sequential loop of `galerkin_matrix!` (linear.jl:268-311).
"""
function device_galerkin_band(B::AbstractBSplineBasis, deriv = (Derivative(0), Derivative(0)); device = Int32(0))
    k = order(B); N = length(B)
    qx, qw = BSplineKit.Galerkin.default_quadrature((B, B))
    qxv = collect(Float64, qx); qwv = collect(Float64, qw)
    m, n = map(d -> Int32(BSplineKit.DifferentialOps.max_order(d)), deriv)
    band = DeviceVector{Float64}(undef, (2k - 1) * N; device)
    GC.@preserve qxv qwv check(ccall((:bsk_galerkin_matrix, lib), Int32,
        (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Cvoid}),
        handle(B; device).ptr, m, n, pointer(qxv), pointer(qwv), Int32(length(qxv)), band.ptr, C_NULL))
    reshape(Array(band), 2k - 1, N)
end

# galerkin_matrix((B1, B2), derivs): two bases sharing their parent (linear.jl:134-199) -> ((l + u + 1) x N2 band store, l, u)
function device_galerkin_band(Bs::Tuple{AbstractBSplineBasis, AbstractBSplineBasis},
                              deriv = (Derivative(0), Derivative(0)); device = Int32(0))
    B1, B2 = Bs
    qx, qw = BSplineKit.Galerkin.default_quadrature(Bs)
    qxv = collect(Float64, qx); qwv = collect(Float64, qw)
    m, n = map(d -> Int32(BSplineKit.DifferentialOps.max_order(d)), deriv)
    l = Ref{Int32}(0); u = Ref{Int32}(0)
    GC.@preserve qxv qwv check(ccall((:bsk_galerkin_matrix_pair, lib), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int64, Ptr{Float64},
         Ptr{Int32}, Ptr{Int32}, Ptr{Cvoid}),
        handle(B1; device).ptr, handle(B2; device).ptr, m, n, pointer(qxv), pointer(qwv), Int32(length(qxv)),
        C_NULL, 0, C_NULL, l, u, C_NULL))
    band = DeviceVector{Float64}(undef, (l[] + u[] + 1) * length(B2); device)
    GC.@preserve qxv qwv check(ccall((:bsk_galerkin_matrix_pair, lib), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int64, Ptr{Float64},
         Ptr{Int32}, Ptr{Int32}, Ptr{Cvoid}),
        handle(B1; device).ptr, handle(B2; device).ptr, m, n, pointer(qxv), pointer(qwv), Int32(length(qxv)),
        C_NULL, 0, band.ptr, l, u, C_NULL))
    reshape(Array(band), Int(l[] + u[] + 1), length(B2)), Int(l[]), Int(u[])
end

# galerkin_matrix(B::PeriodicBSplineBasis) as the reference's SparseMatrixCSC, assembled on the device
function device_galerkin_matrix(B::PeriodicBSplineBasis, deriv = (Derivative(0), Derivative(0)); device = Int32(0))
    w = device_galerkin_band(B, deriv; device); k = order(B); N = length(B)
    I = Int[]; J = Int[]; V = Float64[]
    for j in 1:N, d in -(k - 1):(k - 1)
        push!(I, mod1(j + d, N)); push!(J, j); push!(V, w[k + d, j])
    end
    sparse(I, J, V, N, N)
end

# galerkin_projection(f, B, Derivative(n)): f is evaluated on the host at the nodes the library lists
function device_galerkin_projection(f, B::AbstractBSplineBasis, deriv = Derivative(0); device = Int32(0))
    k = order(B)
    qx, qw = BSplineKit.Galerkin.default_quadrature((B, B))
    qxv = collect(Float64, qx); qwv = collect(Float64, qw)
    np = Ref{Int64}(0)
    GC.@preserve qxv check(ccall((:bsk_galerkin_nodes, lib), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Float64}, Int64, Ptr{Int64}),
        handle(B; device).ptr, pointer(qxv), Int32(length(qxv)), C_NULL, 0, np))
    xs = Vector{Float64}(undef, np[])
    GC.@preserve qxv xs check(ccall((:bsk_galerkin_nodes, lib), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Float64}, Int64, Ptr{Int64}),
        handle(B; device).ptr, pointer(qxv), Int32(length(qxv)), pointer(xs), length(xs), np))
    fx = DeviceVector(Float64.(f.(xs)); device)
    φ = DeviceVector{Float64}(undef, length(B); device)
    GC.@preserve qxv qwv check(ccall((:bsk_galerkin_projection, lib), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Cvoid}),
        handle(B; device).ptr, Int32(BSplineKit.DifferentialOps.max_order(deriv)), pointer(qxv), pointer(qwv),
        Int32(length(qxv)), fx.ptr, length(xs), φ.ptr, C_NULL))
    Array(φ)
end

# ---- extrapolation -------------------------------------------------------------------------------------
extrap_code(::SplineExtrapolations.Flat) = Int32(1)
extrap_code(::SplineExtrapolations.Linear) = Int32(2)
extrap_code(::SplineExtrapolations.Smooth) = Int32(3)

# (f::SplineExtrapolation)(xs) on device points   (SplineExtrapolations.jl:22-101)
function (f::SplineExtrapolations.SplineExtrapolation)(xs::DeviceVector{T}) where {T}
    S = SplineExtrapolations.spline(f)
    out = DeviceVector{T}(undef, length(xs))
    ptrs = Ptr{Cvoid}[out.ptr]
    ds = Int32[0]
    GC.@preserve ptrs ds check(ccall((:bsk_spline_eval_extrap, lib), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Int32, Int32, Ptr{Cvoid}),
        device_spline(S; device = xs.device).ptr, xs.ptr, length(xs), 1, ds, ptrs, 0, extrap_code(SplineExtrapolations.method(f)), C_NULL))
    out
end

"""
    interpolate(x, Y::DeviceMatrix, BSplineOrder(k)) -> (B, F, Y)

Batched `interpolate` (SplineInterpolations.jl:260-273): knots from `make_knots` (host), one
collocation assembly + LU on the device, then every data set of `Y` is solved in place; column
`j` of the result holds the B-spline coefficients of data set `j`.
"""
function SplineInterpolations.interpolate(x::AbstractVector, Y::DeviceMatrix{T}, ord::BSplineOrder{k}) where {k, T <: Union{Float32, Float64}}
    t = SplineInterpolations.make_knots(collect(T, x), ord, nothing)   # :302-326, stays host code
    B = BSplineBasis(ord, t; augment = Val(false))
    xs = DeviceVector(collect(T, x))
    F = lu!(Collocation.collocation_matrix(B, xs))
    ldiv!(F, Y)
    B, F, Y
end

end # module
