"""ctypes binding of libbsplinekit_b200.so (include/bsk.h).

This is the Python stand-in for the Julia `ccall` shim (julia/BSplineKitB200.jl): one
thin call per C-ABI entry point plus status -> exception mapping.  There is NO CPU
fallback: if the shared library is missing, import fails loudly; if no CUDA device is
present, every compute call raises CudaError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libbsplinekit_b200.so")

BSK_F32, BSK_F64 = 0, 1
BASIS_PLAIN, BASIS_PERIODIC, BASIS_RECOMBINED = 0, 1, 2
EVAL_AUTO, EVAL_DEBOOR, EVAL_PP = 0, 1, 2
SOLVE_AUTO, SOLVE_TWOPASS, SOLVE_WINDOWED = 0, 1, 2
LU_EXACT, LU_FAST = 0, 1
EXTRAP_NONE, EXTRAP_FLAT, EXTRAP_LINEAR, EXTRAP_SMOOTH = 0, 1, 2, 3


class ArgumentError(ValueError):
    """Julia ArgumentError (e.g. BSplines/basis.jl:80-82, Splines/spline.jl:303-306)."""


class DimensionMismatch(ValueError):
    """Julia DimensionMismatch (Collocation/matrix.jl:134-138,230-232)."""


class ZeroPivotException(ArithmeticError):
    """LinearAlgebra.ZeroPivotException(i) (Collocation/matrix.jl:150-198)."""

    def __init__(self, index, msg=""):
        super().__init__(msg or "ZeroPivotException(%d)" % index)
        self.info = index


class CudaError(RuntimeError):
    pass


class UnsupportedError(NotImplementedError):
    pass


class RecombDesc(C.Structure):
    _fields_ = [("cl", C.c_int32), ("cr", C.c_int32), ("nl", C.c_int32), ("nr", C.c_int32),
                ("ul_rows", C.c_int32), ("lr_rows", C.c_int32),
                ("ul", C.POINTER(C.c_double)), ("lr", C.POINTER(C.c_double))]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libbsplinekit_b200.so not found at %s — build it with bsplinekit.jl_b200/build.sh "
        "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

_vp, _i32, _i64, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_pp = C.POINTER(C.c_void_p)

# every symbol declared in include/bsk.h, with its argument types
SIGNATURES = {
    "bsk_version": (C.c_char_p, []),
    "bsk_last_error": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "bsk_device_count": (_i32, [C.POINTER(_i32)]),
    "bsk_probe_stream_1r2w": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "bsk_probe_stream_1r2w_tma": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "bsk_probe_fma": (_i32, [_i32, _i64, _vp, C.POINTER(_dbl), _vp]),
    "bsk_probe_gather": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp]),
    "bsk_probe_copy_host": (_i32, [_vp, _vp, _i64, _i32, _pp]),
    "bsk_malloc": (_i32, [_i32, C.c_size_t, _pp]),
    "bsk_free": (_i32, [_i32, _vp]),
    "bsk_memcpy_h2d": (_i32, [_i32, _vp, _vp, C.c_size_t, _vp]),
    "bsk_memcpy_d2h": (_i32, [_i32, _vp, _vp, C.c_size_t, _vp]),
    "bsk_stream_create": (_i32, [_i32, _pp]),
    "bsk_stream_destroy": (_i32, [_i32, _vp]),
    "bsk_stream_sync": (_i32, [_i32, _vp]),
    "bsk_basis_create": (_i32, [_i32, _i32, _i32, _vp, _i64, C.POINTER(RecombDesc), _i32, _pp]),
    "bsk_basis_destroy": (_i32, [_vp]),
    "bsk_basis_length": (_i32, [_vp, C.POINTER(_i64)]),
    "bsk_find_knot_interval": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "bsk_evaluate_all": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "bsk_spline_create": (_i32, [_vp, _vp, _i64, _pp]),
    "bsk_spline_set_coefs_device": (_i32, [_vp, _vp, _i64, _i64, _vp]),
    "bsk_spline_destroy": (_i32, [_vp]),
    "bsk_spline_coefficients": (_i32, [_vp, _vp, _i64]),
    "bsk_spline_order": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i64), C.POINTER(_i64)]),
    "bsk_spline_knots": (_i32, [_vp, _vp, _i64]),
    "bsk_spline_derivative": (_i32, [_vp, _i32, _pp]),
    "bsk_spline_eval": (_i32, [_vp, _vp, _i64, _i32, C.POINTER(_i32), _pp, _i32, _vp]),
    "bsk_spline_eval_extrap": (_i32, [_vp, _vp, _i64, _i32, C.POINTER(_i32), _pp, _i32, _i32, _vp]),
    "bsk_spline_table_info": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_dbl)]),
    "bsk_spline_eval_host": (_i32, [_vp, _vp, _i64, _i32, C.POINTER(_i32), _pp, _i32]),
    "bsk_spline_eval_multi": (_i32, [_vp, _vp, _i64, _i64, _vp, _i64, _vp, _vp]),
    "bsk_collocation_matrix": (_i32, [_vp, _vp, _i64, _i32, _dbl, _vp, C.POINTER(_i64), _vp]),
    "bsk_collocation_rows": (_i32, [_vp, _vp, _i64, _i32, _dbl, _vp, _vp, _vp, _vp]),
    "bsk_collocation_cyclic": (_i32, [_vp, _vp, _i64, _i32, _dbl, _vp, _vp, _vp, C.POINTER(_i64), _vp]),
    "bsk_banded_create": (_i32, [_vp, _i64, _i32, _i32, _i32, _pp]),
    "bsk_banded_create_host": (_i32, [_vp, _i64, _i32, _i32, _i32, _pp]),
    "bsk_banded_lu": (_i32, [_vp, C.POINTER(_i64)]),
    "bsk_banded_lu_mode": (_i32, [_vp, _i32, C.POINTER(_i64)]),
    "bsk_banded_data": (_i32, [_vp, _vp]),
    "bsk_banded_halo": (_i32, [_vp, C.POINTER(_i32)]),
    "bsk_banded_solve": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "bsk_banded_solve_host": (_i32, [_vp, _vp, _i64, _i32]),
    "bsk_banded_matvec": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "bsk_banded_matvec_solve": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "bsk_banded_destroy": (_i32, [_vp]),
    "bsk_cyclic_create": (_i32, [_vp, _vp, _vp, _i64, _i32, _pp]),
    "bsk_cyclic_create_device": (_i32, [_vp, _vp, _vp, _i64, _i32, _pp]),
    "bsk_cyclic_solve": (_i32, [_vp, _vp, _i64, _vp]),
    "bsk_cyclic_destroy": (_i32, [_vp]),
    "bsk_cyclic_banded_create": (_i32, [_vp, _i64, _i32, _i32, _i32, _pp]),
    "bsk_cyclic_banded_solve": (_i32, [_vp, _vp, _i64, _vp]),
    "bsk_cyclic_banded_destroy": (_i32, [_vp]),
    "bsk_collocation_matrix_f32": (_i32, [_vp, _vp, _i64, _i32, C.c_float, _vp, C.POINTER(_i64), _vp]),
    "bsk_collocation_cyclic_f32": (_i32, [_vp, _vp, _i64, _i32, C.c_float, _vp, _vp, _vp, C.POINTER(_i64), _vp]),
    "bsk_banded_f32_create": (_i32, [_vp, _i64, _i32, _i32, _i32, _pp]),
    "bsk_banded_f32_create_host": (_i32, [_vp, _i64, _i32, _i32, _i32, _pp]),
    "bsk_banded_f32_lu": (_i32, [_vp, C.POINTER(_i64)]),
    "bsk_banded_f32_data": (_i32, [_vp, _vp]),
    "bsk_banded_f32_halo": (_i32, [_vp, C.POINTER(_i32)]),
    "bsk_banded_f32_solve": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "bsk_banded_f32_destroy": (_i32, [_vp]),
    "bsk_cyclic_f32_create": (_i32, [_vp, _vp, _vp, _i64, _i32, _pp]),
    "bsk_cyclic_f32_solve": (_i32, [_vp, _vp, _i64, _vp]),
    "bsk_cyclic_f32_destroy": (_i32, [_vp]),
    "bsk_galerkin_nodes": (_i32, [_vp, _vp, _i32, _vp, _i64, C.POINTER(_i64)]),
    "bsk_galerkin_matrix": (_i32, [_vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    "bsk_galerkin_matrix_weighted": (_i32, [_vp, _i32, _i32, _vp, _vp, _i32, _vp, _i64, _vp, _vp]),
    "bsk_galerkin_matrix_pair": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _i64, _vp, C.POINTER(_i32),
                                        C.POINTER(_i32), _vp]),
    "bsk_galerkin_projection": (_i32, [_vp, _i32, _vp, _vp, _i32, _vp, _i64, _vp, _vp]),
    "bsk_event_create": (_i32, [_i32, _pp]),
    "bsk_event_record": (_i32, [_vp, _vp]),
    "bsk_event_elapsed_ms": (_i32, [_vp, _vp, C.POINTER(C.c_float)]),
    "bsk_event_destroy": (_i32, [_vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args


def last_error():
    buf = C.create_string_buffer(512)
    lib.bsk_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(status, info=None):
    """Map a bsk_status to the exception the reference would throw."""
    if status == 0:
        return
    msg = last_error()
    if status == -1:
        raise ArgumentError(msg)
    if status == -2:
        raise DimensionMismatch(msg)
    if status == -3:
        raise CudaError(msg)
    if status == -4:
        raise ZeroPivotException(int(info) if info is not None else -1, msg)
    if status == -5:
        raise MemoryError(msg)
    if status == -6:
        raise UnsupportedError(msg)
    raise RuntimeError("bsk status %d: %s" % (status, msg))


def device_count():
    n = _i32(0)
    st = lib.bsk_device_count(C.byref(n))
    return n.value if st == 0 else 0


def version():
    return lib.bsk_version().decode()
