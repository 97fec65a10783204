"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly the symbols that
include/bsk.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bsk.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bsk_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("bsk_find_knot_interval", "bsk_evaluate_all", "bsk_spline_eval", "bsk_spline_eval_host",
                 "bsk_spline_derivative", "bsk_collocation_matrix", "bsk_banded_lu", "bsk_banded_solve",
                 "bsk_banded_solve_host", "bsk_cyclic_solve", "bsk_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(bk):
    lib = ctypes.CDLL(bk.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", bk.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (bsk_[a-z0-9_]+)", out))
    assert set(declared_symbols()) <= exported
    # nothing undeclared leaks out of the C ABI
    assert exported <= set(declared_symbols()), sorted(exported - set(declared_symbols()))


def test_binding_covers_header(bk):
    from bsplinekit_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_signatures_are_plain_c():
    """No C++ / torch types cross the boundary."""
    src = open(HEADER).read()
    assert "torch" not in src and "std::" not in src and "template" not in src
    assert 'extern "C"' in src


def test_library_is_sm100a(bk):
    out = subprocess.run(["cuobjdump", "-lelf", bk.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", "-arch", "sm_100a", bk.LIB_PATH],
                          capture_output=True, text=True).stdout
    assert "UBLKCP" in sass          # TMA bulk copies in the windowed solve kernel
    assert "DFMA" in sass


def test_version_and_error_text(bk):
    assert "sm_100a" in bk.version()
    from bsplinekit_b200 import _lib
    n = ctypes.c_int64(0)
    st = _lib.lib.bsk_basis_length(None, ctypes.byref(n))
    assert st == -1 and "NULL" in _lib.last_error()
    with pytest.raises(bk.ArgumentError):
        _lib.check(st)


def test_no_cpu_fallback_without_device(bk):
    """On a box without a GPU every compute entry point must refuse to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    assert bk.device_count() == 0
    with pytest.raises(bk.CudaError):
        bk.BSplineBasis(4, np.linspace(0, 1, 8))
    from bsplinekit_b200 import _lib
    out = ctypes.c_void_p()
    kn = np.zeros(8)
    st = _lib.lib.bsk_basis_create(0, 4, 1, kn.ctypes.data_as(ctypes.c_void_p), 8, None, 0, ctypes.byref(out))
    assert st == -3          # BSK_ERR_CUDA
    band = np.eye(3)[None, :, 0].repeat(3, 0).copy(order="F")
    st = _lib.lib.bsk_banded_create_host(band.ctypes.data_as(ctypes.c_void_p), 3, 1, 1, 0, ctypes.byref(out))
    assert st in (-3, -5)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "bsplinekit.jl_b200")
    for dp, _, files in os.walk(pkg):
        if "build" in dp.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh", ".jl")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "README.md", os.path.join(dp, f)


def _julia_ccalls():
    """Every `ccall((:bsk_x, lib), Ret, (ArgTypes...), ...)` of the Julia shim: (symbol, return type, [arg types])."""
    src = open(os.path.join(ROOT, "bsplinekit.jl_b200", "julia", "BSplineKitB200.jl")).read()
    src = re.sub(r"#[^\n]*", "", src)
    out = []
    for m in re.finditer(r"ccall\(\(:(bsk_[a-z0-9_]+),\s*lib\),\s*([A-Za-z0-9_{}]+),\s*\(", src):
        i = m.end()
        depth, j = 1, i
        while depth:                                   # matching parenthesis of the argument-type tuple
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        body = src[i:j - 1]
        args, cur, d = [], "", 0
        for ch in body:                                 # split on top-level commas (Ptr{Ptr{Cvoid}} has none, but be safe)
            if ch in "{(":
                d += 1
            elif ch in "})":
                d -= 1
            if ch == "," and d == 0:
                args.append(cur.strip()); cur = ""
            else:
                cur += ch
        if cur.strip():
            args.append(cur.strip())
        out.append((m.group(1), m.group(2), args))
    return out


def _kind_julia(t):
    if t.startswith("Ptr{") or t in ("Cstring",):
        return "ptr"
    return {"Int32": "i32", "Cint": "i32", "Int64": "i64", "Float64": "f64", "Cdouble": "f64", "Float32": "f32",
            "Csize_t": "size", "UInt64": "size"}[t]


def _kind_ctypes(t):
    if t in (ctypes.c_void_p, ctypes.c_char_p) or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
        return "ptr"
    return {ctypes.c_int32: "i32", ctypes.c_int64: "i64", ctypes.c_double: "f64", ctypes.c_float: "f32",
            ctypes.c_size_t: "size"}[t]


def test_julia_shim_ccalls_match_the_c_abi(bk):
    """The Julia shim cannot be executed here (no Julia toolchain), so its `ccall` signatures are checked
    statically: every symbol exists in include/bsk.h, and arity, argument classes (pointer / Int32 / Int64 /
    Float64 / Float32 / Csize_t) and return type agree with the ctypes table that the tests drive the same
    library through (which test_binding_covers_header ties to the header)."""
    from bsplinekit_b200 import _lib
    calls = _julia_ccalls()
    assert len(calls) >= 30
    declared = set(declared_symbols())
    seen = set()
    for sym, ret, args in calls:
        assert sym in declared, sym
        res, argtypes = _lib.SIGNATURES[sym]
        assert [_kind_julia(a) for a in args] == [_kind_ctypes(a) for a in argtypes], (sym, args)
        assert _kind_julia(ret) == _kind_ctypes(res), (sym, ret)
        seen.add(sym)
    # the hot-path entry points are all bound
    for must in ("bsk_find_knot_interval", "bsk_evaluate_all", "bsk_spline_create", "bsk_spline_eval", "bsk_spline_eval_host",
                 "bsk_collocation_matrix", "bsk_banded_create", "bsk_banded_lu_mode", "bsk_banded_solve",
                 "bsk_cyclic_create", "bsk_cyclic_solve", "bsk_banded_f32_solve", "bsk_spline_eval_multi"):
        assert must in seen, must


def test_julia_shim_caches_handles_per_device():
    src = open(os.path.join(ROOT, "bsplinekit.jl_b200", "julia", "BSplineKitB200.jl")).read()
    assert src.count("get!(_basis_cache, _key(") == 3 and "get!(_spline_cache, _key(S, device))" in src
    assert "get!(_basis_cache, B)" not in src and "get!(_spline_cache, S)" not in src
