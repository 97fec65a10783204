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
