"""A plain C program linked against libbsplinekit_b200.so through include/bsk.h only — the drop-in
boundary as a C / C++ / Julia-ccall caller sees it (SURVEY §8(b) "what calls it": C driver)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_driver", "c_abi_driver.c")
LIBDIR = os.path.join(ROOT, "bsplinekit.jl_b200", "lib")
ORC = os.path.join(ROOT, "oracle")
EXE = os.path.join(ROOT, "tests", "c_driver", "c_abi_driver")


def build():
    subprocess.check_call(["make", "-s", "-C", ORC])
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-std=c11", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
                           "-L", LIBDIR, "-lbsplinekit_b200", "-L", ORC, "-lbsk_oracle", "-lm",
                           "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + ORC])


def test_c_driver_compiles_against_header_only():
    """The header is plain C and the library needs nothing but itself at link time."""
    build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_c_driver_runs():
    build()
    out = subprocess.run([EXE, str(1 << 22)], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stderr
    assert "C-ABI DRIVER PASS" in out.stdout
