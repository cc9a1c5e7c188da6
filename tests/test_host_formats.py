"""The Fortran edit descriptors the command-line twins emulate for their ASCII outputs (trpsig.txt: FORMAT(f9.4, 20e16.7),
src/cdfsigtrp.f90:674; the -print tables: f8.3 / f8.0 / f7.3, :915-918), against what gfortran prints."""
import ctypes as C
import subprocess
from pathlib import Path

import pytest

HERE = Path(__file__).resolve().parent
SRC = HERE / "helpers" / "fmt_check.cpp"
SO = HERE / "helpers" / "libfmt_check.so"


@pytest.fixture(scope="module")
def fmt():
    dep = HERE.parent / "cdftools_b200" / "csrc" / "host" / "host_common.hpp"
    if not SO.exists() or max(SRC.stat().st_mtime, dep.stat().st_mtime) > SO.stat().st_mtime:
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(SO), str(SRC)], check=True)
    L = C.CDLL(str(SO))

    def f(x, w, d):
        b = C.create_string_buffer(64)
        L.fmt_f(C.c_double(x), w, d, b)
        return b.value.decode()

    def e(x):
        b = C.create_string_buffer(64)
        L.fmt_e16_7(C.c_double(x), b)
        return b.value.decode()
    return f, e


def test_e16_7(fmt):
    _, e = fmt
    # gfortran: 0.ddddddd mantissa, two-digit exponent, right-justified in 16
    assert e(0.0) == "   0.0000000E+00"
    assert e(1.0) == "   0.1000000E+01"
    assert e(-16863694.046648856) == "  -0.1686369E+08"
    assert e(123456.75) == "   0.1234568E+06"          # rounded to seven digits
    assert e(0.99999996) == "   0.1000000E+01"         # the rounding carries into the exponent
    assert e(-3.991771e5) == "  -0.3991771E+06"
    assert e(1.5e-7) == "   0.1500000E-06"
    assert all(len(e(x)) == 16 for x in (0.0, 1e30, -1e-30, 7.0))


def test_f_descriptors(fmt):
    f, _ = fmt
    assert f(23.25, 9, 4) == "  23.2500"
    assert f(12.3456, 8, 3) == "  12.346"
    assert f(-0.5, 8, 3) == "  -0.500"
    assert f(0.25, 8, 3) == "   0.250"
    assert f(1234.4, 8, 0) == "   1234."                # F8.0 keeps the decimal point
    assert f(0.0, 8, 0) == "      0."
    assert f(-17.6, 8, 0) == "    -18."
    assert f(123456.789, 8, 3) == "********"            # does not fit
    assert f(-999.9996, 8, 3) == "********"             # -1000.000 is nine characters
    assert f(27.75, 7, 3) == " 27.750"
    assert f(float("nan"), 8, 3) == "     NaN"
    # the optional leading zero goes first when the field is one character short
    assert f(-0.123, 6, 3) == "-0.123" and f(-0.123, 5, 3) == "-.123" and f(0.123, 4, 3) == ".123"
