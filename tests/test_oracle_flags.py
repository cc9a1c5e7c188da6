"""The C oracle must not depend on the optimiser: the same source built with -O0 gives bit-identical results to the -O3
build the tests and the bench use (no reassociation, no contraction -- SURVEY.md App. A arithmetic rules)."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from cdftools_b200 import synth
from util import case_inputs


@pytest.fixture(scope="module")
def o0_lib(tmp_path_factory, oracle_mod):
    d = tmp_path_factory.mktemp("o0")
    so = d / "libcdforacle_O0.so"
    src = oracle_mod._DIR / "cdf_oracle.c"
    subprocess.run(["/usr/bin/gcc", "-O0", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-std=c11", "-shared",
                    "-o", str(so), str(src), "-lm"], check=True, capture_output=True)
    L = C.CDLL(str(so))
    L.oracle_eos_dlr.restype = C.c_double
    L.oracle_eos_dlr.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
    L.oracle_num_threads.restype = C.c_int
    return L


def _both(oracle_mod, o0_lib, fn):
    a = fn()
    saved = oracle_mod._LIB
    oracle_mod._LIB = o0_lib
    try:
        b = fn()
    finally:
        oracle_mod._LIB = saved
    return a, b


@pytest.mark.parametrize("grid", ["TINY", "ODD"])
def test_o0_equals_o3(oracle_mod, o0_lib, grid):
    oracle_mod.lib()
    m = synth.make_mesh(grid)
    ib, e3m = case_inputs(oracle_mod, m, synth)
    v = synth.make_v_record(m, 2, adversarial=True)[:-1]
    t, s = (x[:-1] for x in synth.make_ts_record(m, 2))
    a, b = _both(oracle_mod, o0_lib, lambda: oracle_mod.cdfmoc_record(m.e1v, e3m, ib, v))
    assert np.array_equal(a, b)
    for pref, eos in ((0.0, 0), (2000.0, 0), (1000.0, 1), (0.0, 2)):
        nb, smin, sstp = oracle_mod.default_bins(pref, eos == 2) if eos != 2 else (60, 1020.0, 0.25)
        a, b = _both(oracle_mod, o0_lib, lambda: oracle_mod.cdfmocsig_record(m.e1v, m.e3v_0, ib, v, t, s, 0.0, 0.0, 0.0, pref, eos,
                                                                            smin, sstp, nb))
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    tt = np.linspace(-2, 32, 4001).astype(np.float32)
    ss = np.linspace(0, 41, 4001).astype(np.float32)
    for teos in (False, True):
        a, b = _both(oracle_mod, o0_lib, lambda: oracle_mod.sigmai_dep(tt, ss, 1500.0, teos))
        assert np.array_equal(a, b)
    a, b = _both(oracle_mod, o0_lib, lambda: oracle_mod.sigmantr(tt, ss))
    assert np.array_equal(a, b)
