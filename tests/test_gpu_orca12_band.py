"""BASELINE configs 4 and 5 at full ROW size: a latitude band of the ORCA12 L75 grid (4322 x 3059 x 75; rows are
independent in the reference, so a band is an exact sub-problem -- it is also exactly what one rank of the band-sharded
run computes).  Oracle on sampled rows plus the size-independent properties: linearity, idempotence, psi(bottom) = 0,
indo-pacific = indian + pacific, and (density space) band == the rows of a wider band."""
import numpy as np
import pytest

from cdftools_b200 import synth
from util import assert_psi_close, case_inputs

pytestmark = pytest.mark.gpu
J0, J1 = 1338, 1530          # 192 mid-latitude rows: every basin present


@pytest.fixture(scope="module")
def band():
    m = synth.make_mesh("ORCA12", rows=(J0, J1))
    assert (m.nx, m.ny, m.nz) == (4322, J1 - J0, 75)
    return m


def test_cdfmoc_orca12_band(gpu_lib, oracle_mod, band):
    m = band
    ib, e3m = case_inputs(oracle_mod, m, synth)
    v = synth.make_v_record(m, 0, adversarial=True)[:-1]
    nz, ny, nb = gpu_lib.cdfmoc_setup(m.e1v, e3m, ib)
    outs = []
    for r, rec in enumerate((v, (2.0 * v).astype(np.float32), v)):
        o = np.empty((nz, ny, nb))
        gpu_lib.cdfmoc_submit(r % 3, r, np.ascontiguousarray(rec))
        gpu_lib.cdfmoc_fetch(r % 3, o)
        outs.append(o)
    d1, d2, d1b = outs
    rows = np.array([0, 1, 63, 100, 190, 191])
    ref = oracle_mod.cdfmoc_record(m.e1v[rows], e3m[:, rows], ib[rows], v[:, rows])
    err = assert_psi_close(d1[:, rows], ref, "ORCA12 band, sampled rows")
    assert np.array_equal(d2, 2.0 * d1)                  # linear (power-of-two scaling is exact)
    assert np.array_equal(d1, d1b)                       # idempotent / deterministic
    assert np.all(d1[-1] == 0.0)                         # psi(k = nz) = 0
    assert np.allclose(d1[:, :, 2], d1[:, :, 3] + d1[:, :, 4], rtol=0, atol=1e-7)
    print("ORCA12 band max |dpsi| on sampled rows: %.3e Sv (max |psi| %.1f)" % (err, np.abs(ref).max()))


def test_cdfmocsig_orca12_band(gpu_lib, oracle_mod, band):
    m = band
    ib, _ = case_inputs(oracle_mod, m, synth)
    v = synth.make_v_record(m, 1)[:-1]
    t, s = (x[:-1] for x in synth.make_ts_record(m, 1))
    ny_glob = 3059

    def run(j0, j1):
        c = lambda a: np.ascontiguousarray(a[:, j0:j1])
        ny, nbins, nb = gpu_lib.cdfmocsig_setup(np.ascontiguousarray(m.e1v[j0:j1]), c(m.e3v_0), np.ascontiguousarray(ib[j0:j1]),
                                                m.nz, 158, 30.0, 0.05, 2000.0, 0, j_first_global=J0 + j0, ny_global=ny_glob)
        out = np.empty((ny, nbins, nb))
        gpu_lib.cdfmocsig_submit(0, 0, c(v), c(t), c(s))
        gpu_lib.cdfmocsig_fetch(0, out)
        return out

    whole = run(0, m.ny)
    for j in (1, 64, 100, 190):                          # the oracle skips the first and last row it is given
        w = slice(j - 1, j + 2)
        ref, _ = oracle_mod.cdfmocsig_record(m.e1v[w], m.e3v_0[:, w], ib[w], v[:, w], t[:, w], s[:, w], 0.0, 0.0, 0.0,
                                             2000.0, 0, 30.0, 0.05, 158)
        assert_psi_close(whole[j], ref[1], f"ORCA12 band row {j}")
    part = run(40, 120)                                  # a narrower band gives the same rows, bit for bit
    assert np.array_equal(part, whole[40:120])
    assert np.any(whole[0] != 0.0)                       # row J0 is not a global edge row: it is computed
