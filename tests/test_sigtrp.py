"""cdfsigtrp (src/cdfsigtrp.f90:428-627): the C oracle against its independent NumPy twin bit for
bit, a hand-computed column, conservation properties, and -- on the B200 box -- kernel K7 through the C ABI against the
oracle (every output array bit for bit)."""
import numpy as np
import pytest

from cdftools_b200 import lib as cdflib
from cdftools_b200 import synth
from oracle import np_oracle as npo

# (mode, refdep, teos10, sigma_min, sigma_max, nbins)
SETTINGS = [(0, 0.0, False, 24.0, 28.2, 42), (0, 2000.0, False, 33.0, 37.5, 45), (1, 0.0, False, 24.0, 28.5, 30),
            (2, 0.0, False, -26.0, -1.0, 25), (0, 0.0, True, 23.5, 28.5, 17)]
SECTIONS = [(1, 6), (37, 31), (300, 31), (1200, 75)]


def prepared(oracle_mod, npts, npk, merid=False, spval=0.0, seed=3):
    s = synth.make_section(npts, npk, seed=seed, spval=spval)
    p = oracle_mod.sigtrp_prepare(s["gdept"][0], s["e3w_a"], s["e3w_b"], s["zu"], spval, s["zs_a"], s["zs_b"], spval, s["zt_a"],
                                  s["zt_b"], merid=merid)
    return s, p


def same(a, b):
    return all(np.array_equal(a[k], b[k], equal_nan=True) for k in ("dsigma_lev", "dsig", "dhiso", "dwtrp", "dwtrpbin", "dtrpbin"))


@pytest.mark.parametrize("merid", [False, True])
@pytest.mark.parametrize("spval", [0.0, 9999.0])
def test_prepare_c_equals_numpy(oracle_mod, merid, spval):
    s = synth.make_section(53, 31, seed=5, spval=spval)
    args = (s["gdept"][0], s["e3w_a"], s["e3w_b"], s["zu"], spval, s["zs_a"], s["zs_b"], spval, s["zt_a"], s["zt_b"])
    a, b = oracle_mod.sigtrp_prepare(*args, merid=merid), npo.sigtrp_prepare(*args, merid=merid)
    assert a["nk"] == b["nk"] and a["found"] == b["found"] and a["found"]
    for k in ("ddepu", "zu", "zs", "zt", "zmask"):
        assert np.array_equal(a[k], b[k]), k
    assert a["zmask"].min() == 0.0 and a["zmask"].max() == 1.0
    if spval != 0.0 and not merid:   # the zonal branch does not mask the mean temperature (:555)
        assert np.any(a["zt"][a["zmask"] == 0] != 0)


@pytest.mark.parametrize("setting", SETTINGS)
@pytest.mark.parametrize("sec", SECTIONS[:3])
def test_section_c_equals_numpy(oracle_mod, setting, sec):
    mode, refdep, teos10, smin, smax, nbins = setting
    s, p = prepared(oracle_mod, *sec)
    args = (s["eu"], s["de3"], p["ddepu"], s["gdepw"], p["zu"], p["zt"], p["zs"], p["zmask"], p["nk"], smin, smax, nbins)
    a = oracle_mod.sigtrp_section(*args, mode=mode, refdep=refdep, teos10=teos10)
    b = npo.sigtrp_section(*args, mode=mode, refdep=refdep, teos10=teos10)
    assert same(a, b)
    assert np.any(a["dtrpbin"] != 0) or sec[0] == 1


def test_hand_computed_column(oracle_mod):
    # one column, 4 levels, -temp mode (dsig = -T): T = 10, 6, 2, land.  Class limits -11, -7, -3 (2 bins).
    # gdept = 5, 15, 25, 35 -> ddepu = 0 | 5, 15, 25, 35 (e3w = 10); gdepw = 0, 10, 20, 30; de3 = 10; eu = 1000; u = 0.1, 0.2, 0.3.
    npk = 4
    zt = np.array([[10.0], [6.0], [2.0], [0.0]], np.float32)
    zs = np.array([[35.0], [35.0], [35.0], [0.0]], np.float32)
    zmask = np.array([[1.0], [1.0], [1.0], [0.0]], np.float32)
    zu = np.array([[0.1], [0.2], [0.3], [0.0]], np.float32)
    ddepu = np.array([[0.0], [5.0], [15.0], [25.0], [35.0]])
    gdepw = np.array([0.0, 10.0, 20.0, 30.0], np.float32)
    o = oracle_mod.sigtrp_section(np.array([1000.0], np.float32), np.full((npk, 1), 10.0, np.float32), ddepu, gdepw, zu, zt, zs,
                                  zmask, 4, -11.0, -3.0, 2, mode=2)
    assert np.array_equal(o["dsigma_lev"], [-11.0, -7.0, -3.0])
    # dsig = -10.0001 | -10, -6, -2, -2 + 1e-5.  Level -11: first dsig >= -11 is jk=1, alfa < 0 -> depth 0.
    # Level -7: between -10 (5 m) and -6 (15 m): alfa = 0.75 -> 12.5 m.  Level -3: between -6 (15 m) and -2 (25 m): 22.5 m.
    assert np.allclose(o["dhiso"][:, 0], [0.0, 12.5, 22.5], rtol=0, atol=1e-12)
    u = zu[:, 0].astype(np.float64)
    w1 = 1000.0 * 10.0 * u[0] + 1000.0 * (12.5 - 10.0) * u[1]
    w2 = 1000.0 * 10.0 * u[0] + 1000.0 * 10.0 * u[1] + 1000.0 * (22.5 - 20.0) * u[2]
    assert np.allclose(o["dwtrp"][:, 0], [0.0, w1, w2], rtol=1e-15)
    assert np.allclose(o["dtrpbin"], [w1, w2 - w1], rtol=1e-15)


def test_bins_sum_to_the_section_transport(oracle_mod):
    # class limits wider than every density: the bins hold the whole transport above the deepest velocity level used
    s, p = prepared(oracle_mod, 300, 31)
    o = oracle_mod.sigtrp_section(s["eu"], s["de3"], p["ddepu"], s["gdepw"], p["zu"], p["zt"], p["zs"], p["zmask"], p["nk"], -5.0,
                                  60.0, 13)
    nk = p["nk"]
    total = (s["eu"].astype(np.float64)[None] * s["de3"][: nk - 1].astype(np.float64) * p["zu"][: nk - 1].astype(np.float64)).sum()
    assert abs(o["dtrpbin"].sum() - total) <= 1e-9 * np.abs(o["dwtrpbin"]).sum()
    assert np.all(o["dhiso"][0] == 0.0) and np.array_equal(o["dhiso"][-1], p["ddepu"][-1])
    assert np.all(np.diff(o["dhiso"], axis=0) >= 0)   # deeper limits lie deeper (the land fill keeps columns increasing)


def test_broken_line_depths(oracle_mod):
    # -brk: the w depths come per column (:608); identical columns of gdepw reproduce the 1-D result
    s, p = prepared(oracle_mod, 37, 31)
    args = (s["eu"], s["de3"], p["ddepu"], s["gdepw"], p["zu"], p["zt"], p["zs"], p["zmask"], p["nk"], 24.0, 28.2, 42)
    a = oracle_mod.sigtrp_section(*args)
    brk = np.repeat(s["gdepw"][:, None], 37, axis=1)
    b = oracle_mod.sigtrp_section(*args, ddepw_brk=brk)
    assert same(a, b) and same(b, npo.sigtrp_section(*args, ddepw_brk=brk))


# ---- B200: kernel K7 through the C ABI -----------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("setting", SETTINGS)
@pytest.mark.parametrize("sec", SECTIONS)
def test_gpu_section_matches_oracle(oracle_mod, gpu_lib, setting, sec):
    mode, refdep, teos10, smin, smax, nbins = setting
    s, p = prepared(oracle_mod, *sec, merid=(sec[0] % 2 == 0))
    args = (s["eu"], s["de3"], p["ddepu"], s["gdepw"], p["zu"], p["zt"], p["zs"], p["zmask"], p["nk"], smin, smax, nbins)
    ref = oracle_mod.sigtrp_section(*args, mode=mode, refdep=refdep, teos10=teos10)
    got = cdflib.cdfsigtrp_section(*args, mode=mode, refdep=refdep, teos10=teos10)
    for k in ("dsigma_lev", "dsig", "dhiso", "dwtrp", "dwtrpbin", "dtrpbin"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    assert cdflib.cdfsigtrp_kernel_ms() > 0


@pytest.mark.gpu
def test_gpu_section_nan_and_broken_line(oracle_mod, gpu_lib):
    s, p = prepared(oracle_mod, 300, 31, spval=9999.0)
    zt = p["zt"].copy()
    zt[3, 17] = np.nan
    zt[0, 40] = np.inf
    brk = np.repeat(s["gdepw"][:, None], 300, axis=1) * np.linspace(0.9, 1.1, 300, dtype=np.float32)[None]
    args = (s["eu"], s["de3"], p["ddepu"], None, p["zu"], zt, p["zs"], p["zmask"], p["nk"], 24.0, 28.2, 42)
    ref = oracle_mod.sigtrp_section(s["eu"], s["de3"], p["ddepu"], s["gdepw"], p["zu"], zt, p["zs"], p["zmask"], p["nk"], 24.0, 28.2,
                                    42, ddepw_brk=brk)
    got = cdflib.cdfsigtrp_section(*args, ddepw_brk=brk)
    for k in ("dsig", "dhiso", "dwtrp", "dwtrpbin", "dtrpbin"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k


@pytest.mark.gpu
def test_gpu_section_argument_errors(gpu_lib):
    z = np.zeros((4, 3), np.float32)
    with pytest.raises(cdflib.CdfGpuError):
        cdflib.cdfsigtrp_section(np.ones(3, np.float32), z, np.zeros((5, 3)), np.zeros(4, np.float32), z, z, z, z, 9, 20.0, 30.0, 4)
    with pytest.raises(cdflib.CdfGpuError):
        cdflib.cdfsigtrp_section(np.ones(3, np.float32), z, np.zeros((5, 3)), np.zeros(4, np.float32), z, z, z, z, 4, 20.0, 30.0, 4, mode=7)
