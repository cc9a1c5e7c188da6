"""The fp32 / fp64-FMA tiers of K2's bin function (cdftools_b200/csrc/mocsig_kernel.cuh, constants and error bound from
cdftools_b200/csrc/mocsig_filter.hpp), emulated on the host operation for operation (tests/helpers/filter_emul.cpp):

  * every bin a tier ACCEPTS is the oracle's bin (the reference formula src/cdfmocsig.f90:399-403 on the reference EOS,
    src/eos.f90:842-882) -- what the tiers do not accept goes to the reference chain on the device;
  * the fp32 tier's density stays inside the bound the host derives, whatever the rounding of rsqrt.approx;
  * the fp32 tier accepts the bulk of the cells (otherwise it would be a slow path, not a filter).
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

HERE = Path(__file__).resolve().parent
SRC = HERE / "helpers" / "filter_emul.cpp"
SO = HERE / "helpers" / "libfilter_emul.so"


@pytest.fixture(scope="module")
def emul():
    deps = [SRC, HERE.parent / "cdftools_b200" / "csrc" / "mocsig_filter.hpp", HERE.parent / "include" / "cdf_eos_coeffs.h"]
    if not SO.exists() or any(d.stat().st_mtime > SO.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(SO), str(SRC)], check=True)
    return C.CDLL(str(SO))


def _run(L, teos10, pref, smin, sstp, nbins, T, S, pert):
    info = (C.c_double * 8)()
    L.filter_build(teos10, C.c_float(pref), C.c_float(smin), C.c_float(sstp), nbins, info)
    n = T.size
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ib1, ok1, q1 = np.empty(n, np.int32), np.empty(n, np.uint8), np.empty(n, np.float32)
    ib2, ok2, q2 = np.empty(n, np.int32), np.empty(n, np.uint8), np.empty(n, np.float64)
    L.filter_tier1(p(T), p(S), C.c_long(n), C.c_double(pert), p(ib1), p(ok1), p(q1))
    L.filter_tier2(p(T), p(S), C.c_long(n), p(ib2), p(ok2), p(q2))
    return list(info), ib1, ok1.astype(bool), q1, ib2, ok2.astype(bool), q2


CASES = [  # teos10, pref, (sigmin, sigstp, nbins) or None = the reference's defaults for that depth
    (0, 0.0, None), (0, 1000.0, None), (0, 2000.0, None), (1, 0.0, None), (1, 2000.0, None),
    (0, 0.0, (23.0, 0.05, 104)), (0, 2000.0, (30.0, 0.05, 158)), (0, 3000.0, (35.0, 0.05, 120)), (1, 4000.0, (40.0, 0.02, 400)),
]


@pytest.mark.parametrize("case", CASES)
def test_accepted_bins_are_the_oracles(emul, oracle_mod, case):
    import oracle.np_oracle as npo
    teos10, pref, bins = case
    nbins, smin, sstp = oracle_mod.default_bins(pref, False) if bins is None else (bins[2], bins[0], bins[1])
    rng = np.random.default_rng(11 + int(pref))
    n = 1 << 19
    # the oceanic range, the edges of the fp32 tier's domain box, zeros (land), and values clustered on bin edges
    T = rng.uniform(-4.5, 43.0, n).astype(np.float32)
    S = rng.uniform(0.0, 46.0, n).astype(np.float32)
    S[::17] = 0.0
    T[1::17] = np.float32(-4.0)
    S[2::17] = np.float32(45.0)
    ref, _, _ = npo.mocsig_bins(T.reshape(1, 1, n), S.reshape(1, 1, n), 0.0, pref, teos10, smin, sstp, nbins)
    ref = ref.ravel()
    sig = oracle_mod.sigmai_dep(T.reshape(1, n), S.reshape(1, n), pref, teos10).ravel()
    qref = (sig - float(np.float32(smin))) / float(np.float32(sstp))
    for pert in (0.0, 2.0 ** -22, -2.0 ** -22):
        info, ib1, ok1, q1, ib2, ok2, q2 = _run(emul, teos10, pref, smin, sstp, nbins, T, S, pert)
        assert info[0] == 1.0 and info[1] == 1.0
        assert np.array_equal(ib1[ok1], ref[ok1]), (case, pert, int((ib1[ok1] != ref[ok1]).sum()))
        assert np.array_equal(ib2[ok2], ref[ok2]), (case, pert, int((ib2[ok2] != ref[ok2]).sum()))
        # the bound: |sigma_fp32 - sigma_reference| <= err32 wherever the fp32 tier accepts (q is in bin units)
        err = np.abs(q1.astype(np.float64) - qref)[ok1] * float(np.float32(sstp))
        qround = 2.0 ** -24 * (nbins + 2 + 2 * abs(info[6]) + info[7] / float(np.float32(sstp))) * float(np.float32(sstp))
        assert err.max() <= info[2] + qround, (case, pert, err.max(), info[2])
        assert err.max() < 0.75 * (info[2] + qround)          # and it is not a tight squeeze
        # an fp64 evaluation of the same polynomial differs from the reference chain by rounding only
        wet = S != 0.0
        assert np.abs(q2 - qref)[wet].max() * float(np.float32(sstp)) < 1e-11
        # the filter filters: inside the domain box and the bin range nearly every cell is decided in fp32
        inbox = (T >= -3.9) & (T <= 41.9) & (S >= 1.1) & (S <= 44.9) & (qref > 1.01) & (qref < nbins - 0.01)
        assert ok1[inbox].mean() > 1.0 - 6.0 * info[3]


def test_sigma_stepping_along_a_bin_edge(emul, oracle_mod):
    """Salinities stepping one fp32 ulp at a time across bin edges: the decision flips exactly where the oracle's does."""
    import oracle.np_oracle as npo
    pref, smin, sstp, nbins = 0.0, 23.0, 0.05, 104
    base = np.float32(34.7)
    S = base + np.arange(130000, dtype=np.float32) * np.float32(2.0 ** -18)   # ~0.5 psu in steps of 4e-6
    T = np.full_like(S, np.float32(3.25))
    ref, _, _ = npo.mocsig_bins(T.reshape(1, 1, -1), S.reshape(1, 1, -1), 0.0, pref, 0, smin, sstp, nbins)
    ref = ref.ravel()
    assert len(np.unique(ref)) >= 8 and ref.max() < nbins      # several edges are crossed, none clamped
    info, ib1, ok1, q1, ib2, ok2, q2 = _run(emul, 0, pref, smin, sstp, nbins, T, S, 0.0)
    assert np.array_equal(ib1[ok1], ref[ok1]) and np.array_equal(ib2[ok2], ref[ok2])
    assert (~ok1).sum() > 0 and (~ok2).sum() > 0                # cells on the edges are handed on, not guessed
    assert (~ok1).mean() < 0.01
