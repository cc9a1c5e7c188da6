"""The committed fixtures of tests/golden/ (made by tests/golden/make_golden.py) against the C oracle, its NumPy twin and
-- on the B200 box -- the CUDA path through the C ABI.  The EOS values are the reference's own (comments in src/eos.f90);
the array fixtures are frozen oracle vectors (the Fortran reference cannot be run in this image)."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import np_oracle as npo
from util import assert_psi_close

G = Path(__file__).resolve().parent / "golden"
MOC = sorted(p.name for p in G.glob("moc_*.npz"))
SIG = sorted(p.name for p in G.glob("sig_*.npz"))


def test_fixture_set_is_complete():
    assert len(MOC) == 4 and len(SIG) == 5
    assert (G / "eos_kat.json").exists() and (G / "hand_cdfmoc.json").exists() and (G / "make_golden.py").exists()


def test_eos_known_answers(oracle_mod):
    kat = json.loads((G / "eos_kat.json").read_text())
    for k in kat["reference_comment_values"]:
        if k["eos"] == "NEUTRAL":
            v = oracle_mod.sigmantr(np.array([k["t"]], np.float32), np.array([k["s"]], np.float32))[0] + 1000.0
        else:
            v = oracle_mod.eos_dlr(k["t"], k["s"], k["z"], k["eos"] == "TEOS10")
        assert abs(v - k["rho"]) < 5e-11 * 1028, k["cite"]          # the comments carry 15 significant digits
    for k in kat["derived_sigmai"]:
        t, s = np.array([k["t"]], np.float32), np.array([k["s"]], np.float32)
        assert oracle_mod.sigmai_dep(t, s, k["pref"], k["teos10"])[0] == k["sigma"]
        assert npo.sigmai_dep(t, s, k["pref"], k["teos10"])[0] == k["sigma"]


def hand_case():
    h = json.loads((G / "hand_cdfmoc.json").read_text())
    nx, ny, nz = h["nx"], h["ny"], h["nz"]
    e1v = np.full((ny, nx), h["e1v"], np.float32)
    e3m = np.ones((nz, ny, nx), np.float32) * np.array(h["e3v_levels"], np.float32)[:, None, None]
    ib = np.ones((ny, nx, 1), np.int16)
    ib[:, 0, 0] = ib[:, nx - 1, 0] = 0
    zv = np.zeros((nz - 1, ny, nx), np.float32)
    zv[0, 1], zv[1, 1] = h["zv_row_j2"]["level1"], h["zv_row_j2"]["level2"]
    return h, e1v, e3m, ib, zv


def test_hand_case_oracles(oracle_mod):
    h, e1v, e3m, ib, zv = hand_case()
    for f in (oracle_mod.cdfmoc_record, npo.cdfmoc_record):
        d = f(e1v, e3m, ib, zv)
        assert list(d[:, 1, 0]) == h["psi_sv_row_j2"]
        assert np.all(d[:, [0, 2], :] == 0.0)


@pytest.mark.parametrize("name", MOC)
def test_cdfmoc_fixture_oracles(oracle_mod, name):
    z = np.load(G / name)
    for r, zv in enumerate(z["zv"]):
        for f, fo in ((oracle_mod.cdfmoc_record, oracle_mod.cdfmoc_output), (npo.cdfmoc_record, npo.cdfmoc_output)):
            d = f(z["e1v"], z["e3m"], z["ibmask"], zv)
            assert np.array_equal(d, z["dmoc"][r])
            assert np.array_equal(fo(d), z["out"][r])


def sig_args(z):
    pref, eos, smin, sstp, nbins, spv, spt, sps = z["params"]
    return (z["e1v"], z["e3v"], z["ibmask"], z["zv"], z["zt"], z["zs"], float(spv), float(spt), float(sps),
            float(pref), int(eos), float(np.float32(smin)), float(np.float32(sstp)), int(nbins))


@pytest.mark.parametrize("name", SIG)
def test_cdfmocsig_fixture_oracles(oracle_mod, name):
    z = np.load(G / name)
    a = sig_args(z)
    for faithful in (False, True):
        H, b = oracle_mod.cdfmocsig_record(*a, faithful=faithful)
        assert np.array_equal(b, z["ibin"]) and np.array_equal(H, z["dmoc"])
    H, b = npo.cdfmocsig_record(*a)
    assert np.array_equal(b, z["ibin"]) and np.array_equal(H, z["dmoc"])


# ---------------------------------------------------------------------------------------------- CUDA path (B200 box)
@pytest.mark.gpu
def test_hand_case_gpu(gpu_lib):
    h, e1v, e3m, ib, zv = hand_case()
    nz, ny, nb = gpu_lib.cdfmoc_setup(e1v, e3m, ib)
    d = np.empty((nz, ny, nb), np.float64)
    gpu_lib.cdfmoc_submit(0, 0, zv)
    gpu_lib.cdfmoc_fetch(0, d)
    assert list(d[:, 1, 0]) == h["psi_sv_row_j2"]          # exact: every product and sum is representable
    assert np.all(d[:, [0, 2], :] == 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", MOC)
def test_cdfmoc_fixture_gpu(gpu_lib, name):
    z = np.load(G / name)
    nz, ny, nb = gpu_lib.cdfmoc_setup(z["e1v"], z["e3m"], z["ibmask"])
    for r, zv in enumerate(z["zv"]):
        d = np.empty((nz, ny, nb), np.float64)
        gpu_lib.cdfmoc_submit(r % 3, r, np.ascontiguousarray(zv))
        gpu_lib.cdfmoc_fetch(r % 3, d)
        assert_psi_close(d, z["dmoc"][r], f"{name} rec {r}")


@pytest.mark.gpu
@pytest.mark.parametrize("name", SIG)
def test_cdfmocsig_fixture_gpu(gpu_lib, name):
    import torch
    z = np.load(G / name)
    e1v, e3v, ib, zv, zt, zs, spv, spt, sps, pref, eos, smin, sstp, nbins = sig_args(z)
    ny, nbins, nb = gpu_lib.cdfmocsig_setup(e1v, e3v, ib, e3v.shape[0], nbins, smin, sstp, pref, eos, spv, spt, sps)
    out = np.empty((ny, nbins, nb), np.float64)
    gpu_lib.cdfmocsig_submit(0, 0, zv, zt, zs, None, None)
    gpu_lib.cdfmocsig_fetch(0, out)
    assert_psi_close(out, z["dmoc"], name)
    d_t, d_s = torch.from_numpy(zt).cuda(), torch.from_numpy(zs).cuda()
    d_b = torch.empty(zt.shape, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    gpu_lib.cdfmocsig_bins_device(d_t, d_s, d_b)
    gpu_lib.synchronize()
    assert np.array_equal(d_b.cpu().numpy(), z["ibin"].astype(np.int32)), "sigma-bin assignment must be bit-exact"
