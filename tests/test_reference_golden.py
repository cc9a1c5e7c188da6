"""Oracle AND CUDA path against outputs of the real Fortran reference, whenever they exist.

tests/golden/ref/<case>/ is filled by tools/make_reference_golden.sh on a machine with gfortran + netcdf-fortran (the
build image has neither, so the directory is empty there and these tests skip -- parity then rests on the oracle alone,
DESIGN.md section 2).  A compiled reference under baseline/_ref/bin/{cdfmoc,cdfmocsig} is used the same way: the cases
are run on the spot.  Inputs are regenerated from their seeds (tools/reference_cases.py), never stored.

Tolerances are north_star's: masks / basin indexing / sigma-bin assignment bit-exact -- checked through the streamfunction,
which moves by whole cell transports when a bin differs -- and psi within 1e-9 relative or 1e-6 Sv.  The files hold
REAL(4), so the comparison allows one float32 rounding of psi on top.
"""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
from scipy.io import netcdf_file

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
import reference_cases as rc   # noqa: E402

REF = Path(os.environ["CDFTOOLS_REF_GOLDEN"]) if os.environ.get("CDFTOOLS_REF_GOLDEN") else ROOT / "tests" / "golden" / "ref"
BIN = ROOT / "baseline" / "_ref" / "bin"


def _cases():
    have = sorted(p.parent.name for p in REF.glob("*/case.json"))
    if not have and (BIN / "cdfmoc").exists() and (BIN / "cdfmocsig").exists():
        have = sorted(c for c, v in rc.CASES.items() if (BIN / v[3]).exists())
    return have


CASES = _cases()
pytestmark = pytest.mark.skipif(not CASES, reason="no output of the Fortran reference available (tests/golden/ref is empty and "
                                                  "baseline/_ref/bin holds no cdfmoc / cdfmocsig)")


def _reference_file(case, tmp_path):
    """(path of the reference output, directory holding the regenerated inputs)"""
    d = tmp_path / "inputs"
    d.mkdir()
    rc.write_inputs(case, d)
    stored = REF / case / rc.CASES[case][5]
    if stored.exists():
        return stored, d
    tool, argv, fout = rc.CASES[case][3:6]
    r = subprocess.run([str(BIN / tool)] + argv, cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ref = tmp_path / ("ref_" + fout)
    (d / fout).rename(ref)
    return ref, d


def _variables(path):
    f = netcdf_file(str(path), "r", mmap=False)
    out = {k: np.array(v[:]) for k, v in f.variables.items() if k.startswith(("zomsf", "zoiso", "sigtrp", "sigma_class"))}
    f.close()
    return out


def _close(got, ref, what):
    got, ref = got.astype(np.float64), ref.astype(np.float64)
    tol = np.maximum(1e-9 * np.abs(ref), 1e-6) + np.abs(ref) * 2.0 ** -23   # + one REAL(4) rounding of the stored value
    bad = np.abs(got - ref) > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} values differ, max |d| = {np.abs(got - ref).max():.3e} Sv"


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_the_fortran_reference(oracle_mod, tmp_path, case):
    """The CPU restatement (oracle/) fed with the same arrays as the files."""
    from cdftools_b200 import synth
    from util import case_inputs
    ref_path, d = _reference_file(case, tmp_path)
    ref = _variables(ref_path)
    grid, with_basins, nrec, tool, argv, _ = rc.CASES[case]
    m = synth.make_mesh(grid)
    ib, e3m = case_inputs(oracle_mod, m, synth, with_basins)
    vf = netcdf_file(str(d / "gridV.nc"), "r", mmap=False)
    tf = netcdf_file(str(d / "gridT.nc"), "r", mmap=False)
    for r in range(nrec):
        v = np.array(vf.variables["vomecrty"][r], np.float32)[:-1]
        if tool == "cdfmoc" and "-decomp" not in argv:
            e3 = e3m
            if "-full" in argv:
                e3 = oracle_mod.mask_e3v(np.broadcast_to(m.e3t_1d[:, None, None], m.e3v_0.shape).astype(np.float32), m.vmask.astype(np.float32))
            out = oracle_mod.cdfmoc_output(oracle_mod.cdfmoc_record(m.e1v, e3, ib, v))
            names = ["zomsfglo"] + (["zomsfatl", "zomsfinp", "zomsfind", "zomsfpac", "zomsfinp0"] if with_basins else [])
            for iv, n in enumerate(names):
                _close(out[iv], ref[n][r, :, :, 0], f"{case} rec {r} {n}")
        elif tool == "cdfmocsig" and "-isodep" not in argv:
            t = np.array(tf.variables["votemper"][r], np.float32)[:-1]
            s = np.array(tf.variables["vosaline"][r], np.float32)[:-1]
            ntr = "-ntr" in argv
            pref = float(argv[argv.index("-r") + 1]) if "-r" in argv else 0.0
            eos = 2 if ntr else (1 if "-teos10" in argv else 0)
            if all(x in argv for x in ("-sigmin", "-sigstp", "-nbins")):
                smin, sstp, nbins = (float(argv[argv.index("-sigmin") + 1]), float(argv[argv.index("-sigstp") + 1]),
                                     int(argv[argv.index("-nbins") + 1]))
            else:
                nbins, smin, sstp = oracle_mod.default_bins(pref, ntr)
            psi, _ = oracle_mod.cdfmocsig_record(m.e1v, m.e3v_0, ib, v, t, s, 1.0e20, 1.0e20, 1.0e20, pref, eos, smin, sstp, nbins)
            out = oracle_mod.cdfmocsig_output(psi)
            for iv, n in enumerate(["zomsfglo", "zomsfatl", "zomsfinp", "zomsfind", "zomsfpac"][: psi.shape[2]]):
                _close(out[iv], ref[n][r, :, :, 0], f"{case} rec {r} {n}")
        elif tool == "cdfsigtrp":
            from util import sigtrp_expected
            uf = netcdf_file(str(d / "gridU.nc"), "r", mmap=False)
            u = np.array(uf.variables["vozocrtx"][r], np.float32)
            uf.close()
            t = np.array(tf.variables["votemper"][r], np.float32)
            s = np.array(tf.variables["vosaline"][r], np.float32)
            name = rc.CASES[case][5].replace("_trpsig.nc", "")
            kw = dict(refdep=float(argv[argv.index("-refdep") + 1])) if "-refdep" in argv else {}
            smin, smax, nbins = float(argv[argv.index("-smin") + 1]), float(argv[argv.index("-smax") + 1]), int(argv[argv.index("-nbins") + 1])
            vfull = np.array(vf.variables["vomecrty"][r], np.float32)
            # spval: the V file carries 1e20 on land, the U file 0 (no missing-value attribute)
            _, o = sigtrp_expected(oracle_mod, m, u, np.where(vfull == np.float32(1.0e20), np.float32(0), vfull), t, s,
                                   rc.SIGTRP_SECTIONS[name], smin, smax, nbins, spval=1.0e20, **kw)
            _close((o["dtrpbin"] / 1.e6).astype(np.float32), ref["sigtrp"][0, :, 0, 0], f"{case} sigtrp")
            assert np.array_equal(ref["sigma_class"][0, :, 0, 0], o["dsigma_lev"][:nbins].astype(np.float32))
    vf.close()
    tf.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_twin_matches_the_fortran_reference(tmp_path, case):
    """The CUDA path through the command-line twin: same files in, same variables out."""
    from cdftools_b200 import build
    tools = build.build_host()
    ref_path, d = _reference_file(case, tmp_path)
    tool, argv, fout = rc.CASES[case][3:6]
    if tool == "cdfsigtrp":   # no -o: the output files are named after the sections
        r = subprocess.run([str(tools[tool + "_gpu"])] + argv, cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        ref, got = _variables(ref_path), _variables(d / fout)
    else:
        r = subprocess.run([str(tools[tool + "_gpu"])] + argv + ["-o", "gpu_" + fout], cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        ref, got = _variables(ref_path), _variables(d / ("gpu_" + fout))
    assert set(ref) == set(got), (sorted(ref), sorted(got))
    for n in sorted(ref):
        _close(got[n], ref[n], f"{case} {n}")
