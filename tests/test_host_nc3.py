"""Host side above the C ABI: the self-contained NetCDF-3 reader/writer and the command-line contracts of the C++
twins (exit codes of the reference: STOP 99 usage/missing file/unknown option).  CPU only."""
import json
import subprocess

import numpy as np
import pytest
from scipy.io import netcdf_file

from cdftools_b200 import build, ncfiles, synth


@pytest.fixture(scope="module")
def tools():
    return build.build_host()


@pytest.mark.parametrize("version", [1, 2])
def test_reader_parses_scipy_files(tools, tmp_path, version):
    m = synth.make_mesh("TINY")
    ncfiles.write_mesh(m, tmp_path, version=version)
    recs = ncfiles.write_gridv(m, tmp_path / "gridV.nc", 3, spval=1.0e20, version=version)
    d = json.loads(subprocess.run([tools["nc3dump"], tmp_path / "gridV.nc"], capture_output=True, text=True, check=True).stdout)
    assert d["version"] == version and d["numrecs"] == 3
    assert d["dims"] == {"x": m.nx, "y": m.ny, "depthv": m.nz, "time_counter": 3}
    v = d["vars"]["vomecrty"]
    allv = np.stack(recs).astype(np.float64)
    assert v["nelem"] == allv.size and v["spval"] == pytest.approx(1.0e20)
    assert v["abssum"] == pytest.approx(np.abs(allv).sum(), rel=1e-12)
    assert d["vars"]["time_counter"]["sum"] == pytest.approx(sum(432000.0 * (r + 0.5) for r in range(3)))
    d = json.loads(subprocess.run([tools["nc3dump"], tmp_path / "mask.nc"], capture_output=True, text=True, check=True).stdout)
    assert d["vars"]["vmask"]["sum"] == float(m.vmask.sum())       # NC_BYTE variables
    d = json.loads(subprocess.run([tools["nc3dump"], tmp_path / "mesh_zgr.nc"], capture_output=True, text=True, check=True).stdout)
    assert d["vars"]["e3v_0"]["sum"] == pytest.approx(float(m.e3v_0.astype(np.float64).sum()), rel=1e-12)


def test_reader_unpacks_like_getvar(tools, tmp_path):
    """Packed variables (NC_SHORT + scale_factor / add_offset) and savelog10: getvar applies them in three passes and leaves
    the values that equal the missing value alone (src/cdfio.F90:1599-1602)."""
    from scipy.io import netcdf_file
    f = netcdf_file(str(tmp_path / "packed.nc"), "w", version=2)
    f.createDimension("x", 6)
    raw = np.array([100, -200, 32767, 0, 15, 32767], np.int16)
    v = f.createVariable("vpacked", "h", ("x",))
    v[:] = raw
    v.scale_factor = np.float32(0.01)
    v.add_offset = np.float32(2.5)
    v.missing_value = np.int16(32767)
    w = f.createVariable("vlog", "f", ("x",))
    w[:] = np.array([0.0, 1.0, -1.0, 9999.0, 2.0, 0.5], np.float32)
    w.savelog10 = np.int32(1)
    w.missing_value = np.float32(9999.0)
    f.close()
    d = json.loads(subprocess.run([tools["nc3dump"], tmp_path / "packed.nc"], capture_output=True, text=True, check=True).stdout)
    x = raw.astype(np.float32)
    keep = x == np.float32(32767.0)
    x = np.where(keep, x, x * np.float32(0.01))
    x = np.where(x == np.float32(32767.0), x, x + np.float32(2.5))
    assert d["vars"]["vpacked"]["sum"] == pytest.approx(float(x.astype(np.float64).sum()), rel=1e-12)
    assert d["vars"]["vpacked"]["spval"] == 32767.0
    y = np.array([1.0, 10.0, 0.1, 9999.0, 100.0, 10 ** 0.5], np.float64)
    assert d["vars"]["vlog"]["sum"] == pytest.approx(float(y.sum()), rel=1e-6)


def test_reader_rejects_non_netcdf(tools, tmp_path):
    p = tmp_path / "x.nc"
    p.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    r = subprocess.run([tools["nc3dump"], p], capture_output=True, text=True)
    assert r.returncode == 98 and "not a NetCDF classic file" in r.stdout


@pytest.mark.parametrize("tool", ["cdfmoc_gpu", "cdfmocsig_gpu"])
def test_cli_contracts(tools, tmp_path, tool):
    # no argument: usage, normal termination (cdfmoc.f90:129-209)
    r = subprocess.run([tools[tool]], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and "usage" in r.stdout
    # unknown option: STOP 99 (cdfmoc.f90:232, cdfmocsig.f90:206)
    r = subprocess.run([tools[tool], "-v", "V.nc", "-bogus"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 99 and "unknown option" in r.stdout
    # missing files: STOP 99 (cdfmoc.f90:246, cdfmocsig.f90:224)
    args = ["-v", "V.nc"] if tool == "cdfmoc_gpu" else ["-v", "V.nc", "-t", "T.nc", "-r", "0"]
    r = subprocess.run([tools[tool]] + args, capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 99 and "is missing" in r.stdout


def test_cdfmocsig_needs_three_mandatory_arguments(tools, tmp_path):
    r = subprocess.run([tools["cdfmocsig_gpu"], "-v", "V.nc", "-t", "T.nc"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 99 and "mandatory arguments missing" in r.stdout


def test_cdftransig_cli_contracts(tools, tmp_path):
    """cdftransig_xy3d_gpu: usage, unknown option / bad code / incomplete '-code none' -> STOP 99
    (src/cdftransig_xy3d.f90:183,200-204), missing mesh files -> 99, missing DRAKKAR file set -> 97 (modutils.f90:111),
    and -- with all files in place but no GPU in this container -- the library's own STOP 97."""
    t = tools["cdftransig_xy3d_gpu"]
    run = lambda *a: subprocess.run([t, *a], capture_output=True, text=True, cwd=tmp_path)
    r = run()
    assert r.returncode == 0 and "usage" in r.stdout
    r = run("-c", "X", "-l", "t1", "-bogus")
    assert r.returncode == 99 and "unkown option" in r.stdout          # the reference's own spelling
    r = run("-c", "X", "-l", "t1", "-code", "3000")
    assert r.returncode == 99 and "is not available" in r.stdout
    r = run("-c", "X", "-l", "t1", "-code", "none", "-depref", "2000")
    assert r.returncode == 99 and "individually" in r.stdout
    r = run("-c", "X", "-l", "t1")
    assert r.returncode == 99 and "is missing" in r.stdout              # mesh_zgr.nc / mesh_hgr.nc
    m = synth.make_mesh("TINY")
    ncfiles.write_mesh(m, tmp_path)
    r = run("-c", "X", "-l", "t1")
    assert r.returncode == 97 and "missing gridV" in r.stdout
    ncfiles.write_gridu(m, tmp_path / "X_t1_gridU.nc", 1)
    ncfiles.write_gridv(m, tmp_path / "X_t1_gridV.nc", 1)
    ncfiles.write_gridt(m, tmp_path / "X_t1_gridT.nc", 1)
    r = run("-c", "X", "-l", "t1", "-v")
    import torch
    if not torch.cuda.is_available():                                    # files parsed, bins printed, then no device
        assert r.returncode == 97 and "NBINS    : 93" in r.stdout and "nbins  = 93" in r.stdout and "libcdfgpu status" in r.stdout
    else:
        assert r.returncode == 0
