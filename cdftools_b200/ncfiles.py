"""Writes a synthetic configuration as the NetCDF-3 (64-bit offset) files cdfmoc / cdfmocsig read
(SURVEY.md App. B): mesh_hgr.nc, mesh_zgr.nc (v3.6 naming, cdfio.F90:3323-3331), mask.nc, new_maskglo.nc and the
gridV / gridT record files.  Uses scipy.io.netcdf_file (the only NetCDF writer in this image); reading them back in
the product is done by the self-contained C++ reader (csrc/host/nc3.hpp)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
from scipy.io import netcdf_file

from . import synth


def _new(path, dims, version=2):
    f = netcdf_file(str(path), "w", version=version)
    for n, l in sorted(dims.items(), key=lambda kv: kv[1] is not None):   # scipy wants the unlimited dimension first
        f.createDimension(n, l)
    return f


def e3w_fields(m: synth.Mesh):
    """(e3w_1d (nz), e3w (nz,ny,nx)) float32 for the synthetic mesh: the distance between T levels, thinned like e3v in the
    partial bottom cells (cdfsigtrp reads them, src/cdfsigtrp.f90:418-425)."""
    e3w_1d = np.empty(m.nz, np.float32)
    e3w_1d[0] = 2.0 * m.gdept_1d[0]
    e3w_1d[1:] = m.gdept_1d[1:] - m.gdept_1d[:-1]
    ratio = np.clip(m.e3v_0 / m.e3t_1d[:, None, None].astype(np.float32), 0.05, 1.0)
    return e3w_1d, (e3w_1d[:, None, None] * ratio).astype(np.float32)


def write_mesh(m: synth.Mesh, outdir, with_basins=True, version=2, zgr="v3.6"):
    """zgr: the mesh_zgr flavour SetMeshZgrVersion tells apart (src/cdfio.F90:3310-3335): "v3.6" (e3v_0, *_1d), "v3.0"
    (e3v, gdepw_0 / gdept_0 / e3t_0 as (t,z) columns) or "v2.0" (e3v_ps, gdepw / gdept / e3t as (t,z,1,1))."""
    out = Path(outdir)
    out.mkdir(parents=True, exist_ok=True)
    nz, ny, nx = m.e3v_0.shape
    f = _new(out / "mesh_hgr.nc", {"x": nx, "y": ny, "t": None}, version)
    e2v = (m.e1v * np.float32(0.75)).astype(np.float32)   # the synthetic mesh carries no e2v: any positive field will do
    e2u = (m.e1u * np.float32(0.9)).astype(np.float32)    # likewise e2u (cdftransig_xy3d)
    for name, arr in (("e1v", m.e1v), ("e1u", m.e1u), ("e2v", e2v), ("e2u", e2u), ("gphiv", m.gphiv), ("glamv", m.glamv)):
        v = f.createVariable(name, "f", ("t", "y", "x"))
        v[0] = arr
    for name, arr in (("nav_lon", m.glamv), ("nav_lat", m.gphiv)):   # as every NEMO mesh file carries them
        v = f.createVariable(name, "f", ("y", "x"))
        v[:] = arr
    f.close()
    f = _new(out / "mesh_zgr.nc", {"x": nx, "y": ny, "z": nz, "t": None, **({"x_a": 1, "y_a": 1} if zgr == "v2.0" else {})}, version)
    e3u = (m.e3v_0 * np.float32(1.01)).astype(np.float32)
    e3w_1d, e3w = e3w_fields(m)
    v = f.createVariable({"v3.6": "e3w_0", "v3.0": "e3w", "v2.0": "e3w_ps"}[zgr], "f", ("t", "z", "y", "x")); v[0] = e3w
    if zgr == "v2.0":
        v = f.createVariable("e3w", "f", ("t", "z", "y_a", "x_a")); v[0, :, 0, 0] = e3w_1d
    else:
        v = f.createVariable({"v3.6": "e3w_1d", "v3.0": "e3w_0"}[zgr], "f", ("t", "z")); v[0] = e3w_1d
    if zgr == "v3.6":
        v = f.createVariable("e3v_0", "f", ("t", "z", "y", "x")); v[0] = m.e3v_0
        v = f.createVariable("e3u_0", "f", ("t", "z", "y", "x")); v[0] = e3u
        v = f.createVariable("e3t_0", "f", ("t", "z", "y", "x")); v[0] = m.e3v_0     # presence + rank select 'v3.6'
        for name, arr in (("gdepw_1d", m.gdepw_1d), ("gdept_1d", m.gdept_1d), ("e3t_1d", m.e3t_1d)):
            v = f.createVariable(name, "f", ("t", "z")); v[0] = arr
    elif zgr == "v3.0":
        v = f.createVariable("e3v", "f", ("t", "z", "y", "x")); v[0] = m.e3v_0
        v = f.createVariable("e3u", "f", ("t", "z", "y", "x")); v[0] = e3u
        for name, arr in (("gdepw_0", m.gdepw_1d), ("gdept_0", m.gdept_1d), ("e3t_0", m.e3t_1d)):   # e3t_0 of rank 1: 'v3.0'
            v = f.createVariable(name, "f", ("t", "z")); v[0] = arr
    else:
        v = f.createVariable("e3v_ps", "f", ("t", "z", "y", "x")); v[0] = m.e3v_0
        v = f.createVariable("e3u_ps", "f", ("t", "z", "y", "x")); v[0] = e3u
        for name, arr in (("gdepw", m.gdepw_1d), ("gdept", m.gdept_1d), ("e3t", m.e3t_1d)):        # no e3t_0 at all: 'v2.0'
            v = f.createVariable(name, "f", ("t", "z", "y_a", "x_a")); v[0, :, 0, 0] = arr
    f.close()
    f = _new(out / "mask.nc", {"x": nx, "y": ny, "z": nz, "t": None}, version)
    for name, arr in (("vmask", m.vmask), ("tmask", m.tmask), ("umask", m.umask)):
        v = f.createVariable(name, "b", ("t", "z", "y", "x")); v[0] = arr
    f.close()
    if with_basins:
        f = _new(out / "new_maskglo.nc", {"x": nx, "y": ny}, version)
        for name, arr in (("tmaskatl", m.tmaskatl), ("tmaskind", m.tmaskind), ("tmaskpac", m.tmaskpac)):
            v = f.createVariable(name, "f", ("y", "x")); v[:] = arr
        f.close()


def write_gridv(m: synth.Mesh, path, nrec, spval=0.0, adversarial=False, eiv=False, version=2, first=0, vname="vomecrty"):
    nz, ny, nx = m.e3v_0.shape
    f = _new(path, {"x": nx, "y": ny, "depthv": nz, "time_counter": None}, version)
    f.start_date = np.int32(20260101)
    f.output_frequency = "5d"
    f.CONFIG = m.meta.get("grid", "SYN")
    f.CASE = "B200"
    tc = f.createVariable("time_counter", "d", ("time_counter",))
    tc.units = "seconds since 2026-01-01 00:00:00"
    v = f.createVariable(vname, "f", ("time_counter", "depthv", "y", "x"))
    v.missing_value = np.float32(spval)
    e = None
    if eiv:
        e = f.createVariable("vomeeivv", "f", ("time_counter", "depthv", "y", "x"))
        e.missing_value = np.float32(spval)
    recs = []
    for r in range(nrec):
        tc[r] = 432000.0 * (r + 0.5)
        rec = synth.make_v_record(m, first + r, adversarial=adversarial, spval=spval if spval else None)
        v[r] = rec
        recs.append(rec)
        if e is not None:
            e[r] = (0.01 * synth.make_v_record(m, 1000 + r)).astype(np.float32)
    f.close()
    return recs


def write_gridt(m: synth.Mesh, path, nrec, spval=0.0, version=2, first=0):
    nz, ny, nx = m.e3v_0.shape
    f = _new(path, {"x": nx, "y": ny, "deptht": nz, "time_counter": None}, version)
    tc = f.createVariable("time_counter", "d", ("time_counter",))
    dep = f.createVariable("deptht", "f", ("deptht",))
    dep[:] = m.gdept_1d.astype(np.float32)
    t = f.createVariable("votemper", "f", ("time_counter", "deptht", "y", "x"))
    s = f.createVariable("vosaline", "f", ("time_counter", "deptht", "y", "x"))
    t.missing_value = np.float32(spval)
    s.missing_value = np.float32(spval)
    recs = []
    for r in range(nrec):
        tc[r] = 432000.0 * (first + r + 0.5)
        tt, ss = synth.make_ts_record(m, first + r, spval=spval if spval else None)
        t[r] = tt
        s[r] = ss
        recs.append((tt, ss))
    f.close()
    return recs


def write_vt(m: synth.Mesh, path, nrec, version=2):
    """VT file as cdfvT writes it: vomevt, vomevs (time_counter, depthv, y, x) f32 (only the V components are needed)."""
    nz, ny, nx = m.e3v_0.shape
    f = _new(path, {"x": nx, "y": ny, "depthv": nz, "time_counter": None}, version)
    tc = f.createVariable("time_counter", "d", ("time_counter",))
    vt = f.createVariable("vomevt", "f", ("time_counter", "depthv", "y", "x"))
    vs = f.createVariable("vomevs", "f", ("time_counter", "depthv", "y", "x"))
    recs = []
    for r in range(nrec):
        tc[r] = 432000.0 * (r + 0.5)
        v = synth.make_v_record(m, r)
        t, s = synth.make_ts_record(m, r)
        a, b = (v * t).astype(np.float32), (v * s).astype(np.float32)
        vt[r] = a
        vs[r] = b
        recs.append((a, b))
    f.close()
    return recs


def write_gridu(m: synth.Mesh, path, nrec, first=0, version=2):
    """gridU file: vozocrtx (time_counter, depthu, y, x) f32, zero on land (umask)."""
    nz, ny, nx = m.e3v_0.shape
    f = _new(path, {"x": nx, "y": ny, "depthu": nz, "time_counter": None}, version)
    tc = f.createVariable("time_counter", "d", ("time_counter",))
    u = f.createVariable("vozocrtx", "f", ("time_counter", "depthu", "y", "x"))
    recs = []
    for r in range(nrec):
        tc[r] = 432000.0 * (first + r + 0.5)
        rec = (synth.make_v_record(m, 500 + first + r) * m.umask).astype(np.float32)
        u[r] = rec
        recs.append(rec)
    f.close()
    return recs
