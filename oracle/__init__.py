"""ctypes front-end of the CPU oracle (oracle/cdf_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
The product package (cdftools_b200) must never import this module; tests/test_no_oracle_in_product.py checks.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB = None

EOS80, TEOS10, NEUTRAL = 0, 1, 2


def build(force: bool = False) -> Path:
    so = _DIR / "libcdforacle.so"
    src = _DIR / "cdf_oracle.c"
    hdr = _DIR.parent / "include" / "cdf_eos_coeffs.h"
    if force or not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["make", "-C", str(_DIR), "-B", "libcdforacle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = _DIR / "libcdforacle.so"
        if not so.exists():
            build()
        _LIB = C.CDLL(str(so))
        _LIB.oracle_eos_dlr.restype = C.c_double
        _LIB.oracle_eos_dlr.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
        _LIB.oracle_num_threads.restype = C.c_int
        _LIB.oracle_set_num_threads.restype = None
        _LIB.oracle_set_num_threads.argtypes = [C.c_int]
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def set_num_threads(n: int) -> None:
    """OpenMP threads of the following calls (a launcher may have exported OMP_NUM_THREADS=1, as torchrun does)."""
    lib().oracle_set_num_threads(int(n))


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def eos_dlr(t, s, depth, teos10=False) -> float:
    return float(lib().oracle_eos_dlr(float(t), float(s), float(depth), int(teos10)))


def sigmai_dep(t, s, pref, teos10=False):
    t, s = _f32(t), _f32(s)
    out = np.empty(t.shape, np.float64)
    lib().oracle_sigmai_dep(C.c_size_t(t.size), _p(t, C.c_float), _p(s, C.c_float), C.c_float(pref), int(teos10),
                            _p(out, C.c_double))
    return out


def sigmantr(t, s):
    t, s = _f32(t), _f32(s)
    out = np.empty(t.shape, np.float64)
    lib().oracle_sigmantr(C.c_size_t(t.size), _p(t, C.c_float), _p(s, C.c_float), _p(out, C.c_double))
    return out


def basin_masks(vmask1, atl=None, ind=None, pac=None, zero_edges=True):
    vmask1 = _f32(vmask1)
    ny, nx = vmask1.shape
    nb = 5 if atl is not None else 1
    out = np.zeros((ny, nx, nb), np.int16)
    a, i, p = (_f32(x) if x is not None else None for x in (atl, ind, pac))
    lib().oracle_basin_masks(nx, ny, nb, _p(vmask1, C.c_float), _p(a, C.c_float), _p(i, C.c_float), _p(p, C.c_float),
                             int(zero_edges), _p(out, C.c_int16))
    return out


def mask_e3v(e3v_file, vmask):
    e, v = _f32(e3v_file), _f32(vmask)
    out = np.empty(e.shape, np.float32)
    lib().oracle_mask_e3v(C.c_size_t(e.size), _p(e, C.c_float), _p(v, C.c_float), _p(out, C.c_float))
    return out


def cdfmoc_record(e1v, e3m, ibmask, zv):
    """-> dmoc (nz, ny, nb) float64, scanned (Sv).  cdfmoc.f90:352-388."""
    e1v, e3m, zv = _f32(e1v), _f32(e3m), _f32(zv)
    ibmask = np.ascontiguousarray(ibmask, np.int16)
    nz, ny, nx = e3m.shape
    nb = ibmask.shape[2]
    assert zv.shape == (nz - 1, ny, nx) and e1v.shape == (ny, nx) and ibmask.shape[:2] == (ny, nx)
    out = np.empty((nz, ny, nb), np.float64)
    lib().oracle_cdfmoc_record(nx, ny, nz, nb, _p(e1v, C.c_float), _p(e3m, C.c_float), _p(ibmask, C.c_int16),
                               _p(zv, C.c_float), _p(out, C.c_double))
    return out


def cdfmoc_output(dmoc):
    nz, ny, nb = dmoc.shape
    nv = nb + (1 if nb >= 5 else 0)
    out = np.empty((nv, nz, ny), np.float32)
    d = np.ascontiguousarray(dmoc, np.float64)
    lib().oracle_cdfmoc_output(ny, nz, nb, _p(d, C.c_double), _p(out, C.c_float))
    return out


def sigma_axis(nbins, sigmin, sigstp):
    out = np.empty(nbins, np.float32)
    lib().oracle_sigma_axis(int(nbins), C.c_float(sigmin), C.c_float(sigstp), _p(out, C.c_float))
    return out


def default_bins(pref, lntr=False):
    nb, smin, sstp = C.c_int(), C.c_float(), C.c_float()
    rc = lib().oracle_default_bins(C.c_float(pref), int(lntr), C.byref(nb), C.byref(smin), C.byref(sstp))
    if rc:
        raise SystemExit(rc)
    return nb.value, smin.value, sstp.value


def cdfmocsig_record(e1v, e3v, ibmask, zv, zt, zs, spv, spt, sps, pref, eos, sigmin, sigstp, nbins, zveiv=None,
                     faithful=False, want_bins=True):
    """-> (dmoc (ny, nbins, nb) float64 scanned, ibin (nz-1, ny, nx) int32 or None).  cdfmocsig.f90:366-475."""
    e1v, e3v, zv, zt, zs = (_f32(x) for x in (e1v, e3v, zv, zt, zs))
    ibmask = np.ascontiguousarray(ibmask, np.int16)
    nzm1, ny, nx = zv.shape
    nz = nzm1 + 1
    nb = ibmask.shape[2]
    assert e3v.shape[0] >= nzm1 and e3v.shape[1:] == (ny, nx)
    ze = _f32(zveiv) if zveiv is not None else None
    out = np.empty((ny, nbins, nb), np.float64)
    bins = np.empty((nzm1, ny, nx), np.int32) if want_bins else None
    lib().oracle_cdfmocsig_record(nx, ny, nz, nb, int(nbins), C.c_float(sigmin), C.c_float(sigstp), C.c_float(pref),
                                  int(eos), _p(e1v, C.c_float), _p(e3v, C.c_float), _p(ibmask, C.c_int16),
                                  C.c_float(spv), C.c_float(spt), C.c_float(sps), _p(zv, C.c_float),
                                  _p(zt, C.c_float), _p(zs, C.c_float), _p(ze, C.c_float), int(faithful),
                                  _p(out, C.c_double), _p(bins, C.c_int32))
    return out, bins


def cdfmocsig_output(dmoc):
    ny, nbins, nb = dmoc.shape
    out = np.empty((nb, nbins, ny), np.float32)
    d = np.ascontiguousarray(dmoc, np.float64)
    lib().oracle_cdfmocsig_output(ny, nb, nbins, _p(d, C.c_double), _p(out, C.c_float))
    return out


def maxmoc(rmoc, ijmin, ijmax, ikmin, ikmax):
    """rmoc (nz, ny) float32 -> (ovtmax, ovtmin, (jj,jk)max, (jj,jk)min), 1-based (cdfmaxmoc.f90:158-167)."""
    r = _f32(rmoc)
    ovt = (C.c_float * 2)()
    loc = (C.c_int * 4)()
    lib().oracle_maxmoc(r.shape[1], _p(r, C.c_float), ijmin, ijmax, ikmin, ikmax, ovt, loc)
    return ovt[0], ovt[1], (loc[0], loc[1]), (loc[2], loc[3])


def cdfmocsig_record_isodep(e1v, e3v, ibmask, gdept, zv, zt, zs, spv, spt, sps, pref, eos, sigmin, sigstp, nbins):
    """-> (dmoc, depi), both (ny, nbins, nb) float64.  cdfmocsig.f90:366-475 with -isodep (:423-469)."""
    e1v, e3v, zv, zt, zs, gdept = (_f32(x) for x in (e1v, e3v, zv, zt, zs, gdept))
    ibmask = np.ascontiguousarray(ibmask, np.int16)
    nzm1, ny, nx = zv.shape
    nb = ibmask.shape[2]
    out = np.empty((ny, nbins, nb), np.float64)
    depi = np.empty((ny, nbins, nb), np.float64)
    lib().oracle_cdfmocsig_record_isodep(nx, ny, nzm1 + 1, nb, int(nbins), C.c_float(sigmin), C.c_float(sigstp),
                                         C.c_float(pref), int(eos), _p(e1v, C.c_float), _p(e3v, C.c_float),
                                         _p(ibmask, C.c_int16), C.c_float(spv), C.c_float(spt), C.c_float(sps),
                                         _p(gdept, C.c_float), _p(zv, C.c_float), _p(zt, C.c_float), _p(zs, C.c_float),
                                         _p(out, C.c_double), _p(depi, C.c_double))
    return out, depi


def cdfmoc_decomp_record(e1v, e1u, gphiv, gdept, e3m, ibmask, umask, tmask, zv, zt, zs, teos10=False, dmoc_sh_in=None):
    """cdfmoc -decomp (cdfmoc.f90:390-517) -> dict of (nz,ny,nb) float64 arrays: total, sh, bt, ag."""
    e1v, e1u, gphiv, gdept, e3m, zv, zt, zs = (_f32(x) for x in (e1v, e1u, gphiv, gdept, e3m, zv, zt, zs))
    ibmask = np.ascontiguousarray(ibmask, np.int16)
    um, tm = np.ascontiguousarray(umask, np.int16), np.ascontiguousarray(tmask, np.int16)
    nz, ny, nx = e3m.shape
    nb = ibmask.shape[2]
    outs = {k: np.zeros((nz, ny, nb), np.float64) for k in ("total", "sh", "bt", "ag")}
    if dmoc_sh_in is not None:
        outs["sh"][...] = dmoc_sh_in
    lib().oracle_cdfmoc_decomp_record(nx, ny, nz, nb, int(teos10), _p(e1v, C.c_float), _p(e1u, C.c_float),
                                      _p(gphiv, C.c_float), _p(gdept, C.c_float), _p(e3m, C.c_float),
                                      _p(ibmask, C.c_int16), _p(um, C.c_int16), _p(tm, C.c_int16), _p(zv, C.c_float),
                                      _p(zt, C.c_float), _p(zs, C.c_float), _p(outs["total"], C.c_double),
                                      _p(outs["sh"], C.c_double), _p(outs["bt"], C.c_double), _p(outs["ag"], C.c_double))
    return outs


# ---- sibling tools ---------------------------------------------------------------------
def zonal_masks(msk1, atl=None, ind=None, pac=None):
    """-> zmask (ny, nx, nb) float32, nb = 5 with basins else 1.  cdfzonalsum.f90:286-296."""
    msk1 = _f32(msk1)
    ny, nx = msk1.shape
    nb = 5 if atl is not None else 1
    out = np.zeros((ny, nx, nb), np.float32)
    a, i, p_ = (_f32(x) if x is not None else None for x in (atl, ind, pac))
    lib().oracle_zonal_masks(nx, ny, nb, _p(msk1, C.c_float), _p(a, C.c_float) if nb == 5 else None,
                             _p(i, C.c_float) if nb == 5 else None, _p(p_, C.c_float) if nb == 5 else None,
                             _p(out, C.c_float))
    return out


def zonal_dlsurf(e1, e2):
    e1, e2 = _f32(e1), _f32(e2)
    out = np.empty(e1.shape, np.float64)
    lib().oracle_zonal_dlsurf(C.c_size_t(e1.size), _p(e1, C.c_float), _p(e2, C.c_float), _p(out, C.c_double))
    return out


def zonalsum_record(zmask, zmaskvar, zv, dl_surf, alpha=None):
    """-> dzosum (nb, nk, ny) float64: every level of one record.  cdfzonalsum.f90:309-322."""
    zmask, zmaskvar, zv = _f32(zmask), _f32(zmaskvar), _f32(zv)
    dl = np.ascontiguousarray(dl_surf, np.float64)
    nk, ny, nx = zv.shape
    nb = zmask.shape[2]
    al = _f32(alpha) if alpha is not None else None
    out = np.empty((nb, nk, ny), np.float64)
    tmp = np.empty((nb, ny), np.float64)
    for k in range(nk):
        lib().oracle_zonalsum_level(nx, ny, nb, _p(zmask, C.c_float), _p(zmaskvar[k], C.c_float), _p(zv[k], C.c_float),
                                    _p(dl, C.c_double), _p(al, C.c_float) if al is not None else None, _p(tmp, C.c_double))
        out[:, k, :] = tmp
    return out


def zonalmean_record(zmask, zmaskvar, zv, dl_surf, zspval=0.0, lmax=False):
    """-> (mean (nb, nk, ny) f64, zmax, zmin (nb, nk, ny) f32 or None).  cdfzonalmean.f90:312-344."""
    zmask, zmaskvar, zv = _f32(zmask), _f32(zmaskvar), _f32(zv)
    dl = np.ascontiguousarray(dl_surf, np.float64)
    nk, ny, nx = zv.shape
    nb = zmask.shape[2]
    mean = np.empty((nb, nk, ny), np.float64)
    zmax = np.empty((nb, nk, ny), np.float32) if lmax else None
    zmin = np.empty((nb, nk, ny), np.float32) if lmax else None
    tm, ta, ti = np.empty((nb, ny), np.float64), np.empty((nb, ny), np.float32), np.empty((nb, ny), np.float32)
    for k in range(nk):
        lib().oracle_zonalmean_level(nx, ny, nb, _p(zmask, C.c_float), _p(zmaskvar[k], C.c_float), _p(zv[k], C.c_float),
                                     _p(dl, C.c_double), C.c_float(zspval), int(lmax), _p(tm, C.c_double),
                                     _p(ta, C.c_float), _p(ti, C.c_float))
        mean[:, k, :] = tm
        if lmax:
            zmax[:, k, :] = ta
            zmin[:, k, :] = ti
    return mean, zmax, zmin


def mhst_record(e1v, e3v, vmask1, zvt, zvs, atl=None, pac=None, ind=None, zdim=False):
    """-> heat, salt (nlev, 4, ny) float64 raw zonal sums (glo, atl, pac, ind).  cdfmhst.f90:303-366."""
    e1v, e3v, vmask1, zvt, zvs = _f32(e1v), _f32(e3v), _f32(vmask1), _f32(zvt), _f32(zvs)
    nz, ny, nx = zvt.shape
    assert e3v.shape == (nz, ny, nx)
    nlev = nz if zdim else 1
    heat, salt = np.empty((nlev, 4, ny), np.float64), np.empty((nlev, 4, ny), np.float64)
    a, p_, i = (_f32(x) if x is not None else None for x in (atl, pac, ind))
    ptr = lambda x: _p(x, C.c_float) if x is not None else None
    lib().oracle_mhst_record(nx, ny, nz, _p(e1v, C.c_float), _p(e3v, C.c_float), _p(vmask1, C.c_float), ptr(a), ptr(p_), ptr(i),
                             _p(zvt, C.c_float), _p(zvs, C.c_float), int(zdim), _p(heat, C.c_double), _p(salt, C.c_double))
    return heat, salt


# ---- cdftransig_xy3d --------------------------------------------------------------------
def transig_bins(nbins, ds1min, ds1scal, ds1zoom=999.0, ds1scalmin=999.0):
    """-> (dsigma (nbins), dsig_edge (nbins+1), itab (nsigmax) int32 1-based bins, ds1scalmin used).  cdftransig_xy3d.f90:213,229-262."""
    dsigma, edge = np.empty(nbins, np.float64), np.empty(nbins + 1, np.float64)
    cap = 1 << 16
    itab = np.zeros(cap, np.int32)
    lib().oracle_transig_bins.restype = C.c_int
    n = lib().oracle_transig_bins(int(nbins), C.c_double(ds1min), C.c_double(ds1scal), C.c_double(ds1zoom), C.c_double(ds1scalmin),
                                  _p(dsigma, C.c_double), _p(edge, C.c_double), _p(itab, C.c_int32), cap)
    assert 0 < n <= cap
    return dsigma, edge, itab[:n].copy(), min(ds1scalmin, ds1scal)


def transig_record(e2u, e1v, e3u, e3v, zu, zv, zt, zs, pref, ds1min, ds1scalmin, itab, nbins, dusig, dvsig, masku, maskv,
                   set_masks=True, lperio=False, teos10=False):
    """One frame, all levels, accumulated in place into dusig / dvsig (nbins, ny, nx) float64.  cdftransig_xy3d.f90:396-461."""
    e2u, e1v, e3u, e3v, zu, zv, zt, zs = (_f32(x) for x in (e2u, e1v, e3u, e3v, zu, zv, zt, zs))
    nzm1, ny, nx = zu.shape
    it = np.ascontiguousarray(itab, np.int32)
    assert dusig.dtype == np.float64 and dusig.shape == (nbins, ny, nx) and dusig.flags.c_contiguous
    assert masku.dtype == np.uint8 and masku.shape == (nzm1, ny, nx)
    lib().oracle_transig_record(nx, ny, nzm1, int(teos10), C.c_float(pref), int(lperio), C.c_double(ds1min),
                                C.c_double(ds1scalmin), int(it.size), _p(it, C.c_int32), int(nbins), _p(e2u, C.c_float),
                                _p(e1v, C.c_float), _p(e3u, C.c_float), _p(e3v, C.c_float), _p(zu, C.c_float),
                                _p(zv, C.c_float), _p(zt, C.c_float), _p(zs, C.c_float), int(set_masks),
                                _p(masku, C.c_uint8), _p(maskv, C.c_uint8), _p(dusig, C.c_double), _p(dvsig, C.c_double))


def sigtrp_prepare(gdept1, e3w_a, e3w_b, zu, zspu, zs_a, zs_b, zsps, zt_a, zt_b, merid=False):
    """Section slices (npk, npts) as read from the files -> dict(ddepu (npk+1, npts) f64, zu, zs, zt, zmask f32, nk, found).
    src/cdfsigtrp.f90:428-460 (meridional) / :521-555 (zonal)."""
    e3w_a, e3w_b, zu, zs_a, zs_b, zt_a, zt_b = (_f32(x) for x in (e3w_a, e3w_b, zu, zs_a, zs_b, zt_a, zt_b))
    npk, npts = zu.shape
    out = dict(ddepu=np.empty((npk + 1, npts), np.float64), zu=np.empty((npk, npts), np.float32),
               zs=np.empty((npk, npts), np.float32), zt=np.empty((npk, npts), np.float32),
               zmask=np.empty((npk, npts), np.float32))
    found = C.c_int(0)
    lib().oracle_sigtrp_prepare.restype = C.c_int
    out["nk"] = lib().oracle_sigtrp_prepare(npts, npk, int(merid), C.c_float(gdept1), _p(e3w_a, C.c_float), _p(e3w_b, C.c_float),
                                            _p(zu, C.c_float), C.c_float(zspu), _p(zs_a, C.c_float), _p(zs_b, C.c_float),
                                            C.c_float(zsps), _p(zt_a, C.c_float), _p(zt_b, C.c_float),
                                            _p(out["ddepu"], C.c_double), _p(out["zu"], C.c_float), _p(out["zs"], C.c_float),
                                            _p(out["zt"], C.c_float), _p(out["zmask"], C.c_float), C.byref(found))
    out["found"] = bool(found.value)
    return out


def sigtrp_section(eu, de3, ddepu, gdepw, zu, zt, zs, zmask, nk, dsigma_min, dsigma_max, nbins, mode=0, refdep=0.0,
                   teos10=False, ddepw_brk=None):
    """The compute part of one cdfsigtrp section (src/cdfsigtrp.f90:559-627) -> dict(dsigma_lev, dsig, dhiso, dwtrp,
    dwtrpbin, dtrpbin).  mode 0 sigmai(refdep), 1 neutral density, 2 -temp."""
    eu, de3, gdepw, zu, zt, zs, zmask = (_f32(x) for x in (eu, de3, gdepw, zu, zt, zs, zmask))
    ddepu = np.ascontiguousarray(ddepu, np.float64)
    npk, npts = zu.shape
    assert ddepu.shape == (npk + 1, npts) and de3.shape == (npk, npts) and eu.shape == (npts,) and gdepw.shape == (npk,)
    brk = None if ddepw_brk is None else _f32(ddepw_brk)
    out = dict(dsigma_lev=np.empty(nbins + 1, np.float64), dsig=np.empty((nk + 1, npts), np.float64),
               dhiso=np.empty((nbins + 1, npts), np.float64), dwtrp=np.empty((nbins + 1, npts), np.float64),
               dwtrpbin=np.empty((nbins, npts), np.float64), dtrpbin=np.empty(nbins, np.float64))
    lib().oracle_sigtrp_section(npts, npk, int(nk), _p(eu, C.c_float), _p(de3, C.c_float), _p(ddepu, C.c_double),
                                _p(gdepw, C.c_float), None if brk is None else _p(brk, C.c_float), _p(zu, C.c_float),
                                _p(zt, C.c_float), _p(zs, C.c_float), _p(zmask, C.c_float), int(mode), C.c_float(refdep),
                                int(teos10), C.c_double(dsigma_min), C.c_double(dsigma_max), int(nbins),
                                _p(out["dsigma_lev"], C.c_double), _p(out["dsig"], C.c_double), _p(out["dhiso"], C.c_double),
                                _p(out["dwtrp"], C.c_double), _p(out["dwtrpbin"], C.c_double), _p(out["dtrpbin"], C.c_double))
    return out
