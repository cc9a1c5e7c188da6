#!/usr/bin/env python
"""Generates the committed golden fixtures of tests/golden/.

    python tests/golden/make_golden.py        (CPU only; needs gcc for oracle/libcdforacle.so)

What the fixtures are -- and are not.  The reference (Fortran, NetCDF) cannot be compiled or run in this image and holds
no tests or golden vectors for the cdfmoc / cdfmocsig path, so:

* `eos_kat.json`      : the ONLY known-answer values of the reference itself -- the three check values printed in the
                        comments of src/eos.f90 (:646, :817, :820) -- plus values derived from the formulas at survey
                        time.  These pin the oracle's EOS.
* `hand_cdfmoc.json`  : a 4 x 3 x 3 cdfmoc case whose products and sums are exact in binary, worked by hand from
                        src/cdfmoc.f90:352-388.
* `<case>.npz`        : FROZEN INPUT / OUTPUT vectors of the C oracle (oracle/cdf_oracle.c) on small seeded grids:
                        inputs are stored in full (not regenerated from a seed), outputs are what the oracle returned the
                        day the fixture was made.  They guard the oracle, its NumPy twin and the CUDA path against drift
                        and give the GPU tests an input set that does not depend on the synthetic generator; they are
                        NOT outputs of the Fortran reference ("parity unpinned by the reference" for everything but the
                        EOS -- DESIGN.md section 2).
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

import oracle                                    # noqa: E402
from cdftools_b200 import synth                  # noqa: E402

MOC_CASES = [("TINY", True), ("ODD", True), ("ODD", False), ("SMALL", True)]
# (grid, pref, eos, (sigmin, sigstp, nbins) or None for the defaults of cdfmocsig.f90:265-304)
SIG_CASES = [("TINY", 0.0, 0, None), ("ODD", 2000.0, 0, None), ("TINY", 1000.0, 1, None),
             ("ODD", 0.0, 2, (1020.0, 0.25, 60)), ("SMALL", 0.0, 0, (23.0, 0.05, 104))]


def moc_case(grid, with_basins):
    m = synth.make_mesh(grid)
    ib = oracle.basin_masks(*synth.basin_mask_inputs(m, with_basins), zero_edges=True)
    e3m = oracle.mask_e3v(m.e3v_0, m.vmask.astype(np.float32))
    zv = np.stack([synth.make_v_record(m, r, adversarial=(r == 1))[:-1] for r in range(2)])
    dmoc = np.stack([oracle.cdfmoc_record(m.e1v, e3m, ib, v) for v in zv])
    out = np.stack([oracle.cdfmoc_output(d) for d in dmoc])
    return dict(e1v=m.e1v, e3m=e3m, ibmask=ib, zv=zv, dmoc=dmoc, out=out)


def sig_case(grid, pref, eos, bins):
    m = synth.make_mesh(grid)
    ib = oracle.basin_masks(*synth.basin_mask_inputs(m, True), zero_edges=True)
    nbins, smin, sstp = oracle.default_bins(pref, eos == 2) if bins is None else (bins[2], bins[0], bins[1])
    zv = synth.make_v_record(m, 1, spval=1.0e20)[:-1]
    zt, zs = (x[:-1] for x in synth.make_ts_record(m, 1))
    H, ibin = oracle.cdfmocsig_record(m.e1v, m.e3v_0, ib, zv, zt, zs, 1.0e20, 0.0, 0.0, pref, eos, smin, sstp, nbins)
    return dict(e1v=m.e1v, e3v=m.e3v_0, ibmask=ib, zv=zv, zt=zt, zs=zs, dmoc=H, ibin=ibin.astype(np.int16),
                params=np.array([pref, eos, smin, sstp, nbins, 1.0e20, 0.0, 0.0], np.float64))


def main():
    oracle.build()
    kat = {
        "reference_comment_values": [
            {"cite": "src/eos.f90:820", "eos": "EOS80", "what": "in situ rho", "t": 3.0, "s": 35.5, "z": 3000.0,
             "rho": 1028.35011066567},
            {"cite": "src/eos.f90:817", "eos": "TEOS10", "what": "in situ rho", "t": 3.0, "s": 35.5, "z": 3000.0,
             "rho": 1028.21993233072},
            {"cite": "src/eos.f90:646", "eos": "NEUTRAL", "what": "sigmantr + 1000", "t": 20.0, "s": 35.0,
             "rho": 1024.59416751197},
        ],
        "derived_sigmai": [
            {"t": t, "s": s, "pref": p, "teos10": te,
             "sigma": float(oracle.sigmai_dep(np.array([t], np.float32), np.array([s], np.float32), p, te)[0])}
            for (t, s, p, te) in [(3.0, 35.5, 3000.0, False), (3.0, 35.5, 3000.0, True), (20.0, 35.0, 0.0, False),
                                  (10.0, 35.0, 0.0, False), (2.0, 34.9, 2000.0, False), (-1.5, 33.0, 1000.0, True),
                                  (28.0, 36.5, 0.0, True)]
        ],
    }
    (HERE / "eos_kat.json").write_text(json.dumps(kat, indent=1) + "\n")
    hand = {
        "cite": "src/cdfmoc.f90:352-388",
        "nx": 4, "ny": 3, "nz": 3, "e1v": 2.0, "e3v_levels": [1.0, 0.5, 1.0],
        "mask": "1 everywhere but i=1 and i=nx (cdfmoc.f90:327-328)",
        "zv_row_j2": {"level1": [100.0, 1.0, 2.0, 100.0], "level2": [100.0, 4.0, 8.0, 100.0]},
        "transport_row_j2": {"level1": -6.0, "level2": -12.0},
        "psi_sv_row_j2": [-18.0e-6, -12.0e-6, 0.0],
    }
    (HERE / "hand_cdfmoc.json").write_text(json.dumps(hand, indent=1) + "\n")
    for grid, wb in MOC_CASES:
        np.savez_compressed(HERE / f"moc_{grid}_{'basins' if wb else 'global'}.npz", **moc_case(grid, wb))
    for i, (grid, pref, eos, bins) in enumerate(SIG_CASES):
        np.savez_compressed(HERE / f"sig_{i}_{grid}_p{int(pref)}_eos{eos}.npz", **sig_case(grid, pref, eos, bins))
    for f in sorted(HERE.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
