"""On-box timing of K6 (cdftransig_xy3d) on a synthetic ORCA025 frame: python tools/prof_transig.py [GRID] [code]"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from cdftools_b200 import lib, synth  # noqa: E402

grid = sys.argv[1] if len(sys.argv) > 1 else "ORCA025"
code = sys.argv[2] if len(sys.argv) > 2 else "1000"
noise = float(sys.argv[3]) if len(sys.argv) > 3 else 0.15
pref, nb, smin, scal, zoom, smn = lib.TRANSIG_CODES[code]
m = synth.make_mesh(grid)
lib.load()
lib.init(0, 1)
_, _, itab, scm = lib.transig_bins(nb, smin, scal, zoom, smn)
e3 = m.e3v_0[:-1]
shape = lib.cdftransig_setup(m.e1u, m.e1v, e3, e3, m.nz, nb, pref, smin, scm, itab, lperio=True)
zu = synth.make_v_record(m, 7)[:-1] * m.umask[:-1]
zv = synth.make_v_record(m, 3)[:-1] * m.vmask[:-1]
zt, zs = (x[:-1] for x in synth.make_ts_record(m, 1, noise=noise))
ms = []
for r in range(4):
    lib.cdftransig_record(zu, zv, zt, zs, set_masks=(r == 0))
    ms.append(lib.cdftransig_kernel_ms())
cells = m.nx * m.ny * m.nz
# algorithmic bytes per level-cell: T, S (8) -> dens (8 written, 8 read back + halo), u, v, e3u, e3v (16), masks (2), two fp64 RMWs (32)
bytes_rec = (m.nz - 1) * m.ny * m.nx * (8 + 16 + 16 + 2 + 32)
print(json.dumps({"kernel": "K6 cdftransig_xy3d", "grid": grid, "code": code, "ts_noise_K": noise, "nbins": nb, "ms_per_frame": ms, "cells_per_s": cells / min(ms) * 1e3,
                  "GBps_algorithmic": bytes_rec / min(ms) / 1e6}))
lib.cdftransig_teardown()
