"""Driver for ncu captures / timings of the sibling-tool kernels (K4 zonal sum / mean, K5 mhst) on a synthetic grid:
python tools/prof_siblings.py [grid] [reps]"""
import sys, json
import numpy as np
sys.path.insert(0, ".")
from cdftools_b200 import lib, synth
import oracle
grid = sys.argv[1] if len(sys.argv) > 1 else "ORCA025"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m = synth.make_mesh(grid)
nk = m.nz - 1
rng = np.random.default_rng(0)
zv = (rng.standard_normal((nk, m.ny, m.nx), dtype=np.float32) * 3 + 10)
msk = m.tmask[:nk].astype(np.float32)
zm = oracle.zonal_masks(msk[0], m.tmaskatl, m.tmaskind, m.tmaskpac)
lib.init(0, 3)
shape = lib.cdfzonal_setup(m.e1v, m.e1v, zm, msk)
cells = nk * m.ny * m.nx
res = {}
for r in range(reps):
    lib.cdfzonal_sum(zv, shape); res["zonalsum_ms"] = lib.cdfzonal_kernel_ms()
    lib.cdfzonal_mean(zv, shape, 0.0, False); res["zonalmean_ms"] = lib.cdfzonal_kernel_ms()
    lib.cdfzonal_mean(zv, shape, 0.0, True); res["zonalmean_max_ms"] = lib.cdfzonal_kernel_ms()
lib.cdfzonal_teardown()
dims = lib.cdfmhst_setup(m.e1v, m.e3v_0[:nk], m.vmask[0].astype(np.float32), m.tmaskatl, m.tmaskpac, m.tmaskind)
for r in range(reps):
    lib.cdfmhst_record(zv, zv, dims, False); res["mhst_ms"] = lib.cdfmhst_kernel_ms()
    lib.cdfmhst_record(zv, zv, dims, True); res["mhst_zdim_ms"] = lib.cdfmhst_kernel_ms()
lib.cdfmhst_teardown()
lib.finalize()
res["zonalsum_GBps"] = cells * 8 / res["zonalsum_ms"] / 1e6
res["mhst_GBps"] = cells * 12 / res["mhst_ms"] / 1e6
print(json.dumps({k: round(v, 4) for k, v in res.items()}))
