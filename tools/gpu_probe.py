"""Quick on-box probe: first timing of K1/K2 on device-resident synthetic ORCA025 records (not the bench)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from cdftools_b200 import lib, synth
import oracle

grid = sys.argv[1] if len(sys.argv) > 1 else "ORCA025"
m = synth.make_mesh(grid)
ib = oracle.basin_masks(*synth.basin_mask_inputs(m))
e3m = oracle.mask_e3v(m.e3v_0, m.vmask.astype(np.float32))
lib.init(0, 3)
nz, ny, nb = lib.cdfmoc_setup(m.e1v, e3m, ib)
nrec = 8
g = torch.Generator(device="cuda"); g.manual_seed(1)
vm = torch.from_numpy(m.vmask[:-1].astype(np.float32)).cuda()
recs = [(0.1 * torch.randn((m.nz - 1, m.ny, m.nx), device="cuda", generator=g)) * vm for _ in range(nrec)]
out = torch.empty((nz, ny, nb), dtype=torch.float64, device="cuda")
st = torch.cuda.Stream()
torch.cuda.synchronize()
cells = m.nx * m.ny * m.nz
with torch.cuda.stream(st):
    for r in recs: lib.cdfmoc_compute_device(r, out, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for it in range(5):
        for r in recs: lib.cdfmoc_compute_device(r, out, st)
    e1.record(st)
st.synchronize()
ms = e0.elapsed_time(e1) / (5 * nrec)
bytes_rec = (m.nz - 1) * m.ny * m.nx * 8 + m.ny * m.nx + nb * ny * nz * 8
print(json.dumps({"kernel": "K1", "grid": grid, "ms_per_record": ms, "cells_per_s": cells / ms * 1e3, "GBps": bytes_rec / ms / 1e6, "wet": m.wet_fraction}))
lib.cdfmoc_teardown()
# K2
tm = torch.from_numpy(m.tmask[:-1].astype(np.float32)).cuda()
z = torch.from_numpy(m.gdept_1d[:-1].astype(np.float32)).cuda()[:, None, None]
cl = torch.cos(torch.deg2rad(torch.from_numpy(m.gphiv).cuda()))[None]
for cfg in ((0.0, 23.0, 0.05, 104), (2000.0, 30.0, 0.05, 158)):
    pref, smin, sstp, nbins = cfg
    lib.cdfmocsig_setup(m.e1v, m.e3v_0, ib, m.nz, nbins, smin, sstp, pref, 0)
    for noise in (0.15, 0.0):
        trec, srec = [], []
        for r in range(2):
            t = (1.0 + 24.0 * torch.exp(-z / 1000.0) * cl ** 2 + noise * torch.randn(vm.shape, device="cuda", generator=g)) * tm
            s = (34.2 + 1.2 * cl * torch.exp(-z / 600.0) + 0.5 * (1 - torch.exp(-z / 1500.0)) + 0.2 * noise * torch.randn(vm.shape, device="cuda", generator=g)) * tm
            trec.append(t.float().contiguous()); srec.append(s.float().contiguous())
        o2 = torch.empty((m.ny, nbins, nb), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        with torch.cuda.stream(st):
            for r in range(2): lib.cdfmocsig_compute_device(recs[r], trec[r], srec[r], o2, stream=st)
            e0.record(st)
            for it in range(3):
                for r in range(2): lib.cdfmocsig_compute_device(recs[r], trec[r], srec[r], o2, stream=st)
            e1.record(st)
        st.synchronize()
        ms = e0.elapsed_time(e1) / 6
        bytes_rec = (m.nz - 1) * m.ny * m.nx * 16 + m.ny * m.nx + nb * nbins * ny * 8
        print(json.dumps({"kernel": "K2", "grid": grid, "pref": pref, "nbins": nbins, "noise": noise, "ms_per_record": ms, "cells_per_s": cells / ms * 1e3, "GBps": bytes_rec / ms / 1e6}))
        del trec, srec
    lib.cdfmocsig_teardown()
