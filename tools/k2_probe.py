"""K2-only timing on device-resident synthetic records: python tools/k2_probe.py [grid] [pref] (prints one JSON line per noise level)."""
import sys, json, os
import numpy as np, torch
sys.path.insert(0, ".")
from cdftools_b200 import lib, synth
import oracle
grid = sys.argv[1] if len(sys.argv) > 1 else "ORCA025"
pref = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
vscale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0   # 0: no cell contributes -> times the sweep alone
rows = tuple(int(x) for x in os.environ["ROWS"].split(",")) if os.environ.get("ROWS") else None   # e.g. ROWS=1338,1721: a band
m = synth.make_mesh(grid, rows=rows)
ib = oracle.basin_masks(*synth.basin_mask_inputs(m))
nb = ib.shape[2]
lib.init(0, 3)
lib.set_device_inputs_ready(True)   # resident, synchronised records: consecutive launches may overlap
g = torch.Generator(device="cuda"); g.manual_seed(1)
vm = torch.from_numpy(m.vmask[:-1].astype(np.float32)).cuda()
recs = [(0.1 * vscale * torch.randn((m.nz - 1, m.ny, m.nx), device="cuda", generator=g)) * vm for _ in range(2)]
tm = torch.from_numpy(m.tmask[:-1].astype(np.float32)).cuda()
z = torch.from_numpy(m.gdept_1d[:-1].astype(np.float32)).cuda()[:, None, None]
cl = torch.cos(torch.deg2rad(torch.from_numpy(m.gphiv).cuda()))[None]
nbins, smin, sstp = {0.0: (104, 23.0, 0.05), 2000.0: (158, 30.0, 0.05), 1000.0: (88, 24.0, 0.1)}[pref]
nbins = int(os.environ.get('NBINS', nbins))
eos = int(os.environ.get('EOS', 0))   # 0 EOS80, 1 TEOS10, 2 neutral density
if eos == 2:
    nbins, smin, sstp = oracle.default_bins(0.0, True)
jg = dict(j_first_global=rows[0], ny_global=synth.GRIDS[grid][1]) if rows else {}
lib.cdfmocsig_setup(m.e1v, m.e3v_0, ib, m.nz, nbins, smin, sstp, pref, eos, **jg)
if os.environ.get('ISO'):
    lib.cdfmocsig_set_isodep(m.gdept_1d)
print(json.dumps({"filter": lib.cdfmocsig_filter_info()}))
st = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
cells = m.nx * m.ny * m.nz
for noise in (0.0, 0.15):
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
        for it in range(4):
            for r in range(2): lib.cdfmocsig_compute_device(recs[r], trec[r], srec[r], o2, stream=st)
        e1.record(st)
    st.synchronize()
    ms = e0.elapsed_time(e1) / 8
    ibn = torch.empty(vm.shape, dtype=torch.int32, device="cuda")
    tiers = lib.cdfmocsig_bins_device_stats(trec[0], srec[0], ibn)
    print(json.dumps({"kernel": "K2", "tiers_cells_past32_exact": tiers, "past32_frac": tiers[1] / max(tiers[0], 1), "grid": grid, "pref": pref, "nbins": nbins, "eos": eos, "iso": bool(os.environ.get("ISO")), "noise": noise, "ms_per_record": round(ms, 4),
                      "cells_per_s": cells / ms * 1e3, "checksum": float(o2.abs().sum().item())}))
    del trec, srec
lib.cdfmocsig_teardown()
lib.finalize()
