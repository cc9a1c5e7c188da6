"""Short kernel-only driver for ncu captures: python tools/prof_run.py {k1|k2} [grid] [pref] [noise]"""
import sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from cdftools_b200 import lib, synth
import oracle
which = sys.argv[1]
grid = sys.argv[2] if len(sys.argv) > 2 else "ORCA025"
pref = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
noise = float(sys.argv[4]) if len(sys.argv) > 4 else 0.15
m = synth.make_mesh(grid)
ib = oracle.basin_masks(*synth.basin_mask_inputs(m))
lib.init(0, 3)
g = torch.Generator(device="cuda"); g.manual_seed(1)
vm = torch.from_numpy(m.vmask[:-1].astype(np.float32)).cuda()
recs = [(0.1 * float(os.environ.get("VSCALE", "1")) * torch.randn((m.nz - 1, m.ny, m.nx), device="cuda", generator=g)) * vm for _ in range(3)]
st = torch.cuda.Stream()
if which == "k1":
    e3m = oracle.mask_e3v(m.e3v_0, m.vmask.astype(np.float32))
    nz, ny, nb = lib.cdfmoc_setup(m.e1v, e3m, ib)
    out = torch.empty((nz, ny, nb), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for it in range(3):
            for r in recs: lib.cdfmoc_compute_device(r, out, st)
    st.synchronize()
else:
    nbins, smin, sstp = {0.0: (104, 23.0, 0.05), 2000.0: (158, 30.0, 0.05), 1000.0: (88, 24.0, 0.1)}[pref]
    lib.cdfmocsig_setup(m.e1v, m.e3v_0, ib, m.nz, nbins, smin, sstp, pref, 0)
    tm = torch.from_numpy(m.tmask[:-1].astype(np.float32)).cuda()
    z = torch.from_numpy(m.gdept_1d[:-1].astype(np.float32)).cuda()[:, None, None]
    cl = torch.cos(torch.deg2rad(torch.from_numpy(m.gphiv).cuda()))[None]
    t = ((1.0 + 24.0 * torch.exp(-z / 1000.0) * cl ** 2 + noise * torch.randn(vm.shape, device="cuda", generator=g)) * tm).float().contiguous()
    s = ((34.2 + 1.2 * cl * torch.exp(-z / 600.0) + 0.5 * (1 - torch.exp(-z / 1500.0)) + 0.2 * noise * torch.randn(vm.shape, device="cuda", generator=g)) * tm).float().contiguous()
    o2 = torch.empty((m.ny, nbins, ib.shape[2]), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for it in range(3):
            for r in recs[:2]: lib.cdfmocsig_compute_device(r, t, s, o2, stream=st)
    st.synchronize()
lib.finalize()
