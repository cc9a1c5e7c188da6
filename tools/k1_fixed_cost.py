"""K1 launch time against the number of rows (same row length): t(ny) = F + c*ny separates the per-launch fixed cost
(ramp-up, tail, launch gap) from the streaming rate.  python tools/k1_fixed_cost.py"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from cdftools_b200 import lib, synth  # noqa: E402
import oracle  # noqa: E402

lib.load()
lib.init(0, 1)
g = torch.Generator(device="cuda")
g.manual_seed(1)
st = torch.cuda.Stream()
res = []
for ny in (128, 255, 510, 1021, 2042, 4084):
    m = synth.make_mesh((1442, ny, 75))
    ib = oracle.basin_masks(*synth.basin_mask_inputs(m))
    e3m = oracle.mask_e3v(m.e3v_0, m.vmask.astype(np.float32))
    nz, nyy, nb = lib.cdfmoc_setup(m.e1v, e3m, ib)
    nrec = max(4, min(16, int(8e9 // (74 * ny * 1442 * 4))))
    recs = [0.1 * torch.randn((74, ny, 1442), device="cuda", generator=g) for _ in range(nrec)]
    out = torch.zeros((nz, nyy, nb), dtype=torch.float64, device="cuda")
    best = 1e9
    with torch.cuda.stream(st):
        for r in recs:
            lib.cdfmoc_compute_device(r, out, st)
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for it in range(4):
                for r in recs:
                    lib.cdfmoc_compute_device(r, out, st)
            e1.record(st)
            st.synchronize()
            best = min(best, e0.elapsed_time(e1) / (4 * nrec))
    mb = (74 * ny * 1442 * 8 + ny * 1442 + 5 * ny * 75 * 8) / 1e6
    res.append((ny, best * 1e3, mb))
    print(json.dumps({"ny": ny, "us_per_launch": round(best * 1e3, 2), "MB": round(mb, 1), "GBps": round(mb / best, 1)}), flush=True)
    lib.cdfmoc_teardown()
    del recs
a = np.array(res)
c, f = np.polyfit(a[:, 2], a[:, 1], 1)
print(json.dumps({"fit": "us = F + MB / BW", "F_us": round(float(f), 2), "BW_GBps": round(1e3 / float(c), 1)}))
