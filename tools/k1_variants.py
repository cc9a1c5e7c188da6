"""On-box A/B of the K1 kernel variants ($CDFGPU_K1_VARIANT is read by cdfmoc_gpu_setup): time per record on
device-resident synthetic records (inputs >> L2) and bit-comparison of every variant with variant 0.
    python tools/k1_variants.py [GRID] [variants, comma separated] [nrec]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from cdftools_b200 import lib, synth  # noqa: E402
import oracle  # noqa: E402

grid = sys.argv[1] if len(sys.argv) > 1 else "ORCA025"
variants = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,10,11,12,13,14,15,16").split(",")]
nrec = int(sys.argv[3]) if len(sys.argv) > 3 else 16
m = synth.make_mesh(grid)
ib = oracle.basin_masks(*synth.basin_mask_inputs(m))
e3m = oracle.mask_e3v(m.e3v_0, m.vmask.astype(np.float32))
lib.load()
lib.init(0, 1)
g = torch.Generator(device="cuda")
g.manual_seed(1)
recs = [0.1 * torch.randn((m.nz - 1, m.ny, m.nx), device="cuda", generator=g) for _ in range(nrec)]
bytes_rec = (m.nz - 1) * m.ny * m.nx * 8 + m.ny * m.nx + 5 * m.ny * m.nz * 8
st = torch.cuda.Stream()
ref = None
for v in variants:
    os.environ["CDFGPU_K1_VARIANT"] = str(v)
    for chunk in (os.environ.get("K1_CHUNKS", "0").split(",")):
        if chunk != "0":
            os.environ["CDFGPU_K1_CHUNK"] = chunk
        nz, ny, nb = lib.cdfmoc_setup(m.e1v, e3m, ib)
        out = torch.zeros((nz, ny, nb), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        best = 1e9
        with torch.cuda.stream(st):
            for r in recs[:3]:
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
            out.zero_()
            lib.cdfmoc_compute_device(recs[0], out, st)
        st.synchronize()
        o = out.cpu().numpy()
        if ref is None:
            ref = o
        same = bool(np.array_equal(o, ref))
        print(json.dumps({"variant": v, "chunk": chunk, "grid": grid, "ms_per_record": round(best, 5), "GBps": round(bytes_rec / best / 1e6, 1),
                          "bit_equal_to_first": same, "max_abs_diff": float(np.abs(o - ref).max())}), flush=True)
        lib.cdfmoc_teardown()
