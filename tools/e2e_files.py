"""End-to-end throughput INCLUDING NetCDF I/O: synthetic gridV file on local disk -> cdfmoc_gpu (C++ twin) -> moc.nc.
usage: python tools/e2e_files.py [grid] [nrec]   (run on the GPU box; prints one JSON line)"""
import json, subprocess, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cdftools_b200 import build, ncfiles, synth

grid = sys.argv[1] if len(sys.argv) > 1 else "ORCA025"
nrec = int(sys.argv[2]) if len(sys.argv) > 2 else 6
tools = build.build_host()
m = synth.make_mesh(grid)
with tempfile.TemporaryDirectory(dir="/dev/shm" if Path("/dev/shm").exists() else None) as d:
    ncfiles.write_mesh(m, d)
    ncfiles.write_gridv(m, Path(d) / "gridV.nc", nrec)
    size = (Path(d) / "gridV.nc").stat().st_size
    subprocess.run([tools["cdfmoc_gpu"], "-v", "gridV.nc"], cwd=d, check=True, capture_output=True)   # warm page cache
    t0 = time.perf_counter()
    subprocess.run([tools["cdfmoc_gpu"], "-v", "gridV.nc"], cwd=d, check=True, capture_output=True)
    dt = time.perf_counter() - t0
    cells = nrec * m.nx * m.ny * m.nz
    print(json.dumps({"tool": "cdfmoc_gpu", "grid": grid, "records": nrec, "gridV_bytes": size, "wall_s": dt,
                      "cells_per_s": cells / dt, "note": "whole process: CUDA init, mesh/mask read, setup, record pipeline "
                      "(fread raw big-endian -> pinned -> H2D -> GPU bswap -> K1 -> D2H -> moc.nc), files on tmpfs"}))
