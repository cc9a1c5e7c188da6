#!/bin/bash
# Phase times of the file-based end-to-end run (cdfmoc_gpu on a tmpfs gridV.nc): run on the GPU box.
#   tools/e2e_phases.sh [nrec]   ->  prints the tool's "phase" lines for a few reader-thread counts
set -e
NREC=${1:-12}
D=$(mktemp -d -p /dev/shm)
python - "$D" "$NREC" <<'PY'
import sys
from pathlib import Path
sys.path.insert(0, ".")
from cdftools_b200 import build, ncfiles, synth
build.build_host()
d, nrec = sys.argv[1], int(sys.argv[2])
m = synth.make_mesh("ORCA025")
ncfiles.write_mesh(m, d)
ncfiles.write_gridv(m, Path(d) / "gridV.nc", nrec)
PY
TOOL=$PWD/cdftools_b200/bin/cdfmoc_gpu
cd "$D"
$TOOL -v gridV.nc > /dev/null
for i in 1 2; do
  T0=$(date +%s.%N)
  CDFGPU_TIMING=1 $TOOL -v gridV.nc 2>&1 >/dev/null | grep -E "phase" || true
  echo "wall $(echo "$(date +%s.%N) - $T0" | bc -l 2>/dev/null || python3 -c "import time;print(time.time()-$T0)") s"
  echo --
done
rm -rf "$D"
