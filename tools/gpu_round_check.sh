#!/bin/bash
# One GPU call that refreshes the round's evidence: full GPU suite, the two ORCA025 bench lines, the reference arm,
# ncu launch lists of both bench commands and one `ncu --set full` capture of K2 (exported to CSV: reports are too big to ship).
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py > gpurun_out/bench_r01_k1.json 2> gpurun_out/bench_r01_k1.err
timeout 400 python bench.py --workload cdfmocsig-ORCA025-L75-sigma0-104bins-73rec --steps 3 > gpurun_out/bench_r01_k2.json 2> gpurun_out/bench_r01_k2.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_k1.csv python bench.py --steps 2 --warmup 3 > gpurun_out/lk1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_k2.csv python bench.py --workload cdfmocsig-ORCA025-L75-sigma0-104bins-73rec --steps 1 --warmup 3 > gpurun_out/lk2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mocsig_eos -s 2 -c 1 -o /tmp/k2tab -f python tools/prof_run.py k2 ORCA025 0.0 0.15 > gpurun_out/k2tab.log 2>&1
ncu -i /tmp/k2tab.ncu-rep --page raw --csv > gpurun_out/k2tab.raw.csv
for n in 0.15 0.0; do timeout 300 python tools/prof_transig.py ORCA025 1000 $n 2>&1 | tail -1; done | tee gpurun_out/transig_b.log
tail -c 700 gpurun_out/bench_r01_k2.json
