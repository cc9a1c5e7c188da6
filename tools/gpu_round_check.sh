#!/bin/bash
# One GPU call (gpurun -- tools/gpu_round_check.sh) that refreshes a round's evidence on one B200: full GPU suite, smoke, the
# default bench line (headline + sub-records), the reference arm, the ncu launch list of the bench command and one
# `ncu --set full` capture of K2 exported to CSV (the reports themselves are too big to ship back).
R=${1:-r02}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 500 python bench.py > gpurun_out/bench_${R}_n1.json 2> gpurun_out/bench_${R}_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${R}_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-subrecords > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mocsig_eos_hist -s 3 -c 1 -o gpurun_out/${R}_k2 -f \
    python tools/k2_probe.py ORCA025 0 > gpurun_out/${R}_k2.log 2>&1
ncu -i gpurun_out/${R}_k2.ncu-rep --page raw --csv > gpurun_out/${R}_k2.raw.csv 2>/dev/null
ncu -i gpurun_out/${R}_k2.ncu-rep --page source --csv > gpurun_out/${R}_k2.src.csv 2>/dev/null
python tools/prof_siblings.py ORCA025 3 2>&1 | tail -1
tail -c 600 gpurun_out/bench_${R}_n1.json
