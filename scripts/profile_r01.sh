#!/bin/bash
# ncu evidence for the bench step (run under gpurun, 1 GPU).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
WL=${1:-lev_k8_len128}
# launch list of the bench command (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${WL}.csv \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_${WL}.log 2>&1
# full capture of the hot kernel
ncu --set full --clock-control none --import-source on -k regex:'lev_|hamming|search' -s 3 -c 2 -f -o gpurun_out/prof_${WL} \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_${WL}.log 2>&1
ls -la gpurun_out
