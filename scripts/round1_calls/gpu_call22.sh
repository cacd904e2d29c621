#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'lev_' -s 3 -c 1 -f -o gpurun_out/prof_final_lev_k8_len128 \
    python bench.py --workload lev_k8_len128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_final_lev_k8_len128.log 2>&1
ls -la gpurun_out/prof_final_lev_k8_len128.ncu-rep
