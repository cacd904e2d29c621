#!/bin/bash
# round 2, call 35: ncu --set full of the tiled and the plain duo kernel on the headline workload (why is the tiled one 5 % slower there?)
mkdir -p gpurun_out
for t in 1 0; do
TA_DUO_TILED=$t ncu --set full --clock-control none --import-source on -k regex:lev_bitpar_duo -s 5 -c 1 -f -o gpurun_out/prof_duo_tiled${t}_lev_k8_len128 \
    python bench.py --workload lev_k8_len128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > /dev/null 2>&1
done
ls -la gpurun_out/prof_duo_tiled*.ncu-rep
