#!/bin/bash
# final refresh of every bench line, launch lists and ncu captures for profiles/ (r01c = end of round 1)
mkdir -p gpurun_out
STEPS=100 bash scripts/bench_all.sh > gpurun_out/c15_bench_all.log 2>&1
tail -14 gpurun_out/c15_bench_all.log
python bench.py > gpurun_out/c15_bench_default.json 2> gpurun_out/c15_bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c15_bench_reference.json 2>&1
for WL in lev_k8_len128 search_n32_h4096; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${WL}.csv \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_${WL}.log 2>&1
done
for WL in lev_k8_len128 lev_k16_len128 rdamerau_k16_len512 lev_k8_len128_R; do
ncu --set full --clock-control none --import-source on -k regex:'lev_' -s 3 -c 1 -f -o gpurun_out/prof_${WL} \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_${WL}.log 2>&1
done
WL=search_n32_h4096
ncu --set full --clock-control none --import-source on -k regex:'search' -s 6 -c 2 -f -o gpurun_out/prof_${WL} \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_${WL}.log 2>&1
ls -la gpurun_out | tail -8
