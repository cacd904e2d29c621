#!/bin/bash
# GPU call: refresh every bench line + launch lists + ncu captures for profiles/ (r01b = second half of round 1)
mkdir -p gpurun_out
STEPS=100 bash scripts/bench_all.sh > gpurun_out/c3_bench_all.log 2>&1
tail -12 gpurun_out/c3_bench_all.log
python bench.py > gpurun_out/c3_bench_default.json 2> gpurun_out/c3_bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c3_bench_reference.json 2>&1
tail -c 600 gpurun_out/c3_bench_reference.json
for WL in lev_k8_len128 search_n32_h4096; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${WL}.csv \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_${WL}.log 2>&1
done
for WL in lev_k16_len128 rdamerau_k16_len512 search_n32_h4096; do
ncu --set full --clock-control none --import-source on -k regex:'lev_|search' -s 6 -c 3 -f -o gpurun_out/prof_${WL} \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_${WL}.log 2>&1
done
ls -la gpurun_out | tail -12
