#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c5_pytest.log 2>&1
tail -5 gpurun_out/c5_pytest.log
export STEPS=100
{
echo "== search (pigeonhole filter)"; TA_TRACE_SEARCH=1 STEPS=20 bash scripts/quick_bench.sh search_n32_h4096 2>&1 | tail -3
echo "== search (myers filter)"; TA_SEARCH_FILTER=myers TA_TRACE_SEARCH=1 STEPS=20 bash scripts/quick_bench.sh search_n32_h4096 2>&1 | tail -3
} > gpurun_out/c5_variants.log 2>&1
cat gpurun_out/c5_variants.log
WL=search_n32_h4096
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${WL}.csv \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_${WL}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'search' -s 6 -c 2 -f -o gpurun_out/prof_${WL} \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_${WL}.log 2>&1
tail -4 gpurun_out/launches_${WL}.csv | cut -c1-60,200-400
