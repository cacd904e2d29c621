#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -k "search" ) > gpurun_out/c6_pytest.log 2>&1
tail -3 gpurun_out/c6_pytest.log
{
for t in 128 256 512; do echo "== pigeon threads=$t"; TA_PIGEON_THREADS=$t TA_TRACE_SEARCH=1 STEPS=20 bash scripts/quick_bench.sh search_n32_h4096 2>&1 | tail -2; done
} > gpurun_out/c6_variants.log 2>&1
cat gpurun_out/c6_variants.log
WL=search_n32_h4096
ncu --set full --clock-control none --import-source on -k regex:'search' -s 6 -c 2 -f -o gpurun_out/prof_${WL} \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_${WL}.log 2>&1
