#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -k "search or exp or variant" ) > gpurun_out/c9_pytest.log 2>&1
tail -3 gpurun_out/c9_pytest.log
{
echo "== staged"; TA_TRACE_SEARCH=1 STEPS=20 bash scripts/quick_bench.sh search_n32_h4096 2>&1 | tail -2
echo "== staged 128 threads"; TA_PIGEON_THREADS=128 STEPS=20 bash scripts/quick_bench.sh search_n32_h4096 2>&1 | tail -1
echo "== lane-per-segment"; TA_PIGEON_STAGED=0 STEPS=20 bash scripts/quick_bench.sh search_n32_h4096 2>&1 | tail -1
} > gpurun_out/c9_variants.log 2>&1
cat gpurun_out/c9_variants.log
WL=search_n32_h4096
ncu --set full --clock-control none --import-source on -k regex:'search' -s 6 -c 2 -f -o gpurun_out/prof_${WL}_staged \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_${WL}_staged.log 2>&1
