#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/c4_pytest.log 2>&1
tail -5 gpurun_out/c4_pytest.log
export STEPS=100
{
echo "== duo (default for k<=8)"; bash scripts/quick_bench.sh lev_k8_len128
echo "== duo off"; TA_BLK_DUO=0 bash scripts/quick_bench.sh lev_k8_len128
echo "== search"; TA_TRACE_SEARCH=1 STEPS=20 bash scripts/quick_bench.sh search_n32_h4096 2>&1 | tail -3
} > gpurun_out/c4_variants.log 2>&1
cat gpurun_out/c4_variants.log
for WL in lev_k8_len128 lev_k16_len128 rdamerau_k16_len512; do
ncu --set full --clock-control none --import-source on -k regex:'lev_' -s 3 -c 1 -f -o gpurun_out/prof_${WL} \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_${WL}.log 2>&1
done
ls -la gpurun_out | tail -6
