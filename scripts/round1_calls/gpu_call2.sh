#!/bin/bash
# GPU call: parity suite, block-table kernel variants, search phase trace, one ncu capture.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/c2_pytest.log 2>&1
tail -5 gpurun_out/c2_pytest.log
export STEPS=100
{
echo "== old sliding table"; TA_BITPAR=tab bash scripts/quick_bench.sh lev_k8_len128
echo "== blk defaults"; bash scripts/quick_bench.sh lev_k8_len128 lev_k16_len128 rdamerau_k16_len512 lev_k16_len4096 exp_len1024
for t in 64 96 160 192; do echo "== blk planes=1 threads=$t"; TA_BITPAR_THREADS=$t bash scripts/quick_bench.sh lev_k8_len128; done
echo "== blk planes=0"; TA_BLK_PLANES=0 bash scripts/quick_bench.sh lev_k8_len128 rdamerau_k16_len512
echo "== blk C=8 on k8"; TA_BLK_C=8 bash scripts/quick_bench.sh lev_k8_len128
echo "== search"; TA_TRACE_SEARCH=1 STEPS=20 bash scripts/quick_bench.sh search_n32_h4096 2>&1 | tail -8
} > gpurun_out/c2_variants.log 2>&1
cat gpurun_out/c2_variants.log
ncu --set full --clock-control none --import-source on -k regex:'lev_' -s 3 -c 1 -f -o gpurun_out/prof_blk2_k8 \
    python bench.py --workload lev_k8_len128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_blk2_k8.log 2>&1
ls -la gpurun_out | tail -5
