#!/bin/bash
# GPU call: parity suite, then the block-table kernel variants on the headline workloads, then search.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.log 2>&1
tail -5 gpurun_out/c1_pytest.log
export STEPS=100
{
echo "== old sliding table"; TA_BITPAR=tab bash scripts/quick_bench.sh lev_k8_len128 lev_k16_len128 rdamerau_k16_len512
for p in 1 0; do for m in 0 1 3 7; do
  echo "== blk planes=$p mad=$m"; TA_BLK_PLANES=$p TA_BLK_MAD=$m bash scripts/quick_bench.sh lev_k8_len128 rdamerau_k16_len512
done; done
for t in 64 96 192 256; do echo "== blk planes=1 mad=0 threads=$t"; TA_BLK_PLANES=1 TA_BLK_MAD=0 TA_BITPAR_THREADS=$t bash scripts/quick_bench.sh lev_k8_len128; done
for t in 64 96 128 192; do echo "== blk planes=0 mad=0 threads=$t"; TA_BLK_PLANES=0 TA_BLK_MAD=0 TA_BITPAR_THREADS=$t bash scripts/quick_bench.sh lev_k8_len128; done
echo "== blk C=8 on k8"; TA_BLK_C=8 bash scripts/quick_bench.sh lev_k8_len128
echo "== defaults"; bash scripts/quick_bench.sh lev_k16_len128 exp_len1024 search_n32_h4096 lev_k16_len4096
} > gpurun_out/c1_variants.log 2>&1
cat gpurun_out/c1_variants.log
# one full ncu capture of the default block-table kernel
ncu --set full --clock-control none --import-source on -k regex:'lev_' -s 3 -c 1 -f -o gpurun_out/prof_blk_k8 \
    python bench.py --workload lev_k8_len128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_blk_k8.log 2>&1
ls -la gpurun_out | tail -5
