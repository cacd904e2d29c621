#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c14_pytest.log 2>&1
tail -3 gpurun_out/c14_pytest.log
export STEPS=100
bash scripts/quick_bench.sh lev_k8_len128 lev_k16_len128 lev_k8_len128_R lev_k16_len128_R rdamerau_k16_len512 exp_len1024 lev_k16_len4096 lev_k60_len1024 > gpurun_out/c14_variants.log 2>&1
TA_BLK_DUO=0 bash scripts/quick_bench.sh lev_k8_len128 >> gpurun_out/c14_variants.log 2>&1
cat gpurun_out/c14_variants.log
