#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -k "lev or exp or kat or variant" ) > gpurun_out/c17_pytest.log 2>&1
tail -3 gpurun_out/c17_pytest.log
export STEPS=100
bash scripts/quick_bench.sh lev_k8_len128 lev_k16_len128 lev_k8_len128_R rdamerau_k16_len512 exp_len1024 lev_k16_len4096 > gpurun_out/c17_variants.log 2>&1
cat gpurun_out/c17_variants.log
