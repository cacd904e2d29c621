#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -k "lev_k_mutated or lev_k_random_short or device_resident or full_size" ) > gpurun_out/c20_pytest.log 2>&1
tail -3 gpurun_out/c20_pytest.log
export STEPS=100
bash scripts/quick_bench.sh lev_k8_len128 lev_k8_ragged96_160 lev_k8_len128_R > gpurun_out/c20_variants.log 2>&1
cat gpurun_out/c20_variants.log
