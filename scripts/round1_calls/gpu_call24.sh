#!/bin/bash
mkdir -p gpurun_out
TA_LEN_BUCKETS=1 STEPS=50 bash scripts/quick_bench.sh lev_k8_len128 lev_k16_len128 > gpurun_out/c24_buckets.log 2>&1
cat gpurun_out/c24_buckets.log
