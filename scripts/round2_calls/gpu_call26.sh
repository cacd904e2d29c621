#!/bin/bash
# round 2, call 26: randomised search stress (3 seeds, default dispatch and forced fallback), large-needle Hamming search test
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hamming_search" 2>&1 | tail -3
for seed in 1 2; do timeout 400 python scripts/stress_search.py 100 $seed 2>&1 | tail -3; done | tee gpurun_out/r02_c26_stress.txt
TA_QGRAM_QCAP=1 timeout 300 python scripts/stress_search.py 60 3 2>&1 | tail -3 | tee -a gpurun_out/r02_c26_stress.txt
TA_WAVE_SPLIT=4 TA_SEARCH_FILTER=pigeon timeout 300 python scripts/stress_search.py 60 4 2>&1 | tail -3 | tee -a gpurun_out/r02_c26_stress.txt
