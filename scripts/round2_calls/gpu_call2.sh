#!/bin/bash
# round 2, call 2: first run of the diagonal-extension kernel (lev_fr.cu): parity, then timings against the bit-parallel kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_lev_fr_long_strings or diagonal-extension" 2>&1 | tail -15 > gpurun_out/r02_c2_tests.txt
cat gpurun_out/r02_c2_tests.txt
for wl in lev_k16_len4096 exp_len1024 rdamerau_k16_len512 lev_k8_len128; do
  for fr in 0 1; do
    echo "== $wl TA_FR=$fr" >> gpurun_out/r02_c2_bench.txt
    TA_FR=$fr timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c2_bench.txt
  done
done
cat gpurun_out/r02_c2_bench.txt
