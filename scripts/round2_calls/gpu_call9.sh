#!/bin/bash
# round 2, call 9: diag16 v3 (adds on the FMA pipe), FR v6 (L2 prefetch of the next level), new search workloads, captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_lev_diag16_large_batch or u16-thread-per-pair or test_lev_fr_long_strings or diagonal-extension-kernel-forced or test_search_long or test_kat or hamming_search" 2>&1 | tail -5 > gpurun_out/r02_c9_tests.txt
cat gpurun_out/r02_c9_tests.txt
run() { # name env...
  echo "== $*" >> gpurun_out/r02_c9_bench.txt
  env "${@:2}" timeout 300 python bench.py --workload $1 --steps 20 --warmup 3 --no-cpu-baseline --no-configs 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'], d['roofline']['kernel'], 'e2e_ms', d['e2e']['ms_per_step'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c9_bench.txt
}
run affine_k16_len128 TA_X=1
run rdamerau_k16_len512 TA_FORCE_BAND=1
run lev_k16_len4096 TA_X=1
run exp_len1024 TA_X=1
run search_n32_h4096 TA_X=1
run search_all_n32_h4096 TA_X=1
run search_n64_h4096 TA_X=1
run search_affine_n32_h4096 TA_X=1
cat gpurun_out/r02_c9_bench.txt
ncu --set full --clock-control none --import-source on -k regex:'lev_diag16' -s 3 -c 1 -f -o gpurun_out/prof_diag16v3_affine_k16_len128 \
    python bench.py --workload affine_k16_len128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/prof_diag16_affine.log 2>&1
TA_FORCE_BAND=1 ncu --set full --clock-control none --import-source on -k regex:'lev_diag16' -s 3 -c 1 -f -o gpurun_out/prof_diag16v3_trans_rdamerau_k16_len512 \
    python bench.py --workload rdamerau_k16_len512 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/prof_diag16_trans.log 2>&1
for wl in lev_k16_len4096 exp_len1024; do
ncu --set full --clock-control none --import-source on -k regex:'lev_fr' -s 3 -c 1 -f -o gpurun_out/prof_fr6_${wl} \
    python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/prof_fr6_${wl}.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_search_affine.csv \
    python bench.py --workload search_affine_n32_h4096 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > /dev/null 2>&1
TA_TRACE_SEARCH=1 python bench.py --workload search_n32_h4096 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>&1 | grep "ta search" | tail -3
