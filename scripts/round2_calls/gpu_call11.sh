#!/bin/bash
# round 2, call 11: lane-per-pair diagonal-extension kernel (lev_fr2_kernel) vs the octet-per-pair form; search after the
# weighted pre-filter and the cheaper hit ordering
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_lev_fr_long_strings or diagonal-extension or test_search_random or test_search_planted or test_kat_search or test_search_long" 2>&1 | tail -5 > gpurun_out/r02_c11_tests.txt
cat gpurun_out/r02_c11_tests.txt
run() { # name env...
  echo "== $*" >> gpurun_out/r02_c11_bench.txt
  env "${@:2}" timeout 300 python bench.py --workload $1 --steps 20 --warmup 3 --no-cpu-baseline --no-configs --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'], d['roofline']['kernel'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c11_bench.txt
}
run lev_k16_len4096 TA_X=1
run lev_k16_len4096 TA_FR_KERNEL=octet
run exp_len1024 TA_X=1
run exp_len1024 TA_FR_KERNEL=octet
run rdamerau_k16_len512 TA_FR=1
run rdamerau_k16_len512 TA_FR=0
run lev_k16_len128 TA_FR=1
run lev_k8_len128_R TA_FR=1
run search_n32_h4096 TA_X=1
run search_n64_h4096 TA_X=1
run search_affine_n32_h4096 TA_X=1
run search_affine_n32_h4096 TA_NO_SEARCH_FILTER=1
cat gpurun_out/r02_c11_bench.txt
for wl in lev_k16_len4096 exp_len1024; do
ncu --set full --clock-control none --import-source on -k regex:'lev_fr2' -s 3 -c 1 -f -o gpurun_out/prof_fr2k_${wl} \
    python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/prof_fr2k_${wl}.log 2>&1
done
TA_TRACE_SEARCH=1 python bench.py --workload search_n32_h4096 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>&1 | grep "ta search" | tail -2
