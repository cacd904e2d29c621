#!/bin/bash
# round 2, call 7: thread-per-pair u16 general-cost kernel (lev_diag16.cu): parity, timing, ncu; diagonal-extension v5
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_lev_diag16_large_batch or u16-thread-per-pair or test_lev_fr_long_strings or diagonal-extension or general-band-kernel" 2>&1 | tail -15 > gpurun_out/r02_c7_tests.txt
cat gpurun_out/r02_c7_tests.txt
run() { # name env...
  echo "== $*" >> gpurun_out/r02_c7_bench.txt
  env "${@:2}" timeout 300 python bench.py --workload $1 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'], d['roofline']['kernel'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c7_bench.txt
}
run affine_k16_len128 TA_DIAG16=0
run affine_k16_len128 TA_X=1
run lev_k8_len128 TA_FORCE_BAND=1
run lev_k16_len128 TA_FORCE_BAND=1
run rdamerau_k16_len512 TA_FORCE_BAND=1
run rdamerau_k16_len512 TA_FORCE_BAND=1 TA_DIAG16=0
run lev_k16_len4096 TA_FR=1
run exp_len1024 TA_FR=1
run exp_len1024 TA_FR=0
cat gpurun_out/r02_c7_bench.txt
ncu --set full --clock-control none --import-source on -k regex:'lev_diag16' -s 3 -c 1 -f -o gpurun_out/prof_diag16_affine_k16_len128 \
    python bench.py --workload affine_k16_len128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/prof_diag16_affine.log 2>&1
TA_FORCE_BAND=1 ncu --set full --clock-control none --import-source on -k regex:'lev_diag16' -s 3 -c 1 -f -o gpurun_out/prof_diag16_trans_rdamerau_k16_len512 \
    python bench.py --workload rdamerau_k16_len512 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/prof_diag16_trans.log 2>&1
TA_FR=1 ncu --set full --clock-control none --import-source on -k regex:'lev_fr' -s 3 -c 1 -f -o gpurun_out/prof_fr5_lev_k16_len4096 \
    python bench.py --workload lev_k16_len4096 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/prof_fr5.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
