#!/bin/bash
# round 2, call 8: full GPU suite (long needles, unbounded-k traceback, diag16 v2), timings of the general kernel
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_c8_tests.txt
cat gpurun_out/r02_c8_tests.txt
run() { # name env...
  echo "== $*" >> gpurun_out/r02_c8_bench.txt
  env "${@:2}" timeout 300 python bench.py --workload $1 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'], d['roofline']['kernel'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c8_bench.txt
}
run affine_k16_len128 TA_X=1
run lev_k8_len128 TA_FORCE_BAND=1
run rdamerau_k16_len512 TA_FORCE_BAND=1
cat gpurun_out/r02_c8_bench.txt
ncu --set full --clock-control none --import-source on -k regex:'lev_diag16' -s 3 -c 1 -f -o gpurun_out/prof_diag16v2_affine_k16_len128 \
    python bench.py --workload affine_k16_len128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/prof_diag16_affine.log 2>&1
TA_FORCE_BAND=1 ncu --set full --clock-control none --import-source on -k regex:'lev_diag16' -s 3 -c 1 -f -o gpurun_out/prof_diag16v2_trans_rdamerau_k16_len512 \
    python bench.py --workload rdamerau_k16_len512 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/prof_diag16_trans.log 2>&1
