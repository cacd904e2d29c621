#!/bin/bash
# round 2, call 12: lane-per-pair diagonal-extension kernel with slide queues; two-stage register count of the u16 kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_lev_fr_long_strings or diagonal-extension or test_lev_diag16_large_batch or u16-thread-per-pair" 2>&1 | tail -5 > gpurun_out/r02_c12_tests.txt
cat gpurun_out/r02_c12_tests.txt
run() { # name env...
  echo "== $*" >> gpurun_out/r02_c12_bench.txt
  env "${@:2}" timeout 300 python bench.py --workload $1 --steps 20 --warmup 3 --no-cpu-baseline --no-configs --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'], d['roofline']['kernel'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c12_bench.txt
}
run lev_k16_len4096 TA_X=1
run lev_k16_len4096 TA_FR_KERNEL=octet
run exp_len1024 TA_X=1
run exp_len1024 TA_FR_KERNEL=octet
run affine_k16_len128 TA_X=1
run affine_k16_len128 TA_DIAG16_STAGES=1
run lev_k16_len128 TA_FORCE_BAND=1
run rdamerau_k16_len512 TA_FORCE_BAND=1
cat gpurun_out/r02_c12_bench.txt
for wl in lev_k16_len4096; do
ncu --set full --clock-control none --import-source on -k regex:'lev_fr2' -s 3 -c 1 -f -o gpurun_out/prof_fr2q_${wl} \
    python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/prof_fr2q_${wl}.log 2>&1
done
